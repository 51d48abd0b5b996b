// oracle/oracle_capi.cpp — C entry points over oracle.hpp so that tests/ and bench.py can drive the oracle with ctypes.
// TEST INFRASTRUCTURE ONLY (see oracle.hpp header).  The builder functions deliberately have the same shape as the
// product's host-mirror C API (include/rtc.h, prefix rtc_) so one Python scene description drives both.
#include "oracle.hpp"

#include <chrono>
#include <mutex>

using namespace orc;

namespace {
thread_local std::string g_err;
int fail(const std::exception& e) { g_err = e.what(); return -1; }
Matrix4 m16(const double* m) {
    Matrix4 r;
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) r.v[i][j] = m[i * 4 + j];
    return r;
}
void out16(const Matrix4& m, double* o) {
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) o[i * 4 + j] = m.v[i][j];
}
}  // namespace

extern "C" {

// Plain-C mirror of material.rs:4-14 + pattern.rs:14-19 (same layout as rtc_material in include/rtc.h).
struct orc_material {
    double color[3];
    double ambient, diffuse, specular, shininess, reflective, transparency, refractive_index;
    int32_t pattern_kind;  // -1 none, 0 stripe, 1 gradient, 2 ring, 3 checkers, 4 test
    int32_t _pad;
    double pattern_a[3], pattern_b[3];
    double pattern_transform[16], pattern_inverse[16];
};
struct orc_counters {
    uint64_t primary, shadow, reflect, refract, leaf_tests;
    double seconds;
};

const char* orc_last_error() { return g_err.c_str(); }

// ---- matrix.rs / transformations.rs ------------------------------------------------------------------------------
void orc_translation(double x, double y, double z, double* o) { out16(translation(x, y, z), o); }
void orc_scaling(double x, double y, double z, double* o) { out16(scaling(x, y, z), o); }
void orc_rotation_x(double r, double* o) { out16(rotation_x(r), o); }
void orc_rotation_y(double r, double* o) { out16(rotation_y(r), o); }
void orc_rotation_z(double r, double* o) { out16(rotation_z(r), o); }
void orc_shearing(double xy, double xz, double yx, double yz, double zx, double zy, double* o) {
    out16(shearing(xy, xz, yx, yz, zx, zy), o);
}
int orc_view_transform(const double* from, const double* to, const double* up, double* o) {
    try {
        out16(view_transform(Tuple::point(from[0], from[1], from[2]), Tuple::point(to[0], to[1], to[2]),
                             Tuple::vector(up[0], up[1], up[2])), o);
        return 0;
    } catch (const std::exception& e) { return fail(e); }
}
void orc_matrix_mul(const double* a, const double* b, double* o) { out16(m16(a) * m16(b), o); }
void orc_matrix_transpose(const double* a, double* o) { out16(m16(a).transpose(), o); }
int orc_matrix_inverse(const double* a, double* o) {
    auto inv = m16(a).inverse();
    if (!inv) { g_err = "matrix is not invertible"; return -1; }
    out16(*inv, o);
    return 0;
}
void orc_matrix_mul_tuple(const double* a, const double* t, double* o) {
    Tuple r = m16(a) * Tuple{t[0], t[1], t[2], t[3]};
    o[0] = r.x; o[1] = r.y; o[2] = r.z; o[3] = r.w;
}

// ---- material / pattern -----------------------------------------------------------------------------------------
void orc_material_default(orc_material* m) {
    std::memset(m, 0, sizeof(*m));
    Material d;
    m->color[0] = d.color.red; m->color[1] = d.color.green; m->color[2] = d.color.blue;
    m->ambient = d.ambient; m->diffuse = d.diffuse; m->specular = d.specular; m->shininess = d.shininess;
    m->reflective = d.reflective; m->transparency = d.transparency; m->refractive_index = d.refractive_index;
    m->pattern_kind = -1;
    out16(Matrix4::identity(), m->pattern_transform);
    out16(Matrix4::identity(), m->pattern_inverse);
}
// pattern.rs:63-66
int orc_material_set_pattern_transform(orc_material* m, const double* t) {
    auto inv = m16(t).inverse();
    if (!inv) { g_err = "should be invertible"; return -1; }
    std::memcpy(m->pattern_transform, t, sizeof(double) * 16);
    out16(*inv, m->pattern_inverse);
    return 0;
}
static Material to_material(const orc_material* m) {
    Material r;
    r.color = {m->color[0], m->color[1], m->color[2]};
    r.ambient = m->ambient; r.diffuse = m->diffuse; r.specular = m->specular; r.shininess = m->shininess;
    r.reflective = m->reflective; r.transparency = m->transparency; r.refractive_index = m->refractive_index;
    if (m->pattern_kind >= 0) {
        Pattern p = Pattern::make((PatternKind)m->pattern_kind, {m->pattern_a[0], m->pattern_a[1], m->pattern_a[2]},
                                  {m->pattern_b[0], m->pattern_b[1], m->pattern_b[2]});
        p.transform = m16(m->pattern_transform);
        p.transform_inverse = m16(m->pattern_inverse);
        r.pattern = p;
    }
    return r;
}

// ---- shape.rs ---------------------------------------------------------------------------------------------------
// kind: 0 sphere, 1 plane, 2 cube, 3 cylinder, 4 cone, 5 group
Shape* orc_shape_new(int kind, double minimum, double maximum, int capped) {
    switch (kind) {
        case 0: return new Shape(Shape::sphere());
        case 1: return new Shape(Shape::plane());
        case 2: return new Shape(Shape::cube());
        case 3: return new Shape(Shape::cylinder(minimum, maximum, capped != 0));
        case 4: return new Shape(Shape::cone(minimum, maximum, capped != 0));
        case 5: return new Shape(Shape::group());
    }
    g_err = "unknown shape kind";
    return nullptr;
}
Shape* orc_shape_triangle(const double* p1, const double* p2, const double* p3) {
    try {
        return new Shape(Shape::triangle(Tuple::point(p1[0], p1[1], p1[2]), Tuple::point(p2[0], p2[1], p2[2]),
                                         Tuple::point(p3[0], p3[1], p3[2])));
    } catch (const std::exception& e) { fail(e); return nullptr; }
}
// the book's smooth_triangle (absent from the reference: intersection.rs:381-386, obj_file.rs:295-335 quote its scenarios)
Shape* orc_shape_smooth_triangle(const double* p1, const double* p2, const double* p3, const double* n1, const double* n2,
                                 const double* n3) {
    try {
        return new Shape(Shape::smooth_triangle(Tuple::point(p1[0], p1[1], p1[2]), Tuple::point(p2[0], p2[1], p2[2]),
                                                Tuple::point(p3[0], p3[1], p3[2]), Tuple::vector(n1[0], n1[1], n1[2]),
                                                Tuple::vector(n2[0], n2[1], n2[2]), Tuple::vector(n3[0], n3[1], n3[2])));
    } catch (const std::exception& e) { fail(e); return nullptr; }
}
void orc_shape_free(Shape* s) { delete s; }
int orc_shape_set_transform(Shape* s, const double* m) {
    try { s->set_transform(m16(m)); return 0; } catch (const std::exception& e) { return fail(e); }
}
int orc_shape_set_material(Shape* s, const orc_material* m) {
    try { s->set_material(to_material(m)); return 0; } catch (const std::exception& e) { return fail(e); }
}
// moves `child` into `group`; the child handle is consumed
int orc_shape_push_shape(Shape* group, Shape* child) {
    try {
        group->push_shape(std::move(*child));
        delete child;
        return 0;
    } catch (const std::exception& e) { return fail(e); }
}
static size_t count_leaves(const Shape& s) {
    if (s.kind != Kind::Group) return 1;
    size_t n = 0;
    for (auto& c : s.shapes) n += count_leaves(c);
    return n;
}
uint64_t orc_shape_leaf_count(const Shape* s) { return count_leaves(*s); }

// ---- obj_file.rs ------------------------------------------------------------------------------------------------
Shape* orc_obj_parse_file(const char* path, uint64_t* ignored_lines) {
    try {
        Parser p = Parser::from_obj_file(path);
        if (ignored_lines) *ignored_lines = p.ignored_lines;
        return new Shape(p.obj_to_group());
    } catch (const std::exception& e) { fail(e); return nullptr; }
}
Shape* orc_obj_parse_str(const char* text, uint64_t len, uint64_t* ignored_lines) {
    try {
        Parser p = Parser::from_obj_str(std::string(text, len));
        if (ignored_lines) *ignored_lines = p.ignored_lines;
        return new Shape(p.obj_to_group());
    } catch (const std::exception& e) { fail(e); return nullptr; }
}
// The group obj_to_group() would return for an OBJ holding these vertices and (1-based) triangular faces and no
// `g` line: group{ default_group{ triangles } }.
Shape* orc_mesh_from_arrays(const double* verts, uint64_t nverts, const int32_t* faces, uint64_t nfaces) {
    try {
        Shape def = Shape::group();
        for (uint64_t f = 0; f < nfaces; f++) {
            Tuple p[3];
            for (int k = 0; k < 3; k++) {
                int64_t i = faces[f * 3 + k];
                if (i < 1 || (uint64_t)i > nverts) throw Panic("index out of bounds");
                p[k] = Tuple::point(verts[(i - 1) * 3], verts[(i - 1) * 3 + 1], verts[(i - 1) * 3 + 2]);
            }
            def.push_shape(Shape::triangle(p[0], p[1], p[2]));
        }
        Shape g = Shape::group();
        g.push_shape(std::move(def));
        return new Shape(std::move(g));
    } catch (const std::exception& e) { fail(e); return nullptr; }
}

// The same for an OBJ that also holds `vn` records and `f v//n` faces: face f's corner k uses vertex faces[3f + k] and
// normal face_normals[3f + k] (both 1-based).
Shape* orc_smooth_mesh_from_arrays(const double* verts, uint64_t nverts, const double* normals, uint64_t nnormals,
                                   const int32_t* faces, const int32_t* face_normals, uint64_t nfaces) {
    try {
        Shape def = Shape::group();
        for (uint64_t f = 0; f < nfaces; f++) {
            Tuple p[3], n[3];
            for (int k = 0; k < 3; k++) {
                int64_t i = faces[f * 3 + k], j = face_normals[f * 3 + k];
                if (i < 1 || (uint64_t)i > nverts || j < 1 || (uint64_t)j > nnormals) throw Panic("index out of bounds");
                p[k] = Tuple::point(verts[(i - 1) * 3], verts[(i - 1) * 3 + 1], verts[(i - 1) * 3 + 2]);
                n[k] = Tuple::vector(normals[(j - 1) * 3], normals[(j - 1) * 3 + 1], normals[(j - 1) * 3 + 2]);
            }
            def.push_shape(Shape::smooth_triangle(p[0], p[1], p[2], n[0], n[1], n[2]));
        }
        Shape g = Shape::group();
        g.push_shape(std::move(def));
        return new Shape(std::move(g));
    } catch (const std::exception& e) { fail(e); return nullptr; }
}

// ---- world.rs ---------------------------------------------------------------------------------------------------
World* orc_world_new(const double* light_pos, const double* intensity) {
    World* w = new World();
    w->light = Light{Tuple::point(light_pos[0], light_pos[1], light_pos[2]), Color{intensity[0], intensity[1], intensity[2]}};
    return w;
}
World* orc_world_default() { return new World(World::default_world()); }
void orc_world_free(World* w) { delete w; }
// world.rs:11 with the constant edited (0 = the reference's 5)
int orc_world_set_recursion_limit(World* w, uint32_t limit) {
    w->recursion_limit = limit ? limit : RECURSION_LIMIT;
    return 0;
}
int orc_world_push(World* w, Shape* s) {
    w->objects.push_back(std::move(*s));
    delete s;
    return 0;
}
// World::color_at for a batch of rays (origin xyz, direction xyz per ray) -> rgb f64.  mode: 0 faithful, 1 cached.
int orc_world_color_at(World* w, int mode, const double* rays, uint64_t n, double* rgb) {
    try {
        if (mode == 1) w->build_cache();
        ctx().cached = (mode == 1);
        for (uint64_t i = 0; i < n; i++) {
            const double* r = rays + i * 6;
            Ray ray{Tuple::point(r[0], r[1], r[2]), Tuple::vector(r[3], r[4], r[5])};
            Color c = w->color_at(ray);
            rgb[i * 3] = c.red; rgb[i * 3 + 1] = c.green; rgb[i * 3 + 2] = c.blue;
        }
        ctx().cached = false;
        return 0;
    } catch (const std::exception& e) { ctx().cached = false; return fail(e); }
}

// ---- camera.rs / canvas.rs --------------------------------------------------------------------------------------
Camera* orc_camera_new(uint64_t hsize, uint64_t vsize, double fov) { return new Camera(hsize, vsize, fov); }
void orc_camera_free(Camera* c) { delete c; }
int orc_camera_set_transform(Camera* c, const double* m) {
    try { c->set_transform(m16(m)); return 0; } catch (const std::exception& e) { return fail(e); }
}
void orc_camera_params(const Camera* c, double* pixel_size, double* half_width, double* half_height, double* inv16) {
    *pixel_size = c->pixel_size; *half_width = c->half_width; *half_height = c->half_height;
    out16(c->transform_inverse, inv16);
}
void orc_camera_ray_for_pixel(const Camera* c, uint64_t px, uint64_t py, double* origin4, double* dir4) {
    Ray r = c->ray_for_pixel(px, py);
    origin4[0] = r.origin.x; origin4[1] = r.origin.y; origin4[2] = r.origin.z; origin4[3] = r.origin.w;
    dir4[0] = r.direction.x; dir4[1] = r.direction.y; dir4[2] = r.direction.z; dir4[3] = r.direction.w;
}

// Camera::render over the whole frame (pixel_xy == NULL) or over an explicit pixel list (x,y pairs; the camera keeps
// its full resolution).  out_rgb: 3 f64 per rendered pixel, in frame order / list order.
//   mode 0 = faithful (the reference's algorithm and cost), mode 1 = cached (same pixels, see oracle.hpp).
//   nthreads: 1 reproduces the reference's single thread; >1 splits pixels over std::threads (pixels are independent).
int orc_camera_render(const Camera* cam, World* world, int mode, int nthreads, const uint32_t* pixel_xy,
                      uint64_t npixels, double* out_rgb, orc_counters* counters) {
    try {
        if (mode == 1) world->build_cache();
        const uint64_t total = pixel_xy ? npixels : (uint64_t)cam->hsize * cam->vsize;
        if (nthreads < 1) nthreads = 1;
        std::atomic<uint64_t> next{0};
        std::vector<Counters> cnt(nthreads);
        std::mutex err_mu;
        std::string err;
        const uint64_t chunk = 64;
        auto t0 = std::chrono::steady_clock::now();
        auto worker = [&](int tid) {
            ctx().cached = (mode == 1);
            ctx().counters = &cnt[tid];
            try {
                for (;;) {
                    uint64_t b = next.fetch_add(chunk);
                    if (b >= total) break;
                    uint64_t e = std::min(total, b + chunk);
                    for (uint64_t i = b; i < e; i++) {
                        uint64_t x, y;
                        if (pixel_xy) { x = pixel_xy[2 * i]; y = pixel_xy[2 * i + 1]; }
                        else { x = i % cam->hsize; y = i / cam->hsize; }
                        Ray ray = cam->ray_for_pixel(x, y);
                        cnt[tid].primary++;
                        Color c = world->color_at(ray);
                        out_rgb[3 * i] = c.red; out_rgb[3 * i + 1] = c.green; out_rgb[3 * i + 2] = c.blue;
                    }
                }
            } catch (const std::exception& e) {
                std::lock_guard<std::mutex> lk(err_mu);
                err = e.what();
                next.store(total);
            }
            ctx().cached = false;
            ctx().counters = nullptr;
        };
        if (nthreads == 1) {
            worker(0);
        } else {
            std::vector<std::thread> th;
            for (int t = 0; t < nthreads; t++) th.emplace_back(worker, t);
            for (auto& t : th) t.join();
        }
        auto t1 = std::chrono::steady_clock::now();
        if (!err.empty()) { g_err = err; return -1; }
        if (counters) {
            std::memset(counters, 0, sizeof(*counters));
            for (auto& c : cnt) {
                counters->primary += c.primary; counters->shadow += c.shadow; counters->reflect += c.reflect;
                counters->refract += c.refract; counters->leaf_tests += c.leaf_tests;
            }
            counters->seconds = std::chrono::duration<double>(t1 - t0).count();
        }
        return 0;
    } catch (const std::exception& e) { return fail(e); }
}

// canvas.rs:61-63 over an rgb f64 buffer -> rgba8 (a = 255)
void orc_quantise(const double* rgb, uint64_t npixels, uint8_t* rgba8) {
    for (uint64_t i = 0; i < npixels; i++) {
        rgba8[4 * i] = (uint8_t)Canvas::quantise(rgb[3 * i]);
        rgba8[4 * i + 1] = (uint8_t)Canvas::quantise(rgb[3 * i + 1]);
        rgba8[4 * i + 2] = (uint8_t)Canvas::quantise(rgb[3 * i + 2]);
        rgba8[4 * i + 3] = 255;
    }
}
// canvas.rs:28-58 over an rgb f64 buffer; returns a malloc'd buffer the caller frees with orc_free
char* orc_to_ppm(const double* rgb, uint64_t width, uint64_t height, uint64_t* len) {
    Canvas c(width, height);
    for (uint64_t i = 0; i < width * height; i++) c.pixels[i] = Color{rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2]};
    std::string s = c.to_ppm();
    char* out = (char*)std::malloc(s.size() + 1);
    std::memcpy(out, s.data(), s.size());
    out[s.size()] = 0;
    *len = s.size();
    return out;
}
void orc_free(void* p) { std::free(p); }

}  // extern "C"
