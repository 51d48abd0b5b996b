// host_model.hpp — the reference's host-side scene API (Shape / World / Camera / Canvas / Parser) restated in C++.
//
// In the reference this layer is Rust and stays Rust (BASELINE north_star); no Rust toolchain exists in this image, so
// the same construction semantics live here behind the `rtc_shape_* / rtc_world_* / rtc_camera_* / rtc_canvas_* /
// rtc_obj_*` entry points of include/rtc.h.  Nothing here runs per pixel: it builds the World, marshals it into the
// rtc_scene_desc the CORE boundary takes (exactly what the patched Rust Camera::render would send) and formats PPMs.
// Compiled with -ffp-contract=off: the matrices computed here reach pixels.
#pragma once
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "../../include/rtc.h"
#include "host_math.hpp"

namespace rtc {

// a reference panic!/expect()/assert! -> RTC_ERR_PANIC with the reference's message
struct HostPanic : std::runtime_error {
    using std::runtime_error::runtime_error;
};

// ---- Shape (shape.rs:42-49) -----------------------------------------------------------------------------------------
struct HShape {
    int kind = RTC_SPHERE;
    double minimum = 0., maximum = 0.;
    bool capped = false;
    std::vector<std::unique_ptr<HShape>> children;  // ShapeKind::Group { shapes }
    rtc_triangle_desc tri{};                        // ShapeKind::Triangle payload
    rtc_vertex_normals vn{};                        // RTC_SMOOTH_TRIANGLE: n1, n2, n3
    Mat4 transform = Mat4::identity();
    Mat4 inverse = Mat4::identity();
    bool transformed = false;
    rtc_material material{};
};

inline void material_default(rtc_material* m) {  // material.rs:17-29
    std::memset(m, 0, sizeof(*m));
    m->color[0] = m->color[1] = m->color[2] = 1.0;
    m->ambient = 0.1;
    m->diffuse = 0.9;
    m->specular = 0.9;
    m->shininess = 200.0;
    m->reflective = 0.0;
    m->transparency = 0.0;
    m->refractive_index = 1.0;
    m->pattern_kind = RTC_PATTERN_NONE;
    Mat4 id = Mat4::identity();
    std::memcpy(m->pattern_transform, id.m, sizeof(id.m));
    std::memcpy(m->pattern_inverse, id.m, sizeof(id.m));
}

inline std::unique_ptr<HShape> shape_new(int kind, double minimum, double maximum, bool capped) {
    auto s = std::make_unique<HShape>();
    s->kind = kind;
    if (kind == RTC_CYLINDER || kind == RTC_CONE) {
        s->minimum = minimum;
        s->maximum = maximum;
        s->capped = capped;
    }
    material_default(&s->material);
    return s;
}

// Shape::triangle (shape.rs:171-193): e1 = p2 - p1, e2 = p3 - p1, normal = normalize(e2 x e1)
inline std::unique_ptr<HShape> shape_triangle(const double* p1, const double* p2, const double* p3) {
    auto s = shape_new(RTC_TRIANGLE, 0., 0., false);
    Vec4 a = point(p1[0], p1[1], p1[2]), b = point(p2[0], p2[1], p2[2]), c = point(p3[0], p3[1], p3[2]);
    Vec4 e1 = sub(b, a), e2 = sub(c, a);
    Vec4 n = normalize(cross(e2, e1));
    rtc_triangle_desc& t = s->tri;
    t.p1[0] = a.x; t.p1[1] = a.y; t.p1[2] = a.z;
    t.p2[0] = b.x; t.p2[1] = b.y; t.p2[2] = b.z;
    t.p3[0] = c.x; t.p3[1] = c.y; t.p3[2] = c.z;
    t.e1[0] = e1.x; t.e1[1] = e1.y; t.e1[2] = e1.z;
    t.e2[0] = e2.x; t.e2[1] = e2.y; t.e2[2] = e2.z;
    t.normal[0] = n.x; t.normal[1] = n.y; t.normal[2] = n.z;
    return s;
}

// The book's smooth_triangle(p1, p2, p3, n1, n2, n3) — not in the reference (rtc.h RTC_SMOOTH_TRIANGLE): a triangle that
// also keeps its three vertex normals.
inline std::unique_ptr<HShape> shape_smooth_triangle(const double* p1, const double* p2, const double* p3, const double* n1,
                                                     const double* n2, const double* n3) {
    auto s = shape_triangle(p1, p2, p3);
    s->kind = RTC_SMOOTH_TRIANGLE;
    std::memcpy(s->vn.n1, n1, sizeof(s->vn.n1));
    std::memcpy(s->vn.n2, n2, sizeof(s->vn.n2));
    std::memcpy(s->vn.n3, n3, sizeof(s->vn.n3));
    return s;
}

// set_transform_internal (shape.rs:203-218): groups push the matrix down; a leaf left-multiplies it into its own
inline void push_down_transform(HShape* s, const Mat4& t) {
    if (s->kind == RTC_GROUP) {
        for (auto& c : s->children) push_down_transform(c.get(), t);
        return;
    }
    s->transform = mul(t, s->transform);
    if (!inverse(s->transform, &s->inverse)) throw HostPanic("should be invertible (src/shape.rs:215)");
}
inline void shape_set_transform(HShape* s, const Mat4& t) {  // shape.rs:196-201
    if (s->transformed) throw HostPanic("Can't call set_transform more than once. (src/shape.rs:200)");
    s->transformed = true;
    push_down_transform(s, t);
}
inline void shape_set_material(HShape* s, const rtc_material& m) {  // shape.rs:220-229
    if (s->kind == RTC_GROUP) {
        for (auto& c : s->children) shape_set_material(c.get(), m);
        return;
    }
    s->material = m;
}
inline void shape_push(HShape* group, std::unique_ptr<HShape> child) {  // shape.rs:528-535
    if (group->kind != RTC_GROUP) throw HostPanic("push_shape was called on something that isn't a group (src/shape.rs:533)");
    group->children.push_back(std::move(child));
}
inline uint64_t shape_leaf_count(const HShape* s) {
    if (s->kind != RTC_GROUP) return 1;
    uint64_t n = 0;
    for (auto& c : s->children) n += shape_leaf_count(c.get());
    return n;
}

// ---- Parser (obj_file.rs:5-128) -------------------------------------------------------------------------------------
struct ObjResult {
    std::unique_ptr<HShape> group;  // Parser::obj_to_group()
    uint64_t ignored_lines = 0;
    uint64_t vertex_count = 0;
    uint64_t normal_count = 0;  // `vn` records (obj_file.rs:295-310, commented scenario)
};

namespace detail {
inline bool is_space(char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\f' || c == '\v'; }
// [+-]digits[.digits], at most 15 digits in all, no exponent: the digit string as an integer and the power of ten are both
// exact doubles, so one correctly rounded division gives the correctly rounded value (Clinger's fast path) — what
// str::parse::<f64> and strtod return, without strtod's cost.  Anything else: false, the caller takes the general path.
inline bool parse_f64_fast(const char* b, const char* e, double* out) {
    static const double p10[16] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15};
    const char* p = b;
    bool neg = false;
    if (p < e && (*p == '-' || *p == '+')) neg = *p++ == '-';
    uint64_t m = 0;
    int digits = 0, frac = 0;
    bool dot = false;
    for (; p < e; p++) {
        if (*p >= '0' && *p <= '9') {
            if (++digits > 15) return false;
            m = m * 10 + (uint64_t)(*p - '0');
            if (dot) frac++;
        } else if (*p == '.' && !dot) {
            dot = true;
        } else {
            return false;
        }
    }
    if (digits == 0) return false;
    const double v = frac ? (double)m / p10[frac] : (double)m;
    *out = neg ? -v : v;
    return true;
}
// str::parse::<f64>: decimal/exponent forms, inf/infinity/nan; no hex floats, no trailing junk.  (lb, le) is the line,
// quoted in the panic message only.
inline double parse_f64(const char* b, const char* e, const char* what, const char* lb, const char* le) {
    double fast;
    if (parse_f64_fast(b, e, &fast)) return fast;
    std::string tok(b, e);
    bool bad = tok.empty();
    for (char c : tok)
        if (c == 'x' || c == 'X' || c == 'p' || c == 'P' || c == '(') bad = true;
    char* end = nullptr;
    double v = bad ? 0. : std::strtod(tok.c_str(), &end);
    if (bad || end != tok.c_str() + tok.size())
        throw HostPanic(std::string("vertex ") + what + " should be an f64 in \"" + std::string(lb, le) + "\" (src/obj_file.rs:42-55)");
    return v;
}
// str::parse::<usize>: optional '+', then decimal digits only
inline uint64_t parse_usize(const char* b, const char* e, const char* what, const char* lb, const char* le) {
    const char* p = b;
    if (p < e && *p == '+') p++;
    bool ok = p < e;
    uint64_t v = 0;
    for (; p < e; p++) {
        if (*p < '0' || *p > '9') { ok = false; break; }
        v = v * 10 + (uint64_t)(*p - '0');
    }
    if (!ok) throw HostPanic(std::string("face ") + what + " should be a usize in \"" + std::string(lb, le) + "\" (src/obj_file.rs:57-74)");
    return v;
}
// One corner of a face: `v`, `v/t`, `v//n` or `v/t/n` (obj_file.rs:312-335, commented scenario; the texture index is
// skipped unread).  A plain `v` parses exactly as the reference does.  *n = 0: the corner names no normal.
inline uint64_t parse_corner(const char* b, const char* e, const char* what, const char* lb, const char* le, uint64_t* n) {
    *n = 0;
    const char* s1 = b;
    while (s1 < e && *s1 != '/') s1++;
    const uint64_t v = parse_usize(b, s1, what, lb, le);
    if (s1 == e) return v;
    const char* s2 = s1 + 1;
    while (s2 < e && *s2 != '/') s2++;
    if (s2 < e) *n = parse_usize(s2 + 1, e, what, lb, le);
    return v;
}
}  // namespace detail

inline ObjResult obj_parse(const char* text, size_t len) {
    ObjResult out;
    std::vector<double> verts;  // xyz
    std::vector<double> norms;  // xyz of the `vn` records
    auto default_group = shape_new(RTC_GROUP, 0., 0., false);
    // named groups in first-insertion order (the reference iterates a HashMap: any order); re-declaring a name
    // replaces the group (HashMap::insert, obj_file.rs:101-103)
    std::vector<std::pair<std::string, std::unique_ptr<HShape>>> named;
    HShape* current = nullptr;
    auto vertex = [&](uint64_t one_based, double* p) {
        if (one_based == 0 || one_based > verts.size() / 3) throw HostPanic("index out of bounds (src/obj_file.rs:117)");
        std::memcpy(p, &verts[(one_based - 1) * 3], sizeof(double) * 3);
    };
    auto normal = [&](uint64_t one_based, double* p) {
        if (one_based == 0 || one_based > norms.size() / 3) throw HostPanic("index out of bounds (vn)");
        std::memcpy(p, &norms[(one_based - 1) * 3], sizeof(double) * 3);
    };
    size_t pos = 0;
    std::vector<std::pair<const char*, const char*>> tok;
    while (pos < len) {
        size_t nl = pos;
        while (nl < len && text[nl] != '\n') nl++;
        const char* lb = text + pos;
        const char* le = text + nl;
        pos = nl + 1;
        tok.clear();
        for (const char* p = lb; p < le;) {
            while (p < le && detail::is_space(*p)) p++;
            const char* s = p;
            while (p < le && !detail::is_space(*p)) p++;
            if (p > s) tok.emplace_back(s, p);
        }
        if (tok.empty()) continue;
        const size_t tl = (size_t)(tok[0].second - tok[0].first);
        auto line = [&]() { return std::string(lb, le); };
        if (tl == 1 && *tok[0].first == 'v') {
            static const char* names[3] = {"x", "y", "z"};
            double xyz[3];
            for (int k = 0; k < 3; k++) {
                if (tok.size() < (size_t)k + 2)
                    throw HostPanic(std::string("vertex token to have a ") + names[k] + " in \"" + line() + "\"");
                xyz[k] = detail::parse_f64(tok[k + 1].first, tok[k + 1].second, names[k], lb, le);
            }
            verts.insert(verts.end(), xyz, xyz + 3);
        } else if (tl == 1 && *tok[0].first == 'f') {
            if (tok.size() < 2) throw HostPanic("face should have a v1 in \"" + line() + "\"");
            uint64_t n1i, n2i, n3i;
            uint64_t v1 = detail::parse_corner(tok[1].first, tok[1].second, "v1", lb, le, &n1i);
            if (tok.size() < 3) throw HostPanic("face should have a v2 in \"" + line() + "\"");
            uint64_t v2 = detail::parse_corner(tok[2].first, tok[2].second, "v2", lb, le, &n2i);
            for (size_t k = 3; k < tok.size(); k++) {  // fan triangulation, obj_file.rs:70-94
                uint64_t v3 = detail::parse_corner(tok[k].first, tok[k].second, "v3", lb, le, &n3i);
                double p1[3], p2[3], p3[3];
                vertex(v1, p1);
                vertex(v2, p2);
                vertex(v3, p3);
                if (n1i && n2i && n3i) {  // every corner names a normal: a smooth triangle
                    double a[3], b[3], c[3];
                    normal(n1i, a);
                    normal(n2i, b);
                    normal(n3i, c);
                    shape_push(current ? current : default_group.get(), shape_smooth_triangle(p1, p2, p3, a, b, c));
                } else {
                    shape_push(current ? current : default_group.get(), shape_triangle(p1, p2, p3));
                }
                v2 = v3;
                n2i = n3i;
            }
        } else if (tl == 2 && tok[0].first[0] == 'v' && tok[0].first[1] == 'n') {
            static const char* names[3] = {"x", "y", "z"};
            double xyz[3];
            for (int k = 0; k < 3; k++) {
                if (tok.size() < (size_t)k + 2)
                    throw HostPanic(std::string("normal token to have a ") + names[k] + " in \"" + line() + "\"");
                xyz[k] = detail::parse_f64(tok[k + 1].first, tok[k + 1].second, names[k], lb, le);
            }
            norms.insert(norms.end(), xyz, xyz + 3);
        } else if (tl == 1 && *tok[0].first == 'g') {
            if (tok.size() < 2) throw HostPanic("group should have a name in \"" + line() + "\"");
            std::string name(tok[1].first, tok[1].second);
            current = nullptr;
            for (auto& g : named)
                if (g.first == name) {
                    g.second = shape_new(RTC_GROUP, 0., 0., false);
                    current = g.second.get();
                }
            if (!current) {
                named.emplace_back(name, shape_new(RTC_GROUP, 0., 0., false));
                current = named.back().second.get();
            }
        } else {
            out.ignored_lines++;
        }
    }
    out.vertex_count = verts.size() / 3;
    out.normal_count = norms.size() / 3;
    out.group = shape_new(RTC_GROUP, 0., 0., false);  // obj_to_group, obj_file.rs:120-128
    shape_push(out.group.get(), std::move(default_group));
    for (auto& g : named) shape_push(out.group.get(), std::move(g.second));
    return out;
}

// ---- World -> rtc_scene_desc ----------------------------------------------------------------------------------------
// What the Rust-side Camera::render patch does with its &World: a pre-order walk that copies each
// Shape's kind, its transform with the cached inverse, its material and triangle payload into flat arrays.
struct Marshalled {
    std::vector<rtc_shape_desc> shapes;
    std::vector<rtc_transform_desc> transforms;
    std::vector<rtc_material> materials;
    std::vector<rtc_triangle_desc> triangles;
    std::vector<rtc_vertex_normals> vertex_normals;  // empty unless the world holds a smooth triangle; else like triangles
    rtc_scene_desc desc{};
};

namespace detail {
struct Bytes {
    std::string b;
    bool operator<(const Bytes& o) const { return b < o.b; }
};
struct Marshaller {
    Marshalled& m;
    std::map<Bytes, int32_t> xf, mat;
    // the triangles of a mesh share one transform and one material: remember the last entry and compare bytes before
    // building a map key (a 256-byte string per shape otherwise — most of the marshalling time of a 10 k-triangle world)
    int32_t last_xf = -1, last_mat = -1;
    int32_t transform_id(const HShape* s) {
        rtc_transform_desc t;
        std::memcpy(t.transform, s->transform.m, sizeof(t.transform));
        std::memcpy(t.inverse, s->inverse.m, sizeof(t.inverse));
        if (last_xf >= 0 && std::memcmp(&m.transforms[last_xf], &t, sizeof(t)) == 0) return last_xf;
        Bytes k{std::string((const char*)&t, sizeof(t))};
        auto it = xf.find(k);
        if (it != xf.end()) return last_xf = it->second;
        int32_t id = (int32_t)m.transforms.size();
        m.transforms.push_back(t);
        xf.emplace(std::move(k), id);
        return last_xf = id;
    }
    int32_t material_id(const rtc_material& mm) {
        if (last_mat >= 0 && std::memcmp(&m.materials[last_mat], &mm, sizeof(mm)) == 0) return last_mat;
        Bytes k{std::string((const char*)&mm, sizeof(mm))};
        auto it = mat.find(k);
        if (it != mat.end()) return last_mat = it->second;
        int32_t id = (int32_t)m.materials.size();
        m.materials.push_back(mm);
        mat.emplace(std::move(k), id);
        return last_mat = id;
    }
    void walk(const HShape* s) {
        rtc_shape_desc d;
        std::memset(&d, 0, sizeof(d));
        d.kind = s->kind;
        d.transform = transform_id(s);
        d.capped = s->capped ? 1 : 0;
        d.minimum = s->minimum;
        d.maximum = s->maximum;
        d.material = -1;
        d.triangle = -1;
        if (s->kind == RTC_GROUP) {
            d.child_count = (int32_t)s->children.size();
            m.shapes.push_back(d);
            for (auto& c : s->children) walk(c.get());
            return;
        }
        d.material = material_id(s->material);
        if (s->kind == RTC_TRIANGLE || s->kind == RTC_SMOOTH_TRIANGLE) {
            d.triangle = (int32_t)m.triangles.size();
            m.triangles.push_back(s->tri);
            if (s->kind == RTC_SMOOTH_TRIANGLE) {
                m.vertex_normals.resize(m.triangles.size(), rtc_vertex_normals{});
                m.vertex_normals.back() = s->vn;
            }
        }
        m.shapes.push_back(d);
    }
};
}  // namespace detail

struct HWorld {  // world.rs:13-16
    std::vector<std::unique_ptr<HShape>> objects;
    double light_position[3] = {0, 0, 0};
    double light_intensity[3] = {1, 1, 1};
};

inline void marshal_world(const HWorld& w, Marshalled& out) {
    detail::Marshaller mm{out, {}, {}};
    size_t leaves = 0;
    for (auto& o : w.objects) leaves += shape_leaf_count(o.get());
    out.shapes.reserve(leaves + 16);
    out.triangles.reserve(leaves);
    for (auto& o : w.objects) mm.walk(o.get());
    rtc_scene_desc& d = out.desc;
    d.shapes = out.shapes.data();
    d.shape_count = (uint32_t)out.shapes.size();
    d.root_count = (uint32_t)w.objects.size();
    d.transforms = out.transforms.data();
    d.transform_count = (uint32_t)out.transforms.size();
    d.materials = out.materials.data();
    d.material_count = (uint32_t)out.materials.size();
    d.triangles = out.triangles.data();
    d.triangle_count = (uint32_t)out.triangles.size();
    if (!out.vertex_normals.empty()) out.vertex_normals.resize(out.triangles.size(), rtc_vertex_normals{});
    d.vertex_normals = out.vertex_normals.empty() ? nullptr : out.vertex_normals.data();
    for (int k = 0; k < 3; k++) {
        d.light_position[k] = w.light_position[k];
        d.light_intensity[k] = w.light_intensity[k];
    }
}

// World::default_world (world.rs:26-41)
inline std::unique_ptr<HWorld> world_default() {
    auto w = std::make_unique<HWorld>();
    w->light_position[0] = -10.0; w->light_position[1] = 10.0; w->light_position[2] = -10.0;
    auto s1 = shape_new(RTC_SPHERE, 0., 0., false);
    s1->material.color[0] = 0.8; s1->material.color[1] = 1.0; s1->material.color[2] = 0.6;
    s1->material.diffuse = 0.7;
    s1->material.specular = 0.2;
    auto s2 = shape_new(RTC_SPHERE, 0., 0., false);
    shape_set_transform(s2.get(), scaling(0.5, 0.5, 0.5));
    w->objects.push_back(std::move(s1));
    w->objects.push_back(std::move(s2));
    return w;
}

// ---- Camera (camera.rs:5-46) ----------------------------------------------------------------------------------------
struct HCamera {
    uint64_t hsize = 0, vsize = 0;
    double field_of_view = 0.;
    Mat4 transform = Mat4::identity(), inverse = Mat4::identity();
    double half_width = 0., half_height = 0., pixel_size = 0.;
};
inline std::unique_ptr<HCamera> camera_new(uint64_t hsize, uint64_t vsize, double fov) {  // camera.rs:16-41
    auto c = std::make_unique<HCamera>();
    c->hsize = hsize;
    c->vsize = vsize;
    c->field_of_view = fov;
    double half_view = std::tan(fov / 2.0);
    double aspect = (double)hsize / (double)vsize;
    if (aspect >= 1.0) {
        c->half_width = half_view;
        c->half_height = half_view / aspect;
    } else {
        c->half_width = half_view * aspect;
        c->half_height = half_view;
    }
    c->pixel_size = (c->half_width * 2.0) / (double)hsize;
    return c;
}
inline void camera_set_transform(HCamera* c, const Mat4& t) {  // camera.rs:43-46
    Mat4 inv;
    if (!inverse(t, &inv)) throw HostPanic("should be invertible (src/camera.rs:45)");
    c->transform = t;
    c->inverse = inv;
}

// ---- Canvas (canvas.rs:5-63) ----------------------------------------------------------------------------------------
// canvas.rs:61-63: (c.clamp(0., 1.) * 255.).round() as i32
inline uint8_t quantise_channel(double c) {
    double k = c;
    if (k < 0.) k = 0.;
    else if (k > 1.) k = 1.;
    double r = std::round(k * 255.);
    if (!(r == r)) return 0;
    return (uint8_t)(int)r;
}

// Canvas::to_ppm (canvas.rs:28-58) over quantised RGBA8 rows: "P3\nW H\n255\n", then per row the 3*W channel values
// separated by single spaces with a newline before any token that would push the line past 70 columns, and a newline at
// row end.  Rows are independent (the column counter restarts per row), so the rows are encoded in parallel: every
// thread formats a contiguous block of rows into its own buffer and the blocks are concatenated in order — the bytes are
// the ones the reference's 3*W*H sequential write!() calls produce (at 8K that is 10^8 of them on an unbuffered File).
inline void ppm_rows(const uint8_t* rgba, uint64_t width, uint64_t y0, uint64_t y1, std::string& out) {
    static const struct Digits {
        char d[256][4];
        uint8_t n[256];
        Digits() {
            for (int v = 0; v < 256; v++) n[v] = (uint8_t)std::snprintf(d[v], 4, "%d", v);
        }
    } kDigits;
    out.clear();
    out.reserve((size_t)width * (y1 - y0) * 12 + 16);
    for (uint64_t y = y0; y < y1; y++) {
        size_t len = 0;
        const uint8_t* row = rgba + y * width * 4;
        for (uint64_t x = 0; x < width; x++)
            for (int ch = 0; ch < 3; ch++) {
                const uint8_t v = row[x * 4 + ch];
                const size_t n = kDigits.n[v];
                if (len + n + 1 > 70) {
                    out.push_back('\n');
                    len = 0;
                }
                if (len > 0) {
                    out.push_back(' ');
                    len += 1;
                }
                out.append(kDigits.d[v], n);
                len += n;
            }
        out.push_back('\n');
    }
}

inline std::string ppm_from_rgba8(const uint8_t* rgba, uint64_t width, uint64_t height) {
    std::string head = "P3\n" + std::to_string(width) + " " + std::to_string(height) + "\n255\n";
    unsigned nt = std::thread::hardware_concurrency();
    if (nt == 0) nt = 1;
    if (nt > 32) nt = 32;
    if ((uint64_t)width * height < (1u << 16) || height < 2 * (uint64_t)nt) nt = 1;
    std::vector<std::string> part(nt);
    auto work = [&](unsigned k) {
        ppm_rows(rgba, width, height * k / nt, height * (k + 1) / nt, part[k]);
    };
    if (nt == 1) {
        work(0);
    } else {
        std::vector<std::thread> th;
        for (unsigned k = 1; k < nt; k++) th.emplace_back(work, k);
        work(0);
        for (auto& t : th) t.join();
    }
    size_t total = head.size();
    for (auto& p : part) total += p.size();
    std::string out;
    out.reserve(total);
    out += head;
    for (auto& p : part) out += p;
    return out;
}

}  // namespace rtc
