/* examples/rtc_main.c — the reference's `main.rs` restated in C over the C ABI of include/rtc.h.
 *
 *   usage: rtc_main <filename.ppm> [width-in-px] [--scene hexagon|table|cow|teapot] [--objs DIR]
 *
 * Same command line as the reference binary (main.rs:39-81: positional file name, optional width, the frame is
 * width x width/2, fov 0.785); main.rs hard-codes the cow scene at main.rs:80 and keeps the other three builders as dead
 * code — here --scene selects one.  The scene builders are line-for-line what main.rs:84-397 builds, through the host
 * mirror (rtc_shape_*, rtc_world_*, rtc_camera_*): this is the compiled-language host a Rust maintainer would write
 * with the rtc-sys binding of INTEGRATION.md, in the one compiled language this image has a toolchain for.
 *   build: gcc -O2 -Iinclude examples/rtc_main.c -Lray-tracer-challenge-rust_b200 -lrtc_b200 -lm -o rtc_main
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "rtc.h"

static void die(const char* what) {
    fprintf(stderr, "%s: %s\n", what, rtc_last_error());
    exit(1);
}
#define OK(call) do { if ((call) != RTC_OK) die(#call); } while (0)

typedef struct { double m[16]; } M4;
static M4 mul(M4 a, M4 b) { M4 r; rtc_matrix_mul(a.m, b.m, r.m); return r; }
static M4 translation(double x, double y, double z) { M4 r; rtc_translation(x, y, z, r.m); return r; }
static M4 scaling(double x, double y, double z) { M4 r; rtc_scaling(x, y, z, r.m); return r; }
static M4 rotation_y(double a) { M4 r; rtc_rotation_y(a, r.m); return r; }
static M4 rotation_z(double a) { M4 r; rtc_rotation_z(a, r.m); return r; }

static rtc_camera* camera(unsigned width, double fx, double fy, double fz, double tx, double ty, double tz) {
    rtc_camera* c = rtc_camera_new(width, width / 2, 0.785);
    double from[3] = {fx, fy, fz}, to[3] = {tx, ty, tz}, up[3] = {0.0, 1.0, 0.0};
    M4 v;
    OK(rtc_view_transform(from, to, up, v.m));
    OK(rtc_camera_set_transform(c, v.m));
    return c;
}
static rtc_world* world(void) {
    double pos[3] = {0.0, 6.9, -5.0}, intensity[3] = {1.0, 1.0, 0.9};
    return rtc_world_new(pos, intensity);
}

/* main.rs:84-146 */
static void hexagon_scene(unsigned width, rtc_world** w, rtc_camera** c) {
    const double PI = 3.14159265358979323846;
    *c = camera(width, 8.0, 6.0, -8.0, 0.0, 0.0, 0.0);
    *w = world();
    rtc_shape* hex = rtc_shape_new(RTC_GROUP, 0, 0, 0);
    for (int i = 0; i < 6; i++) {
        rtc_shape* side = rtc_shape_new(RTC_GROUP, 0, 0, 0);
        rtc_shape* corner = rtc_shape_new(RTC_SPHERE, 0, 0, 0);
        OK(rtc_shape_set_transform(corner, mul(translation(0., 0., -1.), scaling(0.25, 0.25, 0.25)).m));
        rtc_shape* edge = rtc_shape_new(RTC_CYLINDER, 0., 1., 1);
        OK(rtc_shape_set_transform(edge, mul(mul(mul(translation(0., 0., -1.), rotation_y(-PI / 6.)), rotation_z(-PI / 2.)),
                                             scaling(0.25, 1., 0.25)).m));
        OK(rtc_shape_push_shape(side, corner));
        OK(rtc_shape_push_shape(side, edge));
        OK(rtc_shape_set_transform(side, rotation_y((double)i * PI / 3.).m));
        OK(rtc_shape_push_shape(hex, side));
    }
    OK(rtc_shape_set_transform(hex, scaling(2.5, 2.5, 2.5).m));
    OK(rtc_world_push(*w, hex));
}

static void cube(rtc_world* w, M4 t, const rtc_material* m) {
    rtc_shape* s = rtc_shape_new(RTC_CUBE, 0, 0, 0);
    OK(rtc_shape_set_transform(s, t.m));
    OK(rtc_shape_set_material(s, m));
    OK(rtc_world_push(w, s));
}
static rtc_material mat(double r, double g, double b) {
    rtc_material m;
    rtc_material_default(&m);
    m.color[0] = r; m.color[1] = g; m.color[2] = b;
    return m;
}
static void pattern(rtc_material* m, int kind, double ar, double ag, double ab, double br, double bg, double bb, M4 t) {
    m->pattern_kind = kind;
    m->pattern_a[0] = ar; m->pattern_a[1] = ag; m->pattern_a[2] = ab;
    m->pattern_b[0] = br; m->pattern_b[1] = bg; m->pattern_b[2] = bb;
    OK(rtc_material_set_pattern_transform(m, t.m));
}

/* main.rs:151-323 */
static void table_scene(unsigned width, rtc_world** wp, rtc_camera** c) {
    *c = camera(width, 8.0, 6.0, -8.0, 0.0, 3.0, 0.0);
    rtc_world* w = *wp = world();
    rtc_material m = mat(1, 1, 1);
    pattern(&m, RTC_PATTERN_CHECKERS, 0, 0, 0, 0.25, 0.25, 0.25, scaling(0.07, 0.07, 0.07));
    m.ambient = 0.25; m.diffuse = 0.7; m.specular = 0.9; m.shininess = 300.0; m.reflective = 0.1;
    cube(w, mul(scaling(20.0, 7.0, 20.0), translation(0.0, 1.0, 0.1)), &m);
    m = mat(1, 1, 1);
    pattern(&m, RTC_PATTERN_CHECKERS, 0.4863, 0.3765, 0.2941, 0.3725, 0.2902, 0.2275, scaling(0.05, 20.0, 0.05));
    m.ambient = 0.1; m.diffuse = 0.7; m.specular = 0.9; m.shininess = 300.0; m.reflective = 0.1;
    cube(w, scaling(10.0, 10.0, 10.0), &m);
    m = mat(1, 1, 1);
    pattern(&m, RTC_PATTERN_STRIPE, 0.5529, 0.4235, 0.3255, 0.6588, 0.5098, 0.4000,
            mul(scaling(0.05, 0.05, 0.05), rotation_y(0.1)));
    m.ambient = 0.1; m.diffuse = 0.7; m.specular = 0.9; m.shininess = 300.0; m.reflective = 0.2;
    cube(w, mul(translation(0.0, 3.1, 0.0), scaling(3.0, 0.1, 2.0)), &m);
    const double legs[4][2] = {{2.7, -1.7}, {2.7, 1.7}, {-2.7, -1.7}, {-2.7, 1.7}};
    for (int i = 0; i < 4; i++) {
        m = mat(0.5529, 0.4235, 0.3255);
        m.ambient = 0.2; m.diffuse = 0.7;
        cube(w, mul(translation(legs[i][0], 1.5, legs[i][1]), scaling(0.1, 1.5, 0.1)), &m);
    }
    m = mat(1.0, 1.0, 0.8);
    m.ambient = 0.0; m.diffuse = 0.3; m.specular = 0.9; m.shininess = 300.0; m.reflective = 0.1; m.transparency = 0.7;
    m.refractive_index = 1.5;
    cube(w, mul(mul(translation(0.0, 3.45001, 0.0), rotation_y(0.2)), scaling(0.25, 0.25, 0.25)), &m);
    m = mat(1.0, 0.5, 0.5); m.reflective = 0.6; m.diffuse = 0.4;
    cube(w, mul(mul(translation(1.0, 3.35, -0.9), rotation_y(-0.4)), scaling(0.15, 0.15, 0.15)), &m);
    m = mat(1.0, 1.0, 0.5);
    cube(w, mul(mul(translation(-1.5, 3.27, 0.3), rotation_y(0.4)), scaling(0.15, 0.7, 0.15)), &m);
    m = mat(0.5, 1.0, 0.5);
    cube(w, mul(mul(translation(0.0, 3.25, 1.0), rotation_y(0.4)), scaling(0.2, 0.05, 0.05)), &m);
    m = mat(0.5, 0.5, 1.0);
    cube(w, mul(mul(translation(-0.6, 3.4, -1.0), rotation_y(0.8)), scaling(0.05, 0.2, 0.05)), &m);
    m = mat(0.5, 1.0, 1.0);
    cube(w, mul(mul(translation(2.0, 3.4, 1.0), rotation_y(0.8)), scaling(0.05, 0.2, 0.05)), &m);
    m = mat(0.7098, 0.2471, 0.2196); m.diffuse = 0.6;
    cube(w, mul(translation(-10.0, 4.0, 1.0), scaling(0.05, 1.0, 1.0)), &m);
    m = mat(0.2667, 0.2706, 0.6902); m.diffuse = 0.6;
    cube(w, mul(translation(-10.0, 3.4, 2.7), scaling(0.05, 0.4, 0.4)), &m);
    m = mat(0.3098, 0.5961, 0.3098); m.diffuse = 0.6;
    cube(w, mul(translation(-10.0, 4.6, 2.7), scaling(0.05, 0.4, 0.4)), &m);
    m = mat(0.3882, 0.2627, 0.1882); m.diffuse = 0.7;
    cube(w, mul(translation(-2.0, 3.5, 9.95), scaling(5.0, 1.5, 0.05)), &m);
    m = mat(0, 0, 0); m.diffuse = 0.0; m.ambient = 0.0; m.specular = 0.0; m.shininess = 300.0; m.reflective = 1.0;
    cube(w, mul(translation(-2.0, 3.5, 9.95), scaling(4.8, 1.4, 0.06)), &m);
}

static rtc_shape* obj(const char* dir, const char* name) {
    char path[1024];
    snprintf(path, sizeof path, "%s/%s", dir, name);
    rtc_shape* s = rtc_obj_parse_file(path, NULL);
    if (!s) die(path);
    return s;
}

/* main.rs:328-363 */
static void cow_scene(unsigned width, const char* objs, rtc_world** w, rtc_camera** c) {
    *c = camera(width, 8.0, 6.0, -8.0, 0.0, 3.0, 0.0);
    *w = world();
    rtc_shape* cow = obj(objs, "cow-nonormals.obj");
    OK(rtc_shape_set_transform(cow, mul(translation(0., 3.5, 0.), scaling(0.5, 0.5, 0.5)).m));
    rtc_material m = mat(1, 1, 1);
    m.ambient = 0.1; m.diffuse = 0.7; m.specular = 0.9; m.shininess = 300.0; m.reflective = 0.2;
    OK(rtc_shape_set_material(cow, &m));
    OK(rtc_world_push(*w, cow));
}

/* main.rs:368-397 */
static void teapot_scene(unsigned width, const char* objs, rtc_world** w, rtc_camera** c) {
    *c = camera(width, 0.0, 4.0, -12.0, 0.0, 0.0, 0.0);
    *w = world();
    rtc_shape* pot = obj(objs, "teapot.obj");
    OK(rtc_shape_set_transform(pot, translation(0., -1.5, 0.).m));
    rtc_material m = mat(1, 1, 1);
    m.pattern_kind = RTC_PATTERN_GRADIENT; /* Pattern::gradient(GREEN, BLUE), identity transform */
    m.pattern_a[0] = 0; m.pattern_a[1] = 1; m.pattern_a[2] = 0;
    m.pattern_b[0] = 0; m.pattern_b[1] = 0; m.pattern_b[2] = 1;
    OK(rtc_shape_set_material(pot, &m));
    OK(rtc_world_push(*w, pot));
}

static void help(void) { printf("usage: rtc_main <filename.ppm> [width-in-px] [--scene hexagon|table|cow|teapot] [--objs DIR]\n"); }

int main(int argc, char** argv) {
    const char *filename = NULL, *scene = "cow", *objs = "objs";
    unsigned width = 400; /* main.rs:77 */
    int positional = 0;
    for (int i = 1; i < argc; i++) {
        if (!strcmp(argv[i], "--scene") && i + 1 < argc) scene = argv[++i];
        else if (!strcmp(argv[i], "--objs") && i + 1 < argc) objs = argv[++i];
        else if (positional == 0) { filename = argv[i]; positional++; }
        else if (positional == 1) {
            char* end = NULL;
            unsigned long v = strtoul(argv[i], &end, 10);
            if (!*argv[i] || *end) { fprintf(stderr, "Error: Second argument not number!\n"); help(); return 0; }
            width = (unsigned)v;
            positional++;
        } else { printf("too many arguments argument!\n"); help(); return 0; }
    }
    if (!filename) { printf("Expected a filename argument!\n"); help(); return 0; }

    rtc_world* w = NULL;
    rtc_camera* c = NULL;
    if (!strcmp(scene, "hexagon")) hexagon_scene(width, &w, &c);
    else if (!strcmp(scene, "table")) table_scene(width, &w, &c);
    else if (!strcmp(scene, "teapot")) teapot_scene(width, objs, &w, &c);
    else cow_scene(width, objs, &w, &c);

    rtc_canvas* canvas = NULL;
    rtc_stats st;
    OK(rtc_camera_render(c, w, 0, 0, &canvas, &st)); /* camera.render(&world), camera.rs:67 — on the B200 */
    uint64_t len = 0;
    char* ppm = rtc_canvas_to_ppm(canvas, &len); /* canvas.to_ppm(&mut file), canvas.rs:28 */
    FILE* f = fopen(filename, "wb");
    if (!f) { printf("Can't open %s\n", filename); return 0; } /* main.rs:142-145 */
    fwrite(ppm, 1, len, f);
    fclose(f);
    fprintf(stderr, "%s %ux%u: %llu rays, kernel %.3f ms, %llu bytes\n", scene, width, width / 2,
            (unsigned long long)(st.primary_rays + st.shadow_rays + st.reflect_rays + st.refract_rays), st.device_ms,
            (unsigned long long)len);
    rtc_free(ppm);
    rtc_canvas_free(canvas);
    rtc_camera_free(c);
    rtc_world_free(w);
    return 0;
}
