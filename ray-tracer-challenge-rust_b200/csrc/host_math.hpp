// host_math.hpp — host-side f64 linear algebra for scene set-up (product code; never runs per pixel).
//
// Everything here feeds device tables whose values reach pixels, so each routine reproduces the reference's operation
// order exactly (cited file:line, relative to /root/reference/) and this translation unit is compiled with
// -ffp-contract=off (no FMA contraction).  Layout is row-major double[16] throughout so the C ABI passes matrices
// without conversion.
#pragma once
#include <cmath>
#include <cstring>

namespace rtc {

constexpr double kEpsilon = 0.00001;  // utils.rs:2

struct Vec4 {
    double x, y, z, w;
};
inline Vec4 point(double x, double y, double z) { return {x, y, z, 1.0}; }
inline Vec4 vector(double x, double y, double z) { return {x, y, z, 0.0}; }
inline Vec4 sub(const Vec4& a, const Vec4& b) { return {a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w}; }
// tuple.rs:43-48 / 50-66 / 75-83 — sums run left to right and include w
inline double magnitude(const Vec4& v) { return std::sqrt(v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w); }
inline Vec4 normalize(const Vec4& v) {
    double m = magnitude(v);
    if (m == 0.0) return {0., 0., 0., 0.};
    return {v.x / m, v.y / m, v.z / m, v.w / m};
}
inline Vec4 cross(const Vec4& a, const Vec4& b) {
    return vector(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}

struct Mat4 {
    double m[16];
    double& at(int r, int c) { return m[r * 4 + c]; }
    double at(int r, int c) const { return m[r * 4 + c]; }
    static Mat4 identity() {
        Mat4 r;
        for (int i = 0; i < 16; i++) r.m[i] = (i % 5 == 0) ? 1.0 : 0.0;
        return r;
    }
    static Mat4 from(const double* p) {
        Mat4 r;
        std::memcpy(r.m, p, sizeof(r.m));
        return r;
    }
    bool bits_equal(const Mat4& o) const { return std::memcmp(m, o.m, sizeof(m)) == 0; }
};

// matrix.rs:187-205: each element accumulates from 0. in column order n = 0..3
inline Mat4 mul(const Mat4& a, const Mat4& b) {
    Mat4 r;
    for (int row = 0; row < 4; row++)
        for (int col = 0; col < 4; col++) {
            double val = 0.;
            for (int n = 0; n < 4; n++) val += a.at(row, n) * b.at(n, col);
            r.at(row, col) = val;
        }
    return r;
}
// matrix.rs:207-227
inline Vec4 mul(const Mat4& a, const Vec4& t) {
    Vec4 r;
    r.x = a.at(0, 0) * t.x + a.at(0, 1) * t.y + a.at(0, 2) * t.z + a.at(0, 3) * t.w;
    r.y = a.at(1, 0) * t.x + a.at(1, 1) * t.y + a.at(1, 2) * t.z + a.at(1, 3) * t.w;
    r.z = a.at(2, 0) * t.x + a.at(2, 1) * t.y + a.at(2, 2) * t.z + a.at(2, 3) * t.w;
    r.w = a.at(3, 0) * t.x + a.at(3, 1) * t.y + a.at(3, 2) * t.z + a.at(3, 3) * t.w;
    return r;
}
inline Mat4 transpose(const Mat4& a) {  // matrix.rs:29-39
    Mat4 r;
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) r.at(j, i) = a.at(i, j);
    return r;
}

// Cofactor expansion exactly as matrix.rs:41-135 evaluates it: a 4x4 determinant is 0. + sum_c m[0][c]*cof(0,c); each
// 3x3 minor is 0. + sum_c s[0][c]*cof3(0,c); each 2x2 is a*d - b*c.  The sub-matrix keeps row/column order.
namespace detail {
inline double det2(double a, double b, double c, double d) { return a * d - b * c; }
inline double det3(const double s[3][3]) {
    double result = 0.;
    for (int col = 0; col < 3; col++) {
        double q[4];
        int k = 0;
        for (int r = 1; r < 3; r++)
            for (int c = 0; c < 3; c++)
                if (c != col) q[k++] = s[r][c];
        double minor = det2(q[0], q[1], q[2], q[3]);
        double cof = (col % 2 == 0) ? minor : -minor;
        result += s[0][col] * cof;
    }
    return result;
}
inline double cofactor4(const Mat4& a, int row, int col) {
    double s[3][3];
    int rr = 0;
    for (int r = 0; r < 4; r++) {
        if (r == row) continue;
        int cc = 0;
        for (int c = 0; c < 4; c++) {
            if (c == col) continue;
            s[rr][cc++] = a.at(r, c);
        }
        rr++;
    }
    double minor = det3(s);
    return ((row + col) % 2 == 0) ? minor : -minor;
}
}  // namespace detail

inline double determinant(const Mat4& a) {
    double result = 0.;
    for (int col = 0; col < 4; col++) result += a.at(0, col) * detail::cofactor4(a, 0, col);
    return result;
}
// matrix.rs:138-157: None when |det| < EPSILON; element (col,row) = cofactor(row,col) / det (true division)
inline bool inverse(const Mat4& a, Mat4* out) {
    double det = determinant(a);
    if (std::fabs(det - 0.) < kEpsilon) return false;
    for (int row = 0; row < 4; row++)
        for (int col = 0; col < 4; col++) out->at(col, row) = detail::cofactor4(a, row, col) / det;
    return true;
}

// transformations.rs:4-93
inline Mat4 translation(double x, double y, double z) {
    Mat4 r = Mat4::identity();
    r.at(0, 3) = x; r.at(1, 3) = y; r.at(2, 3) = z;
    return r;
}
inline Mat4 scaling(double x, double y, double z) {
    Mat4 r = Mat4::identity();
    r.at(0, 0) = x; r.at(1, 1) = y; r.at(2, 2) = z;
    return r;
}
inline Mat4 rotation_x(double rad) {
    Mat4 r = Mat4::identity();
    double c = std::cos(rad), s = std::sin(rad);
    r.at(1, 1) = c; r.at(2, 2) = c; r.at(1, 2) = -s; r.at(2, 1) = s;
    return r;
}
inline Mat4 rotation_y(double rad) {
    Mat4 r = Mat4::identity();
    double c = std::cos(rad), s = std::sin(rad);
    r.at(0, 0) = c; r.at(2, 2) = c; r.at(0, 2) = s; r.at(2, 0) = -s;
    return r;
}
inline Mat4 rotation_z(double rad) {
    Mat4 r = Mat4::identity();
    double c = std::cos(rad), s = std::sin(rad);
    r.at(0, 0) = c; r.at(1, 1) = c; r.at(0, 1) = -s; r.at(1, 0) = s;
    return r;
}
inline Mat4 shearing(double xy, double xz, double yx, double yz, double zx, double zy) {
    Mat4 r = Mat4::identity();
    r.at(0, 1) = xy; r.at(0, 2) = xz; r.at(1, 0) = yx; r.at(1, 2) = yz; r.at(2, 0) = zx; r.at(2, 1) = zy;
    return r;
}
inline Mat4 view_transform(const Vec4& from, const Vec4& to, const Vec4& up) {
    Vec4 forward = normalize(sub(to, from));
    Vec4 upn = normalize(up);
    Vec4 left = cross(forward, upn);
    Vec4 true_up = cross(left, forward);
    Mat4 o;
    for (double& v : o.m) v = 0.;
    o.at(0, 0) = left.x; o.at(0, 1) = left.y; o.at(0, 2) = left.z;
    o.at(1, 0) = true_up.x; o.at(1, 1) = true_up.y; o.at(1, 2) = true_up.z;
    o.at(2, 0) = -forward.x; o.at(2, 1) = -forward.y; o.at(2, 2) = -forward.z;
    o.at(3, 3) = 1.;
    return mul(o, translation(-from.x, -from.y, -from.z));
}

}  // namespace rtc
