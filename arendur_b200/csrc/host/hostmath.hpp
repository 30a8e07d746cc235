// Host-side f32 linear algebra for scene set-up (product code; not shared with oracle/).
// Mirrors the cgmath 0.14 operations arendur uses at LOAD time only:
//   Matrix4 * Matrix4, invert, transform_point (homogeneous divide), transform_vector,
//   TransformExt::transform_norm (src/geometry/transform.rs:53-59).
// Everything per-ray/per-hit runs on the GPU (csrc/kernels).
#pragma once
#include <cmath>
#include <cstring>

namespace arnhost {

struct Vec3 { float x, y, z; };
struct Mat4 {                      // column-major: c[col][row]
    float c[4][4];
    static Mat4 identity() { Mat4 m; std::memset(&m, 0, sizeof m); for (int i = 0; i < 4; i++) m.c[i][i] = 1.f; return m; }
    static Mat4 from_array(const float* p) { Mat4 m; std::memcpy(m.c, p, 64); return m; }
    static Mat4 translation(float x, float y, float z) { Mat4 m = identity(); m.c[3][0] = x; m.c[3][1] = y; m.c[3][2] = z; return m; }
    static Mat4 scale(float x, float y, float z) { Mat4 m = identity(); m.c[0][0] = x; m.c[1][1] = y; m.c[2][2] = z; return m; }
    void to_array(float* p) const { std::memcpy(p, c, 64); }
    bool is_identity() const { Mat4 i = identity(); return std::memcmp(c, i.c, 64) == 0; }
};

// M * (v0,v1,v2,v3): columns scaled by the vector's components and summed left to right
inline void mul_vec4(const Mat4& m, const float v[4], float out[4]) {
    for (int r = 0; r < 4; r++) out[r] = m.c[0][r] * v[0] + m.c[1][r] * v[1] + m.c[2][r] * v[2] + m.c[3][r] * v[3];
}
inline Mat4 operator*(const Mat4& a, const Mat4& b) {
    Mat4 o;
    for (int j = 0; j < 4; j++) mul_vec4(a, b.c[j], o.c[j]);
    return o;
}
inline Vec3 transform_vector(const Mat4& m, Vec3 v) {
    float in[4] = {v.x, v.y, v.z, 0.f}, o[4]; mul_vec4(m, in, o); return Vec3{o[0], o[1], o[2]};
}
inline Vec3 transform_point(const Mat4& m, Vec3 p) {
    float in[4] = {p.x, p.y, p.z, 1.f}, o[4]; mul_vec4(m, in, o);
    float iw = 1.f / o[3];
    return Vec3{o[0] * iw, o[1] * iw, o[2] * iw};
}
inline Mat4 transpose(const Mat4& m) { Mat4 t; for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) t.c[i][j] = m.c[j][i]; return t; }

namespace detail {
// determinant of the 3x3 matrix with columns (a, b, c)
inline float det3(const float a[3], const float b[3], const float c[3]) {
    return a[0] * (b[1] * c[2] - c[1] * b[2]) - b[0] * (a[1] * c[2] - c[1] * a[2]) + c[0] * (a[1] * b[2] - b[1] * a[2]);
}
inline void drop_row(const float col[4], int row, float out[3]) { int k = 0; for (int r = 0; r < 4; r++) if (r != row) out[k++] = col[r]; }
}  // namespace detail

inline float determinant(const Mat4& m) {
    float acc = 0.f;
    for (int col = 0; col < 4; col++) {
        // minor of element (col, row 0).  cgmath builds it as Matrix3::new(self[a][1], self[b][1], self[c][1], self[a][2], ...) for the
        // other columns a < b < c: the TRANSPOSED minor (column j = row j + 1 of the three columns).  Same determinant, but not
        // the same roundings — and inv_det scales every entry of the inverse, hence every transformed normal.
        int other[3], k = 0;
        for (int c2 = 0; c2 < 4; c2++) if (c2 != col) other[k++] = c2;
        float cols[3][3];
        for (int j = 0; j < 3; j++) for (int q = 0; q < 3; q++) cols[j][q] = m.c[other[q]][j + 1];
        float d = detail::det3(cols[0], cols[1], cols[2]);
        float term = m.c[col][0] * d;
        if (col == 0) acc = term; else if (col & 1) acc = acc - term; else acc = acc + term;
    }
    return acc;
}
// cgmath SquareMatrix::invert: cofactors of the transpose scaled by 1/det; fails if det ~ 0
// (ulps_eq!(det, 0) — |det| <= f32::EPSILON).
inline bool invert(const Mat4& m, Mat4* out) {
    float det = determinant(m);
    if (std::fabs(det) <= 1.1920929e-7f) return false;
    float inv_det = 1.f / det;
    Mat4 t = transpose(m);
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) {
        float cols[3][3]; int k = 0;
        for (int c2 = 0; c2 < 4; c2++) if (c2 != i) detail::drop_row(t.c[c2], j, cols[k++]);
        float d = detail::det3(cols[0], cols[1], cols[2]);
        float sign = ((i + j) & 1) ? -1.f : 1.f;
        out->c[i][j] = d * sign * inv_det;
    }
    return true;
}
inline Vec3 normalize(Vec3 v) {
    float inv = 1.f / std::sqrt(v.x * v.x + v.y * v.y + v.z * v.z);
    return Vec3{v.x * inv, v.y * inv, v.z * inv};
}
inline Vec3 transform_norm(const Mat4& m, Vec3 n) {
    Mat4 inv; if (!invert(m, &inv)) return n;
    return normalize(transform_vector(transpose(inv), n));
}

}  // namespace arnhost
