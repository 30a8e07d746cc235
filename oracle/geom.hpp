// ORACLE — test infrastructure only.  CPU restatement of arendur's geometry layer.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may build or call this code; the product never links it.
//
// Restates (paths relative to the arendur tree):
//   src/geometry/float.rs, foundamental.rs, ray.rs, bbox.rs, transform.rs, interaction.rs
// and the cgmath 0.14 semantics they lean on (SURVEY.md Appendix C — cgmath's source
// is NOT in the reference tree; its documented behaviour is restated here:
// PARITY UNPINNED for those pieces).
//
// Compile with -O2 -ffp-contract=off -fno-fast-math: all arithmetic is scalar IEEE f32,
// evaluated left to right exactly as written in the Rust source, no FMA contraction.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>

namespace orc {

typedef float Float;

// ---- geometry/float.rs -------------------------------------------------------
inline Float clampf(Float f, Float mn, Float mx) {           // float.rs:13-19
    if (f < mn) return mn; else if (f < mx) return f; else return mx;
}
inline Float epsilon() { return std::numeric_limits<float>::epsilon(); }     // :22-24
inline Float machine_epsilon() { return epsilon() * 0.5f; }                  // :27-29
inline Float eb_term(Float n) {                                              // :33-36
    Float ne = n * machine_epsilon();
    return ne / (1.f - ne);
}
inline Float infinity() { return std::numeric_limits<float>::infinity(); }
inline Float pi() { return 3.14159265358979323846f; }
inline Float frac_1_pi() { return 0.318309886183790671537767526745028724f; }
inline Float frac_pi_2() { return 1.57079632679489661923132169163975144f; }
inline Float frac_pi_4() { return 0.785398163397448309615660845819875721f; }

// Transcendentals.  Rust's f32::{sin,cos,tan,acos,atan2,ln,exp,powf} call the platform libm,
// whose results are only defined up to its own error bound.  The oracle AND the GPU evaluate them
// as correctly-rounded f32 (computed in f64, rounded once), so both sides agree bit for bit
// (up to ~1e-8 double-rounding cases) and stay within 0.5 ulp of any faithful libm.  DESIGN.md.
inline Float fsin(Float x) { return (Float)std::sin((double)x); }
inline Float fcos(Float x) { return (Float)std::cos((double)x); }
inline Float ftan(Float x) { return (Float)std::tan((double)x); }
inline Float facos(Float x) { return (Float)std::acos((double)x); }
inline Float fatan2(Float y, Float x) { return (Float)std::atan2((double)y, (double)x); }
inline Float flog(Float x) { return (Float)std::log((double)x); }
inline Float fexp(Float x) { return (Float)std::exp((double)x); }
inline Float fpow(Float a, Float b) { return (Float)std::pow((double)a, (double)b); }

inline uint32_t to_bits(Float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
inline Float from_bits(uint32_t u) { Float f; std::memcpy(&f, &u, 4); return f; }
inline bool sign_positive(Float f) { return (to_bits(f) >> 31) == 0; }

inline Float next_up(Float f) {                                              // :104-117
    if (std::isinf(f) && sign_positive(f)) return f;
    else if (f == -0.f) return 0.f;       // note: also true for +0 (as in the source)
    else {
        uint32_t t = to_bits(f);
        return sign_positive(f) ? from_bits(t + 1) : from_bits(t - 1);
    }
}
inline Float next_down(Float f) {                                            // :120-133
    if (std::isinf(f) && !sign_positive(f)) return f;
    else if (f == 0.f) return -0.f;
    else {
        uint32_t t = to_bits(f);
        return !sign_positive(f) ? from_bits(t + 1) : from_bits(t - 1);
    }
}
// Rust f32::max / f32::min (IEEE maxNum/minNum)
inline Float fmax_(Float a, Float b) { return std::fmax(a, b); }
inline Float fmin_(Float a, Float b) { return std::fmin(a, b); }
// Rust f32::signum: 1.0 for +0 and positives, -1.0 for -0 and negatives, NaN for NaN
inline Float signum(Float x) { if (std::isnan(x)) return x; return sign_positive(x) ? 1.f : -1.f; }

// ---- cgmath vectors ------------------------------------------------------------
struct V2 { Float x, y; };
struct V3 {
    Float x, y, z;
    Float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
    Float& at(int i) { return i == 0 ? x : (i == 1 ? y : z); }
};
inline V3 v3(Float x, Float y, Float z) { V3 r = {x, y, z}; return r; }
inline V2 v2(Float x, Float y) { V2 r = {x, y}; return r; }
inline V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline V3 operator-(V3 a) { return v3(-a.x, -a.y, -a.z); }
inline V3 operator*(V3 a, Float s) { return v3(a.x * s, a.y * s, a.z * s); }
inline V3 operator*(Float s, V3 a) { return v3(s * a.x, s * a.y, s * a.z); }
inline V3 operator/(V3 a, Float s) { return v3(a.x / s, a.y / s, a.z / s); }
inline V3 mul_elem(V3 a, V3 b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline Float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }   // (x+y)+z
inline Float magnitude2(V3 a) { return dot(a, a); }
inline Float magnitude(V3 a) { return std::sqrt(magnitude2(a)); }
inline V3 normalize(V3 a) { return a * (1.f / magnitude(a)); }               // normalize_to(1)
inline V3 cross(V3 a, V3 b) {
    return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
inline V2 operator+(V2 a, V2 b) { return v2(a.x + b.x, a.y + b.y); }
inline V2 operator-(V2 a, V2 b) { return v2(a.x - b.x, a.y - b.y); }
inline V2 operator*(V2 a, Float s) { return v2(a.x * s, a.y * s); }
inline V2 operator*(Float s, V2 a) { return v2(s * a.x, s * a.y); }
inline bool isnan3(V3 a) { return std::isnan(a.x) || std::isnan(a.y) || std::isnan(a.z); }
inline bool isinf3(V3 a) { return std::isinf(a.x) || std::isinf(a.y) || std::isinf(a.z); }

// approx::relative_eq! with default epsilon = max_relative = f32::EPSILON
inline bool relative_eq(Float a, Float b) {
    if (a == b) return true;
    if (std::isinf(a) || std::isinf(b)) return false;
    Float d = std::fabs(a - b);
    if (d <= epsilon()) return true;
    Float aa = std::fabs(a), ab = std::fabs(b);
    Float largest = ab > aa ? ab : aa;
    return d <= largest * epsilon();
}
inline bool relative_eq(V3 a, V3 b) {
    return relative_eq(a.x, b.x) && relative_eq(a.y, b.y) && relative_eq(a.z, b.z);
}

// ---- cgmath Matrix4 (column-major, m[c*4+r]) -------------------------------------
struct M4 {
    Float m[16];
    Float at(int c, int r) const { return m[c * 4 + r]; }
};
inline M4 m4_identity() {
    M4 r; for (int i = 0; i < 16; i++) r.m[i] = (i % 5 == 0) ? 1.f : 0.f; return r;
}
inline M4 m4_from_cols(const Float* p) { M4 r; std::memcpy(r.m, p, 64); return r; }
inline M4 m4_from_translation(V3 v) {
    M4 r = m4_identity(); r.m[12] = v.x; r.m[13] = v.y; r.m[14] = v.z; return r;
}
inline M4 m4_from_nonuniform_scale(Float x, Float y, Float z) {
    M4 r = m4_identity(); r.m[0] = x; r.m[5] = y; r.m[10] = z; return r;
}
inline M4 m4_transpose(const M4& a) {
    M4 r; for (int c = 0; c < 4; c++) for (int rr = 0; rr < 4; rr++) r.m[c * 4 + rr] = a.m[rr * 4 + c];
    return r;
}
// Matrix4 * Vector4: a*v0 + b*v1 + c*v2 + d*v3 (columns scaled, summed left to right)
inline void m4_mul_v4(const M4& M, const Float v[4], Float out[4]) {
    for (int r = 0; r < 4; r++)
        out[r] = M.at(0, r) * v[0] + M.at(1, r) * v[1] + M.at(2, r) * v[2] + M.at(3, r) * v[3];
}
inline M4 m4_mul(const M4& L, const M4& R) {
    M4 o;
    for (int j = 0; j < 4; j++) {
        Float v[4] = {R.at(j, 0), R.at(j, 1), R.at(j, 2), R.at(j, 3)};
        Float c[4]; m4_mul_v4(L, v, c);
        for (int r = 0; r < 4; r++) o.m[j * 4 + r] = c[r];
    }
    return o;
}
// Transform::transform_vector: (M * (v,0)).xyz
inline V3 transform_vector(const M4& M, V3 v) {
    Float in[4] = {v.x, v.y, v.z, 0.f}, o[4]; m4_mul_v4(M, in, o); return v3(o[0], o[1], o[2]);
}
// Transform::transform_point: Point3::from_homogeneous(M * (p,1)) = xyz * (1/w)
inline V3 transform_point(const M4& M, V3 p) {
    Float in[4] = {p.x, p.y, p.z, 1.f}, o[4]; m4_mul_v4(M, in, o);
    Float iw = 1.f / o[3];
    return v3(o[0] * iw, o[1] * iw, o[2] * iw);
}
inline Float det3(Float a00, Float a01, Float a02,   // column 0
                  Float a10, Float a11, Float a12,   // column 1
                  Float a20, Float a21, Float a22) { // column 2   (a[c][r])
    return a00 * (a11 * a22 - a21 * a12) - a10 * (a01 * a22 - a21 * a02) + a20 * (a01 * a12 - a11 * a02);
}
inline Float m4_determinant(const M4& s) {
    Float m0 = det3(s.at(1,1), s.at(2,1), s.at(3,1), s.at(1,2), s.at(2,2), s.at(3,2), s.at(1,3), s.at(2,3), s.at(3,3));
    Float m1 = det3(s.at(0,1), s.at(2,1), s.at(3,1), s.at(0,2), s.at(2,2), s.at(3,2), s.at(0,3), s.at(2,3), s.at(3,3));
    Float m2 = det3(s.at(0,1), s.at(1,1), s.at(3,1), s.at(0,2), s.at(1,2), s.at(3,2), s.at(0,3), s.at(1,3), s.at(3,3));
    Float m3 = det3(s.at(0,1), s.at(1,1), s.at(2,1), s.at(0,2), s.at(1,2), s.at(2,2), s.at(0,3), s.at(1,3), s.at(2,3));
    return s.at(0,0) * m0 - s.at(1,0) * m1 + s.at(2,0) * m2 - s.at(3,0) * m3;
}
// SquareMatrix::invert for Matrix4 (cofactors of the transpose times 1/det); false if det ~ 0
inline bool m4_invert(const M4& s, M4* out) {
    Float det = m4_determinant(s);
    if (std::fabs(det) <= epsilon()) return false;   // ulps_eq!(det, 0): |det| <= f32::EPSILON
    Float inv_det = 1.f / det;
    M4 t = m4_transpose(s);
    for (int i = 0; i < 4; i++) {
        for (int j = 0; j < 4; j++) {
            // columns of t except column i, each with row j removed
            int cols[3], n = 0; for (int c = 0; c < 4; c++) if (c != i) cols[n++] = c;
            Float a[3][3];
            for (int c = 0; c < 3; c++) { int rr = 0; for (int r = 0; r < 4; r++) if (r != j) a[c][rr++] = t.at(cols[c], r); }
            Float d = det3(a[0][0], a[0][1], a[0][2], a[1][0], a[1][1], a[1][2], a[2][0], a[2][1], a[2][2]);
            Float sign = ((i + j) & 1) ? -1.f : 1.f;
            out->m[i * 4 + j] = d * sign * inv_det;
        }
    }
    return true;
}
// TransformExt::transform_norm (transform.rs:53-59): inverse-transpose, then normalize.
inline V3 transform_norm(const M4& M, V3 n) {
    M4 inv; m4_invert(M, &inv);
    return normalize(transform_vector(m4_transpose(inv), n));
}

// ---- geometry/bbox.rs ------------------------------------------------------------
inline Float partial_min(Float a, Float b) { return a < b ? a : b; }
inline Float partial_max(Float a, Float b) { return a > b ? a : b; }
struct BBox3 {
    V3 pmin, pmax;
    static BBox3 make(V3 p, V3 q) {                                          // bbox.rs:244-257
        BBox3 b; b.pmin = v3(partial_min(p.x, q.x), partial_min(p.y, q.y), partial_min(p.z, q.z));
        b.pmax = v3(partial_max(p.x, q.x), partial_max(p.y, q.y), partial_max(p.z, q.z)); return b;
    }
    BBox3 extend(V3 p) const {                                               // :294-308
        BBox3 b; b.pmin = v3(partial_min(pmin.x, p.x), partial_min(pmin.y, p.y), partial_min(pmin.z, p.z));
        b.pmax = v3(partial_max(pmax.x, p.x), partial_max(pmax.y, p.y), partial_max(pmax.z, p.z)); return b;
    }
    BBox3 unite(const BBox3& o) const {                                      // :311-325
        BBox3 b; b.pmin = v3(partial_min(pmin.x, o.pmin.x), partial_min(pmin.y, o.pmin.y), partial_min(pmin.z, o.pmin.z));
        b.pmax = v3(partial_max(pmax.x, o.pmax.x), partial_max(pmax.y, o.pmax.y), partial_max(pmax.z, o.pmax.z)); return b;
    }
    V3 diagonal() const { return pmax - pmin; }                              // :394-396
    Float surface_area() const {                                            // :399-411
        Float dx = pmax.x - pmin.x, dy = pmax.y - pmin.y, dz = pmax.z - pmin.z;
        if (dx < 0.f) dx = 0.f; if (dy < 0.f) dy = 0.f; if (dz < 0.f) dz = 0.f;
        return 2.f * (dx * dy + dx * dz + dy * dz);
    }
    int max_extent() const {                                                 // :432-441
        V3 d = diagonal();
        if (d.x > d.y && d.x > d.z) return 0; else if (d.y > d.z) return 1; else return 2;
    }
    // apply_transform (bbox.rs:481-499): pmin as a point, the diagonal as a vector (quirk A-9)
    BBox3 apply_transform(const M4& t) const {
        V3 p = transform_point(t, pmin);
        V3 d = transform_vector(t, diagonal());
        return BBox3::make(p, p + d);
    }
};

// BBox2<T> (bbox.rs:21-232, iterator :594-634) — the tile / pixel-box algebra
template <typename T> struct BBox2 {
    T x0, y0, x1, y1;   // pmin, pmax
    static BBox2 make(T px, T py, T qx, T qy) {                              // :33-44
        BBox2 b; b.x0 = px < qx ? px : qx; b.y0 = py < qy ? py : qy; b.x1 = px > qx ? px : qx; b.y1 = py > qy ? py : qy; return b;
    }
    void corner(int i, T* x, T* y) const { *x = (i & 1) == 0 ? x0 : x1; *y = (i & 2) == 0 ? y0 : y1; }   // :47-62
    BBox2 extend(T px, T py) const {                                          // :66-77
        BBox2 b; b.x0 = x0 < px ? x0 : px; b.y0 = y0 < py ? y0 : py; b.x1 = x1 > px ? x1 : px; b.y1 = y1 > py ? y1 : py; return b;
    }
    BBox2 unite(const BBox2& o) const {                                       // :81-92
        BBox2 b; b.x0 = x0 < o.x0 ? x0 : o.x0; b.y0 = y0 < o.y0 ? y0 : o.y0; b.x1 = x1 > o.x1 ? x1 : o.x1; b.y1 = y1 > o.y1 ? y1 : o.y1; return b;
    }
    bool intersect(const BBox2& o, BBox2* r) const {                          // :96-112
        r->x0 = x0 > o.x0 ? x0 : o.x0; r->y0 = y0 > o.y0 ? y0 : o.y0; r->x1 = x1 < o.x1 ? x1 : o.x1; r->y1 = y1 < o.y1 ? y1 : o.y1;
        return !(r->x0 > r->x1 || r->y0 > r->y1);
    }
    bool overlap(const BBox2& o) const { return (x1 >= o.x0 && x0 <= o.x1) && (y1 >= o.y0 && y0 <= o.y1); }  // :116-119
    bool contain(T px, T py) const { return (px <= x1 && px >= x0) && (py <= y1 && py >= y0); }              // :123-126
    bool contain_lb(T px, T py) const { return (px < x1 && px >= x0) && (py < y1 && py >= y0); }             // :131-134
    BBox2 expand_by(T d) const { BBox2 b; b.x0 = x0 + (-d); b.y0 = y0 + (-d); b.x1 = x1 + d; b.y1 = y1 + d; return b; }  // :138-144
    T surface_area() const { T dx = x1 - x0, dy = y1 - y0; return (dx >= 0 && dy >= 0) ? dx * dy : 0; }     // :164-173
    int max_extent() const { return (x1 - x0) > (y1 - y0) ? 0 : 1; }                                         // :177-184
    void lerp(Float tx, Float ty, T* ox, T* oy) const {                                                      // :187-201
        *ox = (T)(tx * (Float)x1 + (1.0f - tx) * (Float)x0); *oy = (T)(ty * (Float)y1 + (1.0f - ty) * (Float)y0);
    }
};

// ---- geometry/ray.rs ---------------------------------------------------------------
enum Perm { PERM_XZ, PERM_YZ, PERM_ZZ };
struct Stc { Perm perm; V3 neg_o; V3 shear; };                               // ray.rs:171-176
struct RawRay {
    V3 origin, dir; Float tmax; Stc stc;
};
inline V3 permxz(V3 p) { return v3(p.y, p.z, p.x); }                          // :251-253
inline V3 permyz(V3 p) { return v3(p.z, p.x, p.y); }                          // :256-258
inline Stc stc_from(V3 o, V3 direction) {                                    // :190-222
    Stc s; s.neg_o = -o;
    Float ax = std::fabs(direction.x), ay = std::fabs(direction.y), az = std::fabs(direction.z);
    V3 d;
    if (ax > ay && ax > az) { s.perm = PERM_XZ; d = v3(direction.y, direction.z, direction.x); }
    else if (ay > az)       { s.perm = PERM_YZ; d = v3(direction.z, direction.x, direction.y); }
    else                    { s.perm = PERM_ZZ; d = direction; }
    s.shear = v3(-d.x / d.z, -d.y / d.z, 1.f / d.z);
    return s;
}
inline void stc_apply(const Stc& s, V3 p0, V3 p1, V3 p2, V3* q0, V3* q1, V3* q2) { // :224-235
    V3 a = p0 + s.neg_o, b = p1 + s.neg_o, c = p2 + s.neg_o;
    if (s.perm == PERM_XZ) { a = permxz(a); b = permxz(b); c = permxz(c); }
    else if (s.perm == PERM_YZ) { a = permyz(a); b = permyz(b); c = permyz(c); }
    a.x += s.shear.x * a.z; a.y += s.shear.y * a.z;
    b.x += s.shear.x * b.z; b.y += s.shear.y * b.z;
    c.x += s.shear.x * c.z; c.y += s.shear.y * c.z;
    *q0 = a; *q1 = b; *q2 = c;
}
inline RawRay ray_new(V3 o, V3 d, Float tmax) {                               // :72-84
    RawRay r; r.origin = o; r.dir = d; r.tmax = tmax; r.stc = stc_from(o, d); return r;
}
inline RawRay ray_from_od(V3 o, V3 d) { return ray_new(o, d, infinity()); }   // :87-90
inline RawRay ray_spawn(V3 origin, V3 destination) {                          // :93-98
    V3 du = destination - origin; Float tmax = magnitude(du);
    return ray_new(origin, du / tmax, tmax);
}
inline void ray_set_tmax(RawRay& r, Float t) { r.tmax = t; r.stc = stc_from(r.origin, r.dir); } // :136-139
inline V3 ray_evaluate(const RawRay& r, Float t) { return r.origin + r.dir * t; } // :42-44
inline RawRay ray_apply_transform(const RawRay& r, const M4& t) {             // :154-161
    return ray_new(transform_point(t, r.origin), transform_vector(t, r.dir), r.tmax);
}

// BBox3f::construct_ray_cache / intersect_ray_cached (bbox.rs:583-592, 549-580)
struct RayCache { V3 o; V3 inv; bool neg[3]; Float tmax; };
inline RayCache construct_ray_cache(const RawRay& r) {
    RayCache c; c.o = r.origin; c.inv = v3(1.f / r.dir.x, 1.f / r.dir.y, 1.f / r.dir.z);
    c.neg[0] = c.inv.x < 0.f; c.neg[1] = c.inv.y < 0.f; c.neg[2] = c.inv.z < 0.f; c.tmax = r.tmax; return c;
}
inline bool intersect_ray_cached(const BBox3& b, const RayCache& c) {
    const Float k = 1.f + 2.f * eb_term(3.f);
    Float t0 = ((c.neg[0] ? b.pmax : b.pmin).x - c.o.x) * c.inv.x;
    Float t1 = ((!c.neg[0] ? b.pmax : b.pmin).x - c.o.x) * c.inv.x;
    Float ty0 = ((c.neg[1] ? b.pmax : b.pmin).y - c.o.y) * c.inv.y;
    Float ty1 = ((!c.neg[1] ? b.pmax : b.pmin).y - c.o.y) * c.inv.y;
    t1 *= k; ty1 *= k;
    if (t0 > ty1 || ty0 > t1) return false;
    if (ty0 > t0) t0 = ty0;
    if (ty1 < t1) t1 = ty1;
    Float tz0 = ((c.neg[2] ? b.pmax : b.pmin).z - c.o.z) * c.inv.z;
    Float tz1 = ((!c.neg[2] ? b.pmax : b.pmin).z - c.o.z) * c.inv.z;
    tz1 *= k;
    if (t0 > tz1 || tz0 > t1) return false;
    if (tz0 > t0) t0 = tz0;
    if (tz1 < t1) t1 = tz1;
    return t0 < c.tmax && t1 > 0.f;
}

// ---- geometry/interaction.rs ---------------------------------------------------------
struct DuvInfo { V3 dpdu, dpdv, dndu, dndv; };
struct InteractInfo { V3 pos, pos_err, wo, norm; };
struct SurfaceInteraction {
    InteractInfo basic; V2 uv; DuvInfo duv; V3 shading_norm; DuvInfo shading_duv;
    int primitive_hit;   // component index, -1 = None
};
inline SurfaceInteraction si_new(V3 pos, V3 perr, V3 wo, V2 uv, DuvInfo duv) {   // :133-162
    SurfaceInteraction s;
    V3 norm = normalize(cross(duv.dpdu, duv.dpdv));
    s.basic.pos = pos; s.basic.pos_err = perr; s.basic.wo = wo; s.basic.norm = norm;
    s.uv = uv; s.duv = duv; s.shading_norm = norm; s.shading_duv = duv; s.primitive_hit = -1;
    return s;
}
inline void si_set_shading(SurfaceInteraction& s, DuvInfo duv, bool orient_norm_by_shading) { // :167-182
    s.duv = duv;                                      // quirk A-7: writes duv, not shading_duv
    V3 norm = normalize(cross(duv.dpdu, duv.dpdv));
    if (dot(s.basic.norm, norm) < 0.f) {
        if (orient_norm_by_shading) norm = -norm; else s.basic.norm = -s.basic.norm;
    }
    s.shading_norm = norm;
}
inline DuvInfo duv_apply_transform(const DuvInfo& d, const M4& t) {              // :90-99
    DuvInfo r; r.dpdu = transform_vector(t, d.dpdu); r.dpdv = transform_vector(t, d.dpdv);
    r.dndu = transform_norm(t, d.dndu); r.dndv = transform_norm(t, d.dndv); return r;
}
inline SurfaceInteraction si_apply_transform(const SurfaceInteraction& s, const M4& t) { // :190-201, :34-43
    SurfaceInteraction r;
    r.basic.pos = transform_point(t, s.basic.pos);
    r.basic.pos_err = transform_vector(t, s.basic.pos_err);
    r.basic.wo = transform_vector(t, s.basic.wo);
    r.basic.norm = transform_norm(t, s.basic.norm);
    r.uv = s.uv;
    r.duv = duv_apply_transform(s.duv, t);
    r.shading_norm = transform_norm(t, s.shading_norm);
    r.shading_duv = duv_apply_transform(s.shading_duv, t);
    r.primitive_hit = s.primitive_hit;
    return r;
}
inline V3 offset_towards(const InteractInfo& b, V3 dir) {                          // :45-72
    V3 nabs = v3(std::fabs(b.norm.x), std::fabs(b.norm.y), std::fabs(b.norm.z));
    Float edn = dot(nabs, b.pos_err);
    V3 offset = edn * b.norm;
    if (dot(dir, b.norm) <= 0.f) offset = -offset;
    V3 ret = b.pos + offset;
    if (offset.x > 0.f) ret.x = next_up(ret.x); else if (offset.x < 0.f) ret.x = next_down(ret.x);
    if (offset.y > 0.f) ret.y = next_up(ret.y); else if (offset.y < 0.f) ret.y = next_down(ret.y);
    if (offset.z > 0.f) ret.z = next_up(ret.z); else if (offset.z < 0.f) ret.z = next_down(ret.z);
    return ret;
}
// spawn_ray_differential (:236-251) without the differentials (unused with constant textures)
inline RawRay si_spawn_ray(const SurfaceInteraction& s, V3 dir) {
    return ray_from_od(offset_towards(s.basic, dir), dir);
}

// ---- geometry/foundamental.rs `normal` helpers (:205-309) -------------------------------
namespace nrm {
inline Float cos_theta(V3 n) { return n.z; }
inline Float cos2_theta(V3 n) { return n.z * n.z; }
inline Float sin2_theta(V3 n) { return std::fabs(1.f - cos2_theta(n)); }
inline Float sin_theta(V3 n) { return std::sqrt(sin2_theta(n)); }
inline Float tan_theta(V3 n) { return sin_theta(n) / cos_theta(n); }
inline Float tan2_theta(V3 n) { return sin2_theta(n) / cos2_theta(n); }
inline Float cos_phi(V3 n) { Float st = sin_theta(n); return st == 0.f ? 1.f : clampf(n.x / st, -1.f, 1.f); }
inline Float sin_phi(V3 n) { Float st = sin_theta(n); return st == 0.f ? 0.f : clampf(n.y / st, -1.f, 1.f); }
inline Float cos2_phi(V3 n) { Float c = cos_phi(n); return c * c; }
inline Float sin2_phi(V3 n) { Float s = sin_phi(n); return s * s; }
inline V3 reflect(V3 wo, V3 n) { return -wo + 2.f * dot(wo, n) * n; }
inline bool refract(V3 wo, V3 n, Float eta, V3* out) {                             // :278-292
    Float ct = dot(wo, n);
    Float s2 = 1.f - ct * ct;
    Float s2t = eta * eta * fmax_(s2, 0.f);
    if (s2t >= 1.f) return false;
    Float ctt = std::sqrt(1.f - s2t);
    *out = -eta * wo + (eta * ct - ctt) * n;
    return true;
}
inline void get_basis_from(V3 dir, V3* u, V3* v) {                                 // :296-305
    V3 up = v3(0.f, 0.f, 1.f);
    if (relative_eq(up, dir)) up = v3(0.f, 1.f, 0.f);
    *u = normalize(cross(up, dir));
    *v = normalize(cross(dir, *u));
}
}  // namespace nrm

// Sphericalf::to_vec (foundamental.rs:151-163)
inline V3 spherical_to_vec(Float theta, Float phi) {
    Float st = fsin(theta), ct = fcos(theta), sp = fsin(phi), cp = fcos(phi);
    return v3(st * cp, st * sp, ct);
}

}  // namespace orc
