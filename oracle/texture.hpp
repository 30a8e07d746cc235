// ORACLE (test infrastructure) — image textures, UV mapping, image-plane differentials and bump mapping.
// CPU restatement of src/texturing/textures/image.rs (MipMap look-up: :405-527), src/texturing/mappings.rs:14-31,
// src/geometry/interaction.rs:204-251,308-325 (compute_dxy, spawn_ray_differential, solve_over_constrained_2x3),
// src/geometry/ray.rs:262-291 (RayDifferential), src/material/mod.rs:42-86 (add_bumping).
//
// Parity unpinned: the reference has no test, fixture or scene that uses an image texture.  Third-party semantics restated:
//   * `image` 0.12 builds the pyramid (decode + Lanczos3 resize): outside the reference tree, so the levels are INPUT here;
//   * cgmath 0.14 `Matrix2::invert` (det == 0 -> None, else the adjugate divided by det element-wise) and `Matrix2 * Vector2`
//     (column 0 * x + column 1 * y) as published;
//   * `f32 as usize` of a negative value (triangle_filter's `s.floor() as usize`, image.rs:431-432) was undefined in 2017 rustc;
//     restated as x86-64's cvttss2si behaviour (two's-complement wrap), which with power-of-two levels and Repeat wrapping is
//     the ordinary wrap-around.  Never used as the product; see the header of oracle_api.cpp.
#pragma once
#include <cstdint>
#include <vector>
#include "geom.hpp"

namespace orc {

struct DxyInfo { V3 dpdx, dpdy; Float dudx, dvdx, dudy, dvdy; };                  // interaction.rs:264-282
inline DxyInfo dxy_default() { DxyInfo d; d.dpdx = d.dpdy = v3(0, 0, 0); d.dudx = d.dvdx = d.dudy = d.dvdy = 0.f; return d; }

struct RayDifferential { RawRay ray; bool has_diffs; RawRay rx, ry; };            // ray.rs:262-267
inline void scale_differentials(RayDifferential& r, Float s) {                    // ray.rs:282-291
    V3 origin = r.ray.origin, dir = r.ray.dir;
    if (r.has_diffs) {
        r.rx.origin = origin + (r.rx.origin - origin) * s;
        r.ry.origin = origin + (r.ry.origin - origin) * s;
        r.rx.dir = dir + (r.rx.dir - dir) * s;
        r.ry.dir = dir + (r.ry.dir - dir) * s;
    }
}

// cgmath Matrix2::new(c0r0, c0r1, c1r0, c1r1).invert().map(|m| m * v)
inline bool m2_solve(Float c0r0, Float c0r1, Float c1r0, Float c1r1, V2 v, V2* out) {
    Float det = c0r0 * c1r1 - c1r0 * c0r1;
    if (det == 0.f) return false;
    Float i00 = c1r1 / det, i01 = -c0r1 / det, i10 = -c1r0 / det, i11 = c0r0 / det;   // inverse columns (i00, i01), (i10, i11)
    *out = v2(i00 * v.x + i10 * v.y, i01 * v.x + i11 * v.y);
    return true;
}
inline V2 solve_over_constrained_2x3(V3 abc, V3 m0, V3 m1, V3 n) {                // interaction.rs:308-325, None -> (0, 0)
    V2 r = v2(0.f, 0.f); bool ok;
    if (std::fabs(n.x) > std::fabs(n.y) && std::fabs(n.x) > std::fabs(n.z)) ok = m2_solve(m0.y, m1.y, m0.z, m1.z, v2(abc.y, abc.z), &r);
    else if (std::fabs(n.y) > std::fabs(n.z)) ok = m2_solve(m0.x, m1.x, m0.z, m1.z, v2(abc.x, abc.z), &r);
    else ok = m2_solve(m0.x, m1.x, m0.y, m1.y, v2(abc.x, abc.y), &r);
    return ok ? r : v2(0.f, 0.f);
}
inline DxyInfo compute_dxy(const SurfaceInteraction& si, const RayDifferential& rd) {   // interaction.rs:204-224
    if (!rd.has_diffs) return dxy_default();
    V3 n = si.basic.norm, pos = si.basic.pos;
    Float d = dot(n, pos);
    Float tx = (d - dot(n, rd.rx.origin)) / dot(n, rd.rx.dir);
    V3 px = ray_evaluate(rd.rx, tx);
    Float ty = (d - dot(n, rd.ry.origin)) / dot(n, rd.ry.dir);
    V3 py = ray_evaluate(rd.ry, ty);
    DxyInfo o; o.dpdx = px - pos; o.dpdy = py - pos;
    V2 dudxy = solve_over_constrained_2x3(o.dpdx, si.duv.dpdu, si.duv.dpdv, n);
    V2 dvdxy = solve_over_constrained_2x3(o.dpdy, si.duv.dpdu, si.duv.dpdv, n);
    o.dudx = dudxy.x; o.dudy = dudxy.y; o.dvdx = dvdxy.x; o.dvdy = dvdxy.y;         // sic: the field pairing of the source (:218-221)
    return o;
}
inline RayDifferential spawn_ray_differential(const SurfaceInteraction& si, V3 dir, const DxyInfo* dxy) {   // :236-251
    V3 pos = offset_towards(si.basic, dir);
    RayDifferential r; r.ray = ray_from_od(pos, dir); r.has_diffs = dxy != nullptr;
    if (dxy) { r.rx = ray_from_od(pos + dxy->dpdx, dir); r.ry = ray_from_od(pos + dxy->dpdy, dir); }
    return r;
}

// ---- MipMap (image.rs)
struct TexView { const arn_texture* t; const Float* texels; };
struct Texel { Float c[3]; };
inline Texel tx_zero() { Texel z; z.c[0] = z.c[1] = z.c[2] = 0.f; return z; }
inline Texel tx_mul(Texel a, Float f) { a.c[0] = a.c[0] * f; a.c[1] = a.c[1] * f; a.c[2] = a.c[2] * f; return a; }   // mul_float :543-551
inline Texel tx_add(Texel a, Texel b) { a.c[0] = a.c[0] + b.c[0]; a.c[1] = a.c[1] + b.c[1]; a.c[2] = a.c[2] + b.c[2]; return a; }   // add_two :554-559
inline Texel tx_lerp(Texel a, Texel b, Float t) {                                  // approx_lerp :530-540
    Texel r; for (int k = 0; k < 3; k++) r.c[k] = a.c[k] * (1.f - t) + b.c[k] * t; return r;
}
inline Texel tx_fetch(const TexView& v, uint32_t level, uint64_t x, uint64_t y) {
    const arn_texture& t = *v.t;
    const Float* p = v.texels + t.level_offset[level] + ((size_t)y * t.level_w[level] + (size_t)x) * t.channels;
    Texel r = tx_zero();
    for (uint32_t k = 0; k < t.channels; k++) r.c[k] = p[k];
    return r;
}
inline Texel texel_usize(const TexView& v, uint32_t level, uint64_t x, uint64_t y) {   // MipMap::texel (:381-403), p: Point2<usize>
    uint64_t dx = v.t->level_w[level], dy = v.t->level_h[level];
    if (x >= dx || y >= dy) {
        switch (v.t->wrapping) {
        case ARN_WRAP_BLACK: return tx_zero();
        case ARN_WRAP_CLAMP: x = x >= dx ? dx - 1 : x; y = y >= dy ? dy - 1 : y; break;
        default: x = x % dx; y = y % dy; break;
        }
    }
    return tx_fetch(v, level, x, y);
}
inline Texel texel_isize(const TexView& v, uint32_t level, int64_t x, int64_t y) {     // MipMap::texel_isize (:352-378)
    uint64_t dx = v.t->level_w[level], dy = v.t->level_h[level];
    uint64_t ux = (uint64_t)x, uy = (uint64_t)y;
    if (ux >= dx || uy >= dy) {
        switch (v.t->wrapping) {
        case ARN_WRAP_BLACK: return tx_zero();
        case ARN_WRAP_CLAMP: ux = ux >= dx ? dx - 1 : ux; uy = uy >= dy ? dy - 1 : uy; break;      // sic: a negative index clamps to the FAR edge
        default: { int64_t rx = x % (int64_t)dx, ry = y % (int64_t)dy; ux = (uint64_t)(rx < 0 ? -rx : rx); uy = (uint64_t)(ry < 0 ? -ry : ry); break; }   // sic: |remainder|
        }
    }
    return tx_fetch(v, level, ux, uy);
}
inline uint64_t f2usize(Float f) { return (uint64_t)(int64_t)f; }                  // see the header: x86-64 semantics of `as usize`
inline Texel triangle_filter(const TexView& v, uint32_t level, V2 st) {             // :427-445
    Float nx = (Float)v.t->level_w[level], ny = (Float)v.t->level_h[level];
    Float s = st.x * nx - 0.5f, t = st.y * ny - 0.5f;
    uint64_t s0 = f2usize(std::floor(s)), t0 = f2usize(std::floor(t));
    Float ds = s - std::floor(s), dt = t - std::floor(t);
    return tx_add(tx_add(tx_mul(texel_usize(v, level, s0, t0), (1.f - ds) * (1.f - dt)), tx_mul(texel_usize(v, level, s0, t0 + 1), (1.f - ds) * dt)),
                  tx_add(tx_mul(texel_usize(v, level, s0 + 1, t0), ds * (1.f - dt)), tx_mul(texel_usize(v, level, s0 + 1, t0 + 1), ds * dt)));
}
inline Float flog2(Float x) { return (Float)std::log2((double)x); }
inline Float find_level(const TexView& v, Float width) {                            // :522-527 (sic: (levels - 1) * log2(width))
    Float w = flog2(fmax_(width, 1e-8f));
    return (Float)(v.t->n_levels - 1) * w;
}
inline Float ewa_weight(uint32_t i) {                                               // WEIGHT_LUT (:609-621)
    const Float alpha = 2.f;
    Float r2 = (Float)i / (Float)(128 - 1);
    return fexp(-alpha * r2) - fexp(-alpha);
}
inline Texel ewa_filter(const TexView& v, uint32_t level, V2 st, V2 dstmaj, V2 dstmin) {   // :478-519
    const arn_texture& t = *v.t;
    if (level >= t.n_levels) return texel_usize(v, t.n_levels - 1, 0, 0);
    Float nxf = (Float)t.level_w[level], nyf = (Float)t.level_h[level];
    Float s = st.x * nxf - 0.5f, tt0 = st.y * nyf - 0.5f;
    Float dmins = dstmin.x * nxf, dmint = dstmin.y * nyf, dmajs = dstmaj.x * nxf, dmajt = dstmaj.y * nyf;
    Float a = dmint * dmint + dmajt * dmajt + 1.f;
    Float b = -2.f * (dmins * dmint + dmajs * dmajt);
    Float c = dmins * dmins + dmajs * dmajs + 1.f;
    Float inv_f = 1.f / (a * c - b * b * 0.25f);
    a *= inv_f; b *= inv_f; c *= inv_f;
    Float det = -b * b + 4.f * a * c;
    Float inv2_det = 1.f / det * 2.f;
    Float usqrt = std::sqrt(det * c), vsqrt = std::sqrt(det * a);
    int64_t s0 = (int64_t)std::ceil(s - inv2_det * usqrt), s1 = (int64_t)std::ceil(s + inv2_det * usqrt);
    int64_t t0 = (int64_t)std::ceil(tt0 - inv2_det * vsqrt), t1 = (int64_t)std::ceil(tt0 + inv2_det * vsqrt);
    // GUARD (deviation, both sides): a degenerate footprint makes the source loop over up to 2^64 texels; the box is cut to
    // +-ARN_EWA_MAX_RADIUS texels around (s, t).  Never reached by footprints below 64 texels.
    { const int64_t R = 64, cs = (int64_t)std::floor(s), ct = (int64_t)std::floor(tt0);
      if (!(s0 >= cs - R)) s0 = cs - R; if (!(s1 <= cs + R)) s1 = cs + R; if (!(t0 >= ct - R)) t0 = ct - R; if (!(t1 <= ct + R)) t1 = ct + R; }
    Texel sum = tx_zero(); Float sumwt = 0.f;
    for (int64_t it = t0; it < t1 + 1; it++) {
        Float tt = (Float)it - tt0;
        for (int64_t is = s0; is < s1 + 1; is++) {
            Float ss = (Float)is - s;
            Float square_radius = a * ss * ss + b * ss * tt + c * tt * tt;
            if (square_radius < 1.f) {
                uint64_t idx = f2usize(square_radius * 128.f);
                if (idx > 127) idx = 127;
                Float weight = ewa_weight((uint32_t)idx);
                sum = tx_add(sum, tx_mul(texel_isize(v, level, is, it), weight));
                sumwt += weight;
            }
        }
    }
    return tx_mul(sum, 1.f / sumwt);
}
inline Texel look_up(const TexView& v, V2 st, V2 dst0, V2 dst1) {                    // :447-476 with look_up_tri :411-425
    const arn_texture& t = *v.t;
    if (t.trilinear) {
        Float width = fmax_(fmax_(fmax_(dst0.x, dst0.y), dst1.x), dst1.y);
        Float level = find_level(v, width);
        if (level < 0.f) return triangle_filter(v, 0, st);
        if (level >= (Float)(t.n_levels - 1)) return triangle_filter(v, t.n_levels - 1, st);
        Float fl = std::floor(level); uint32_t flu = (uint32_t)fl; Float delta = level - fl;
        return tx_lerp(triangle_filter(v, flu, st), triangle_filter(v, flu + 1, st), delta);
    }
    V2 dstmin, dstmaj;
    if (dst0.x * dst0.x + dst0.y * dst0.y < dst1.x * dst1.x + dst1.y * dst1.y) { dstmin = dst0; dstmaj = dst1; } else { dstmin = dst1; dstmaj = dst0; }
    Float minor = std::sqrt(dstmin.x * dstmin.x + dstmin.y * dstmin.y), major = std::sqrt(dstmaj.x * dstmaj.x + dstmaj.y * dstmaj.y);
    if (minor == 0.f) return triangle_filter(v, 0, st);
    if (minor * t.max_aniso < major) { Float scale = major / (minor * t.max_aniso); minor *= scale; dstmin = dstmin * scale; }
    Float level = fmax_(find_level(v, minor), 0.f);
    Float fl = std::floor(level); Float delta = level - fl; uint32_t lv = (uint32_t)fl;
    return tx_lerp(ewa_filter(v, lv, st, dstmaj, dstmin), ewa_filter(v, lv + 1, st, dstmaj, dstmin), delta);
}
// ImageTexture::evaluate (image.rs:66-68,84-86) through UVMapping::map (mappings.rs:21-30)
inline Texel texture_evaluate(const TexView& v, V2 uv, const DxyInfo& dxy) {
    const arn_texture& t = *v.t;
    V2 p = v2(uv.x * t.scale_u + t.shift_u, uv.y * t.scale_v + t.shift_v);
    V2 dpdx = v2(t.scale_u * dxy.dudx, t.scale_v * dxy.dvdx), dpdy = v2(t.scale_u * dxy.dudy, t.scale_v * dxy.dvdy);
    return look_up(v, p, dpdx, dpdy);
}

// add_bumping (material/mod.rs:42-86).  The shifted interaction only differs in uv for a UV-mapped texture.
inline void add_bumping(SurfaceInteraction& si, const DxyInfo& dxy, const TexView& bump) {
    Float du = 0.5f * (std::fabs(dxy.dudx) + std::fabs(dxy.dudy));
    if (du == 0.f) du = 0.0005f;
    Float displacement_u = texture_evaluate(bump, v2(si.uv.x + du, si.uv.y), dxy).c[0];
    Float dv = 0.5f * (std::fabs(dxy.dvdx) + std::fabs(dxy.dvdy));
    if (dv == 0.f) dv = 0.0005f;
    Float displacement_v = texture_evaluate(bump, v2(si.uv.x + du, si.uv.y + dv), dxy).c[0];   // sic: `sie` keeps the u shift (mod.rs:63-66)
    Float displacement = texture_evaluate(bump, si.uv, dxy).c[0];
    V3 dpdu = si.shading_duv.dpdu + (displacement_u - displacement) / du * si.shading_norm + displacement * si.shading_duv.dndu;
    V3 dpdv = si.shading_duv.dpdv + (displacement_v - displacement) / dv * si.shading_norm + displacement * si.shading_duv.dndv;
    DuvInfo d; d.dpdu = dpdu; d.dpdv = dpdv; d.dndu = si.shading_duv.dndu; d.dndv = si.shading_duv.dndv;
    si_set_shading(si, d, false);
}

}  // namespace orc
