// Image textures, UV mapping, image-plane differentials and bump mapping on the device: the shade
// kernel's textured instance.  Replaces, per hit: SurfaceInteraction::compute_dxy + solve_over_constrained_2x3
// (src/geometry/interaction.rs:204-224,308-325), UVMapping::map (src/texturing/mappings.rs:21-30), MipMap::look_up /
// look_up_tri / triangle_filter / ewa_filter / find_level / texel (src/texturing/textures/image.rs:352-527), add_bumping
// (src/material/mod.rs:42-86), spawn_ray_differential (interaction.rs:236-251).  Arithmetic order follows the Rust source
// (and oracle/texture.hpp, which lists the third-party semantics assumed); pyramid levels arrive ready-made (arn_texture).
#pragma once
#include "shade.cuh"

namespace arn {

struct DxyInfo { float3 dpdx, dpdy; float dudx, dvdx, dudy, dvdy; };
ARN_NOINL float cr_log2f(float x) { return (float)log2((double)x); }

ARN_DEV bool m2_solve(float c0r0, float c0r1, float c1r0, float c1r1, float2 v, float2& out) {     // cgmath Matrix2::invert + Matrix2 * Vector2
    float det = c0r0 * c1r1 - c1r0 * c0r1;
    if (det == 0.f) return false;
    float i00 = c1r1 / det, i01 = -c0r1 / det, i10 = -c1r0 / det, i11 = c0r0 / det;
    out = f2(i00 * v.x + i10 * v.y, i01 * v.x + i11 * v.y);
    return true;
}
ARN_DEV float2 solve_2x3(float3 abc, float3 m0, float3 m1, float3 n) {
    float2 r = f2(0.f, 0.f); bool ok;
    if (fabsf(n.x) > fabsf(n.y) && fabsf(n.x) > fabsf(n.z)) ok = m2_solve(m0.y, m1.y, m0.z, m1.z, f2(abc.y, abc.z), r);
    else if (fabsf(n.y) > fabsf(n.z)) ok = m2_solve(m0.x, m1.x, m0.z, m1.z, f2(abc.x, abc.z), r);
    else ok = m2_solve(m0.x, m1.x, m0.y, m1.y, f2(abc.x, abc.y), r);
    return ok ? r : f2(0.f, 0.f);
}
// pos / n = basic.{pos, norm}; duv_dpdu / duv_dpdv = si.duv (the SHADING frame for triangles, quirk A-7); rx / ry = the offset rays
ARN_DEV DxyInfo compute_dxy(float3 pos, float3 n, float3 duv_dpdu, float3 duv_dpdv, float3 rxo, float3 rxd, float3 ryo, float3 ryd) {
    float d = dot(n, pos);
    float tx = (d - dot(n, rxo)) / dot(n, rxd);
    float3 px = rxo + rxd * tx;
    float ty = (d - dot(n, ryo)) / dot(n, ryd);
    float3 py = ryo + ryd * ty;
    DxyInfo o; o.dpdx = px - pos; o.dpdy = py - pos;
    float2 dudxy = solve_2x3(o.dpdx, duv_dpdu, duv_dpdv, n), dvdxy = solve_2x3(o.dpdy, duv_dpdu, duv_dpdv, n);
    o.dudx = dudxy.x; o.dudy = dudxy.y; o.dvdx = dvdxy.x; o.dvdy = dvdxy.y;       // sic (interaction.rs:218-221)
    return o;
}

struct TexView { const arn_texture* t; const float* __restrict__ texels; };
ARN_DEV float3 tx_fetch(const TexView& v, uint32_t level, unsigned long long x, unsigned long long y) {
    const arn_texture& t = *v.t;
    const float* p = v.texels + t.level_offset[level] + ((size_t)y * t.level_w[level] + (size_t)x) * t.channels;
    return t.channels == 3u ? f3(__ldg(p), __ldg(p + 1), __ldg(p + 2)) : f3(__ldg(p), 0.f, 0.f);
}
ARN_DEV float3 texel_usize(const TexView& v, uint32_t level, unsigned long long x, unsigned long long y) {
    unsigned long long dx = v.t->level_w[level], dy = v.t->level_h[level];
    if (x >= dx || y >= dy) {
        if (v.t->wrapping == ARN_WRAP_BLACK) return f3(0.f, 0.f, 0.f);
        if (v.t->wrapping == ARN_WRAP_CLAMP) { x = x >= dx ? dx - 1 : x; y = y >= dy ? dy - 1 : y; }
        else { x = x % dx; y = y % dy; }
    }
    return tx_fetch(v, level, x, y);
}
ARN_DEV float3 texel_isize(const TexView& v, uint32_t level, long long x, long long y) {
    unsigned long long dx = v.t->level_w[level], dy = v.t->level_h[level];
    unsigned long long ux = (unsigned long long)x, uy = (unsigned long long)y;
    if (ux >= dx || uy >= dy) {
        if (v.t->wrapping == ARN_WRAP_BLACK) return f3(0.f, 0.f, 0.f);
        if (v.t->wrapping == ARN_WRAP_CLAMP) { ux = ux >= dx ? dx - 1 : ux; uy = uy >= dy ? dy - 1 : uy; }
        else { long long rx = x % (long long)dx, ry = y % (long long)dy; ux = (unsigned long long)(rx < 0 ? -rx : rx); uy = (unsigned long long)(ry < 0 ? -ry : ry); }
    }
    return tx_fetch(v, level, ux, uy);
}
ARN_DEV unsigned long long f2usize(float f) { return (unsigned long long)(long long)f; }
ARN_NOINL float3 triangle_filter(const TexView& v, uint32_t level, float2 st) {
    float nx = (float)v.t->level_w[level], ny = (float)v.t->level_h[level];
    float s = st.x * nx - 0.5f, t = st.y * ny - 0.5f;
    unsigned long long s0 = f2usize(floorf(s)), t0 = f2usize(floorf(t));
    float ds = s - floorf(s), dt = t - floorf(t);
    return (texel_usize(v, level, s0, t0) * ((1.f - ds) * (1.f - dt)) + texel_usize(v, level, s0, t0 + 1) * ((1.f - ds) * dt))
         + (texel_usize(v, level, s0 + 1, t0) * (ds * (1.f - dt)) + texel_usize(v, level, s0 + 1, t0 + 1) * (ds * dt));
}
ARN_DEV float find_level(const TexView& v, float width) { return (float)(v.t->n_levels - 1) * cr_log2f(fmaxf(width, 1e-8f)); }
ARN_DEV float3 tx_lerp(float3 a, float3 b, float t) { return f3(a.x * (1.f - t) + b.x * t, a.y * (1.f - t) + b.y * t, a.z * (1.f - t) + b.z * t); }
ARN_NOINL float3 ewa_filter(const TexView& v, uint32_t level, float2 st, float2 dstmaj, float2 dstmin) {
    const arn_texture& t = *v.t;
    if (level >= t.n_levels) return texel_usize(v, t.n_levels - 1, 0, 0);
    float nxf = (float)t.level_w[level], nyf = (float)t.level_h[level];
    float s = st.x * nxf - 0.5f, tt0 = st.y * nyf - 0.5f;
    float dmins = dstmin.x * nxf, dmint = dstmin.y * nyf, dmajs = dstmaj.x * nxf, dmajt = dstmaj.y * nyf;
    float a = dmint * dmint + dmajt * dmajt + 1.f;
    float b = -2.f * (dmins * dmint + dmajs * dmajt);
    float c = dmins * dmins + dmajs * dmajs + 1.f;
    float inv_f = 1.f / (a * c - b * b * 0.25f);
    a *= inv_f; b *= inv_f; c *= inv_f;
    float det = -b * b + 4.f * a * c;
    float inv2_det = 1.f / det * 2.f;
    float usqrt = sqrtf(det * c), vsqrt = sqrtf(det * a);
    long long s0 = (long long)ceilf(s - inv2_det * usqrt), s1 = (long long)ceilf(s + inv2_det * usqrt);
    long long t0 = (long long)ceilf(tt0 - inv2_det * vsqrt), t1 = (long long)ceilf(tt0 + inv2_det * vsqrt);
    {   // GUARD (deviation shared with the oracle): cut degenerate footprints to +-64 texels around (s, t)
        const long long R = 64, cs = (long long)floorf(s), ct = (long long)floorf(tt0);
        if (!(s0 >= cs - R)) s0 = cs - R; if (!(s1 <= cs + R)) s1 = cs + R; if (!(t0 >= ct - R)) t0 = ct - R; if (!(t1 <= ct + R)) t1 = ct + R;
    }
    float3 sum = f3(0.f, 0.f, 0.f); float sumwt = 0.f;
    for (long long it = t0; it < t1 + 1; it++) {
        float tt = (float)it - tt0;
        for (long long is = s0; is < s1 + 1; is++) {
            float ss = (float)is - s;
            float square_radius = a * ss * ss + b * ss * tt + c * tt * tt;
            if (square_radius < 1.f) {
                unsigned long long idx = f2usize(square_radius * 128.f);
                if (idx > 127) idx = 127;
                float r2 = (float)(uint32_t)idx / (float)(128 - 1);
                float weight = cr_expf(-2.f * r2) - cr_expf(-2.f);             // WEIGHT_LUT[idx] (image.rs:609-621)
                sum = sum + texel_isize(v, level, is, it) * weight;
                sumwt += weight;
            }
        }
    }
    return sum * (1.f / sumwt);
}
ARN_NOINL float3 texture_evaluate(const TexView& v, float2 uv, const DxyInfo& dxy) {
    const arn_texture& t = *v.t;
    float2 st = f2(uv.x * t.scale_u + t.shift_u, uv.y * t.scale_v + t.shift_v);
    float2 dst0 = f2(t.scale_u * dxy.dudx, t.scale_v * dxy.dvdx), dst1 = f2(t.scale_u * dxy.dudy, t.scale_v * dxy.dvdy);
    if (t.trilinear) {
        float width = fmaxf(fmaxf(fmaxf(dst0.x, dst0.y), dst1.x), dst1.y);
        float level = find_level(v, width);
        if (level < 0.f) return triangle_filter(v, 0, st);
        if (level >= (float)(t.n_levels - 1)) return triangle_filter(v, t.n_levels - 1, st);
        float fl = floorf(level); uint32_t flu = (uint32_t)fl; float delta = level - fl;
        return tx_lerp(triangle_filter(v, flu, st), triangle_filter(v, flu + 1, st), delta);
    }
    float2 dstmin, dstmaj;
    if (dst0.x * dst0.x + dst0.y * dst0.y < dst1.x * dst1.x + dst1.y * dst1.y) { dstmin = dst0; dstmaj = dst1; } else { dstmin = dst1; dstmaj = dst0; }
    float minor = sqrtf(dstmin.x * dstmin.x + dstmin.y * dstmin.y), major = sqrtf(dstmaj.x * dstmaj.x + dstmaj.y * dstmaj.y);
    if (minor == 0.f) return triangle_filter(v, 0, st);
    if (minor * t.max_aniso < major) { float scale = major / (minor * t.max_aniso); minor *= scale; dstmin = f2(dstmin.x * scale, dstmin.y * scale); }
    float level = fmaxf(find_level(v, minor), 0.f);
    float fl = floorf(level); float delta = level - fl; uint32_t lv = (uint32_t)fl;
    return tx_lerp(ewa_filter(v, lv, st, dstmaj, dstmin), ewa_filter(v, lv + 1, st, dstmaj, dstmin), delta);
}

// add_bumping (material/mod.rs:42-86) for a UV-mapped displacement texture: updates the shading normal and, through
// set_shading(.., false) (interaction.rs:167-182), possibly the side of the geometric normal
ARN_DEV void add_bumping(Surf& s, const SurfTex& x, const DxyInfo& dxy, const TexView& bump) {
    float du = 0.5f * (fabsf(dxy.dudx) + fabsf(dxy.dudy));
    if (du == 0.f) du = 0.0005f;
    float displacement_u = texture_evaluate(bump, f2(x.uv.x + du, x.uv.y), dxy).x;
    float dv = 0.5f * (fabsf(dxy.dvdx) + fabsf(dxy.dvdy));
    if (dv == 0.f) dv = 0.0005f;
    float displacement_v = texture_evaluate(bump, f2(x.uv.x + du, x.uv.y + dv), dxy).x;      // sic: the u shift stays (mod.rs:63-66)
    float displacement = texture_evaluate(bump, x.uv, dxy).x;
    float3 dpdu = s.dpdu + (displacement_u - displacement) / du * s.ns + displacement * x.sh_dndu;
    float3 dpdv = x.sh_dpdv + (displacement_v - displacement) / dv * s.ns + displacement * x.sh_dndv;
    float3 norm = normalize(cross(dpdu, dpdv));
    if (dot(s.ng, norm) < 0.f) s.ng = -s.ng;
    s.ns = norm;
}

// roughness_to_alpha (bxdf/microfacet.rs:57-63) for a textured roughness (constant roughness: evaluated once on the host)
ARN_DEV float roughness_to_alpha_dev(float roughness) {
    float r = fmaxf(roughness, 1e-3f);
    float x = cr_logf(r);
    return 1.62142f + 0.819955f * x + 0.1734f * x * x + 0.0171201f * x * x * x + 0.000640711f * x * x * x * x;
}

// Material parameters of a hit with the scene's image textures applied (bump first: material/matte.rs:46-48 and alike)
ARN_DEV arn_material textured_material(const DevScene& sc, uint32_t mat, Surf& s, const SurfTex& x, const DxyInfo& dxy) {
    arn_material m = sc.materials[mat];
    if (m.bump_tex) { TexView v; v.t = &sc.textures[m.bump_tex - 1]; v.texels = sc.texels; add_bumping(s, x, dxy, v); }
    if (m.kd_tex) { TexView v; v.t = &sc.textures[m.kd_tex - 1]; v.texels = sc.texels; float3 t = texture_evaluate(v, x.uv, dxy); m.kd[0] = t.x; m.kd[1] = t.y; m.kd[2] = t.z; }
    if (m.ks_tex) { TexView v; v.t = &sc.textures[m.ks_tex - 1]; v.texels = sc.texels; float3 t = texture_evaluate(v, x.uv, dxy); m.ks[0] = t.x; m.ks[1] = t.y; m.ks[2] = t.z; }
    if (m.aux_tex) {
        TexView v; v.t = &sc.textures[m.aux_tex - 1]; v.texels = sc.texels;
        float a = texture_evaluate(v, x.uv, dxy).x;
        if (m.type == ARN_MAT_MATTE) m.sigma = a; else { m.roughness = a; m.alpha = roughness_to_alpha_dev(a); }
    }
    return m;
}

}  // namespace arn
