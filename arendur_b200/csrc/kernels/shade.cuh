// Device-side appearance code for the shade kernel (K3): surface-interaction reconstruction
// for the FINAL hit only, BxDFs, the Bsdf mixture, the four materials, sphere area lights.
//
// Replaces, per hit: TriangleInstance::intersect_ray after acceptance + computedpduv +
// compute_shading_at (src/shape/triangle.rs:453-484,308-376), Sphere::intersect_ray's
// differential part (src/shape/sphere.rs:250-290), SurfaceInteraction::{new,set_shading,
// apply_transform} (src/geometry/interaction.rs:133-201), Bsdf (src/material/bsdf.rs),
// the BxDFs (src/bxdf/*.rs) and materials (src/material/*.rs) with constant textures,
// the area-light methods of ShapedPrimitive / TransformedComposable
// (src/component/shape.rs:74-168, transformed.rs:103-158).
// Arithmetic order follows the Rust source; see dev_math.cuh for the rounding contract.
#pragma once
#include "traverse.cuh"

namespace arn {

#define BXDF_REFLECTION 0x01u
#define BXDF_TRANSMISSION 0x02u
#define BXDF_DIFFUSE 0x04u
#define BXDF_GLOSSY 0x08u
#define BXDF_SPECULAR 0x10u
#define BXDF_ALL 0x1fu

// what the integrator reads from a SurfaceInteraction
struct Surf {
    float3 pos, perr, wo, ng;     // basic.{pos,pos_err,wo,norm}
    float3 ns;                    // shading_norm
    float3 dpdu;                  // shading_duv.dpdu == the GEOMETRIC dpdu (quirk A-7)
};

ARN_DEV float3 ld3(const float* __restrict__ p, uint32_t i) { return f3(__ldg(p + 3 * i), __ldg(p + 3 * i + 1), __ldg(p + 3 * i + 2)); }
ARN_DEV float2 ld2(const float* __restrict__ p, uint32_t i) { return f2(__ldg(p + 2 * i), __ldg(p + 2 * i + 1)); }

// cgmath Matrix3::look_at(dir, up).{x, z} (triangle.rs:321-323)
ARN_DEV void look_at_xz(float3 dir, float3 up, float3& cx, float3& cz) {
    float3 d = normalize(dir);
    float3 side = normalize(cross(up, d));
    float3 u = normalize(cross(d, side));
    cx = f3(side.x, u.x, d.x); cz = f3(side.z, u.z, d.z);
}
ARN_DEV void computedpduv(float3 p0, float3 p1, float3 p2, float2 uv0, float2 uv1, float2 uv2, float3& dpdu, float3& dpdv) {
    float2 duv02 = f2(uv0.x - uv2.x, uv0.y - uv2.y), duv12 = f2(uv1.x - uv2.x, uv1.y - uv2.y);
    float3 dp02 = p0 - p2, dp12 = p1 - p2;
    float determinant = duv02.x * duv12.y - duv02.y * duv12.x;
    if (determinant == 0.f) {
        float3 up = cross(dp02, p0 - p1);
        look_at_xz(dp02, up, dpdu, dpdv);
    } else {
        float inv = 1.f / determinant;
        dpdu = (duv12.y * dp02 - duv02.y * dp12) * inv;
        dpdv = (-duv12.x * dp02 + duv02.x * dp12) * inv;
    }
}
ARN_DEV void get_basis_from(float3 dir, float3& u, float3& v) {         // foundamental.rs:296-305
    float3 up = f3(0.f, 0.f, 1.f);
    if (relative_eq(up.x, dir.x) && relative_eq(up.y, dir.y) && relative_eq(up.z, dir.z)) up = f3(0.f, 1.f, 0.f);
    u = normalize(cross(up, dir));
    v = normalize(cross(dir, u));
}

// what the textured shade instance needs beyond Surf (kernels/shade_tex.cuh)
struct SurfTex {
    float2 uv;                    // si.uv
    float3 duv_dpdu, duv_dpdv;    // si.duv: the shading frame for triangles (set_shading writes duv, quirk A-7), the (transformed) geometric one for spheres
    float3 sh_dpdv, sh_dndu, sh_dndv;   // si.shading_duv beyond dpdu (= Surf.dpdu): geometric dpdv; dn = 0 for triangles, Weingarten for spheres
};

// triangle hit -> Surf (triangle.rs:453-484 with interaction.rs:133-182)
ARN_DEV void surf_triangle(const DevScene& sc, uint32_t tri, float b0, float b1, float b2, float3 raydir, Surf& s, SurfTex* x = nullptr) {
    uint32_t i0 = __ldg(&sc.indices[3 * tri]), i1 = __ldg(&sc.indices[3 * tri + 1]), i2 = __ldg(&sc.indices[3 * tri + 2]);
    float3 p0 = ld3(sc.positions, i0), p1 = ld3(sc.positions, i1), p2 = ld3(sc.positions, i2);
    arn_mesh mesh = sc.meshes[__ldg(&sc.tri_mesh[tri])];
    float2 uv0, uv1, uv2;
    if (mesh.has_uvs) { uv0 = ld2(sc.uvs, i0); uv1 = ld2(sc.uvs, i1); uv2 = ld2(sc.uvs, i2); }
    else { uv0 = f2(0.f, 0.f); uv1 = f2(1.f, 0.f); uv2 = f2(1.f, 1.f); }
    s.pos = b0 * p0 + b1 * p1 + b2 * p2;
    s.perr = gamma_n(7.f) * f3(fabsf(b0 * p0.x) + fabsf(b1 * p1.x) + fabsf(b2 * p2.x),
                               fabsf(b0 * p0.y) + fabsf(b1 * p1.y) + fabsf(b2 * p2.y),
                               fabsf(b0 * p0.z) + fabsf(b1 * p1.z) + fabsf(b2 * p2.z));
    float3 dpdu, dpdv; computedpduv(p0, p1, p2, uv0, uv1, uv2, dpdu, dpdv);
    s.wo = -raydir;
    s.ng = normalize(cross(dpdu, dpdv));
    s.dpdu = dpdu;
    // compute_shading_at (triangle.rs:333-376)
    float3 shading_normal;
    if (mesh.has_normals) {
        float3 n0 = ld3(sc.normals, i0), n1 = ld3(sc.normals, i1), n2 = ld3(sc.normals, i2);
        shading_normal = normalize(b0 * n0 + b1 * n1 + b2 * n2);
    } else shading_normal = normalize(cross(p2 - p0, p1 - p0));
    float3 st = normalize(dpdu);
    float3 sbt = cross(st, shading_normal);
    if (length2(sbt) > 0.f) { sbt = normalize(sbt); st = cross(sbt, shading_normal); }
    else get_basis_from(shading_normal, st, sbt);
    // set_shading(.., true) (interaction.rs:167-182)
    float3 n = normalize(cross(st, sbt));
    if (dot(s.ng, n) < 0.f) n = -n;
    s.ns = n;
    if (x) {
        float2 a = f2(b0 * uv0.x, b0 * uv0.y), b = f2(b1 * uv1.x, b1 * uv1.y), c = f2(b2 * uv2.x, b2 * uv2.y);
        x->uv = f2((a.x + b.x) + c.x, (a.y + b.y) + c.y);                // uvhit (triangle.rs:465)
        x->duv_dpdu = st; x->duv_dpdv = sbt;
        x->sh_dpdv = dpdv; x->sh_dndu = f3(0.f, 0.f, 0.f); x->sh_dndv = f3(0.f, 0.f, 0.f);
    }
}

// sphere hit (local refined point p) -> Surf in world space
// (sphere.rs:250-290, interaction.rs:133-162,190-201, transformed.rs:73-83)
ARN_DEV void surf_sphere(const DevSphere& sp, float3 p, float3 raydir_after, Surf& s, SurfTex* x = nullptr) {
    float phimax = sp.phimax;
    float thetadelta = sp.thetamax - sp.thetamin;
    float theta = cr_acosf(p.z / sp.radius);
    float inv_z_radius = 1.f / sqrtf(p.x * p.x + p.y * p.y);
    float cos_phi = p.x * inv_z_radius, sin_phi = p.y * inv_z_radius;
    float3 dpdu = f3(-phimax * p.y, phimax * p.x, 0.f);
    float3 dpdv = thetadelta * f3(p.z * cos_phi, p.z * sin_phi, -sp.radius * cr_sinf(theta));
    float3 n = normalize(cross(dpdu, dpdv));
    if (sp.has_transform) {
        s.pos = xform_point(sp.local_parent, p);
        s.ng = normalize(xform_vector_T(sp.parent_local, n));      // transform_norm: inverse-transpose, normalised
        s.ns = s.ng;
        s.dpdu = xform_vector(sp.local_parent, dpdu);
    } else { s.pos = p; s.ng = n; s.ns = n; s.dpdu = dpdu; }
    if (x) {                                                        // uv and the Weingarten dn (sphere.rs:252-279), transformed like the rest
        float phi = cr_atan2f(p.y, p.x);
        if (phi < 0.f) phi += 2.f * ARN_PI;
        x->uv = f2(phi / phimax, (theta - sp.thetamin) / thetadelta);
        float3 dppduu = -phimax * phimax * f3(p.x, p.y, 0.f);
        float3 dppduv = thetadelta * p.z * phimax * f3(-sin_phi, cos_phi, 0.f);
        float3 dppdvv = -thetadelta * thetadelta * f3(p.x, p.y, p.z);
        float e = dot(dpdu, dpdu), f = dot(dpdu, dpdv), g = dot(dpdv, dpdv);
        float ee = dot(n, dppduu), ff = dot(n, dppduv), gg = dot(n, dppdvv);
        float inv = 1.f / (e * g - f * f);
        float3 dndu = (ff * f - ee * g) * inv * dpdu + (ee * f - ff * e) * inv * dpdv;
        float3 dndv = (gg * f - ff * g) * inv * dpdu + (ff * f - gg * e) * inv * dpdv;
        if (sp.has_transform) {                                     // DuvInfo::apply_transform (interaction.rs:92-101)
            x->sh_dpdv = xform_vector(sp.local_parent, dpdv);
            x->sh_dndu = normalize(xform_vector_T(sp.parent_local, dndu)); x->sh_dndv = normalize(xform_vector_T(sp.parent_local, dndv));
        } else { x->sh_dpdv = dpdv; x->sh_dndu = dndu; x->sh_dndv = dndv; }
        x->duv_dpdu = s.dpdu; x->duv_dpdv = x->sh_dpdv;
    }
    s.perr = f3(0.f, 0.f, 0.f);                                     // "FIXME: wrong" (sphere.rs:281-282)
    s.wo = -raydir_after;       // local_parent * (-(parent_local * d)) == -(ray direction after the round trip)
}

// InteractInfo::offset_towards (interaction.rs:45-72)
ARN_DEV float3 offset_towards(const Surf& s, float3 dir) {
    float3 nabs = f3(fabsf(s.ng.x), fabsf(s.ng.y), fabsf(s.ng.z));
    float edn = dot(nabs, s.perr);
    float3 offset = edn * s.ng;
    if (dot(dir, s.ng) <= 0.f) offset = -offset;
    float3 ret = s.pos + offset;
    if (offset.x > 0.f) ret.x = next_up(ret.x); else if (offset.x < 0.f) ret.x = next_down(ret.x);
    if (offset.y > 0.f) ret.y = next_up(ret.y); else if (offset.y < 0.f) ret.y = next_down(ret.y);
    if (offset.z > 0.f) ret.z = next_up(ret.z); else if (offset.z < 0.f) ret.z = next_down(ret.z);
    return ret;
}

// ---------------------------------------------------------------- `normal` helpers
ARN_DEV float cos_theta(float3 n) { return n.z; }
ARN_DEV float cos2_theta(float3 n) { return n.z * n.z; }
ARN_DEV float sin2_theta(float3 n) { return fabsf(1.f - cos2_theta(n)); }
ARN_DEV float sin_theta(float3 n) { return sqrtf(sin2_theta(n)); }
ARN_DEV float tan_theta(float3 n) { return sin_theta(n) / cos_theta(n); }
ARN_DEV float tan2_theta(float3 n) { return sin2_theta(n) / cos2_theta(n); }
ARN_DEV float cos_phi(float3 n) { float st = sin_theta(n); return st == 0.f ? 1.f : clampf(n.x / st, -1.f, 1.f); }
ARN_DEV float sin_phi(float3 n) { float st = sin_theta(n); return st == 0.f ? 0.f : clampf(n.y / st, -1.f, 1.f); }
ARN_DEV float cos2_phi(float3 n) { float c = cos_phi(n); return c * c; }
ARN_DEV float sin2_phi(float3 n) { float s = sin_phi(n); return s * s; }
ARN_DEV bool refract(float3 wo, float3 n, float eta, float3& out) {   // foundamental.rs:278-292
    float ct = dot(wo, n);
    float s2 = 1.f - ct * ct;
    float s2t = eta * eta * fmaxf(s2, 0.f);
    if (s2t >= 1.f) return false;
    float ctt = sqrtf(1.f - s2t);
    out = -eta * wo + (eta * ct - ctt) * n;
    return true;
}

// ---------------------------------------------------------------- warps (sample/mod.rs)
ARN_DEV float2 sample_concentric_disk(float2 u) {
    float2 w = f2(2.f * u.x - 1.f, 2.f * u.y - 1.f);
    if (w.x == 0.f && w.y == 0.f) return f2(0.f, 0.f);
    float r, theta;
    if (fabsf(w.x) > fabsf(w.y)) { r = w.x; theta = ARN_PI_4 * (w.y / w.x); }
    else { r = w.y; theta = ARN_PI_2 - ARN_PI_4 * (w.x / w.y); }
    float st, ct; cr_sincosf(theta, st, ct);
    return f2(r * ct, r * st);
}
ARN_DEV float3 sample_cosw_hemisphere(float2 u) {
    float2 d = sample_concentric_disk(u);
    float z = sqrtf(fabsf(1.f - d.x * d.x - d.y * d.y));
    return f3(d.x, d.y, z);
}
ARN_DEV float power_heuristic(float pdff, float pdfg) {
    float f = 1.f * pdff, g = 1.f * pdfg;
    return (f * f) / (f * f + g * g);
}

// ---------------------------------------------------------------- microfacet (bxdf/microfacet.rs)
ARN_DEV float powi5(float x) { float x2 = x * x; float x4 = x2 * x2; return x * x4; }   // llvm.powi(x, 5)
ARN_DEV float erf_inv(float x) {
    x = fminf(fmaxf(x, -0.99999f), 0.99999f);
    float w = -cr_logf((1.f - x) * (1.f + x));
    float p;
    if (w < 5.f) {
        w = w - 2.5f;
        p = 2.81022636e-08f; p = 3.43273939e-07f + p * w; p = -3.5233877e-06f + p * w; p = -4.39150654e-06f + p * w;
        p = 0.00021858087f + p * w; p = -0.00125372503f + p * w; p = -0.00417768164f + p * w; p = 0.246640727f + p * w;
        p = 1.50140941f + p * w;
    } else {
        w = sqrtf(w) - 3.f;
        p = -0.000200214257f; p = 0.000100950558f + p * w; p = 0.00134934322f + p * w; p = -0.00367342844f + p * w;
        p = 0.00573950773f + p * w; p = -0.0076224613f + p * w; p = 0.00943887047f + p * w; p = 1.00167406f + p * w;
        p = 2.83297682f + p * w;
    }
    return p * x;
}
ARN_DEV float erf_approx(float x) {
    const float A1 = 0.254829592f, A2 = -0.28449673f, A3 = 1.421413741f, A4 = -1.453152027f, A5 = 1.061405429f, P = 0.3275911f;
    float sign = signum(x);
    x = x * sign;
    float t = 1.f / (1.f + P * x);
    float y = 1.f - (((((A5 * t + A4) * t) + A3) * t + A2) * t + A1) * t * cr_expf(-x * x);
    return sign * y;
}
template <bool BECK> ARN_DEV float dist_D(float ax, float ay, float3 wh) {
    float c2t = cos2_theta(wh), t2t = tan2_theta(wh);
    if (BECK) {
        float c2p = cos2_phi(wh), s2p = sin2_phi(wh);
        return cr_expf(-t2t * (c2p / (ax * ax) + s2p / (ay * ay))) / (ARN_PI * ax * ay * c2t * c2t);
    }
    if (isinf(t2t)) return 0.f;
    float c2p = cos2_phi(wh), s2p = sin2_phi(wh);
    float last_term = 1.f + t2t * (c2p / (ax * ax) + s2p / (ay * ay));
    return 1.f / (ARN_PI * ax * ay * c2t * c2t * last_term * last_term);
}
template <bool BECK> ARN_DEV float dist_lambda(float ax, float ay, float3 w) {
    if (BECK) {
        float tant = fabsf(tan_theta(w));
        if (isinf(tant) || isnan(tant)) return 0.f;
        float alpha = sqrtf(cos2_phi(w) * ax * ax + sin2_phi(w) * ay * ay);
        float a = 1.f / (alpha * tant);
        if (a >= 1.6f) return 0.f;
        return (1.f - 1.259f * a + 0.396f * a * a) / (3.535f * a + 2.181f * a * a);
    }
    float tabs = fabsf(tan_theta(w));
    if (isinf(tabs)) return 0.f;
    float alpha = sqrtf(cos2_phi(w) * ax * ax + sin2_phi(w) * ay * ay);
    float term = alpha * tabs;
    return (-1.f + sqrtf(1.f + term * term)) * 0.5f;
}
template <bool BECK> ARN_DEV float dist_visible(float ax, float ay, float3 w) { return 1.f / (1.f + dist_lambda<BECK>(ax, ay, w)); }
template <bool BECK> ARN_DEV float dist_visible_both(float ax, float ay, float3 w0, float3 w1) {
    return 1.f / (1.f + dist_lambda<BECK>(ax, ay, w0) + dist_lambda<BECK>(ax, ay, w1));
}
template <bool BECK> ARN_DEV float dist_pdf(float ax, float ay, float3 wo, float3 wh) {
    return dist_D<BECK>(ax, ay, wh) * dist_visible<BECK>(ax, ay, wo) * fabsf(dot(wo, wh)) / fabsf(cos_theta(wo));
}
static __device__ __noinline__ float3 sample_wh_beckmann(float3 wo, float2 u, float ax, float ay) {
    float3 ws = normalize(f3(ax * wo.x, ay * wo.y, wo.z));
    float ct = fabsf(cos_theta(ws));
    float sx, sy;
    if (ct > 0.9999f) {
        float r = sqrtf(-cr_logf(u.x));
        float phi = 2.f * u.y * ARN_PI;
        float sphi, cphi; cr_sincosf(phi, sphi, cphi);
        sx = r * cphi; sy = r * sphi;
    } else {
        float st = sqrtf(fmaxf(1.f - ct * ct, 0.f));
        float tant = st / ct, cott = ct / st;
        float a = -1.f;
        float c = erf_approx(cott);
        float ux = fmaxf(u.x, 1e-6f);
        float theta = cr_acosf(ct);
        float fit = 1.f + theta * (-0.876f + theta * (0.4265f - 0.0594f * theta));
        float b = c - (1.f + c) * cr_powf(1.f - ux, fit);
        float sqrt_pi_inv = 1.f / sqrtf(ARN_PI);
        float norm = 1.f / (1.f + c + sqrt_pi_inv * tant * cr_expf(-cott * cott));
        for (int it = 1; it < 10; it++) {
            if (b < a || b > c) b = 0.5f * (a + c);
            float inv = erf_inv(b);
            float value = norm * (1.f + b + sqrt_pi_inv * tant * cr_expf(-inv * inv)) - ux;
            if (fabsf(value) < 1e-5f) break;
            float derivation = norm * (1.f - inv * tant);
            if (value > 0.f) c = b; else a = b;
            b -= value / derivation;
        }
        sx = erf_inv(b);
        sy = erf_inv(2.f * fmaxf(u.y, 1e-6f) - 1.f);
    }
    float cp = cos_phi(ws), sp = sin_phi(ws);
    float rot = cp * sx - sp * sy;
    sy = sp * sx + cp * sy;
    sx = rot;
    sx *= ax; sy *= ay;
    return normalize(f3(-sx, -sy, 1.f)) * signum(wo.z);
}
static __device__ __noinline__ float3 sample_wh_trowbridge(float3 wo_in, float2 u, float ax, float ay) {
    float3 wo = wo_in.z < 0.f ? -wo_in : wo_in;
    float3 ws = normalize(f3(ax * wo.x, ay * wo.y, wo.z));
    float ct = fabsf(cos_theta(ws));
    float sx, sy;
    if (ct > 0.9999f) {
        float r = sqrtf(u.x / (1.f - u.x));
        float phi = 2.f * u.y * ARN_PI;
        float sphi, cphi; cr_sincosf(phi, sphi, cphi);
        sx = r * cphi; sy = r * sphi;
    } else {
        float st = sqrtf(fmaxf(1.f - ct * ct, 0.f));
        float tant = st / ct, cott = ct / st;
        float g1 = 2.f / (1.f + sqrtf(1.f + 1.f / (cott * cott)));
        float a = 2.f * u.y / g1 - 1.f;
        float tmp = fminf(1.f / (a * a - 1.f), 1e10f);
        float d = sqrtf(fmaxf(tant * tant * tmp * tmp - (a * a - tant * tant) * tmp, 0.f));
        float sx1 = tant * tmp - d, sx2 = tant * tmp + d;
        float sxx = (a < 0.f || sx2 > cott) ? sx1 : sx2;
        float s, uy;
        if (u.y > 0.5f) { s = 1.f; uy = 2.f * (u.y - 0.5f); } else { s = -1.f; uy = 2.f * (0.5f - u.y); }
        float z = (uy * (uy * (uy * 0.27385f - 0.73369f) + 0.46341f)) / (uy * (uy * (uy * 0.093073f + 0.309420f) - 1.000000f) + 0.597999f);
        sx = sxx; sy = s * z * (1.f + sxx * sxx);
    }
    float cp = cos_phi(ws), sp = sin_phi(ws);
    float rot = cp * sx - sp * sy;
    sy = sp * sx + cp * sy;
    sx = rot;
    sx *= ax; sy *= ay;
    float3 wh = normalize(f3(-sx, -sy, 1.f));
    return wo_in.z < 0.f ? -wh : wh;
}
template <bool BECK> ARN_DEV float3 dist_sample_wh(float ax, float ay, float3 wo, float2 u) {
    return BECK ? sample_wh_beckmann(wo, u, ax, ay) : sample_wh_trowbridge(wo, u, ax, ay);
}

ARN_DEV float fresnel_dielectric(float cti, float etai, float etat) {       // fresnel.rs:16-37
    if (cti < 0.f) { float t = etai; etai = etat; etat = t; cti = -cti; }
    float s2i = fmaxf(1.f - cti * cti, 0.f);
    float eta = etai / etat;
    float s2t = eta * eta * s2i;
    if (s2t >= 1.f) return 1.f;
    float ctt = sqrtf(1.f - s2t);
    float etci = etat * cti, eict = etai * ctt;
    float r_para = (etci - eict) / (etci + eict);
    float eici = etai * cti, etct = etat * ctt;
    float r_perp = (eici - etct) / (eici + etct);
    return (r_para * r_para + r_perp * r_perp) * 0.5f;
}

// ---------------------------------------------------------------- BxDF lobes
enum LobeKind { LOBE_LAMBERT_R = 0, LOBE_LAMBERT_T, LOBE_OREN_NAYAR, LOBE_FRESNEL, LOBE_TS_R, LOBE_TS_T, LOBE_AS_BECK, LOBE_AS_TROW };
struct Lobe { int kind; float3 a, b; float c0, c1, alpha; float lam; };   // lam: Lambda(wo) of the lobe's microfacet distribution for THIS hit's outgoing direction (bsdf_prepare)
struct Sampled { float3 f, wi; float pdf; uint32_t type; };

ARN_DEV uint32_t lobe_type(int k) {
    switch (k) {
    case LOBE_LAMBERT_R: case LOBE_OREN_NAYAR: return BXDF_REFLECTION | BXDF_DIFFUSE;
    case LOBE_LAMBERT_T: return BXDF_TRANSMISSION | BXDF_DIFFUSE;
    case LOBE_FRESNEL: return BXDF_REFLECTION | BXDF_TRANSMISSION | BXDF_SPECULAR;
    case LOBE_TS_T: return BXDF_TRANSMISSION | BXDF_GLOSSY;
    default: return BXDF_REFLECTION | BXDF_GLOSSY;       // TS_R, Ashikhmin–Shirley
    }
}
// Lobe-kind masks of the shade kernel instances: a kernel that only ever meets some lobe kinds (its hits were sorted
// by material class) tells the compiler so, and the switches below shrink to those cases.  Same code, fewer
// instructions to fetch: the generic instance is 228 KB of SASS and instruction-cache bound.
#define LOBES_ALL 0xFFu
#define LOBES_PLASTIC (1u << LOBE_AS_BECK)                                                   /* material/plastic.rs:40-63 */
#define LOBES_GLASS ((1u << LOBE_FRESNEL) | (1u << LOBE_TS_R) | (1u << LOBE_TS_T))           /* material/glass.rs:42-80 */
template <uint32_t M> ARN_DEV int known_kind(int k) {
    if (M == LOBES_ALL) return k;
    int r = 0;
#pragma unroll
    for (int j = 7; j >= 0; j--) if ((M >> j) & 1u) r = j;          // lowest kind in the mask
#pragma unroll
    for (int j = 0; j < 8; j++) if (((M >> j) & 1u) && k == j) r = j;
    return r;
}
template <bool BECK> ARN_DEV float as_pdf(const Lobe& x, float3 wo, float3 wi) {        // microfacet.rs:613-623
    if (wo.z * wi.z < 0.f) return 0.f;
    float3 wh = normalize(wo + wi);
    return 0.5f * (dist_pdf<BECK>(x.alpha, x.alpha, wo, wh) / (4.f * dot(wo, wh)) + fabsf(cos_theta(wi)) * ARN_INV_PI);
}
template <bool BECK> ARN_DEV float3 as_eval(const Lobe& x, float3 wo, float3 wi) {      // :573-595
    float3 wh = wo + wi;
    if (relative_eq(length2(wh), 0.f)) return grey(0.f);
    wh = normalize(wh);
    float to = 1.f - powi5(1.f - 0.5f * fabsf(cos_theta(wo)));
    float ti = 1.f - powi5(1.f - 0.5f * fabsf(cos_theta(wi)));
    float3 diffuse = (28.f / (23.f * ARN_PI)) * x.a * (grey(1.f) - x.b) * to * ti;
    float cost = dot(wi, wh);
    float3 schlick = x.b + powi5(1.f - cost) * (grey(1.f) - x.b);
    float3 specular = dist_D<BECK>(x.alpha, x.alpha, wh) * schlick
        / (4.f * fabsf(dot(wi, wh)) * fmaxf(fabsf(cos_theta(wi)), fabsf(cos_theta(wo))));
    return diffuse + specular;
}
template <uint32_t M> ARN_NOINL float lobe_pdf(const Lobe& x, float3 wo, float3 wi) {
    switch (known_kind<M>(x.kind)) {
    case LOBE_LAMBERT_R: case LOBE_OREN_NAYAR: return wo.z * wi.z > 0.f ? fabsf(cos_theta(wi)) * ARN_INV_PI : 0.f;
    case LOBE_LAMBERT_T: return wo.z * wi.z >= 0.f ? 0.f : fabsf(cos_theta(wi)) * ARN_INV_PI;
    case LOBE_FRESNEL: return 0.f;
    case LOBE_TS_R: {
        if (wo.z * wi.z <= 0.f) return 0.f;
        float3 wh = normalize(wo + wi);
        return dist_pdf<false>(x.alpha, x.alpha, wo, wh) / (4.f * dot(wo, wh));
    }
    case LOBE_TS_T: {
        if (wo.z * wi.z > 0.f) return 0.f;
        float eta = wo.z > 0.f ? x.c1 / x.c0 : x.c0 / x.c1;
        float3 wh = normalize(wo + wi * eta);
        if (any_inf(wh) || any_nan(wh)) return 1.f;
        float sqrt_denom = dot(wo, wh) + eta * dot(wi, wh);
        float dhdi = eta * eta * fabsf(dot(wi, wh)) / (sqrt_denom * sqrt_denom);
        return dist_pdf<false>(x.alpha, x.alpha, wo, wh) * dhdi;
    }
    case LOBE_AS_BECK: return as_pdf<true>(x, wo, wi);
    default: return as_pdf<false>(x, wo, wi);
    }
}
template <uint32_t M> ARN_NOINL float3 lobe_eval(const Lobe& x, float3 wo, float3 wi) {
    switch (known_kind<M>(x.kind)) {
    case LOBE_LAMBERT_R: case LOBE_LAMBERT_T: return x.a * ARN_INV_PI;
    case LOBE_OREN_NAYAR: {
        float sti = sin_theta(wi), sto = sin_theta(wo);
        float max_cos = 0.f;
        if (sti > 1e-4f || sto > 1e-4f) {
            float spi = sin_phi(wi), spo = sin_phi(wo), cpi = cos_phi(wi), cpo = cos_phi(wo);
            max_cos = fmaxf(max_cos, cpi * cpo + spi * spo);
        }
        float ci = fabsf(cos_theta(wi)), co = fabsf(cos_theta(wo));
        float sin_a, tan_b;
        if (ci > co) { sin_a = sto; tan_b = sti / ci; } else { sin_a = sti; tan_b = sto / co; }
        return x.a * ARN_INV_PI * (x.c0 + x.c1 * max_cos * sin_a * tan_b);
    }
    case LOBE_FRESNEL: return grey(0.f);
    case LOBE_TS_R: {
        float3 wh = normalize(wo + wi);
        if (any_nan(wh)) return grey(0.f);
        return x.a * dist_D<false>(x.alpha, x.alpha, wh) * dist_visible_both<false>(x.alpha, x.alpha, wo, wi)
             * grey(fresnel_dielectric(dot(wi, wh), x.c0, x.c1)) / (4.f * fabsf(wo.z) * fabsf(wi.z));
    }
    case LOBE_TS_T: {
        if (wo.z * wi.z > 0.f) return grey(0.f);
        float eta = wo.z > 0.f ? x.c1 / x.c0 : x.c0 / x.c1;
        float3 wh = normalize(wo + wi * eta);
        if (any_inf(wh) || any_nan(wh)) return grey(1.f);
        if (wh.z < 0.f) wh = -wh;
        float cosoh = dot(wo, wh);
        float3 f = grey(fresnel_dielectric(cosoh, x.c0, x.c1));
        float cosih = dot(wi, wh);
        float sqrt_denom = cosoh + eta * cosih;
        return x.a * dist_D<false>(x.alpha, x.alpha, wh) * dist_visible_both<false>(x.alpha, x.alpha, wo, wi)
             * (grey(1.f) - f) * fabsf(cosih) * fabsf(cosoh)
             / (fabsf(cos_theta(wo)) * fabsf(cos_theta(wi)) * sqrt_denom * sqrt_denom);
    }
    case LOBE_AS_BECK: return as_eval<true>(x, wo, wi);
    default: return as_eval<false>(x, wo, wi);
    }
}
// lobe_eval + lobe_pdf of ONE lobe for the same (wo, wi) in one call: the half vector, D(wh) and Lambda(wo) the two
// share are computed once.  Every value is produced by the same expression as in lobe_eval / lobe_pdf, so the
// results are bit-identical to calling them separately (D is even in wh: its inputs are squares of wh's components).
// `want_f` = false skips the value (Bsdf::evaluate only adds lobes on the matching side, bsdf.rs:82-98).
template <bool BECK> ARN_DEV void as_eval_pdf(const Lobe& x, float3 wo, float3 wi, bool want_f, float3& f, float& pdf) {
    float3 whs = wo + wi;
    float3 wh = normalize(whs);
    const bool need_pdf = !(wo.z * wi.z < 0.f);
    const bool need_f = want_f && !relative_eq(length2(whs), 0.f);
    float D = 0.f;
    if (need_pdf || need_f) D = dist_D<BECK>(x.alpha, x.alpha, wh);
    pdf = 0.f; f = grey(0.f);
    if (need_pdf) {
        float dp = D * (1.f / (1.f + x.lam)) * fabsf(dot(wo, wh)) / fabsf(cos_theta(wo));      // dist_visible(wo) = 1 / (1 + Lambda(wo)), Lambda(wo) once per hit
        pdf = 0.5f * (dp / (4.f * dot(wo, wh)) + fabsf(cos_theta(wi)) * ARN_INV_PI);
    }
    if (need_f) {
        float to = 1.f - powi5(1.f - 0.5f * fabsf(cos_theta(wo)));
        float ti = 1.f - powi5(1.f - 0.5f * fabsf(cos_theta(wi)));
        float3 diffuse = (28.f / (23.f * ARN_PI)) * x.a * (grey(1.f) - x.b) * to * ti;
        float cost = dot(wi, wh);
        float3 schlick = x.b + powi5(1.f - cost) * (grey(1.f) - x.b);
        float3 specular = D * schlick / (4.f * fabsf(dot(wi, wh)) * fmaxf(fabsf(cos_theta(wi)), fabsf(cos_theta(wo))));
        f = diffuse + specular;
    }
}
template <uint32_t M> ARN_NOINL void lobe_eval_pdf(const Lobe& x, float3 wo, float3 wi, bool want_f, float3& f, float& pdf) {
    const int kind = known_kind<M>(x.kind);
    if (kind == LOBE_TS_R || kind == LOBE_TS_T) {
        // Both Torrance–Sparrow lobes in ONE code path: a glass BSDF samples one of them per path, so lanes of a warp hold
        // either; the distribution terms D(wh), Lambda(wo), Lambda(wi) — most of the arithmetic — then run converged.  Every
        // lane evaluates exactly the expressions of its own lobe (microfacet.rs:372-430 / :433-533), in the same order.
        const bool is_t = kind == LOBE_TS_T;
        pdf = 0.f; f = grey(0.f);
        float eta = 1.f; float3 wh; bool need_pdf = true, need_f = want_f;
        if (is_t) {
            if (wo.z * wi.z > 0.f) return;
            eta = wo.z > 0.f ? x.c1 / x.c0 : x.c0 / x.c1;
            wh = normalize(wo + wi * eta);
            if (any_inf(wh) || any_nan(wh)) { pdf = 1.f; if (want_f) f = grey(1.f); return; }
        } else {
            wh = normalize(wo + wi);
            need_pdf = !(wo.z * wi.z <= 0.f);
            need_f = want_f && !any_nan(wh);
            if (!(need_pdf || need_f)) return;
        }
        const float D = dist_D<false>(x.alpha, x.alpha, wh);
        const float lo = x.lam;                                           // Lambda(wo): once per hit (bsdf_prepare), not once per lobe evaluation
        float li = 0.f;
        if (need_f) li = dist_lambda<false>(x.alpha, x.alpha, wi);
        if (is_t) {
            {
                float sqrt_denom = dot(wo, wh) + eta * dot(wi, wh);
                float dhdi = eta * eta * fabsf(dot(wi, wh)) / (sqrt_denom * sqrt_denom);
                pdf = (D * (1.f / (1.f + lo)) * fabsf(dot(wo, wh)) / fabsf(cos_theta(wo))) * dhdi;
            }
            if (want_f) {
                if (wh.z < 0.f) wh = -wh;
                float cosoh = dot(wo, wh);
                float3 fr = grey(fresnel_dielectric(cosoh, x.c0, x.c1));
                float cosih = dot(wi, wh);
                float sqrt_denom = cosoh + eta * cosih;
                f = x.a * D * (1.f / (1.f + lo + li)) * (grey(1.f) - fr) * fabsf(cosih) * fabsf(cosoh)
                  / (fabsf(cos_theta(wo)) * fabsf(cos_theta(wi)) * sqrt_denom * sqrt_denom);
            }
        } else {
            if (need_pdf) pdf = (D * (1.f / (1.f + lo)) * fabsf(dot(wo, wh)) / fabsf(cos_theta(wo))) / (4.f * dot(wo, wh));
            if (need_f) f = x.a * D * (1.f / (1.f + lo + li)) * grey(fresnel_dielectric(dot(wi, wh), x.c0, x.c1)) / (4.f * fabsf(wo.z) * fabsf(wi.z));
        }
        return;
    }
    switch (kind) {
    case LOBE_AS_BECK: as_eval_pdf<true>(x, wo, wi, want_f, f, pdf); return;
    case LOBE_AS_TROW: as_eval_pdf<false>(x, wo, wi, want_f, f, pdf); return;
    case LOBE_FRESNEL: pdf = 0.f; f = grey(0.f); return;                 // specular: no value, no density (fresnel.rs:150-161)
    default:
        pdf = lobe_pdf<M>(x, wo, wi);
        f = want_f ? lobe_eval<M>(x, wo, wi) : grey(0.f);
        return;
    }
}
template <bool BECK> ARN_DEV Sampled as_sample(const Lobe& x, float3 wo, float2 u) {      // microfacet.rs:597-611
    Sampled r; r.type = BXDF_REFLECTION | BXDF_GLOSSY;
    float3 wi;
    if (u.x < 0.5f) {
        u.x *= 2.f;
        float3 wh = dist_sample_wh<BECK>(x.alpha, x.alpha, wo, u);
        wi = normalize(2.f * wh * dot(wo, wh) - wo);
        if (wo.z * wi.z <= 0.f) { r.wi = wi; as_eval_pdf<BECK>(x, wo, wi, false, r.f, r.pdf); return r; }
    } else {
        u.x = (1.f - u.x) * 2.f;
        wi = sample_cosw_hemisphere(u);
        if (wi.z < 0.f) wi.z = -wi.z;
    }
    r.wi = wi; as_eval_pdf<BECK>(x, wo, wi, true, r.f, r.pdf);
    return r;
}
template <uint32_t M> ARN_NOINL Sampled lobe_sample(const Lobe& x, float3 wo, float2 u) {
    Sampled r; r.type = lobe_type(known_kind<M>(x.kind));
    const int kind = known_kind<M>(x.kind);
    if (kind == LOBE_TS_R || kind == LOBE_TS_T) {
        // one code path for both Torrance–Sparrow lobes (microfacet.rs:408-421 / :493-511): the half-vector sample and the
        // value + density evaluation run converged for lanes that drew either lobe; per lane the expressions are the lobe's own
        const float3 wh = dist_sample_wh<false>(x.alpha, x.alpha, wo, u);
        float3 wi; bool ok;
        if (kind == LOBE_TS_R) {
            r.pdf = (dist_D<false>(x.alpha, x.alpha, wh) * (1.f / (1.f + x.lam)) * fabsf(dot(wo, wh)) / fabsf(cos_theta(wo))) / (4.f * dot(wo, wh));   // dist_pdf(wo, wh) / (4 wo.wh)
            wi = normalize(2.f * wh * dot(wo, wh) - wo);
            r.wi = wi; r.f = grey(0.f);
            ok = !(wo.z * wi.z <= 0.f);
        } else {
            const float eta = wo.z > 0.f ? x.c0 / x.c1 : x.c1 / x.c0;
            ok = refract(wo, wh, eta, wi);
            if (ok) r.wi = wi; else { r.f = grey(0.f); r.wi = f3(0.f, 0.f, 0.f); r.pdf = 0.f; }
        }
        if (ok) {
            float3 fv; float pv;
            lobe_eval_pdf<M>(x, wo, wi, true, fv, pv);
            r.f = fv;
            if (kind == LOBE_TS_T) r.pdf = pv;
        }
        return r;
    }
    switch (kind) {
    case LOBE_LAMBERT_R: case LOBE_OREN_NAYAR: {
        float3 wi = sample_cosw_hemisphere(u);
        if (wo.z < 0.f) wi.z = -wi.z;
        lobe_eval_pdf<M>(x, wo, wi, true, r.f, r.pdf); r.wi = wi; return r;
    }
    case LOBE_LAMBERT_T: {
        float3 wi = sample_cosw_hemisphere(u);
        if (wo.z > 0.f) wi.z = -wi.z;
        lobe_eval_pdf<M>(x, wo, wi, true, r.f, r.pdf); r.wi = wi; return r;
    }
    case LOBE_FRESNEL: {                                                     // fresnel.rs:163-196
        float ct = cos_theta(wo);
        float f = fresnel_dielectric(ct, x.c0, x.c1);
        if (u.x < f) {
            r.wi = f3(-wo.x, -wo.y, wo.z); r.pdf = f;
            r.f = r.pdf * x.a / fabsf(ct);
            r.type = BXDF_REFLECTION | BXDF_SPECULAR; return r;
        }
        float pdf = 1.f - f;
        float etai, etao; float3 n;
        if (ct > 0.f) { etai = x.c0; etao = x.c1; n = f3(0.f, 0.f, 1.f); } else { etai = x.c1; etao = x.c0; n = f3(0.f, 0.f, -1.f); }
        float eta = etai / etao;
        float3 wt;
        r.type = BXDF_TRANSMISSION | BXDF_SPECULAR; r.pdf = pdf;
        if (refract(wo, n, eta, wt)) { r.f = x.b * eta * eta * pdf / fabsf(wt.z); r.wi = wt; }
        else { r.f = grey(0.f); r.wi = f3(0.f, 0.f, 0.f); }
        return r;
    }
    case LOBE_AS_BECK: return as_sample<true>(x, wo, u);
    default: return as_sample<false>(x, wo, u);
    }
}

// ---------------------------------------------------------------- Bsdf (material/bsdf.rs)
struct Bsdf { float3 ns, ng, ts, bs; Lobe lobe[3]; int n; float3 wo_l; };   // wo_l: the hit's outgoing direction in the local frame (bsdf_prepare)

ARN_DEV float3 to_local(const Bsdf& b, float3 v) { return f3(dot(v, b.ts), dot(v, b.bs), dot(v, b.ns)); }
ARN_DEV float3 to_parent(const Bsdf& b, float3 v) {
    return f3(dot(v, f3(b.ts.x, b.bs.x, b.ns.x)), dot(v, f3(b.ts.y, b.bs.y, b.ns.y)), dot(v, f3(b.ts.z, b.bs.z, b.ns.z)));
}
// materials with constant textures (material/{matte,plastic,glass,translucent}.rs)
ARN_DEV void bsdf_build(const arn_material& m, const Surf& s, Bsdf& b) {
    b.ts = normalize(s.dpdu); b.ns = s.ns; b.bs = normalize(cross(b.ns, b.ts)); b.ng = s.ng; b.n = 0;
    float3 kd = f3(m.kd[0], m.kd[1], m.kd[2]), ks = f3(m.ks[0], m.ks[1], m.ks[2]);
    Lobe z; z.kind = 0; z.a = grey(0.f); z.b = grey(0.f); z.c0 = 0.f; z.c1 = 0.f; z.alpha = m.alpha; z.lam = 0.f;
    switch (m.type) {
    case ARN_MAT_MATTE: {
        float sig = clampf(m.sigma, 0.f, 90.f);
        if (!is_black(kd)) {
            Lobe x = z; x.a = kd;
            if (sig == 0.f) x.kind = LOBE_LAMBERT_R;
            else {
                x.kind = LOBE_OREN_NAYAR;
                float sigma2 = sig * sig;
                x.c0 = 1.f - (sigma2 / (2.f * (sigma2 + 0.33f)));
                x.c1 = (0.45f * sigma2) / (sigma2 + 0.09f);
            }
            b.lobe[b.n++] = x;
        }
        break; }
    case ARN_MAT_PLASTIC: {
        Lobe x = z; x.kind = LOBE_AS_BECK;
        x.a = f3(clampf(kd.x, 0.f, 1.f), clampf(kd.y, 0.f, 1.f), clampf(kd.z, 0.f, 1.f));
        x.b = f3(clampf(ks.x, 0.f, 1.f), clampf(ks.y, 0.f, 1.f), clampf(ks.z, 0.f, 1.f));
        b.lobe[b.n++] = x;
        break; }
    case ARN_MAT_GLASS: {
        if (!is_black(ks)) { Lobe x = z; x.kind = LOBE_FRESNEL; x.a = ks; x.b = ks; x.c0 = 1.f; x.c1 = m.eta; b.lobe[b.n++] = x; }
        if (!is_black(kd)) {
            Lobe r = z; r.kind = LOBE_TS_R; r.a = kd; r.c0 = 1.f; r.c1 = m.eta; b.lobe[b.n++] = r;
            Lobe t = z; t.kind = LOBE_TS_T; t.a = kd; t.c0 = 1.f; t.c1 = m.eta; b.lobe[b.n++] = t;
        }
        break; }
    default: {  // ARN_MAT_TRANSLUCENT
        if (!relative_eq(m.dissolve, 0.f)) {
            Lobe x = z; x.kind = LOBE_AS_TROW;
            float3 d = kd * m.dissolve, sp = ks * m.dissolve;
            x.a = f3(clampf(d.x, 0.f, 1.f), clampf(d.y, 0.f, 1.f), clampf(d.z, 0.f, 1.f));
            x.b = f3(clampf(sp.x, 0.f, 1.f), clampf(sp.y, 0.f, 1.f), clampf(sp.z, 0.f, 1.f));
            b.lobe[b.n++] = x;
        }
        if (!is_black(kd)) { Lobe x = z; x.kind = LOBE_LAMBERT_T; x.a = kd * (1.f - m.dissolve); b.lobe[b.n++] = x; }
        break; }
    }
}
// Bsdf::evaluate with BXDF_ALL (bsdf.rs:82-98)
ARN_DEV float3 bsdf_eval(const Bsdf& b, float3 wow, float3 wiw) {
    float3 wo = normalize(to_local(b, wow)), wi = normalize(to_local(b, wiw));
    bool is_reflection = dot(wow, b.ng) * dot(wiw, b.ng) > 0.f;
    float3 ret = grey(0.f);
    for (int i = 0; i < b.n; i++) {
        uint32_t k = lobe_type(b.lobe[i].kind);
        if ((is_reflection && (k & BXDF_REFLECTION)) || (!is_reflection && (k & BXDF_TRANSMISSION))) ret = ret + lobe_eval<LOBES_ALL>(b.lobe[i], wo, wi);
    }
    return ret;
}
// Bsdf::pdf with BXDF_ALL (bsdf.rs:205-222)
ARN_DEV float bsdf_pdf(const Bsdf& b, float3 wow, float3 wiw) {
    float3 wo = normalize(to_local(b, wow)), wi = normalize(to_local(b, wiw));
    if (wo.z == 0.f) return 0.f;
    float pdfsum = 0.f;
    for (int i = 0; i < b.n; i++) pdfsum += fmaxf(lobe_pdf<LOBES_ALL>(b.lobe[i], wo, wi), 0.f);
    return b.n == 0 ? pdfsum : pdfsum / (float)b.n;
}
// Bsdf::evaluate + Bsdf::pdf for the same pair of directions (the NEE light sample needs both, scene.rs:98-101)
// The microfacet lobes of one Bsdf share (alpha, distribution), and every evaluation / sampling call of a hit uses the same
// outgoing direction: Lambda(wo) and the local wo are computed ONCE per hit instead of once per lobe evaluation (up to eight
// times for glass).  Same function of the same arguments: the bits do not change.
template <uint32_t M> ARN_DEV void bsdf_prepare(Bsdf& b, float3 wow) {
    b.wo_l = normalize(to_local(b, wow));
    bool need = false, beck = false;
    for (int i = 0; i < b.n; i++) {
        const int k = known_kind<M>(b.lobe[i].kind);
        if (k == LOBE_TS_R || k == LOBE_TS_T || k == LOBE_AS_TROW) need = true;
        if (k == LOBE_AS_BECK) { need = true; beck = true; }
    }
    float lam = 0.f;
    if (need) lam = beck ? dist_lambda<true>(b.lobe[0].alpha, b.lobe[0].alpha, b.wo_l) : dist_lambda<false>(b.lobe[0].alpha, b.lobe[0].alpha, b.wo_l);
    for (int i = 0; i < b.n; i++) b.lobe[i].lam = lam;
}
// `wow` must be the direction bsdf_prepare was given (the hit's wo)
template <uint32_t M> ARN_DEV void bsdf_eval_pdf(const Bsdf& b, float3 wow, float3 wiw, float3& f, float& pdf) {
    float3 wo = b.wo_l, wi = normalize(to_local(b, wiw));
    bool is_reflection = dot(wow, b.ng) * dot(wiw, b.ng) > 0.f;
    f = grey(0.f);
    float pdfsum = 0.f;
    for (int i = 0; i < b.n; i++) {
        uint32_t k = lobe_type(known_kind<M>(b.lobe[i].kind));
        bool match = (is_reflection && (k & BXDF_REFLECTION)) || (!is_reflection && (k & BXDF_TRANSMISSION));
        float3 fi; float pi;
        lobe_eval_pdf<M>(b.lobe[i], wo, wi, match, fi, pi);
        if (match) f = f + fi;
        pdfsum += fmaxf(pi, 0.f);
    }
    pdf = wo.z == 0.f ? 0.f : (b.n == 0 ? pdfsum : pdfsum / (float)b.n);
}
// Bsdf::evaluate_sampled with BXDF_ALL (bsdf.rs:100-145)
template <uint32_t M> ARN_NOINL Sampled bsdf_sample(const Bsdf& b, float3 wow, float2 u) {
    Sampled ret; ret.f = grey(0.f); ret.wi = f3(0.f, 1.f, 0.f); ret.pdf = 0.f; ret.type = 0;
    int match_count = b.n;
    if (match_count == 0) return ret;
    float3 wo = b.wo_l;                                    // = normalize(to_local(b, wow)), see bsdf_prepare
    int idx = (int)floorf(u.x * (float)match_count); if (idx > match_count - 1) idx = match_count - 1;
    Sampled s = lobe_sample<M>(b.lobe[idx], wo, u);
    if (s.pdf == 0.f) return ret;
    bool is_specular = (lobe_type(known_kind<M>(b.lobe[idx].kind)) & BXDF_SPECULAR) != 0;
    ret = s; ret.type = s.type & BXDF_ALL;
    float3 wi = ret.wi;
    ret.wi = to_parent(b, wi);
    if (match_count == 1 || is_specular) return ret;
    ret.f = grey(0.f);
    bool is_reflection = dot(wow, b.ng) * dot(ret.wi, b.ng) > 0.f;
    float pdfsum = 0.f;
    for (int k = 0; k < b.n; k++) {
        uint32_t t = lobe_type(known_kind<M>(b.lobe[k].kind));
        if ((t & ret.type) && ((is_reflection && (t & BXDF_REFLECTION)) || (!is_reflection && (t & BXDF_TRANSMISSION)))) {
            float3 fk; float pk;
            lobe_eval_pdf<M>(b.lobe[k], wo, wi, true, fk, pk);
            ret.f = ret.f + fk;
            pdfsum += fmaxf(pk, 0.f);
        }
    }
    ret.pdf = pdfsum / (float)match_count;
    return ret;
}

// ---------------------------------------------------------------- single diffuse lobe fast path
// Matte materials build at most ONE lobe (Lambert or Oren–Nayar).  The diffuse shade kernel uses
// these inlined forms of Bsdf::{evaluate, pdf, evaluate_sampled}; the arithmetic is the generic code's
// (lobe_eval / lobe_pdf / lobe_sample cases LAMBERT_R / OREN_NAYAR), so results are bit-identical.
ARN_DEV float3 diffuse_lobe_eval(const Lobe& x, float3 wo, float3 wi) {
    if (x.kind == LOBE_LAMBERT_R) return x.a * ARN_INV_PI;
    float sti = sin_theta(wi), sto = sin_theta(wo);
    float max_cos = 0.f;
    if (sti > 1e-4f || sto > 1e-4f) {
        float spi = sin_phi(wi), spo = sin_phi(wo), cpi = cos_phi(wi), cpo = cos_phi(wo);
        max_cos = fmaxf(max_cos, cpi * cpo + spi * spo);
    }
    float ci = fabsf(cos_theta(wi)), co = fabsf(cos_theta(wo));
    float sin_a, tan_b;
    if (ci > co) { sin_a = sto; tan_b = sti / ci; } else { sin_a = sti; tan_b = sto / co; }
    return x.a * ARN_INV_PI * (x.c0 + x.c1 * max_cos * sin_a * tan_b);
}
ARN_DEV float diffuse_lobe_pdf(float3 wo, float3 wi) { return wo.z * wi.z > 0.f ? fabsf(cos_theta(wi)) * ARN_INV_PI : 0.f; }
template <bool DIFFUSE> ARN_DEV float3 bsdf_eval_k(const Bsdf& b, float3 wow, float3 wiw) {
    if (!DIFFUSE) return bsdf_eval(b, wow, wiw);
    float3 wo = normalize(to_local(b, wow)), wi = normalize(to_local(b, wiw));
    bool is_reflection = dot(wow, b.ng) * dot(wiw, b.ng) > 0.f;
    float3 ret = grey(0.f);
    if (b.n > 0 && is_reflection) ret = ret + diffuse_lobe_eval(b.lobe[0], wo, wi);
    return ret;
}
template <bool DIFFUSE> ARN_DEV float bsdf_pdf_k(const Bsdf& b, float3 wow, float3 wiw) {
    if (!DIFFUSE) return bsdf_pdf(b, wow, wiw);
    float3 wo = normalize(to_local(b, wow)), wi = normalize(to_local(b, wiw));
    if (wo.z == 0.f) return 0.f;
    float pdfsum = 0.f;
    if (b.n > 0) pdfsum += fmaxf(diffuse_lobe_pdf(wo, wi), 0.f);
    return b.n == 0 ? pdfsum : pdfsum / (float)b.n;
}
template <bool DIFFUSE, uint32_t M> ARN_DEV Sampled bsdf_sample_k(const Bsdf& b, float3 wow, float2 u) {
    if (!DIFFUSE) return bsdf_sample<M>(b, wow, u);
    Sampled ret; ret.f = grey(0.f); ret.wi = f3(0.f, 1.f, 0.f); ret.pdf = 0.f; ret.type = 0;
    if (b.n == 0) return ret;
    float3 wo = normalize(to_local(b, wow));
    float3 wi = sample_cosw_hemisphere(u);
    if (wo.z < 0.f) wi.z = -wi.z;
    float pdf = diffuse_lobe_pdf(wo, wi);
    if (pdf == 0.f) return ret;
    ret.f = diffuse_lobe_eval(b.lobe[0], wo, wi); ret.pdf = pdf; ret.type = BXDF_REFLECTION | BXDF_DIFFUSE;
    ret.wi = to_parent(b, wi);
    return ret;
}

// ---------------------------------------------------------------- sphere area light
ARN_DEV float3 sphere_emission(const DevSphere& sp) { return f3(sp.emission[0], sp.emission[1], sp.emission[2]); }
ARN_DEV float sphere_area(const DevSphere& sp) { return sp.phimax * sp.radius * (sp.zmax - sp.zmin); }
// Light::evaluate_path: emission if the shape is re-hit from pos+dir going back (shape.rs:91-103)
ARN_NOINL float3 light_le(const DevSphere& sp, float3 pos, float3 dir) {
    if (!sp.emissive) return grey(0.f);
    if (sp.has_transform) { pos = xform_point(sp.parent_local, pos); dir = xform_vector(sp.parent_local, dir); }
    float3 p = pos + dir;
    float t; float3 q;
    return sphere_test(sp, p, -dir, ARN_INF, t, q) ? sphere_emission(sp) : grey(0.f);
}
struct LightSample { float3 radiance, pfrom, pto; float pdf; };
// evaluate_sampled (shape.rs:108-130 via transformed.rs:120-124) with Shape::sample_wrt (shape/mod.rs:52-64)
ARN_NOINL LightSample light_sample(const DevSphere& sp, float3 pos, float2 u) {
    if (sp.has_transform) pos = xform_point(sp.parent_local, pos);
    float phi = u.x * sp.phimax;                                            // Sphere::sample (sphere.rs:304-311)
    float theta = u.y * (sp.thetamax - sp.thetamin) + sp.thetamin;
    float st, ct, sph, cph; cr_sincosf(theta, st, ct); cr_sincosf(phi, sph, cph);
    float3 dir = f3(st * cph, st * sph, ct);
    float3 lp = dir * sp.radius;
    float lpdf = 1.f / sphere_area(sp);
    float3 wi = lp - pos;
    float distance2 = length2(wi);
    if (relative_eq(distance2, 0.f)) lpdf = 0.f;
    else {
        float3 w = wi / sqrtf(distance2);
        lpdf *= distance2 / fabsf(dot(dir, w));
        if (isinf(lpdf)) lpdf = 0.f;
    }
    LightSample ls; ls.radiance = grey(0.f); ls.pdf = lpdf; ls.pfrom = lp; ls.pto = pos;
    if (sp.emissive) {
        float3 ldir = pos - lp;
        if (dot(ldir, dir) > 0.f) {
            float t; float3 q;
            if (sphere_test(sp, pos, -ldir, ARN_INF, t, q)) ls.radiance = sphere_emission(sp);
        }
    }
    if (sp.has_transform) { ls.pfrom = xform_point(sp.local_parent, ls.pfrom); ls.pto = xform_point(sp.local_parent, ls.pto); }
    return ls;
}
// PointLight / SpotLight / DistantLight::evaluate_sampled (lighting/pointlights.rs:50-61,181-194 with
// falloff :147-158; lighting/distantlight.rs:68-80).  pdf is 1 for all three.
ARN_NOINL LightSample analytic_sample(const arn_analytic_light& l, float3 pos) {
    LightSample ls; ls.pdf = 1.f; ls.pto = pos;
    float3 intensity = f3(l.intensity[0], l.intensity[1], l.intensity[2]);
    if (l.type == ARN_LIGHT_POINT) {
        ls.pfrom = f3(l.pos[0], l.pos[1], l.pos[2]);
        ls.radiance = intensity / length2(ls.pto - ls.pfrom);
    } else if (l.type == ARN_LIGHT_SPOT) {
        ls.pfrom = f3(l.pos[0], l.pos[1], l.pos[2]);
        float3 dir = ls.pto - ls.pfrom;
        float mag2 = length2(dir);
        float cos_theta = xform_vector(l.parent_local, dir / sqrtf(mag2)).z;
        float falloff;
        if (cos_theta < l.cost) falloff = 0.f;
        else if (cos_theta > l.cosf) falloff = 1.f;
        else {
            float delta = (cos_theta - l.cost) / (l.cosf - l.cost);
            float delta2 = delta * delta;
            falloff = delta2 * delta2;
        }
        ls.radiance = intensity * falloff / mag2;
    } else {
        ls.radiance = intensity;
        ls.pfrom = pos + f3(l.dir[0], l.dir[1], l.dir[2]) * (-2.0f * l.world_radius);
    }
    return ls;
}

// Light::pdf = Shape::pdf_wrt (shape/mod.rs:67-75): needs the local hit normal
ARN_NOINL float light_pdf(const DevSphere& sp, float3 pos, float3 wi) {
    if (sp.has_transform) { pos = xform_point(sp.parent_local, pos); wi = xform_vector(sp.parent_local, wi); }
    float t; float3 p;
    if (!sphere_test(sp, pos, wi, ARN_INF, t, p)) return 0.f;
    float thetadelta = sp.thetamax - sp.thetamin;
    float theta = cr_acosf(p.z / sp.radius);
    float inv_z_radius = 1.f / sqrtf(p.x * p.x + p.y * p.y);
    float cphi = p.x * inv_z_radius, sphi = p.y * inv_z_radius;
    float3 dpdu = f3(-sp.phimax * p.y, sp.phimax * p.x, 0.f);
    float3 dpdv = thetadelta * f3(p.z * cphi, p.z * sphi, -sp.radius * cr_sinf(theta));
    float3 n = normalize(cross(dpdu, dpdv));
    return length2(p - pos) / (fabsf(dot(wi, n)) * sphere_area(sp));
}

}  // namespace arn
