// ORACLE — test infrastructure only (see geom.hpp header).
// CPU restatement of arendur's sampling, BxDFs, BSDF and materials:
//   src/sample/{mod,distribution,filters}.rs, src/bxdf/{mod,lambertian,oren_nayar,fresnel,
//   microfacet}.rs, src/material/{bsdf,matte,plastic,glass,translucent}.rs, src/spectrum/mod.rs
#pragma once
#include "geom.hpp"
#include "texture.hpp"
#include "../include/arn.h"

namespace orc {

// ------------------------------------------------------------------ spectrum (spectrum/mod.rs)
typedef V3 RGB;   // RGBSpectrumf = Vector3<f32> wrapper; ops are element-wise
inline RGB rgb(Float r, Float g, Float b) { return v3(r, g, b); }
inline RGB grey(Float n) { return v3(n, n, n); }
inline RGB operator*(RGB a, RGB b) { return mul_elem(a, b); }
inline bool is_black(RGB s) { return s.x == 0.f && s.y == 0.f && s.z == 0.f; }   // :113-115
inline Float rgb_y(RGB s) { return 0.212671f * s.x + 0.715160f * s.y + 0.072169f * s.z; } // into_xyz().y :288-294
inline bool rgb_valid(RGB s) {                                                   // :303-307
    return !(std::isnan(s.x) || std::isnan(s.y) || std::isnan(s.z))
        && !(std::isinf(s.x) || std::isinf(s.y) || std::isinf(s.z))
        && s.x >= 0.f && s.y >= 0.f && s.z >= 0.f;
}
inline RGB rgb_clamp(RGB s, Float lo, Float hi) { return v3(clampf(s.x, lo, hi), clampf(s.y, lo, hi), clampf(s.z, lo, hi)); }

// ------------------------------------------------------------------ ParitySampler
// Counter-based sampler satisfying the `Sampler` contract (sample/mod.rs:22-94).  The
// reference's StrataSampler is OS-seeded and, through the idim-never-reset quirk
// (SURVEY.md Appendix A-17), marginally i.i.d. uniform; this is the deterministic
// equivalent shared bit-for-bit with the GPU (32-bit integer hash, (h >> 8) * 2^-24).
inline uint32_t mix32(uint32_t h) {
    h ^= h >> 16; h *= 0x7feb352dU; h ^= h >> 15; h *= 0x846ca68bU; h ^= h >> 16; return h;
}
inline Float u01(uint32_t h) { return (Float)(h >> 8) * (1.0f / 16777216.0f); }
// Hash permutation of [0, l): Kensler, "Correlated Multi-Jittered Sampling" (Pixar TM 13-01), permute(): cycle walking over the
// next power of two; pure 32-bit integer arithmetic, so the device computes the same index.
inline uint32_t hash_permute(uint32_t i, uint32_t l, uint32_t p) {
    uint32_t w = l - 1;
    w |= w >> 1; w |= w >> 2; w |= w >> 4; w |= w >> 8; w |= w >> 16;
    do {
        i ^= p; i *= 0xe170893du; i ^= p >> 16; i ^= (i & w) >> 4; i ^= p >> 8; i *= 0x0929eb3fu; i ^= p >> 23; i ^= (i & w) >> 1;
        i *= 1u | p >> 27; i *= 0x6935fa69u; i ^= (i & w) >> 11; i *= 0x74dcb303u; i ^= (i & w) >> 2; i *= 0x9e501cc3u;
        i ^= (i & w) >> 2; i *= 0xc860a3dfu; i &= w; i ^= i >> 5;
    } while (i >= l);
    return (i + p) % l;
}
// stratum index + jitter, kept strictly below the next index (f32 addition may round x + u up to x + 1 when u is within half an ulp(x) of 1)
inline Float strat_offset(uint32_t x, Float u) {
    Float t = (Float)x + u, hi = (Float)(x + 1u);
    return t < hi ? t : from_bits(to_bits(hi) - 1u);
}
struct ParitySampler {
    uint32_t seed = 0, spp = 1;
    uint32_t px = 0, py = 0, isample = 0, k1 = 0, k2 = 0, i1d = 0, i2d = 0;
    uint32_t mode = 0, sampledx = 1, sampledy = 1, ndim = 0, kpix = 0;     // ARN_SAMPLER_STRATIFIED (include/arn.h): spp = sampledx * sampledy
    void rekey() {
        kpix = mix32(mix32(mix32(seed) + px) + py);
        uint32_t key = mix32(kpix + isample);
        k1 = mix32(key ^ 0xA511E9B3u); k2 = mix32(key ^ 0x63D83595u); i1d = 0; i2d = 0;
    }
    void start_pixel(uint32_t x, uint32_t y) { px = x; py = y; isample = 0; rekey(); }
    void set_sample_index(uint32_t s) { isample = s; rekey(); }
    Float next() {
        const uint32_t d = i1d;
        Float u = u01(mix32(k1 + (i1d++)));
        if (mode == 1u && d < ndim) {
            const uint32_t n = sampledx * sampledy;
            u = strat_offset(hash_permute(isample, n, mix32(kpix ^ (0x1D000000u + d))), u) / (Float)n;
        }
        return u;
    }
    V2 next_2d() {
        const uint32_t d = i2d;
        V2 r = v2(u01(mix32(k2 + 2 * i2d)), u01(mix32(k2 + 2 * i2d + 1))); i2d++;
        if (mode == 1u && d < ndim) {
            const uint32_t c = hash_permute(isample, sampledx * sampledy, mix32(kpix ^ (0x2D000000u + d)));
            r = v2(strat_offset(c / sampledy, r.x) / (Float)sampledx, strat_offset(c % sampledy, r.y) / (Float)sampledy);
        }
        return r;
    }
    bool next_sample() { isample++; if (isample >= spp) return false; rekey(); return true; }
};

// ------------------------------------------------------------------ warps (sample/mod.rs)
inline V2 sample_concentric_disk(V2 u) {                                          // :165-177
    V2 w = (2.f * u) - v2(1.f, 1.f);
    if (w.x == 0.f && w.y == 0.f) return v2(0.f, 0.f);
    Float r, theta;
    if (std::fabs(w.x) > std::fabs(w.y)) { r = w.x; theta = frac_pi_4() * (w.y / w.x); }
    else { r = w.y; theta = frac_pi_2() - frac_pi_4() * (w.x / w.y); }
    return r * v2(fcos(theta), fsin(theta));
}
inline V3 sample_cosw_hemisphere(V2 u) {                                          // :203-207
    V2 d = sample_concentric_disk(u);
    Float z = std::sqrt(std::fabs(1.f - d.x * d.x - d.y * d.y));
    return v3(d.x, d.y, z);
}
inline Float power_heuristic(Float pdff, Float pdfg) {                            // :243-247, nf = ng = 1
    Float f = 1.f * pdff, g = 1.f * pdfg;
    return (f * f) / (f * f + g * g);
}

// Distribution1D::sample_discrete (sample/distribution.rs:99-118,126-150)
inline void sample_discrete(const Float* func, const Float* cdf, uint32_t n, Float integral, Float u,
                            uint32_t* offset, Float* pdf) {
    if (u == 0.f) u += epsilon();
    // binary_search_by(..).unwrap_or_else(|v| v) - 1  ==  (first index with cdf >= u) - 1
    uint32_t lo = 0, hi = n + 1;
    while (lo < hi) { uint32_t mid = lo + (hi - lo) / 2; if (cdf[mid] < u) lo = mid + 1; else hi = mid; }
    *offset = lo - 1;
    *pdf = integral > 0.f ? func[*offset] / integral : 0.f;
}
// Distribution1D::new (:25-63)
inline void distribution_new(const Float* func, uint32_t n, Float* cdf, Float* integral) {
    cdf[0] = 0.f;
    for (uint32_t i = 0; i < n; i++) cdf[i + 1] = cdf[i] + func[i];
    Float fi = cdf[n];
    if (fi == 0.f) { for (uint32_t i = 1; i < n + 1; i++) cdf[i] = (Float)i / (Float)(n + 1); }
    else { for (uint32_t i = 1; i < n + 1; i++) cdf[i] /= fi; }
    *integral = fi;
}

// LanczosSincFilter (sample/filters.rs:193-241), evaluated with SIGNED offsets (quirk A-4)
inline Float lanczos_sinc1(Float x) { if (x < 1.0e-5f) return 1.f; Float xpi = x * pi(); return fsin(xpi) / xpi; }
inline Float lanczos_sinc(Float x, Float inv_tau) { return lanczos_sinc1(x * inv_tau) * lanczos_sinc1(x); }
inline Float lanczos_evaluate(V2 p, Float inv_tau) { return lanczos_sinc(p.x, inv_tau) * lanczos_sinc(p.y, inv_tau); }
// MitchellFilter::mitchell_1d (filters.rs:154-169)
inline Float mitchell_1d(Float x, Float b, Float c) {
    const Float INV_SIX = 1.0f / 6.0f;
    if (x > 1.0f) {
        return (-b - 6.0f * c) * x * x * x + (6.0f * b + 30.0f * c) * x * x - (12.0f * b + 48.0f * c) * x + (8.0f * b + 24.0f * c) * INV_SIX;
    } else {
        return (12.0f - 9.0f * b - 6.0f * c) * x * x * x + (-18.0f - 12.0f * b + 6.0f * c) * x * x + (6.0f - 2.0f * b) * INV_SIX;
    }
}
// Filter::evaluate_unsafe of the film's filter (sample/filters.rs), `p` = pixel centre - sample position (signed)
inline Float filter_evaluate(const arn_film& film, V2 p) {
    Float rx = film.filter_radius_x, ry = film.filter_radius_y;
    switch (film.filter_kind) {
    case ARN_FILTER_BOX: return 1.0f;                                                              // :55-57
    case ARN_FILTER_TRIANGLE: return (rx - std::fabs(p.x)) * (ry - std::fabs(p.y));                // :81-83
    case ARN_FILTER_GAUSSIAN: {                                                                    // :101-126
        Float neg_alpha = -film.filter_a;
        Float ex = neg_alpha * rx * rx, ey = neg_alpha * ry * ry;                                  // sic: the exponent, not its exponential
        Float gx = fexp(neg_alpha * p.x * p.x) - ex;
        Float gy = fexp(neg_alpha * p.y * p.y) - ey;
        return gx * gy;
    }
    case ARN_FILTER_MITCHELL: {                                                                    // :140-186
        Float ix = 1.0f / rx, iy = 1.0f / ry;
        Float mx = 2.0f * (ix * p.x), my = 2.0f * (iy * p.y);
        return mitchell_1d(std::fabs(mx), film.filter_a, film.filter_b) * mitchell_1d(std::fabs(my), film.filter_a, film.filter_b);
    }
    default: {                                                                                     // :189-240
        Float tau = film.filter_a > 0.f ? film.filter_a : 3.0f;
        return lanczos_evaluate(p, 1.0f / tau);
    }
    }
}

// ------------------------------------------------------------------ BxDFs
enum { BXDF_REFLECTION = 0x01, BXDF_TRANSMISSION = 0x02, BXDF_DIFFUSE = 0x04, BXDF_GLOSSY = 0x08,
       BXDF_SPECULAR = 0x10, BXDF_ALL = 0x1f };                                   // bxdf/mod.rs:119-131
enum BxdfKind { BX_LAMBERT_R, BX_LAMBERT_T, BX_OREN_NAYAR, BX_FRESNEL, BX_TS_R, BX_TS_T, BX_ASHIKHMIN };
enum DistKind { DIST_BECKMANN, DIST_TROWBRIDGE };

struct Bxdf {
    BxdfKind kind;
    RGB a, b;              // reflectance / transmittance / diffuse, specular
    Float c0, c1;          // OrenNayar coef_a, coef_b; eta0, eta1
    DistKind dist; Float ax, ay;
};
struct Sampled { RGB f; V3 wi; Float pdf; uint32_t type; };

inline uint32_t bxdf_type(const Bxdf& x) {
    switch (x.kind) {
    case BX_LAMBERT_R: case BX_OREN_NAYAR: return BXDF_REFLECTION | BXDF_DIFFUSE;   // lambertian.rs:30-32, oren_nayar.rs:38-40
    case BX_LAMBERT_T: return BXDF_TRANSMISSION | BXDF_DIFFUSE;                     // lambertian.rs:67-69
    case BX_FRESNEL: return BXDF_REFLECTION | BXDF_TRANSMISSION | BXDF_SPECULAR;    // fresnel.rs:154-156
    case BX_TS_R: case BX_ASHIKHMIN: return BXDF_REFLECTION | BXDF_GLOSSY;          // microfacet.rs:391-393,569-571
    case BX_TS_T: return BXDF_TRANSMISSION | BXDF_GLOSSY;                           // microfacet.rs:456-458
    }
    return 0;
}
inline bool bxdf_is(const Bxdf& x, uint32_t t) { return (bxdf_type(x) & t) != 0; }  // intersects, bxdf/mod.rs:23-25

// microfacet helpers (bxdf/microfacet.rs)
inline Float roughness_to_alpha(Float roughness) {                                // :57-63
    Float x = flog(fmax_(roughness, 1e-3f));
    return 1.62142f + 0.819955f * x + 0.1734f * x * x + 0.0171201f * x * x * x + 0.000640711f * x * x * x * x;
}
inline Float erf_inv(Float x) {                                                    // :313-343
    x = fmin_(fmax_(x, -0.99999f), 0.99999f);
    Float w = -flog((1.f - x) * (1.f + x));
    Float p;
    if (w < 5.f) {
        w = w - 2.5f;
        p = 2.81022636e-08f; p = 3.43273939e-07f + p * w; p = -3.5233877e-06f + p * w; p = -4.39150654e-06f + p * w;
        p = 0.00021858087f + p * w; p = -0.00125372503f + p * w; p = -0.00417768164f + p * w; p = 0.246640727f + p * w;
        p = 1.50140941f + p * w;
    } else {
        w = std::sqrt(w) - 3.f;
        p = -0.000200214257f; p = 0.000100950558f + p * w; p = 0.00134934322f + p * w; p = -0.00367342844f + p * w;
        p = 0.00573950773f + p * w; p = -0.0076224613f + p * w; p = 0.00943887047f + p * w; p = 1.00167406f + p * w;
        p = 2.83297682f + p * w;
    }
    return p * x;
}
inline Float erf_approx(Float x) {                                                 // :346-365
    const Float A1 = 0.254829592f, A2 = -0.28449673f, A3 = 1.421413741f, A4 = -1.453152027f, A5 = 1.061405429f, P = 0.3275911f;
    Float sign = signum(x);
    x = x * sign;
    Float t = 1.f / (1.f + P * x);
    Float y = 1.f - (((((A5 * t + A4) * t) + A3) * t + A2) * t + A1) * t * fexp(-x * x);
    return sign * y;
}
inline Float dist_D(DistKind k, Float ax, Float ay, V3 wh) {
    Float cos2_theta = nrm::cos2_theta(wh), tan2_theta = nrm::tan2_theta(wh);
    if (k == DIST_BECKMANN) {                                                      // :84-93
        Float cos2_phi = nrm::cos2_phi(wh), sin2_phi = nrm::sin2_phi(wh);
        return fexp(-tan2_theta * (cos2_phi / (ax * ax) + sin2_phi / (ay * ay))) / (pi() * ax * ay * cos2_theta * cos2_theta);
    }
    if (std::isinf(tan2_theta)) return 0.f;                                        // :145-158
    Float cos2_phi = nrm::cos2_phi(wh), sin2_phi = nrm::sin2_phi(wh);
    Float last_term = 1.f + tan2_theta * (cos2_phi / (ax * ax) + sin2_phi / (ay * ay));
    return 1.f / (pi() * ax * ay * cos2_theta * cos2_theta * last_term * last_term);
}
inline Float dist_lambda(DistKind k, Float ax, Float ay, V3 w) {
    if (k == DIST_BECKMANN) {                                                      // :96-123
        Float tant = std::fabs(nrm::tan_theta(w));
        if (std::isinf(tant) || std::isnan(tant)) return 0.f;
        Float alpha = std::sqrt(nrm::cos2_phi(w) * ax * ax + nrm::sin2_phi(w) * ay * ay);
        Float a = 1.f / (alpha * tant);
        if (a >= 1.6f) return 0.f;
        return (1.f - 1.259f * a + 0.396f * a * a) / (3.535f * a + 2.181f * a * a);
    }
    Float tabs = std::fabs(nrm::tan_theta(w));                                      // :161-170
    if (std::isinf(tabs)) return 0.f;
    Float alpha = std::sqrt(nrm::cos2_phi(w) * ax * ax + nrm::sin2_phi(w) * ay * ay);
    Float term = alpha * tabs;
    return (-1.f + std::sqrt(1.f + term * term)) * 0.5f;
}
inline Float dist_visible(DistKind k, Float ax, Float ay, V3 w) { return 1.f / (1.f + dist_lambda(k, ax, ay, w)); }          // :32-34
inline Float dist_visible_both(DistKind k, Float ax, Float ay, V3 w0, V3 w1) {                                            // :40-42
    return 1.f / (1.f + dist_lambda(k, ax, ay, w0) + dist_lambda(k, ax, ay, w1));
}
inline Float dist_pdf(DistKind k, Float ax, Float ay, V3 wo, V3 wh) {                                                     // :48-51
    return dist_D(k, ax, ay, wh) * dist_visible(k, ax, ay, wo) * std::fabs(dot(wo, wh)) / std::fabs(nrm::cos_theta(wo));
}
inline V3 sample_wh_beckmann(V3 wo, V2 u, Float ax, Float ay) {                                                            // :181-258
    V3 wo_stretched = normalize(v3(ax * wo.x, ay * wo.y, wo.z));
    Float cos_theta = std::fabs(nrm::cos_theta(wo_stretched));
    Float sx, sy;
    if (cos_theta > 0.9999f) {
        Float r = std::sqrt(-flog(u.x));
        Float phi = 2.f * u.y * pi();
        sx = r * fcos(phi); sy = r * fsin(phi);
    } else {
        Float sin_theta = std::sqrt(fmax_(1.f - cos_theta * cos_theta, 0.f));
        Float tan_theta = sin_theta / cos_theta;
        Float cot_theta = cos_theta / sin_theta;
        Float a = -1.f;
        Float c = erf_approx(cot_theta);
        Float ux = fmax_(u.x, 1e-6f);
        Float theta = facos(cos_theta);
        Float fit = 1.f + theta * (-0.876f + theta * (0.4265f - 0.0594f * theta));
        Float b = c - (1.f + c) * fpow(1.f - ux, fit);
        Float sqrt_pi_inv = 1.f / std::sqrt(pi());
        Float norm = 1.f / (1.f + c + sqrt_pi_inv * tan_theta * fexp(-cot_theta * cot_theta));
        for (int it = 1; it < 10; it++) {
            if (b < a || b > c) b = 0.5f * (a + c);
            Float inv = erf_inv(b);
            Float value = norm * (1.f + b + sqrt_pi_inv * tan_theta * fexp(-inv * inv)) - ux;
            if (std::fabs(value) < 1e-5f) break;
            Float derivation = norm * (1.f - inv * tan_theta);
            if (value > 0.f) c = b; else a = b;
            b -= value / derivation;
        }
        sx = erf_inv(b);
        sy = erf_inv(2.f * fmax_(u.y, 1e-6f) - 1.f);
    }
    Float cos_phi = nrm::cos_phi(wo_stretched), sin_phi = nrm::sin_phi(wo_stretched);
    Float rotation_tmp = cos_phi * sx - sin_phi * sy;
    sy = sin_phi * sx + cos_phi * sy;
    sx = rotation_tmp;
    sx *= ax; sy *= ay;
    return normalize(v3(-sx, -sy, 1.f)) * signum(wo.z);
}
inline V3 sample_wh_trowbridge_pos(V3 wo, V2 u, Float ax, Float ay) {                                                      // :260-309
    V3 wo_stretched = normalize(v3(ax * wo.x, ay * wo.y, wo.z));
    Float cos_theta = std::fabs(nrm::cos_theta(wo_stretched));
    Float sx, sy;
    if (cos_theta > 0.9999f) {
        Float r = std::sqrt(u.x / (1.f - u.x));
        Float phi = 2.f * u.y * pi();
        sx = r * fcos(phi); sy = r * fsin(phi);
    } else {
        Float sin_theta = std::sqrt(fmax_(1.f - cos_theta * cos_theta, 0.f));
        Float tan_theta = sin_theta / cos_theta;
        Float cot_theta = cos_theta / sin_theta;
        Float g1 = 2.f / (1.f + std::sqrt(1.f + 1.f / (cot_theta * cot_theta)));
        Float a = 2.f * u.y / g1 - 1.f;
        Float tmp = fmin_(1.f / (a * a - 1.f), 1e10f);
        Float d = std::sqrt(fmax_(tan_theta * tan_theta * tmp * tmp - (a * a - tan_theta * tan_theta) * tmp, 0.f));
        Float sx1 = tan_theta * tmp - d;
        Float sx2 = tan_theta * tmp + d;
        Float sxx = (a < 0.f || sx2 > cot_theta) ? sx1 : sx2;
        Float s, uy;
        if (u.y > 0.5f) { s = 1.f; uy = 2.f * (u.y - 0.5f); } else { s = -1.f; uy = 2.f * (0.5f - u.y); }
        Float z = (uy * (uy * (uy * 0.27385f - 0.73369f) + 0.46341f)) / (uy * (uy * (uy * 0.093073f + 0.309420f) - 1.000000f) + 0.597999f);
        sx = sxx; sy = s * z * (1.f + sxx * sxx);
    }
    Float cos_phi = nrm::cos_phi(wo_stretched), sin_phi = nrm::sin_phi(wo_stretched);
    Float rotation_tmp = cos_phi * sx - sin_phi * sy;
    sy = sin_phi * sx + cos_phi * sy;
    sx = rotation_tmp;
    sx *= ax; sy *= ay;
    return normalize(v3(-sx, -sy, 1.f));
}
inline V3 dist_sample_wh(DistKind k, Float ax, Float ay, V3 wo, V2 u) {
    if (k == DIST_BECKMANN) return sample_wh_beckmann(wo, u, ax, ay);             // :126-128
    V3 won = wo.z < 0.f ? -wo : wo;                                               // :173-178
    V3 wh = sample_wh_trowbridge_pos(won, u, ax, ay);
    return wo.z < 0.f ? -wh : wh;
}

// fresnel (bxdf/fresnel.rs:16-37)
inline Float fresnel_dielectric(Float cos_theta_i, Float etai, Float etat) {
    if (cos_theta_i < 0.f) { Float t = etai; etai = etat; etat = t; cos_theta_i = -cos_theta_i; }
    Float sin2_theta_i = fmax_(1.f - cos_theta_i * cos_theta_i, 0.f);
    Float eta = etai / etat;
    Float sin2_theta_t = eta * eta * sin2_theta_i;
    if (sin2_theta_t >= 1.f) return 1.f;
    Float cos_theta_t = std::sqrt(1.f - sin2_theta_t);
    Float etci = etat * cos_theta_i, eict = etai * cos_theta_t;
    Float r_para = (etci - eict) / (etci + eict);
    Float eici = etai * cos_theta_i, etct = etat * cos_theta_t;
    Float r_perp = (eici - etct) / (eici + etct);
    return (r_para * r_para + r_perp * r_perp) * 0.5f;
}
// f32::powi(5): llvm.powi with a constant exponent expands by binary decomposition
// (SelectionDAG ExpandPowI; compiler-rt __powisf2 multiplies in the same order): x * (x^2)^2.
inline Float powi5(Float x) { Float x2 = x * x; Float x4 = x2 * x2; return x * x4; }
inline RGB schlick_fresnel(Float cost, RGB s) { return s + powi5(1.f - cost) * (grey(1.f) - s); }      // microfacet.rs:626-629

// Bxdf::pdf
inline Float bxdf_pdf(const Bxdf& x, V3 wo, V3 wi) {
    switch (x.kind) {
    case BX_LAMBERT_R: case BX_OREN_NAYAR:                                          // default, bxdf/mod.rs:77-83
        return wo.z * wi.z > 0.f ? std::fabs(nrm::cos_theta(wi)) * frac_1_pi() : 0.f;
    case BX_LAMBERT_T:                                                              // lambertian.rs:96-102
        return wo.z * wi.z >= 0.f ? 0.f : std::fabs(nrm::cos_theta(wi)) * frac_1_pi();
    case BX_FRESNEL: return 0.f;                                                    // fresnel.rs:199-201
    case BX_TS_R: {                                                                 // microfacet.rs:423-429
        if (wo.z * wi.z <= 0.f) return 0.f;
        V3 wh = normalize(wo + wi);
        return dist_pdf(x.dist, x.ax, x.ay, wo, wh) / (4.f * dot(wo, wh));
    }
    case BX_TS_T: {                                                                 // :514-532
        if (wo.z * wi.z > 0.f) return 0.f;
        Float eta = wo.z > 0.f ? x.c1 / x.c0 : x.c0 / x.c1;
        V3 wh = normalize(wo + wi * eta);
        if (isinf3(wh) || isnan3(wh)) return 1.f;
        Float sqrt_denom = dot(wo, wh) + eta * dot(wi, wh);
        Float dhdi = eta * eta * std::fabs(dot(wi, wh)) / (sqrt_denom * sqrt_denom);
        return dist_pdf(x.dist, x.ax, x.ay, wo, wh) * dhdi;
    }
    case BX_ASHIKHMIN: {                                                            // :613-623
        if (wo.z * wi.z < 0.f) return 0.f;
        V3 wh = normalize(wo + wi);
        return 0.5f * (dist_pdf(x.dist, x.ax, x.ay, wo, wh) / (4.f * dot(wo, wh)) + std::fabs(nrm::cos_theta(wi)) * frac_1_pi());
    }
    }
    return 0.f;
}
// Bxdf::evaluate
inline RGB bxdf_eval(const Bxdf& x, V3 wo, V3 wi) {
    switch (x.kind) {
    case BX_LAMBERT_R: case BX_LAMBERT_T: return x.a * frac_1_pi();                  // lambertian.rs:35-37,72-74
    case BX_OREN_NAYAR: {                                                           // oren_nayar.rs:42-60
        Float sin_theta_i = nrm::sin_theta(wi), sin_theta_o = nrm::sin_theta(wo);
        Float max_cos = 0.f;
        if (sin_theta_i > 1e-4f || sin_theta_o > 1e-4f) {
            Float sin_phi_i = nrm::sin_phi(wi), sin_phi_o = nrm::sin_phi(wo);
            Float cos_phi_i = nrm::cos_phi(wi), cos_phi_o = nrm::cos_phi(wo);
            max_cos = fmax_(max_cos, cos_phi_i * cos_phi_o + sin_phi_i * sin_phi_o);
        }
        Float ci = std::fabs(nrm::cos_theta(wi)), co = std::fabs(nrm::cos_theta(wo));
        Float sin_a, tan_b;
        if (ci > co) { sin_a = sin_theta_o; tan_b = sin_theta_i / ci; } else { sin_a = sin_theta_i; tan_b = sin_theta_o / co; }
        return x.a * frac_1_pi() * (x.c0 + x.c1 * max_cos * sin_a * tan_b);
    }
    case BX_FRESNEL: return grey(0.f);                                              // fresnel.rs:159-161
    case BX_TS_R: {                                                                 // microfacet.rs:395-406
        V3 wh = normalize(wo + wi);
        if (isnan3(wh)) return grey(0.f);
        return x.a * dist_D(x.dist, x.ax, x.ay, wh) * dist_visible_both(x.dist, x.ax, x.ay, wo, wi)
             * grey(fresnel_dielectric(dot(wi, wh), x.c0, x.c1)) / (4.f * std::fabs(wo.z) * std::fabs(wi.z));
    }
    case BX_TS_T: {                                                                 // :460-491
        if (wo.z * wi.z > 0.f) return grey(0.f);
        Float eta = wo.z > 0.f ? x.c1 / x.c0 : x.c0 / x.c1;
        V3 wh = normalize(wo + wi * eta);
        if (isinf3(wh) || isnan3(wh)) return grey(1.f);
        if (wh.z < 0.f) wh = -wh;
        Float cosoh = dot(wo, wh);
        RGB f = grey(fresnel_dielectric(cosoh, x.c0, x.c1));
        Float cosih = dot(wi, wh);
        Float sqrt_denom = cosoh + eta * cosih;
        return x.a * dist_D(x.dist, x.ax, x.ay, wh) * dist_visible_both(x.dist, x.ax, x.ay, wo, wi)
             * (grey(1.f) - f) * std::fabs(cosih) * std::fabs(cosoh)
             / (std::fabs(nrm::cos_theta(wo)) * std::fabs(nrm::cos_theta(wi)) * sqrt_denom * sqrt_denom);
    }
    case BX_ASHIKHMIN: {                                                            // :573-595
        V3 wh = wo + wi;
        if (relative_eq(magnitude2(wh), 0.f)) return grey(0.f);
        wh = normalize(wh);
        Float to = 1.f - powi5(1.f - 0.5f * std::fabs(nrm::cos_theta(wo)));
        Float ti = 1.f - powi5(1.f - 0.5f * std::fabs(nrm::cos_theta(wi)));
        RGB diffuse = (28.f / (23.f * pi())) * x.a * (grey(1.f) - x.b) * to * ti;
        RGB specular = dist_D(x.dist, x.ax, x.ay, wh) * schlick_fresnel(dot(wi, wh), x.b)
            / (4.f * std::fabs(dot(wi, wh)) * fmax_(std::fabs(nrm::cos_theta(wi)), std::fabs(nrm::cos_theta(wo))));
        return diffuse + specular;
    }
    }
    return grey(0.f);
}
// Bxdf::evaluate_sampled
inline Sampled bxdf_sample(const Bxdf& x, V3 wo, V2 u) {
    Sampled r; r.type = bxdf_type(x);
    switch (x.kind) {
    case BX_LAMBERT_R: case BX_OREN_NAYAR: {                                        // default, bxdf/mod.rs:42-48
        V3 wi = sample_cosw_hemisphere(u);
        if (wo.z < 0.f) wi.z = -wi.z;
        r.pdf = bxdf_pdf(x, wo, wi); r.f = bxdf_eval(x, wo, wi); r.wi = wi; return r;
    }
    case BX_LAMBERT_T: {                                                            // lambertian.rs:87-93
        V3 wi = sample_cosw_hemisphere(u);
        if (wo.z > 0.f) wi.z = -wi.z;
        r.pdf = bxdf_pdf(x, wo, wi); r.f = bxdf_eval(x, wo, wi); r.wi = wi; return r;
    }
    case BX_FRESNEL: {                                                              // fresnel.rs:163-196
        Float cos_theta = nrm::cos_theta(wo);
        Float f = fresnel_dielectric(cos_theta, x.c0, x.c1);
        if (u.x < f) {
            r.wi = v3(-wo.x, -wo.y, wo.z); r.pdf = f;
            r.f = r.pdf * x.a / std::fabs(cos_theta);
            r.type = BXDF_REFLECTION | BXDF_SPECULAR; return r;
        }
        Float pdf = 1.f - f;
        Float etai, etao; V3 n;
        if (cos_theta > 0.f) { etai = x.c0; etao = x.c1; n = v3(0.f, 0.f, 1.f); }
        else { etai = x.c1; etao = x.c0; n = v3(0.f, 0.f, -1.f); }
        Float eta = etai / etao;
        V3 wt;
        r.type = BXDF_TRANSMISSION | BXDF_SPECULAR; r.pdf = pdf;
        if (nrm::refract(wo, n, eta, &wt)) { r.f = x.b * eta * eta * pdf / std::fabs(wt.z); r.wi = wt; }
        else { r.f = grey(0.f); r.wi = v3(0, 0, 0); }
        return r;
    }
    case BX_TS_R: {                                                                 // microfacet.rs:408-421
        V3 wh = dist_sample_wh(x.dist, x.ax, x.ay, wo, u);
        r.pdf = dist_pdf(x.dist, x.ax, x.ay, wo, wh) / (4.f * dot(wo, wh));
        V3 wi = normalize(2.f * wh * dot(wo, wh) - wo);
        r.wi = wi;
        r.f = (wo.z * wi.z <= 0.f) ? grey(0.f) : bxdf_eval(x, wo, wi);
        return r;
    }
    case BX_TS_T: {                                                                 // :493-511
        V3 wh = dist_sample_wh(x.dist, x.ax, x.ay, wo, u);
        Float eta = wo.z > 0.f ? x.c0 / x.c1 : x.c1 / x.c0;
        V3 wi;
        if (nrm::refract(wo, wh, eta, &wi)) { r.pdf = bxdf_pdf(x, wo, wi); r.f = bxdf_eval(x, wo, wi); r.wi = wi; }
        else { r.f = grey(0.f); r.wi = v3(0, 0, 0); r.pdf = 0.f; }
        return r;
    }
    case BX_ASHIKHMIN: {                                                            // :597-611
        V3 wi;
        if (u.x < 0.5f) {
            u.x *= 2.f;
            V3 wh = dist_sample_wh(x.dist, x.ax, x.ay, wo, u);
            wi = normalize(2.f * wh * dot(wo, wh) - wo);
            if (wo.z * wi.z <= 0.f) { r.f = grey(0.f); r.wi = wi; r.pdf = bxdf_pdf(x, wo, wi); return r; }
        } else {
            u.x = (1.f - u.x) * 2.f;
            wi = sample_cosw_hemisphere(u);
            if (wi.z < 0.f) wi.z = -wi.z;
        }
        r.f = bxdf_eval(x, wo, wi); r.wi = wi; r.pdf = bxdf_pdf(x, wo, wi); return r;
    }
    }
    return r;
}

// ------------------------------------------------------------------ Bsdf (material/bsdf.rs)
struct Bsdf {
    Float eta; V3 ns, ng, ts, bs; Bxdf bx[8]; int n;
};
inline Bsdf bsdf_new(const SurfaceInteraction& si, Float eta) {                     // :36-46
    Bsdf b; b.eta = eta;
    b.ts = normalize(si.shading_duv.dpdu);
    b.ns = si.shading_norm;
    b.bs = normalize(cross(b.ns, b.ts));
    b.ng = si.basic.norm; b.n = 0; return b;
}
inline int bsdf_have_n(const Bsdf& b, uint32_t kind) { int c = 0; for (int i = 0; i < b.n; i++) if (bxdf_is(b.bx[i], kind)) c++; return c; } // :54-63
inline V3 parent_to_local(const Bsdf& b, V3 v) { return v3(dot(v, b.ts), dot(v, b.bs), dot(v, b.ns)); }  // :67-69
inline V3 local_to_parent(const Bsdf& b, V3 v) {                                     // :72-79
    return v3(dot(v, v3(b.ts.x, b.bs.x, b.ns.x)), dot(v, v3(b.ts.y, b.bs.y, b.ns.y)), dot(v, v3(b.ts.z, b.bs.z, b.ns.z)));
}
inline RGB bsdf_evaluate(const Bsdf& b, V3 wow, V3 wiw, uint32_t types) {            // :82-98
    V3 wo = normalize(parent_to_local(b, wow)), wi = normalize(parent_to_local(b, wiw));
    bool is_reflection = dot(wow, b.ng) * dot(wiw, b.ng) > 0.f;
    RGB ret = grey(0.f);
    for (int i = 0; i < b.n; i++) {
        const Bxdf& x = b.bx[i]; uint32_t k = bxdf_type(x);
        if (bxdf_is(x, types) && ((is_reflection && (k & BXDF_REFLECTION)) || (!is_reflection && (k & BXDF_TRANSMISSION))))
            ret = ret + bxdf_eval(x, wo, wi);
    }
    return ret;
}
inline Float bsdf_pdf(const Bsdf& b, V3 wow, V3 wiw, uint32_t types) {               // :205-222
    V3 wo = normalize(parent_to_local(b, wow)), wi = normalize(parent_to_local(b, wiw));
    if (wo.z == 0.f) return 0.f;
    Float pdfsum = 0.f; int match_count = 0;
    for (int i = 0; i < b.n; i++) if (bxdf_is(b.bx[i], types)) { match_count++; pdfsum += fmax_(bxdf_pdf(b.bx[i], wo, wi), 0.f); }
    return match_count == 0 ? pdfsum : pdfsum / (Float)match_count;
}
inline Sampled bsdf_evaluate_sampled(const Bsdf& b, V3 wow, V2 u, uint32_t types) {  // :100-145
    int match_count = bsdf_have_n(b, types);
    Sampled ret; ret.f = grey(0.f); ret.wi = v3(0.f, 1.f, 0.f); ret.pdf = 0.f; ret.type = 0;
    if (match_count == 0) return ret;
    V3 wo = normalize(parent_to_local(b, wow));
    int idx = (int)std::floor(u.x * (Float)match_count); if (idx > match_count - 1) idx = match_count - 1;
    int i = 0; bool is_specular = false;
    for (int k = 0; k < b.n; k++) {
        const Bxdf& x = b.bx[k];
        if (i == idx) {
            is_specular = bxdf_is(x, BXDF_SPECULAR);
            Sampled s = bxdf_sample(x, wo, u);
            if (s.pdf == 0.f) { Sampled z; z.f = grey(0.f); z.wi = v3(0.f, 1.f, 0.f); z.pdf = 0.f; z.type = 0; return z; }
            ret = s; ret.type = s.type & types;
        }
        if (bxdf_is(x, types)) i++;
    }
    V3 wi = ret.wi;
    ret.wi = local_to_parent(b, wi);
    if (match_count == 1 || is_specular) return ret;
    ret.f = grey(0.f);
    bool is_reflection = dot(wow, b.ng) * dot(ret.wi, b.ng) > 0.f;
    Float pdfsum = 0.f;
    for (int k = 0; k < b.n; k++) {
        const Bxdf& x = b.bx[k];
        if (bxdf_is(x, ret.type) && ((is_reflection && bxdf_is(x, BXDF_REFLECTION)) || (!is_reflection && bxdf_is(x, BXDF_TRANSMISSION)))) {
            ret.f = ret.f + bxdf_eval(x, wo, wi);
            pdfsum += fmax_(bxdf_pdf(x, wo, wi), 0.f);
        }
    }
    ret.pdf = pdfsum / (Float)match_count;
    return ret;
}

// ------------------------------------------------------------------ materials (material/*.rs), constant textures
inline Bxdf mk_bxdf(BxdfKind k) { Bxdf x; std::memset(&x, 0, sizeof x); x.kind = k; return x; }
// Textured form (material/{matte,plastic,glass,translucent}.rs with ImageTexture parameters): bump first, then every texture is
// evaluated on the (bumped) interaction; `tex` = the scene's texture table (null: constants only).
struct TexTable { const arn_texture* textures; const Float* texels; };
inline Bsdf compute_scattering(const arn_material& m_in, SurfaceInteraction& si, const DxyInfo* dxy = nullptr, const TexTable* tex = nullptr) {
    arn_material m = m_in;
    if (tex && dxy) {
        auto view = [&](uint32_t id) { TexView v; v.t = &tex->textures[id - 1]; v.texels = tex->texels; return v; };
        if (m.bump_tex) add_bumping(si, *dxy, view(m.bump_tex));                    // matte.rs:46-48 and alike
        // evaluation order as in the sources: Matte kd, sigma; Plastic diffuse, specular, roughness; Glass specular, diffuse, roughness
        if (m.kd_tex) { Texel t = texture_evaluate(view(m.kd_tex), si.uv, *dxy); m.kd[0] = t.c[0]; m.kd[1] = t.c[1]; m.kd[2] = t.c[2]; }
        if (m.ks_tex) { Texel t = texture_evaluate(view(m.ks_tex), si.uv, *dxy); m.ks[0] = t.c[0]; m.ks[1] = t.c[1]; m.ks[2] = t.c[2]; }
        if (m.aux_tex) {
            Float a = texture_evaluate(view(m.aux_tex), si.uv, *dxy).c[0];
            if (m.type == ARN_MAT_MATTE) m.sigma = a; else { m.roughness = a; m.alpha = roughness_to_alpha(a); }
        }
    }
    Bsdf b = bsdf_new(si, 1.f);
    RGB kd = rgb(m.kd[0], m.kd[1], m.kd[2]), ks = rgb(m.ks[0], m.ks[1], m.ks[2]);
    switch (m.type) {
    case ARN_MAT_MATTE: {                                                          // matte.rs:39-64
        Float sig = clampf(m.sigma, 0.f, 90.f);
        if (!is_black(kd)) {
            if (sig == 0.f) { Bxdf x = mk_bxdf(BX_LAMBERT_R); x.a = kd; b.bx[b.n++] = x; }
            else {                                                                 // OrenNayer::new, oren_nayar.rs:31-41
                Bxdf x = mk_bxdf(BX_OREN_NAYAR); x.a = kd;
                Float sigma2 = sig * sig;
                x.c0 = 1.f - (sigma2 / (2.f * (sigma2 + 0.33f)));
                x.c1 = (0.45f * sigma2) / (sigma2 + 0.09f);
                b.bx[b.n++] = x;
            }
        }
        break; }
    case ARN_MAT_PLASTIC: {                                                        // plastic.rs:40-63
        Bxdf x = mk_bxdf(BX_ASHIKHMIN); x.a = rgb_clamp(kd, 0.f, 1.f); x.b = rgb_clamp(ks, 0.f, 1.f);
        x.dist = DIST_BECKMANN; x.ax = m.alpha; x.ay = m.alpha; b.bx[b.n++] = x;
        break; }
    case ARN_MAT_GLASS: {                                                          // glass.rs:42-80
        if (!is_black(ks)) { Bxdf x = mk_bxdf(BX_FRESNEL); x.a = ks; x.b = ks; x.c0 = 1.f; x.c1 = m.eta; b.bx[b.n++] = x; }
        if (!is_black(kd)) {
            Bxdf r = mk_bxdf(BX_TS_R); r.a = kd; r.dist = DIST_TROWBRIDGE; r.ax = m.alpha; r.ay = m.alpha; r.c0 = 1.f; r.c1 = m.eta; b.bx[b.n++] = r;
            Bxdf t = mk_bxdf(BX_TS_T); t.a = kd; t.dist = DIST_TROWBRIDGE; t.ax = m.alpha; t.ay = m.alpha; t.c0 = 1.f; t.c1 = m.eta; b.bx[b.n++] = t;
        }
        break; }
    case ARN_MAT_TRANSLUCENT: {                                                    // translucent.rs:42-75
        if (!relative_eq(m.dissolve, 0.f)) {
            Bxdf x = mk_bxdf(BX_ASHIKHMIN); x.a = rgb_clamp(kd * m.dissolve, 0.f, 1.f); x.b = rgb_clamp(ks * m.dissolve, 0.f, 1.f);
            x.dist = DIST_TROWBRIDGE; x.ax = m.alpha; x.ay = m.alpha; b.bx[b.n++] = x;
        }
        if (!is_black(kd)) { Bxdf x = mk_bxdf(BX_LAMBERT_T); x.a = kd * (1.f - m.dissolve); b.bx[b.n++] = x; }
        break; }
    }
    return b;
}

}  // namespace orc
