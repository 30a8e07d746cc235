// Device-side f32 helpers for the sm_100a path-tracing kernels.
// The whole library is compiled with --fmad=false, IEEE division and square root
// (nvcc defaults -prec-div=true -prec-sqrt=true -ftz=false), so every expression below
// rounds exactly like the reference's scalar Rust (no FMA contraction) — this is what
// makes closest-hit primitive ids bit-exact (SURVEY.md §7 "Hard parts").
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "cr_math.cuh"

#define ARN_DEV __device__ __forceinline__
#define ARN_INF __int_as_float(0x7f800000)
// one out-of-line copy per translation unit: keeps the shade kernel inside the instruction cache
#define ARN_NOINL static __device__ __noinline__

namespace arn {

ARN_DEV float3 f3(float x, float y, float z) { return make_float3(x, y, z); }
ARN_DEV float3 operator+(float3 a, float3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
ARN_DEV float3 operator-(float3 a, float3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
ARN_DEV float3 operator-(float3 a) { return f3(-a.x, -a.y, -a.z); }
ARN_DEV float3 operator*(float3 a, float s) { return f3(a.x * s, a.y * s, a.z * s); }
ARN_DEV float3 operator*(float s, float3 a) { return f3(s * a.x, s * a.y, s * a.z); }
ARN_DEV float3 operator*(float3 a, float3 b) { return f3(a.x * b.x, a.y * b.y, a.z * b.z); }
ARN_DEV float3 operator/(float3 a, float s) { return f3(a.x / s, a.y / s, a.z / s); }
ARN_DEV float dot(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
ARN_DEV float length2(float3 a) { return dot(a, a); }
ARN_DEV float length(float3 a) { return sqrtf(dot(a, a)); }
ARN_DEV float3 normalize(float3 a) { return a * (1.f / length(a)); }      // cgmath normalize_to(1)
ARN_DEV float3 cross(float3 a, float3 b) {
    return f3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
ARN_DEV float2 f2(float x, float y) { return make_float2(x, y); }
ARN_DEV bool any_nan(float3 a) { return isnan(a.x) || isnan(a.y) || isnan(a.z); }
ARN_DEV bool any_inf(float3 a) { return isinf(a.x) || isinf(a.y) || isinf(a.z); }
ARN_DEV bool is_black(float3 a) { return a.x == 0.f && a.y == 0.f && a.z == 0.f; }
ARN_DEV float3 grey(float v) { return f3(v, v, v); }
ARN_DEV float axis_of(float3 v, int a) { return a == 0 ? v.x : (a == 1 ? v.y : v.z); }

// geometry/float.rs
ARN_DEV float clampf(float f, float lo, float hi) { return f < lo ? lo : (f < hi ? f : hi); }
#define ARN_EPS 1.1920929e-7f                    /* f32::EPSILON */
#define ARN_PI 3.14159265358979323846f
#define ARN_INV_PI 0.318309886183790671537767526745028724f
#define ARN_PI_2 1.57079632679489661923132169163975144f
#define ARN_PI_4 0.785398163397448309615660845819875721f
// eb_term(n) = n*eps_m / (1 - n*eps_m), eps_m = 2^-24 (float.rs:27-36); compile-time folded
ARN_DEV constexpr float gamma_n(float n) { return (n * 5.9604645e-8f) / (1.f - n * 5.9604645e-8f); }

ARN_DEV float next_up(float f) {                 // float.rs:104-117
    if (isinf(f) && f > 0.f) return f;
    if (f == 0.f) return 0.f;                    // `f == -0.` is true for both zeros
    uint32_t t = __float_as_uint(f);
    return (t >> 31) == 0 ? __uint_as_float(t + 1) : __uint_as_float(t - 1);
}
ARN_DEV float next_down(float f) {               // float.rs:120-133
    if (isinf(f) && f < 0.f) return f;
    if (f == 0.f) return -0.f;
    uint32_t t = __float_as_uint(f);
    return (t >> 31) != 0 ? __uint_as_float(t + 1) : __uint_as_float(t - 1);
}
ARN_DEV float signum(float x) { if (isnan(x)) return x; return (__float_as_uint(x) >> 31) ? -1.f : 1.f; }
ARN_DEV bool relative_eq(float a, float b) {     // approx::relative_eq!, eps = max_relative = f32::EPSILON
    if (a == b) return true;
    if (isinf(a) || isinf(b)) return false;
    float d = fabsf(a - b);
    if (d <= ARN_EPS) return true;
    float la = fabsf(a), lb = fabsf(b);
    return d <= (lb > la ? lb : la) * ARN_EPS;
}

// Transcendentals that can steer a discrete decision (ray geometry, lobe / branch choice) are
// evaluated as correctly-rounded f32 via f64, exactly like the oracle (oracle/geom.hpp): the
// f32 libdevice versions differ from libm by an ulp, which flips ~1 % of the paths through
// borderline shadow-ray self-intersections (the reference pulls shadow-ray ends in by only
// 2*eps).  sin / cos / log / exp / pow take a short f64 kernel first and the library routine only
// when the f32 rounding of that value is not certain (cr_math.cuh: same bits by construction).
ARN_NOINL float cr_sinf(float x) { return cr_sinf_fast(x); }
ARN_NOINL float cr_cosf(float x) { return cr_cosf_fast(x); }
// sin and cos of one argument share the range reduction (bit-identical to the separate calls)
ARN_NOINL void cr_sincosf(float x, float& s, float& c) { cr_sincosf_fast(x, s, c); }
ARN_NOINL float cr_acosf(float x) { return (float)acos((double)x); }
ARN_NOINL float cr_atan2f(float y, float x) { return (float)atan2((double)y, (double)x); }
ARN_NOINL float cr_logf(float x) { return cr_logf_fast(x); }
ARN_NOINL float cr_expf(float x) { return cr_expf_fast(x); }
ARN_NOINL float cr_powf(float a, float b) { return cr_powf_fast(a, b); }

// column-major 4x4 (cgmath): transform_point with homogeneous divide, transform_vector
struct Mat4 { float m[16]; };
ARN_DEV float3 xform_vector(const float* __restrict__ m, float3 v) {
    return f3(m[0] * v.x + m[4] * v.y + m[8] * v.z + m[12] * 0.f,
              m[1] * v.x + m[5] * v.y + m[9] * v.z + m[13] * 0.f,
              m[2] * v.x + m[6] * v.y + m[10] * v.z + m[14] * 0.f);
}
ARN_DEV float3 xform_point(const float* __restrict__ m, float3 p) {
    float x = m[0] * p.x + m[4] * p.y + m[8] * p.z + m[12] * 1.f;
    float y = m[1] * p.x + m[5] * p.y + m[9] * p.z + m[13] * 1.f;
    float z = m[2] * p.x + m[6] * p.y + m[10] * p.z + m[14] * 1.f;
    float w = m[3] * p.x + m[7] * p.y + m[11] * p.z + m[15] * 1.f;
    float iw = 1.f / w;
    return f3(x * iw, y * iw, z * iw);
}
// transform_vector by the TRANSPOSE of m (rows of m act as columns): used for transform_norm,
// where inverse_transpose(local_parent) = transpose(parent_local)
ARN_DEV float3 xform_vector_T(const float* __restrict__ m, float3 v) {
    return f3(m[0] * v.x + m[1] * v.y + m[2] * v.z + m[3] * 0.f,
              m[4] * v.x + m[5] * v.y + m[6] * v.z + m[7] * 0.f,
              m[8] * v.x + m[9] * v.y + m[10] * v.z + m[11] * 0.f);
}

// ---- ParitySampler (DESIGN.md "Sampler"): identical integer hash on CPU oracle and GPU
ARN_DEV uint32_t mix32(uint32_t h) {
    h ^= h >> 16; h *= 0x7feb352dU; h ^= h >> 15; h *= 0x846ca68bU; h ^= h >> 16; return h;
}
ARN_DEV float u01(uint32_t h) { return (float)(h >> 8) * (1.0f / 16777216.0f); }
struct Sampler {
    uint32_t k1, k2, i1d, i2d, key;
    ARN_DEV void init(uint32_t seed, uint32_t px, uint32_t py, uint32_t s, uint32_t n1, uint32_t n2) {
        init_key(mix32(mix32(mix32(mix32(seed) + px) + py) + s), n1, n2);
    }
    // the per-sample key travels with the path instead of (pixel, sample index): one word, no re-hashing per bounce
    ARN_DEV void init_key(uint32_t sample_key, uint32_t n1, uint32_t n2) {
        key = sample_key;
        k1 = mix32(key ^ 0xA511E9B3u); k2 = mix32(key ^ 0x63D83595u); i1d = n1; i2d = n2;
    }
    ARN_DEV float next() { return u01(mix32(k1 + (i1d++))); }
    ARN_DEV float2 next_2d() { float2 r = f2(u01(mix32(k2 + 2 * i2d)), u01(mix32(k2 + 2 * i2d + 1))); i2d++; return r; }
};

// ---- ARN_SAMPLER_STRATIFIED (include/arn.h): the first `ndim` 1-D / 2-D draws of a sample are placed in the stratum a hash
// permutation of [0, spp) assigns to that sample — Kensler's permute() ("Correlated Multi-Jittered Sampling", Pixar TM 13-01):
// cycle walking over the next power of two, 32-bit integer arithmetic only, so the oracle computes the same index.
ARN_DEV uint32_t hash_permute(uint32_t i, uint32_t l, uint32_t p) {
    uint32_t w = l - 1;
    w |= w >> 1; w |= w >> 2; w |= w >> 4; w |= w >> 8; w |= w >> 16;
    do {
        i ^= p; i *= 0xe170893du; i ^= p >> 16; i ^= (i & w) >> 4; i ^= p >> 8; i *= 0x0929eb3fu; i ^= p >> 23; i ^= (i & w) >> 1;
        i *= 1u | p >> 27; i *= 0x6935fa69u; i ^= (i & w) >> 11; i *= 0x74dcb303u; i ^= (i & w) >> 2; i *= 0x9e501cc3u;
        i ^= (i & w) >> 2; i *= 0xc860a3dfu; i &= w; i ^= i >> 5;
    } while (i >= l);
    return (i + p) % l;
}
// stratum index + jitter, kept strictly below the next index (f32 addition may round x + u up to x + 1)
ARN_DEV float strat_offset(uint32_t x, float u) {
    const float t = (float)x + u, hi = (float)(x + 1u);
    return t < hi ? t : __uint_as_float(__float_as_uint(hi) - 1u);
}
// pix = px | py << 16, s = the sample's index in the pixel, d = index of the draw within the sample, u = the parity sampler's draw
ARN_NOINL float strat_1d(uint32_t seed, uint32_t sdx, uint32_t sdy, uint32_t pix, uint32_t s, uint32_t d, float u) {
    const uint32_t kpix = mix32(mix32(mix32(seed) + (pix & 0xffffu)) + (pix >> 16)), n = sdx * sdy;
    return strat_offset(hash_permute(s, n, mix32(kpix ^ (0x1D000000u + d))), u) / (float)n;
}
ARN_NOINL float2 strat_2d(uint32_t seed, uint32_t sdx, uint32_t sdy, uint32_t pix, uint32_t s, uint32_t d, float2 u) {
    const uint32_t kpix = mix32(mix32(mix32(seed) + (pix & 0xffffu)) + (pix >> 16));
    const uint32_t c = hash_permute(s, sdx * sdy, mix32(kpix ^ (0x2D000000u + d)));
    return f2(strat_offset(c / sdy, u.x) / (float)sdx, strat_offset(c % sdy, u.y) / (float)sdy);
}

}  // namespace arn
