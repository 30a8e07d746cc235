// Correctly-rounded f32 transcendentals without the f64 library routine on the fast path.
//
// The parity contract (dev_math.cuh, oracle/geom.hpp): sin cos log exp pow ... of an f32 argument are the f64 library value
// rounded once to f32.  The f64 library routines (libdevice) are long — argument reduction for huge arguments, full 53-bit
// accuracy — and dominated the shade kernels' instruction footprint.  Here each function is first evaluated by a short f64
// kernel whose error is far below what deciding an f32 rounding needs:
//     y = short f64 evaluation, relative error < 2^-44 (Taylor / table kernels below, ~20 DFMA);
//     if y's 29 dropped mantissa bits are farther than 2048 units (2^-41 relative) from the round-to-nearest tie pattern,
//     (float)y is the f32 nearest to the true value AND to the library's f64 value (which is within 1 ulp_f64 of it):
//     return it;  otherwise (7.6e-6 of the calls), or outside the kernel's domain, call the library routine as before.
// So every result is bit-identical to `(float)f((double)x)` by construction.  tests/cpp/test_cr_math.cpp sweeps all 2^32
// arguments of the one-argument functions on the host (same source, glibc as the library); arn_selftest_math does the same
// on the device against libdevice (tests/test_gpu_round2.py).
//
// The file compiles as plain C++ too (host sweep): ARN_CR_HOST selects <cmath> for the library calls.
#pragma once
#include <stdint.h>
#ifdef __CUDACC__
#define ARN_CR_FN __device__ __forceinline__
#define ARN_CR_SLOW static __device__ __noinline__
#else
#include <cmath>
#include <cstring>
#define ARN_CR_FN static inline
#define ARN_CR_SLOW static
#endif

namespace arn {

// ---- bit access
ARN_CR_FN uint64_t cr_bits(double y) {
#ifdef __CUDACC__
    return (uint64_t)__double_as_longlong(y);
#else
    uint64_t u; std::memcpy(&u, &y, 8); return u;
#endif
}
ARN_CR_FN double cr_from_bits(uint64_t u) {
#ifdef __CUDACC__
    return __longlong_as_double((long long)u);
#else
    double y; std::memcpy(&y, &u, 8); return y;
#endif
}
ARN_CR_FN double cr_fma(double a, double b, double c) {
#ifdef __CUDACC__
    return __fma_rn(a, b, c);
#else
    return std::fma(a, b, c);
#endif
}
ARN_CR_FN double cr_rint(double a) {
#ifdef __CUDACC__
    return rint(a);
#else
    return std::nearbyint(a);
#endif
}

// Is (float)y certain, given that y carries a relative error below 2^-41?  True when the result is a normal f32 and the 29
// mantissa bits the conversion drops are not within 2048 units of the tie pattern 1000...0.
#define ARN_CR_BAND 2048u
ARN_CR_FN bool cr_round_certain(double y, float& out) {
    out = (float)y;                                              // round to nearest even
    const uint64_t u = cr_bits(y);
    const uint32_t ex = (uint32_t)(u >> 52) & 0x7ffu;
    const uint32_t low = (uint32_t)u & 0x1fffffffu;
    const uint32_t dist = (low - 0x10000000u + ARN_CR_BAND) & 0x1fffffffu;      // <= 2 * BAND  <=>  within BAND of the tie
    return ex >= 1023u - 126u && ex <= 1023u + 126u && dist > 2u * ARN_CR_BAND;
}

// ---- library fallbacks (one out-of-line copy per kernel; the only place the f64 routines are instantiated)
ARN_CR_SLOW float cr_slow_sin(float x) { return (float)sin((double)x); }
ARN_CR_SLOW float cr_slow_cos(float x) { return (float)cos((double)x); }
ARN_CR_SLOW void cr_slow_sincos(float x, float& s, float& c) {
#ifdef __CUDACC__
    double ds, dc; sincos((double)x, &ds, &dc); s = (float)ds; c = (float)dc;
#else
    s = (float)std::sin((double)x); c = (float)std::cos((double)x);
#endif
}
ARN_CR_SLOW float cr_slow_log(float x) { return (float)log((double)x); }
ARN_CR_SLOW float cr_slow_exp(float x) { return (float)exp((double)x); }
ARN_CR_SLOW float cr_slow_pow(float a, float b) { return (float)pow((double)a, (double)b); }

// ---- sin / cos: |x| <= 64, r = x - k pi/2 by two FMAs (x is an exact f64, so r keeps full relative accuracy even next to a
// multiple of pi/2), Taylor polynomials on |r| <= pi/4 (truncation < 5e-17)
ARN_CR_FN void cr_sincos_kernel(float xf, double& s, double& c) {
    const double x = (double)xf;
    const double kd = cr_rint(x * 0.63661977236758138);                          // 2 / pi
    const int k = (int)kd;
    double r = cr_fma(-kd, 1.5707963267948966, x);                               // pi/2 hi
    r = cr_fma(-kd, 6.123233995736766e-17, r);                                   // pi/2 lo
    const double z = r * r;
    double ps = 2.8114572543455206e-15;                                          // 1/17!
    ps = cr_fma(ps, z, -7.6471637318198164e-13);                                 // -1/15!
    ps = cr_fma(ps, z, 1.6059043836821613e-10);                                  // 1/13!
    ps = cr_fma(ps, z, -2.5052108385441720e-08);                                 // -1/11!
    ps = cr_fma(ps, z, 2.7557319223985893e-06);                                  // 1/9!
    ps = cr_fma(ps, z, -1.9841269841269841e-04);                                 // -1/7!
    ps = cr_fma(ps, z, 8.3333333333333332e-03);                                  // 1/5!
    ps = cr_fma(ps, z, -1.6666666666666666e-01);                                 // -1/3!
    const double sr = cr_fma(r * z, ps, r);
    double pc = 1.5619206968586225e-16;                                          // 1/18!
    pc = cr_fma(pc, z, -4.7794773323873853e-14);                                 // -1/16!
    pc = cr_fma(pc, z, 1.1470745597729725e-11);                                  // 1/14!
    pc = cr_fma(pc, z, -2.0876756987868100e-09);                                 // -1/12!
    pc = cr_fma(pc, z, 2.7557319223985888e-07);                                  // 1/10!
    pc = cr_fma(pc, z, -2.4801587301587302e-05);                                 // -1/8!
    pc = cr_fma(pc, z, 1.3888888888888889e-03);                                  // 1/6!
    pc = cr_fma(pc, z, -4.1666666666666664e-02);                                 // -1/4!
    pc = cr_fma(pc, z, 0.5);
    const double cr = cr_fma(-z, pc, 1.0);
    const bool swap = k & 1;
    s = swap ? cr : sr; c = swap ? sr : cr;
    if (k & 2) s = -s;
    if ((k + 1) & 2) c = -c;
}
ARN_CR_FN bool cr_trig_domain(float x) { return x >= -64.f && x <= 64.f; }          // NaN fails
ARN_CR_FN void cr_sincosf_fast(float x, float& s, float& c) {
    if (cr_trig_domain(x)) {
        double ds, dc; cr_sincos_kernel(x, ds, dc);
        float fs, fc;
        const bool ok = cr_round_certain(ds, fs) & cr_round_certain(dc, fc);
        if (ok) { s = fs; c = fc; return; }
    }
    cr_slow_sincos(x, s, c);
}
ARN_CR_FN float cr_sinf_fast(float x) {
    if (cr_trig_domain(x)) { double ds, dc; cr_sincos_kernel(x, ds, dc); float f; if (cr_round_certain(ds, f)) return f; }
    return cr_slow_sin(x);
}
ARN_CR_FN float cr_cosf_fast(float x) {
    if (cr_trig_domain(x)) { double ds, dc; cr_sincos_kernel(x, ds, dc); float f; if (cr_round_certain(dc, f)) return f; }
    return cr_slow_cos(x);
}

// ---- exp: |x| <= 80, x = k ln2 + r, Taylor to r^13 on |r| <= ln2 / 2 (truncation < 5e-18), scaled by 2^k in the exponent field
ARN_CR_FN double cr_exp_kernel(double x) {
    const double kd = cr_rint(x * 1.4426950408889634);                           // 1 / ln 2
    double r = cr_fma(-kd, 0.69314718055994529, x);                              // ln2 hi
    r = cr_fma(-kd, 2.3190468138462996e-17, r);                                  // ln2 lo
    double p = 1.6059043836821613e-10;                                           // 1/13!
    p = cr_fma(p, r, 2.0876756987868100e-09);                                    // 1/12!
    p = cr_fma(p, r, 2.5052108385441720e-08);                                    // 1/11!
    p = cr_fma(p, r, 2.7557319223985888e-07);                                    // 1/10!
    p = cr_fma(p, r, 2.7557319223985893e-06);                                    // 1/9!
    p = cr_fma(p, r, 2.4801587301587302e-05);                                    // 1/8!
    p = cr_fma(p, r, 1.9841269841269841e-04);                                    // 1/7!
    p = cr_fma(p, r, 1.3888888888888889e-03);                                    // 1/6!
    p = cr_fma(p, r, 8.3333333333333332e-03);                                    // 1/5!
    p = cr_fma(p, r, 4.1666666666666664e-02);                                    // 1/4!
    p = cr_fma(p, r, 1.6666666666666666e-01);                                    // 1/3!
    p = cr_fma(p, r, 0.5);
    p = cr_fma(p, r, 1.0);
    p = cr_fma(p, r, 1.0);
    return cr_from_bits(cr_bits(p) + ((uint64_t)(int64_t)(int)kd << 52));        // p in [0.7, 1.42): adding k to the exponent cannot overflow for |k| <= 116
}
ARN_CR_FN float cr_expf_fast(float x) {
    if (x >= -80.f && x <= 80.f) { float f; if (cr_round_certain(cr_exp_kernel((double)x), f)) return f; }
    return cr_slow_exp(x);
}

// ---- log: x a positive normal f32 = 2^e m, m in [1, 2); c = round(128 m) / 128, r = m / c - 1 through the tabulated 1 / c,
// log x = e ln2 + log c + log1p(r), |r| <= 2^-8: series to r^7 (truncation 2^-59 relative to r); no term cancels another:
// [1, 2) has e = 0, [0.5, 1) reads log(c / 2) from the table, and elsewhere |e ln2| >= 2 |log c|
#ifdef __CUDACC__
__device__
#else
static
#endif
const unsigned long long cr_log_table[129][3] = {
#include "cr_log_table.inc"
};
ARN_CR_FN bool cr_log_kernel(double x, double& y) {                                // x > 0, normal (as f32)
    const uint64_t u = cr_bits(x);
    int e = (int)((u >> 52) & 0x7ffu) - 1023;
    const double m = cr_from_bits((u & 0x000fffffffffffffULL) | 0x3ff0000000000000ULL);
    int i = (int)(m * 128.0 + 0.5);                                              // 128 .. 256
    double mm = m;
    if (i == 256) { i = 128; e += 1; mm = m * 0.5; }                             // c = 2: fold into the exponent (exact)
    // x in [0.5, 1) (e = -1): the table's third column holds log(c / 2) rounded once, instead of -ln2 + log c cancelling
    const int col = e == -1 ? 2 : 1;
    if (e == -1) e = 0;
#ifdef __CUDACC__
    const double inv_c = __longlong_as_double((long long)__ldg(&cr_log_table[i - 128][0]));
    const double log_c = __longlong_as_double((long long)__ldg(&cr_log_table[i - 128][col]));
#else
    const double inv_c = cr_from_bits(cr_log_table[i - 128][0]), log_c = cr_from_bits(cr_log_table[i - 128][col]);
#endif
    const double r = cr_fma(mm, inv_c, -1.0);
    double p = 1.0 / 7.0;
    p = cr_fma(p, r, -1.0 / 6.0);
    p = cr_fma(p, r, 0.2);
    p = cr_fma(p, r, -0.25);
    p = cr_fma(p, r, 1.0 / 3.0);
    p = cr_fma(p, r, -0.5);
    p = cr_fma(p * r, r, r);                                                     // r + r^2 * (...)
    const double ed = (double)e;
    y = cr_fma(ed, 0.69314718055994529, log_c) + cr_fma(ed, 2.3190468138462996e-17, p);
    return true;
}
ARN_CR_FN bool cr_log_domain(float x) { return x >= 1.17549435e-38f && x <= 3.4028234e38f; }   // positive normal, finite; NaN fails
ARN_CR_FN float cr_logf_fast(float x) {
    if (cr_log_domain(x)) {
        if (x == 1.0f) return 0.0f;
        double y; cr_log_kernel((double)x, y);
        float f; if (cr_round_certain(y, f)) return f;
    }
    return cr_slow_log(x);
}

// ---- pow(a, b) = exp(b log a) for a > 0: the exponent's absolute error is |b log a| times log's relative error (< 2^-50),
// so the fast path is taken for |b log a| <= 8 only (relative error of the result < 2^-46)
ARN_CR_FN float cr_powf_fast(float a, float b) {
    if (cr_log_domain(a) && b >= -1.0e4f && b <= 1.0e4f && a != 1.0f && b != 0.0f) {
        double l; cr_log_kernel((double)a, l);
        const double t = (double)b * l;
        if (t >= -8.0 && t <= 8.0) { float f; if (cr_round_certain(cr_exp_kernel(t), f)) return f; }
    }
    return cr_slow_pow(a, b);
}

}  // namespace arn
