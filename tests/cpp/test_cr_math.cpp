// Host sweep of kernels/cr_math.cuh (the same source the device compiles): for every f32 bit pattern the fast path either
// declines (falls back to the library) or returns exactly (float)f((double)x).  usage: test_cr_math [stride] [threads]
// stride 1 = all 2^32 arguments (about a minute per function on 8 cores); the test suite runs a strided sweep.
#include <cstdio>
#include <cstdlib>
#include <atomic>
#include <thread>
#include <vector>
#include <cmath>
#include <cstring>
#include "../../arendur_b200/csrc/kernels/cr_math.cuh"
using namespace arn;

static inline float f_of(uint32_t b) { float f; std::memcpy(&f, &b, 4); return f; }
static inline uint32_t b_of(float f) { uint32_t b; std::memcpy(&b, &f, 4); return b; }
static inline bool same(float a, float b) { return b_of(a) == b_of(b) || (a != a && b != b); }

int main(int argc, char** argv) {
    const uint64_t stride = argc > 1 ? std::strtoull(argv[1], nullptr, 10) : 1;
    const int nt = argc > 2 ? std::atoi(argv[2]) : (int)std::thread::hardware_concurrency();
    std::atomic<uint64_t> bad[5], fast[5], total[5];
    for (int i = 0; i < 5; i++) { bad[i] = 0; fast[i] = 0; total[i] = 0; }
    auto work = [&](int tid) {
        uint64_t lb[5] = {0, 0, 0, 0, 0}, lf[5] = {0, 0, 0, 0, 0}, lt[5] = {0, 0, 0, 0, 0};
        uint32_t rng = 0x9E3779B9u * (uint32_t)(tid + 1);
        for (uint64_t v = (uint64_t)tid * stride; v < (1ull << 32); v += stride * (uint64_t)nt) {
            const float x = f_of((uint32_t)v);
            // fast-path-only evaluation: replicate the wrappers' decisions without their fallbacks
            if (cr_trig_domain(x)) {
                double ds, dc; cr_sincos_kernel(x, ds, dc); float fs, fc;
                lt[0]++; lt[1]++;
                if (cr_round_certain(ds, fs)) { lf[0]++; if (!same(fs, (float)std::sin((double)x))) { if (lb[0]++ < 3) std::printf("sin mismatch x=%a fast=%a lib=%a\n", x, fs, (float)std::sin((double)x)); } }
                if (cr_round_certain(dc, fc)) { lf[1]++; if (!same(fc, (float)std::cos((double)x))) { if (lb[1]++ < 3) std::printf("cos mismatch x=%a fast=%a lib=%a\n", x, fc, (float)std::cos((double)x)); } }
            }
            if (x >= -80.f && x <= 80.f) {
                float f; lt[2]++;
                if (cr_round_certain(cr_exp_kernel((double)x), f)) { lf[2]++; if (!same(f, (float)std::exp((double)x))) { if (lb[2]++ < 3) std::printf("exp mismatch x=%a fast=%a lib=%a\n", x, f, (float)std::exp((double)x)); } }
            }
            if (cr_log_domain(x) && x != 1.0f) {
                double y; cr_log_kernel((double)x, y); float f; lt[3]++;
                if (cr_round_certain(y, f)) { lf[3]++; if (!same(f, (float)std::log((double)x))) { if (lb[3]++ < 3) std::printf("log mismatch x=%a fast=%a lib=%a\n", x, f, (float)std::log((double)x)); } }
                // pow: this base with a pseudo-random exponent in [-4, 4] (and the sampler's own range [0.4, 1.1])
                rng = rng * 1664525u + 1013904223u;
                const float b = (rng & 1u) ? -4.f + 8.f * (float)(rng >> 8) * (1.f / 16777216.f) : 0.4f + 0.7f * (float)(rng >> 8) * (1.f / 16777216.f);
                if (b != 0.f) {
                    const double t = (double)b * y;
                    lt[4]++;
                    if (t >= -8.0 && t <= 8.0) { float g; if (cr_round_certain(cr_exp_kernel(t), g)) { lf[4]++; if (!same(g, (float)std::pow((double)x, (double)b))) { if (lb[4]++ < 3) std::printf("pow mismatch a=%a b=%a fast=%a lib=%a\n", x, b, g, (float)std::pow((double)x, (double)b)); } } }
                }
            }
            // the public wrappers must agree with the library everywhere, fallbacks included (spot check on the same argument)
            if ((v & 0xfff) == 0) {
                float s, c; cr_sincosf_fast(x, s, c);
                if (!same(s, (float)std::sin((double)x)) || !same(c, (float)std::cos((double)x)) || !same(cr_expf_fast(x), (float)std::exp((double)x)) || !same(cr_logf_fast(x), (float)std::log((double)x))
                    || !same(cr_sinf_fast(x), (float)std::sin((double)x)) || !same(cr_cosf_fast(x), (float)std::cos((double)x))) { lb[0]++; std::printf("wrapper mismatch x=%a\n", x); }
            }
        }
        for (int i = 0; i < 5; i++) { bad[i] += lb[i]; fast[i] += lf[i]; total[i] += lt[i]; }
    };
    std::vector<std::thread> th;
    for (int t = 0; t < nt; t++) th.emplace_back(work, t);
    for (auto& t : th) t.join();
    const char* names[5] = {"sin", "cos", "exp", "log", "pow"};
    uint64_t nbad = 0;
    for (int i = 0; i < 5; i++) {
        std::printf("%s: %llu in-domain arguments, fast path decided %llu (%.6f %% fell back), mismatches %llu\n", names[i], (unsigned long long)total[i].load(),
                    (unsigned long long)fast[i].load(), 100.0 * (double)(total[i] - fast[i]) / (double)(total[i] ? total[i].load() : 1), (unsigned long long)bad[i].load());
        nbad += bad[i];
    }
    std::printf(nbad ? "FAILED\n" : "OK\n");
    return nbad ? 1 : 0;
}
