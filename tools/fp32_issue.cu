// FP32 issue-rate micro-benchmark (MEASURED_PEAKS.json has no FP32 figure).
// Measures sustained thread-instructions/s of (a) FMUL+FADD pairs — what a --fmad=false parity build
// issues — and (b) FFMA, on every SM with full occupancy and 8 independent chains per thread.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp32_issue tools/fp32_issue.cu
#include <cstdio>
#include <cuda_runtime.h>
template <bool FMA>
__global__ void __launch_bounds__(1024) k(float* out, int iters, float a, float b) {
    float x[8];
    for (int i = 0; i < 8; i++) x[i] = threadIdx.x * 1e-3f + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (FMA) { x[i] = __fmaf_rn(x[i], a, b); x[i] = __fmaf_rn(x[i], a, b); }
            else { x[i] = __fmul_rn(x[i], a); x[i] = __fadd_rn(x[i], b); }
        }
    }
    float s = 0.f;
    for (int i = 0; i < 8; i++) s += x[i];
    if (s == 123.456f) out[0] = s;
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    float* d; cudaMalloc(&d, 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 1 << 16, grid = p.multiProcessorCount * 2;
    for (int mode = 0; mode < 2; mode++) {
        float best = 1e30f;
        for (int rep = 0; rep < 5; rep++) {
            cudaEventRecord(e0);
            if (mode) k<true><<<grid, 1024>>>(d, iters, 0.999f, 1e-3f); else k<false><<<grid, 1024>>>(d, iters, 0.999f, 1e-3f);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        double inst = (double)grid * 1024 * iters * 16;
        printf("{\"kernel\": \"%s\", \"sms\": %d, \"thread_inst_per_s\": %.4e, \"ms\": %.3f}\n", mode ? "ffma" : "fmul+fadd", p.multiProcessorCount, inst / (best * 1e-3), best);
    }
    return 0;
}
