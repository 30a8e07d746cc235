// Device BVH build: a linear BVH over component bounds, emitted in the reference's own
// flattened layout — pre-order 32-byte nodes (`LinearNode`, component/bvh.rs:136-146,219-243) + the ordered
// component list — so that every traversal kernel, the 4-wide collapse and the oracle take it unchanged.
//
// It replaces the host `BVH::new` (bvh.rs:58-79,246-465) where start-up time matters more than reproducing the
// reference's topology: the tree is NOT the reference's SAH tree.  Hits are the same set of ray/primitive tests
// with the same arithmetic, so `t` and the winning primitive are identical except where two primitives tie in
// `t` (the first one met wins, bvh.rs:111) or where a transformed-sphere hit rewrites the traversal ray
// (bvh.rs:108-113) before a later test — both depend on visiting order.  tests/test_gpu_bvh.py states that rule.
//
// Pipeline (all on the context's stream):
//   1. centroid bounds of all components              k_lbvh_bounds   (block reduce + ordered-int atomics)
//   2. 63-bit Morton key per component centroid       k_lbvh_keys
//   3. radix sort (key, component)                    cub::DeviceRadixSort (library; the sort is not the product)
//   4. hierarchy from sorted keys (Karras 2012)       k_lbvh_hierarchy (one thread per interior node)
//   5. bottom-up bounds + subtree sizes               k_lbvh_refit     (one thread per leaf, second arrival climbs)
//   6. pre-order index of every node + emit           k_lbvh_emit      (walk to the root; offset = 1 + size(first child))
// Split axis stored in an interior node = axis of the highest Morton bit in which its two children differ; the
// first child holds the smaller coordinates on that axis, which is what `dir_is_neg[axis]` ordering assumes.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "../../../include/arn.h"

namespace arn {

struct LbvhBuf {
    const float* bounds;        // n * 6 (pmin, pmax), component order
    unsigned long long* keys;   // n
    uint32_t* vals;             // n: component index
    int* cbounds;               // 6 ordered ints: centroid min xyz, max xyz
    // hierarchy: node id = interior i in [0, n-1), leaf j -> (n-1) + j
    uint32_t* left; uint32_t* right;    // per interior node
    uint32_t* parent;                   // per node
    uint32_t* size;                     // per node: nodes in the subtree
    uint32_t* axis;                     // per interior node
    uint32_t* flag;                     // per interior node: arrivals
    float* bb;                          // per node: 6 floats
    uint32_t n;
};

__device__ __forceinline__ int float_ordered(float f) { int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }
__device__ __forceinline__ float ordered_float(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

__global__ void k_lbvh_init(LbvhBuf b) {
    if (threadIdx.x < 3) { b.cbounds[threadIdx.x] = 0x7fffffff; b.cbounds[3 + threadIdx.x] = (int)0x80000000; }
}

__global__ void __launch_bounds__(256) k_lbvh_bounds(LbvhBuf b) {
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < b.n; i += gridDim.x * blockDim.x) {
        const float* p = b.bounds + 6 * (size_t)i;
        for (int k = 0; k < 3; k++) {
            float c = 0.5f * p[k] + 0.5f * p[3 + k];          // BBox3::centroid-like; only used to order components
            lo[k] = fminf(lo[k], c); hi[k] = fmaxf(hi[k], c);
        }
    }
    for (int k = 0; k < 3; k++) {
        for (int off = 16; off > 0; off >>= 1) { lo[k] = fminf(lo[k], __shfl_down_sync(0xffffffffu, lo[k], off)); hi[k] = fmaxf(hi[k], __shfl_down_sync(0xffffffffu, hi[k], off)); }
        if ((threadIdx.x & 31) == 0) {
            if (lo[k] <= hi[k]) { atomicMin(&b.cbounds[k], float_ordered(lo[k])); atomicMax(&b.cbounds[3 + k], float_ordered(hi[k])); }
        }
    }
}

__device__ __forceinline__ unsigned long long expand21(uint32_t v) {
    unsigned long long x = v & 0x1fffffull;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}

__global__ void __launch_bounds__(256) k_lbvh_keys(LbvhBuf b) {
    float lo[3], inv[3];
    for (int k = 0; k < 3; k++) {
        lo[k] = ordered_float(b.cbounds[k]);
        float ext = ordered_float(b.cbounds[3 + k]) - lo[k];
        inv[k] = ext > 0.f ? 2097151.0f / ext : 0.f;          // 2^21 - 1 cells per axis
    }
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < b.n; i += gridDim.x * blockDim.x) {
        const float* p = b.bounds + 6 * (size_t)i;
        uint32_t q[3];
        for (int k = 0; k < 3; k++) {
            float c = 0.5f * p[k] + 0.5f * p[3 + k];
            float g = (c - lo[k]) * inv[k];
            q[k] = (uint32_t)fminf(fmaxf(g, 0.f), 2097151.0f);   // NaN centroids land in cell 0
        }
        b.keys[i] = (expand21(q[0]) << 2) | (expand21(q[1]) << 1) | expand21(q[2]);
        b.vals[i] = i;
    }
}

// common-prefix length of sorted keys i and j, index as the tie-breaker (Karras 2012, §4)
__device__ __forceinline__ int lbvh_delta(const unsigned long long* __restrict__ keys, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    unsigned long long a = keys[i], c = keys[j];
    if (a == c) return 64 + __clz((unsigned)i ^ (unsigned)j);
    return __clzll((long long)(a ^ c));
}

__global__ void __launch_bounds__(256) k_lbvh_hierarchy(LbvhBuf b, const unsigned long long* __restrict__ keys) {
    const int n = (int)b.n;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n - 1; i += gridDim.x * blockDim.x) {
        int d = (lbvh_delta(keys, n, i, i + 1) - lbvh_delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
        int dmin = lbvh_delta(keys, n, i, i - d);
        int lmax = 2;
        while (lbvh_delta(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
        int l = 0;
        for (int t = lmax >> 1; t >= 1; t >>= 1) if (lbvh_delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
        int j = i + l * d;
        int dnode = lbvh_delta(keys, n, i, j);
        int s = 0;
        for (int t = (l + 1) >> 1; ; t = (t + 1) >> 1) {
            if (lbvh_delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
            if (t == 1) break;
        }
        int gamma = i + s * d + (d < 0 ? d : 0);
        int lo = i < j ? i : j, hi = i < j ? j : i;
        uint32_t lc = (lo == gamma) ? (uint32_t)(n - 1 + gamma) : (uint32_t)gamma;
        uint32_t rc = (hi == gamma + 1) ? (uint32_t)(n - 1 + gamma + 1) : (uint32_t)(gamma + 1);
        b.left[i] = lc; b.right[i] = rc;
        b.parent[lc] = (uint32_t)i; b.parent[rc] = (uint32_t)i;
        // split axis: the children first differ in Morton bit 63 - dsplit; x owns bits = 2 (mod 3), y 1, z 0
        int dsplit = lbvh_delta(keys, n, gamma, gamma + 1);
        uint32_t ax = 0;
        if (dsplit < 64) { int bit = 63 - dsplit; ax = (bit % 3 == 2) ? 0u : (bit % 3 == 1 ? 1u : 2u); }
        b.axis[i] = ax;
        b.flag[i] = 0;
        if (i == 0) b.parent[0] = 0xffffffffu;
    }
}

__global__ void __launch_bounds__(256) k_lbvh_refit(LbvhBuf b) {
    const uint32_t n = b.n;
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
        uint32_t id = n - 1 + j;
        const float* p = b.bounds + 6 * (size_t)b.vals[j];
        float* o = b.bb + 6 * (size_t)id;
        for (int k = 0; k < 6; k++) o[k] = p[k];
        b.size[id] = 1;
        if (n == 1) { b.parent[id] = 0xffffffffu; return; }
        uint32_t cur = b.parent[id];
        for (;;) {
            __threadfence();
            if (atomicAdd(&b.flag[cur], 1u) == 0u) break;           // the first child to arrive stops here
            uint32_t l = b.left[cur], r = b.right[cur];
            const volatile float* a = b.bb + 6 * (size_t)l; const volatile float* c = b.bb + 6 * (size_t)r;
            float* q = b.bb + 6 * (size_t)cur;
            // BuildNode::to_interior (bvh.rs:210-219): union of the children's bounds
            for (int k = 0; k < 3; k++) { float x = a[k], y = c[k]; q[k] = x < y ? x : y; }
            for (int k = 3; k < 6; k++) { float x = a[k], y = c[k]; q[k] = x > y ? x : y; }
            b.size[cur] = 1u + ((volatile uint32_t*)b.size)[l] + ((volatile uint32_t*)b.size)[r];
            uint32_t up = b.parent[cur];
            if (up == 0xffffffffu) break;
            cur = up;
        }
    }
}

// pre-order position of node `id`: every step up adds 1 (first child) or 1 + size(first sibling) (second child)
__global__ void __launch_bounds__(256) k_lbvh_emit(LbvhBuf b, arn_node* __restrict__ out, uint32_t* __restrict__ order) {
    const uint32_t n = b.n, total = 2 * n - 1;
    for (uint32_t id = blockIdx.x * blockDim.x + threadIdx.x; id < total; id += gridDim.x * blockDim.x) {
        uint32_t pre = 0, cur = id;
        for (;;) {
            uint32_t p = b.parent[cur];
            if (p == 0xffffffffu) break;
            pre += (b.right[p] == cur) ? 1u + b.size[b.left[p]] : 1u;
            cur = p;
        }
        arn_node nd;
        const float* q = b.bb + 6 * (size_t)id;
        nd.bmin[0] = q[0]; nd.bmin[1] = q[1]; nd.bmin[2] = q[2]; nd.bmax[0] = q[3]; nd.bmax[1] = q[4]; nd.bmax[2] = q[5];
        if (id >= n - 1) {                       // leaf: one ordered slot
            uint32_t j = id - (n - 1);
            nd.offset = j; nd.len_axis = (1u << 2) | 3u;
            order[j] = b.vals[j];
        } else { nd.offset = 1u + b.size[b.left[id]]; nd.len_axis = b.axis[id]; }
        out[pre] = nd;
    }
}

}  // namespace arn
