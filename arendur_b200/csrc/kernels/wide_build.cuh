// Order-preserving 4-wide collapse of the pre-order binary nodes, on the device (layout and reasoning: traverse.cuh,
// traverse4).  Wide node of interior node n = records (n.first's children | n.first itself if it is a leaf,
// n.second's likewise); a record keeps the child's bounds and its reference words.  Built level by level: level k
// reads the (binary node, wide index) pairs level k-1 queued, writes their four records and queues the interior
// grandchildren with freshly allocated wide indices — wide nodes end up in breadth-first order (the top of the
// tree is contiguous).  One launch per level, bounds read from device counters, no host round trip.
#pragma once
#include "traverse.cuh"

namespace arn {

__device__ __forceinline__ bool wb_is_leaf(const arn_node* __restrict__ nodes, uint32_t i) { return (nodes[i].len_axis >> 2) != 0; }
__device__ __forceinline__ uint32_t wb_axes(const arn_node* __restrict__ nodes, uint32_t i) {
    uint32_t a = i + 1, b = i + nodes[i].offset;
    return (nodes[i].len_axis & 3u) | ((wb_is_leaf(nodes, a) ? 0u : (nodes[a].len_axis & 3u)) << 2) | ((wb_is_leaf(nodes, b) ? 0u : (nodes[b].len_axis & 3u)) << 4);
}

// counts[level] pairs in `in`; appends to `out` / counts[level + 1]; *wide_count = wide nodes allocated so far
__global__ void __launch_bounds__(256) k_wide_level(const arn_node* __restrict__ nodes, arn_node* __restrict__ wide, const uint2* __restrict__ in,
                                                    uint2* __restrict__ out, uint32_t* __restrict__ counts, int level, uint32_t* __restrict__ wide_count) {
    const uint32_t n = counts[level];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint2 job = in[i];
        const uint32_t pair[2] = {job.x + 1, job.x + nodes[job.x].offset};
        arn_node rec[4];
#pragma unroll
        for (int k = 0; k < 4; k++) { rec[k].bmin[0] = rec[k].bmin[1] = rec[k].bmin[2] = ARN_INF; rec[k].bmax[0] = rec[k].bmax[1] = rec[k].bmax[2] = -ARN_INF; rec[k].offset = 0; rec[k].len_axis = ARN_W_EMPTY; }   // empty slot: inverted infinite bounds fail every slab test
#pragma unroll
        for (int g = 0; g < 2; g++) {
            const uint32_t ch = pair[g];
            const uint32_t cand[2] = {wb_is_leaf(nodes, ch) ? ch : ch + 1, wb_is_leaf(nodes, ch) ? 0xffffffffu : ch + nodes[ch].offset};
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const uint32_t c = cand[h];
                if (c == 0xffffffffu) continue;
                arn_node r = nodes[c];
                if ((r.len_axis >> 2) != 0) r.len_axis = ((r.len_axis >> 2) << 8) | ARN_W_LEAF;      // offset stays the first slot
                else {
                    const uint32_t w = atomicAdd(wide_count, 1u);
                    const uint32_t slot = atomicAdd(&counts[level + 1], 1u);
                    out[slot] = make_uint2(c, w);
                    r.len_axis = (wb_axes(nodes, c) << 2) | ARN_W_INNER; r.offset = w;
                }
                rec[2 * g + h] = r;
            }
        }
        float4* dst = reinterpret_cast<float4*>(wide + 4 * (size_t)job.y);
        const float4* src = reinterpret_cast<const float4*>(rec);
#pragma unroll
        for (int k = 0; k < 8; k++) dst[k] = src[k];
    }
}

__global__ void k_wide_begin(uint2* frontier, uint32_t* counts, uint32_t* wide_count, int root_is_interior) {
    for (int i = 0; i < ARN_STACK + 2; i++) counts[i] = 0;
    *wide_count = root_is_interior ? 1u : 0u;
    if (root_is_interior) { frontier[0] = make_uint2(0u, 0u); counts[0] = 1; }
}

// Ordered 48-byte primitive slots (layout: traverse.cuh): slot k = component order[k]; a triangle carries its three
// world-space vertices + the component id, a sphere slot only the id with the sphere bit.
__global__ void __launch_bounds__(256) k_build_slots(const uint32_t* __restrict__ order, const uint32_t* __restrict__ prims, const uint32_t* __restrict__ indices,
                                                     const float* __restrict__ positions, float4* __restrict__ slots, uint32_t n) {
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const uint32_t comp = order[k], ref = prims[comp];
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a, c = a;
        if (ref & ARN_PRIM_SPHERE) a.w = __uint_as_float(comp | ARN_PRIM_SPHERE);
        else {
            const float* p0 = positions + 3 * (size_t)indices[3 * (size_t)ref];
            const float* p1 = positions + 3 * (size_t)indices[3 * (size_t)ref + 1];
            const float* p2 = positions + 3 * (size_t)indices[3 * (size_t)ref + 2];
            a = make_float4(p0[0], p0[1], p0[2], __uint_as_float(comp));
            b = make_float4(p1[0], p1[1], p1[2], 0.f); c = make_float4(p2[0], p2[1], p2[2], 0.f);
        }
        slots[3 * (size_t)k] = a; slots[3 * (size_t)k + 1] = b; slots[3 * (size_t)k + 2] = c;
    }
}

}  // namespace arn
