// Compressed 8-wide collapse of the pre-order binary nodes, on the device (walk: traverse.cuh, traverse8).
//
// An 8-wide node = the binary node N with three levels removed: its up-to-eight descendants at depth 3 (a leaf met earlier keeps
// the first slot of its subtree).  Slot index = (b2 b1 b0), b = 0 for a first child — so the reference's visiting order for a ray
// (near child by the sign of d[split axis], component/bvh.rs:118-125) is a permutation of the slots that depends on the ray's sign
// octant only; the eight permutations are precomputed per node.  Child boxes are quantised to 8 bits per plane on a grid anchored
// at the node's own box, rounded OUTWARD: interior culling only has to be conservative (traverse.cuh, slab_cull), the reference's
// slab test runs on the leaf's exact bounds, which live in front of the leaf's primitives in the leaf blob.
//
//   node (128 B = one line; 8 x 16 B):
//     q0  origin x, y, z (f32: the node's bmin), meta = ex | ey << 8 | ez << 16 | leafmask << 24   (plane = origin + q * 2^(e - 127))
//     q1  visiting order for the ray sign octants 0..3, q5 for octants 4..7: nibble k of word o = slot visited k-th
//     q2  lo.x[8] | lo.y[8]      q3  lo.z[8] | hi.x[8]      q4  hi.y[8] | hi.z[8]      (one byte per slot; empty slot: lo = 255, hi = 0)
//     q6  child reference of slots 0..3, q7 of slots 4..7: 8-wide node index (interior) or float4 index into the leaf blob (leaf)
//   leaf blob (float4 units): [bmin.xyz, bmax.x] [bmax.yz, count bits, -] then `count` primitive records of 3 float4 (as `tris`).
#pragma once
#include "traverse.cuh"

namespace arn {

#define ARN_CW8_NONE 0xffffffffu

__global__ void __launch_bounds__(256) k_cw8_leaf_sizes(const arn_node* __restrict__ nodes, uint32_t n_nodes, uint32_t* __restrict__ sizes) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_nodes; i += gridDim.x * blockDim.x) {
        const uint32_t len = nodes[i].len_axis >> 2;
        sizes[i] = len ? 2u + 3u * len : 0u;
    }
}
__global__ void __launch_bounds__(256) k_cw8_leaf_blob(const arn_node* __restrict__ nodes, uint32_t n_nodes, const uint32_t* __restrict__ leaf_off,
                                                       const float4* __restrict__ tris, float4* __restrict__ blob) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_nodes; i += gridDim.x * blockDim.x) {
        const arn_node nd = nodes[i];
        const uint32_t len = nd.len_axis >> 2;
        if (!len) continue;
        float4* dst = blob + leaf_off[i];
        dst[0] = make_float4(nd.bmin[0], nd.bmin[1], nd.bmin[2], nd.bmax[0]);
        dst[1] = make_float4(nd.bmax[1], nd.bmax[2], __uint_as_float(len), 0.f);
        for (uint32_t k = 0; k < 3u * len; k++) dst[2 + k] = tris[3 * (size_t)nd.offset + k];
    }
}

__device__ __forceinline__ uint32_t cw8_axis(const arn_node* __restrict__ nodes, uint32_t i) { return nodes[i].len_axis & 3u; }
__device__ __forceinline__ bool cw8_leaf(const arn_node* __restrict__ nodes, uint32_t i) { return (nodes[i].len_axis >> 2) != 0; }

// One level of the breadth-first collapse.  `out_nodes` == nullptr: counting pass (indices are allocated, nothing is written).
__global__ void __launch_bounds__(256) k_cw8_level(const arn_node* __restrict__ nodes, const uint32_t* __restrict__ leaf_off, uint4* __restrict__ out_nodes,
                                                   const uint2* __restrict__ in, uint2* __restrict__ out, uint32_t* __restrict__ counts, int level, uint32_t* __restrict__ node_count) {
    const uint32_t n = counts[level];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint2 job = in[i];
        const uint32_t N = job.x;
        uint32_t slot[8];
#pragma unroll
        for (int k = 0; k < 8; k++) slot[k] = ARN_CW8_NONE;
        uint32_t ax_c[2] = {0u, 0u}, ax_g[4] = {0u, 0u, 0u, 0u};
        const uint32_t c[2] = {N + 1, N + nodes[N].offset};
#pragma unroll
        for (int b2 = 0; b2 < 2; b2++) {
            if (cw8_leaf(nodes, c[b2])) { slot[b2 * 4] = c[b2]; continue; }
            ax_c[b2] = cw8_axis(nodes, c[b2]);
            const uint32_t g[2] = {c[b2] + 1, c[b2] + nodes[c[b2]].offset};
#pragma unroll
            for (int b1 = 0; b1 < 2; b1++) {
                if (cw8_leaf(nodes, g[b1])) { slot[b2 * 4 + b1 * 2] = g[b1]; continue; }
                ax_g[b2 * 2 + b1] = cw8_axis(nodes, g[b1]);
                slot[b2 * 4 + b1 * 2] = g[b1] + 1; slot[b2 * 4 + b1 * 2 + 1] = g[b1] + nodes[g[b1]].offset;
            }
        }
        const arn_node P = nodes[N];
        uint32_t ex[3]; double scale[3];
#pragma unroll
        for (int a = 0; a < 3; a++) {
            const double ext = (double)P.bmax[a] - (double)P.bmin[a];
            int e = -126;
            if (ext > 0.0) { e = ilogb(ext / 255.0); if (ldexp(1.0, e) < ext / 255.0) e++; }
            e = e < -126 ? -126 : (e > 127 ? 127 : e);
            ex[a] = (uint32_t)(e + 127); scale[a] = ldexp(1.0, e);
        }
        uint32_t refs[8]; uint32_t leafmask = 0;
        unsigned long long plane[6] = {0ull, 0ull, 0ull, 0ull, 0ull, 0ull};        // lo x, y, z, hi x, y, z
#pragma unroll
        for (int k = 0; k < 8; k++) {
            uint32_t q[6] = {255u, 255u, 255u, 0u, 0u, 0u};
            refs[k] = 0u;
            if (slot[k] != ARN_CW8_NONE) {
                const arn_node ch = nodes[slot[k]];
#pragma unroll
                for (int a = 0; a < 3; a++) {
                    double lo = floor(((double)ch.bmin[a] - (double)P.bmin[a]) / scale[a]), hi = ceil(((double)ch.bmax[a] - (double)P.bmin[a]) / scale[a]);
                    lo = lo < 0.0 ? 0.0 : (lo > 255.0 ? 255.0 : lo); hi = hi < 0.0 ? 0.0 : (hi > 255.0 ? 255.0 : hi);
                    q[a] = (uint32_t)lo; q[3 + a] = (uint32_t)hi;
                }
                if ((ch.len_axis >> 2) != 0) { leafmask |= 1u << k; refs[k] = leaf_off[slot[k]]; }
                else {
                    const uint32_t w = atomicAdd(node_count, 1u);
                    const uint32_t s = atomicAdd(&counts[level + 1], 1u);
                    out[s] = make_uint2(slot[k], w);
                    refs[k] = w;
                }
            }
#pragma unroll
            for (int p = 0; p < 6; p++) plane[p] |= (unsigned long long)q[p] << (8 * k);
        }
        if (!out_nodes) continue;
        // visiting order per ray sign octant: slot = (q2 ^ s(N), q1 ^ s(child), q0 ^ s(grandchild)), s(x) = the octant's sign bit of x's split axis
        uint32_t perm[8];
        const uint32_t aN = P.len_axis & 3u;
#pragma unroll
        for (uint32_t o = 0; o < 8; o++) {
            uint32_t w = 0;
#pragma unroll
            for (uint32_t q = 0; q < 8; q++) {
                const uint32_t b2 = (q >> 2) ^ ((o >> aN) & 1u);
                const uint32_t b1 = ((q >> 1) & 1u) ^ ((o >> ax_c[b2]) & 1u);
                const uint32_t b0 = (q & 1u) ^ ((o >> ax_g[b2 * 2 + b1]) & 1u);
                w |= (b2 * 4u + b1 * 2u + b0) << (4u * q);
            }
            perm[o] = w;
        }
        uint4* dst = out_nodes + 8 * (size_t)job.y;
        dst[0] = make_uint4(__float_as_uint(P.bmin[0]), __float_as_uint(P.bmin[1]), __float_as_uint(P.bmin[2]), ex[0] | (ex[1] << 8) | (ex[2] << 16) | (leafmask << 24));
        dst[1] = make_uint4(perm[0], perm[1], perm[2], perm[3]);
        dst[2] = make_uint4((uint32_t)plane[0], (uint32_t)(plane[0] >> 32), (uint32_t)plane[1], (uint32_t)(plane[1] >> 32));
        dst[3] = make_uint4((uint32_t)plane[2], (uint32_t)(plane[2] >> 32), (uint32_t)plane[3], (uint32_t)(plane[3] >> 32));
        dst[4] = make_uint4((uint32_t)plane[4], (uint32_t)(plane[4] >> 32), (uint32_t)plane[5], (uint32_t)(plane[5] >> 32));
        dst[5] = make_uint4(perm[4], perm[5], perm[6], perm[7]);
        dst[6] = make_uint4(refs[0], refs[1], refs[2], refs[3]);
        dst[7] = make_uint4(refs[4], refs[5], refs[6], refs[7]);
    }
}

__global__ void k_cw8_begin(uint2* frontier, uint32_t* counts, uint32_t* node_count) {
    for (int i = 0; i < ARN_STACK + 2; i++) counts[i] = 0;
    *node_count = 1u; frontier[0] = make_uint2(0u, 0u); counts[0] = 1;
}

}  // namespace arn
