// Lane-refilling trace: the warp-cooperative form of the walk (north_star item 2).
//
// k_trace gives every lane ONE ray and lets the warp run until its longest ray is done: on incoherent rays 6-9 of 32 lanes
// are active in the interior-node loop (profiles/r02_*).  Here a warp owns a stream of rays instead:
//   k_ray_setup   (coherent, 32/32 lanes) every ray of the bounce — path, shadow and light rays, one dense index space — gets its
//                 traversal constants computed once (1/d, shear, the conservative-slab constants): an 80-byte record;
//   k_trace_refill persistent warps; every iteration the warp votes (two ballots) and runs ONE phase — an interior step or a
//                 leaf visit — for the lanes standing on that kind of node, whichever group is larger; when at least
//                 ARN_REFILL_MIN lanes are idle they take the next rays of the stream (ONE atomic per refill, five 128-bit
//                 loads per lane).  No per-ray set-up, no queue bookkeeping inside the walk;
//   k_classify    (coherent) hit records -> path state, shading-class queues (ballot + shared-memory staging), occlusion /
//                 light-hit flags.
// The per-ray arithmetic is that of traverse2 / leaf_prims / slab (traverse.cuh): same leaves, same order, same bits.
#pragma once
#include "wavefront.cuh"

namespace arn {

#ifndef ARN_REFILL_MIN
#define ARN_REFILL_MIN 8
#endif

struct TraceBuf {                 // capacity: 3 rays per path slot
    float4* rs;                   // 5 float4 per ray, SoA planes of `cap` entries: (o, tmax) (1/d, bits) (shear, d.z) (oi0, d.x) (oi1, d.y)
    float4* res;                  // per ray: (component index bits, a, b, c)
    uint32_t cap;
};
#define ARN_RS_ANY 4u             /* bits: kz (0..2) | any << 2 | regular << 3 | negbits << 4 */
#define ARN_RS_REGULAR 8u
#define ARN_CNT_CURSOR 20         /* q.counts[20]: next ray of the stream */

// dense ray index -> (kind, path id, record pointers)
struct RayRef { uint32_t kind, pid; };
ARN_DEV RayRef ray_ref(const PathBuf& pb, const Queues& q, uint32_t par, uint32_t gi, uint32_t n_ext, uint32_t n_sh) {
    RayRef rr;
    if (gi < n_ext) { rr.kind = 0u; rr.pid = __ldcs(&q.active[par][gi]); }
    else if (gi < n_ext + n_sh) { rr.kind = 1u; rr.pid = __ldcs(&q.shadow[gi - n_ext]); }
    else { rr.kind = 2u; rr.pid = __ldcs(&q.mis[gi - n_ext - n_sh]); }
    return rr;
}

__global__ void __launch_bounds__(ARN_BLOCK) k_ray_setup(const __grid_constant__ DevScene sc, PathBuf pb, Queues q, TraceBuf tb, int j) {
    const uint32_t par = (uint32_t)j & 1u;
    const uint32_t n_ext = *cnt_active(q.counts, par), n_sh = *cnt_nee(q.counts, par, 1), n_mis = *cnt_nee(q.counts, par, 2);
    const uint32_t n = n_ext + n_sh + n_mis;
    if (blockIdx.x == 0 && threadIdx.x == 0) q.counts[ARN_CNT_CURSOR] = 0u;
    for (uint32_t gi = blockIdx.x * blockDim.x + threadIdx.x; gi < n; gi += gridDim.x * blockDim.x) {
        const RayRef rr = ray_ref(pb, q, par, gi, n_ext, n_sh);
        const float4* __restrict__ rk = rr.kind == 0u ? pb.ray : (rr.kind == 1u ? pb.sh : pb.mis);
        const float4 o = __ldcs(&rk[2 * rr.pid]), d = __ldcs(&rk[2 * rr.pid + 1]);
        TravRay r; trav_init(r, f3(o.x, o.y, o.z), f3(d.x, d.y, d.z), rr.kind == 1u ? o.w : ARN_INF);
        CullRay c; cull_setup(sc, r, c);
        const uint32_t bits = (uint32_t)r.kz | (rr.kind == 1u ? ARN_RS_ANY : 0u) | (ray_is_regular(sc, r) ? ARN_RS_REGULAR : 0u) | (c.negbits << 4);
        const size_t cap = tb.cap;
        __stcs(&tb.rs[gi], make_float4(r.o.x, r.o.y, r.o.z, r.tmax));
        __stcs(&tb.rs[cap + gi], make_float4(r.inv.x, r.inv.y, r.inv.z, __uint_as_float(bits)));
        __stcs(&tb.rs[2 * cap + gi], make_float4(r.shear.x, r.shear.y, r.shear.z, r.d.z));
        __stcs(&tb.rs[3 * cap + gi], make_float4(c.oi0.x, c.oi0.y, c.oi0.z, r.d.x));
        __stcs(&tb.rs[4 * cap + gi], make_float4(c.oi1.x, c.oi1.y, c.oi1.z, r.d.y));
    }
}

// One lane's traversal state between refills.  `node` = the record the lane stands on: interior (to expand) or leaf.
__global__ void __launch_bounds__(ARN_BLOCK, ARN_TRAV_MINB) k_trace_refill(const __grid_constant__ DevScene sc, Queues q, TraceBuf tb, int j) {
    const uint32_t par = (uint32_t)j & 1u;
    const uint32_t n = *cnt_active(q.counts, par) + *cnt_nee(q.counts, par, 1) + *cnt_nee(q.counts, par, 2);
    uint32_t* cursor = q.counts + ARN_CNT_CURSOR;
    const unsigned lane = threadIdx.x & 31u;
    const size_t cap = tb.cap;
    bool busy = false, exhausted = false, any = false;
    uint32_t slot = 0;                                // ray of the stream this lane works on
    TravRay r; CullRay c; HitRec h;
    uint2 stack[ARN_STACK];
    int sp = 0;
    uint32_t idx = 0, offset = 0, len_axis = 4u;
    r.tmax = 0.f; r.kz = 2; r.o = r.co = r.d = r.inv = r.shear = f3(0.f, 0.f, 0.f); c.oi0 = c.oi1 = r.o; c.negbits = 0; h.prim = -1; h.a = h.b = h.c = 0.f;
    unsigned exhausted_mask = 0u;                     // warp-uniform copy of the lanes' `exhausted` flags
    for (;;) {
        // Every iteration the warp runs ONE phase — an interior step or a leaf visit — for the lanes that stand on that kind of
        // node, whichever group is larger: no lane waits for the longest descent of the warp (the while-while form's loss).
        const bool is_int = busy && (len_axis >> 2) == 0;
        const unsigned m_busy = __ballot_sync(0xffffffffu, busy), m_int = __ballot_sync(0xffffffffu, is_int);
        const unsigned m_idle = ~m_busy & ~exhausted_mask;
        const int n_int = __popc(m_int), n_leaf = __popc(m_busy) - n_int;
        // ---- refill: when enough lanes are idle (or nobody works), the idle lanes take the next rays of the stream
        if (m_idle && (__popc(m_idle) >= ARN_REFILL_MIN || m_busy == 0u)) {
            const int leader = __ffs(m_idle) - 1;
            uint32_t base = 0;
            if ((int)lane == leader) base = atomicAdd(cursor, (uint32_t)__popc(m_idle));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (!busy && !exhausted) {
                const uint32_t my = base + (uint32_t)__popc(m_idle & ((1u << lane) - 1u));
                if (my >= n) exhausted = true;
                else {
                    const float4 a0 = __ldcs(&tb.rs[my]), a1 = __ldcs(&tb.rs[cap + my]), a2 = __ldcs(&tb.rs[2 * cap + my]);
                    const float4 a3 = __ldcs(&tb.rs[3 * cap + my]), a4 = __ldcs(&tb.rs[4 * cap + my]);
                    const uint32_t bits = __float_as_uint(a1.w);
                    r.o = f3(a0.x, a0.y, a0.z); r.co = r.o; r.tmax = a0.w;
                    r.inv = f3(a1.x, a1.y, a1.z); r.kz = (int)(bits & 3u); any = (bits & ARN_RS_ANY) != 0u;
                    r.shear = f3(a2.x, a2.y, a2.z); r.d = f3(a3.w, a4.w, a2.w);
                    c.oi0 = f3(a3.x, a3.y, a3.z); c.oi1 = f3(a4.x, a4.y, a4.z); c.negbits = bits >> 4;
                    h.prim = -1; h.a = h.b = h.c = 0.f;
                    sp = 0; slot = my;
                    if (!(bits & ARN_RS_REGULAR)) {                 // axis-parallel / degenerate direction: the exact walk, at once
                        if (any) traverse_exact_any(sc, r, h); else traverse_exact_closest(sc, r, h);
                    } else {
                        const Node8 nd = ld_node(sc.nodes);
                        float lo;
                        if (slab_cull(nd.q0, nd.q1, r, c, lo)) { idx = 0; offset = __float_as_uint(nd.q1.z); len_axis = __float_as_uint(nd.q1.w); busy = true; }
                    }
                    if (!busy) { __stcs(&tb.res[slot], make_float4(__int_as_float(h.prim), h.a, h.b, h.c)); __stcs(&tb.rs[slot], make_float4(r.d.x, r.d.y, r.d.z, 0.f)); }      // ended at the root
                }
            }
            exhausted_mask = __ballot_sync(0xffffffffu, exhausted);
            continue;
        }
        if (m_busy == 0u) break;                          // nobody works and nobody can fetch: the stream is done
        if (n_int >= n_leaf) {
            // ---- interior step: cull both children (first child = idx + 1, second = idx + offset) conservatively
            if (is_int) {
                const uint32_t ia = idx + 1, ib = idx + offset;
                const Node8 a = ld_node(sc.nodes + 2 * ia), b = ld_node(sc.nodes + 2 * ib);
                float la, lb;
                const bool ha = slab_cull(a.q0, a.q1, r, c, la), hb = slab_cull(b.q0, b.q1, r, c, lb);
                const bool first_b = (c.negbits >> (len_axis & 3u)) & 1u;
                if (ha && hb) {
                    stack[sp++] = first_b ? make_uint2(ia, __float_as_uint(la)) : make_uint2(ib, __float_as_uint(lb));
                    idx = first_b ? ib : ia;
                    offset = __float_as_uint(first_b ? b.q1.z : a.q1.z); len_axis = __float_as_uint(first_b ? b.q1.w : a.q1.w);
                } else if (ha || hb) {
                    idx = ha ? ia : ib;
                    offset = __float_as_uint(ha ? a.q1.z : b.q1.z); len_axis = __float_as_uint(ha ? a.q1.w : b.q1.w);
                } else if (!trav_pop(sc, r, stack, sp, idx, offset, len_axis)) {
                    busy = false;
                    __stcs(&tb.res[slot], make_float4(__int_as_float(h.prim), h.a, h.b, h.c));
                    __stcs(&tb.rs[slot], make_float4(r.d.x, r.d.y, r.d.z, 0.f));       // the direction the ray leaves the traversal with (plane 0 of its set-up record is consumed)
                }
            }
        } else {
            // ---- leaf visit: the reference's slab test on the leaf, its primitives, then the next node
            if (busy && !is_int) {
                const Node8 nd = ld_node(sc.nodes + 2 * idx);
                float t0;
                bool done = false;
                if (slab(nd.q0, nd.q1, r, t0) && t0 < r.tmax) done = leaf_prims(sc, offset, len_axis >> 2, r, h, any);
                if (done || !trav_pop(sc, r, stack, sp, idx, offset, len_axis)) {
                    busy = false;
                    __stcs(&tb.res[slot], make_float4(__int_as_float(h.prim), h.a, h.b, h.c));
                    __stcs(&tb.rs[slot], make_float4(r.d.x, r.d.y, r.d.z, 0.f));       // the direction the ray leaves the traversal with (plane 0 of its set-up record is consumed)
                }
            }
        }
        __syncwarp();
    }
}

// hit records -> path state, class queues, occlusion / light flags (replaces the tail of k_trace)
__global__ void __launch_bounds__(ARN_BLOCK) k_classify(const __grid_constant__ DevScene sc, PathBuf pb, Queues q, TraceBuf tb, int j) {
    const uint32_t par = (uint32_t)j & 1u;
    const uint32_t n_ext = *cnt_active(q.counts, par), n_sh = *cnt_nee(q.counts, par, 1), n_mis = *cnt_nee(q.counts, par, 2);
    const uint32_t n = n_ext + n_sh + n_mis, n_round = (n + 31u) & ~31u;
    __shared__ uint32_t stage_rows[ARN_NCLS][ARN_BLOCK / 32][64];
    WarpStage st[ARN_NCLS];
#pragma unroll
    for (int k = 0; k < ARN_NCLS; k++) { st[k].row = stage_rows[k][threadIdx.x >> 5]; st[k].fill = 0; }
    if (blockIdx.x == 0 && threadIdx.x < 9) {          // empty the idle counter set (see cnt_* in wavefront.cuh), as k_trace does
        uint32_t* z = threadIdx.x == 0 ? cnt_active(q.counts, par ^ 1u) : (threadIdx.x < 6 ? cnt_cls(q.counts, par ^ 1u, threadIdx.x - 1) : cnt_nee(q.counts, par ^ 1u, threadIdx.x - 6));
        *z = 0u;
    }
    for (uint32_t gi = blockIdx.x * blockDim.x + threadIdx.x; gi < n_round; gi += gridDim.x * blockDim.x) {
        int cls = -1; uint32_t pid = 0;
        if (gi < n) {
            const RayRef rr = ray_ref(pb, q, par, gi, n_ext, n_sh);
            pid = rr.pid;
            const float4 res = __ldcs(&tb.res[gi]);
            const int prim = __float_as_int(res.x);
            if (rr.kind == 0u) {
                __stcs(&pb.hit[pid], res);
                if (prim >= 0) {
                    uint32_t ref = sc.prims[prim], mat;
                    if (ref & ARN_PRIM_SPHERE) mat = sc.spheres[ref & ~ARN_PRIM_SPHERE].material;
                    else mat = sc.meshes[sc.tri_mesh[ref]].material;
                    {   // `*ray = iray`: accepted hits on transformed spheres leave the ray round-tripped (sphere_slot), see k_trace
                        const float4 d4 = pb.ray[2 * pid + 1], nd = __ldcs(&tb.rs[gi]);
                        if (nd.x != d4.x || nd.y != d4.y || nd.z != d4.z) pb.ray[2 * pid + 1] = make_float4(nd.x, nd.y, nd.z, d4.w);
                    }
                    cls = shading_class(sc.materials[mat]);
                }
            } else if (rr.kind == 1u) {
                __stcs(&pb.occluded[pid], prim >= 0 ? 1u : 0u);
            } else {
                const float4 d4 = __ldcs(&pb.mis[2 * pid + 1]);
                const float3 wi = f3(d4.x, d4.y, d4.z);
                const uint32_t lcomp = __float_as_uint(d4.w);
                uint32_t okl = 0;
                if (prim >= 0 && (uint32_t)prim == lcomp) {                // ptr::eq(light, hit.as_light()) (scene.rs:149)
                    const DevSphere& sp = sc.spheres[sc.prims[lcomp] & ~ARN_PRIM_SPHERE];
                    float3 pos = f3(res.y, res.z, res.w);
                    if (sp.has_transform) pos = xform_point(sp.local_parent, pos);
                    okl = is_black(light_le(sp, pos, -wi)) ? 0u : 1u;       // lsi.le(-wi)
                }
                __stcs(&pb.mis_ok[pid], okl);
            }
        }
#pragma unroll
        for (int k = 0; k < ARN_NCLS; k++) stage_push(st[k], cls == k, pid, q.cls[k], cnt_cls(q.counts, par, k));
    }
#pragma unroll
    for (int k = 0; k < ARN_NCLS; k++) stage_flush(st[k], q.cls[k], cnt_cls(q.counts, par, k));
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        atomicAdd(&q.stats[0], (unsigned long long)n_ext);
        atomicAdd(&q.stats[1], (unsigned long long)n_sh);
        atomicAdd(&q.stats[2], (unsigned long long)n_mis);
        if (j != 0) atomicAdd(&q.stats[4], (unsigned long long)n);
    }
}

}  // namespace arn
