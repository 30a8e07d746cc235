// Wavefront path-tracing kernels for sm_100a (K1..K7 of SURVEY.md §2.1).
//
// One "wave" = up to W camera samples processed bounce by bounce:
//   k_generate -> k_trace -> [ k_shade -> k_trace -> k_resolve ] x max_depth -> k_accumulate
// k_trace handles, in ONE launch, the path rays of the next bounce plus the shadow rays and the
// BSDF-sampled light rays of the current one (three warp-uniform segments of one ray queue).
// Stages are separate kernels linked by compacted queues of path ids; the queues are built
// in k_shade with warp ballots + shared-memory staging (one global atomic per block
// iteration).  All launches are grid-stride over DEVICE-side counters, so a whole wave is
// enqueued without a host round trip.
//
// Replaces: PTRenderer::render's per-sample loop and calculate_lighting
// (src/renderer/pt.rs:55-176), Scene::uniform_sample_one_light / evaluate_direct
// (src/renderer/scene.rs:58-167), Sampler::get_camera_sample (src/sample/mod.rs:38-43),
// PerspecCam::generate_path_differential (src/filming/perspective.rs:292-320),
// FilmTile::add_sample (src/filming/film.rs:297-319).
#pragma once
#include "shade.cuh"
#include "shade_tex.cuh"

namespace arn {

#ifndef ARN_BLOCK
#define ARN_BLOCK 256
#endif
#ifndef ARN_TRAV_MINB
#define ARN_TRAV_MINB 3
#endif
#ifndef ARN_SHADE_MINB
#define ARN_SHADE_MINB 2
#endif
#ifndef ARN_SHADE_MINB_DIFFUSE
#define ARN_SHADE_MINB_DIFFUSE 3
#endif

// Path state of a wave, SoA over path slots (capacity W).  Streams are laid out by WRITER and in 32-byte records where one
// stage reads or writes two float4 of the same path, so that a scattered path id costs one DRAM sector per record
// instead of one per float4 / per 4-byte word (the matte shade instance is bound by exactly this traffic):
struct PathBuf {
    float4* ray;                 // 2 per slot: (origin xyz, sampler key) (direction xyz, st); st = bounces | specular << 8 | n1d << 16 | n2d << 24
    float4* beta;                // throughput rgb
    float4* L;                   // radiance rgb
    float2* pfilm;               // film position of the camera sample
    uint32_t* pix;               // x | y << 16          (film stages only: the shade stages carry the sampler key instead)
    uint32_t* smp;               // sample index         (diagnostics only)
    float4* hit;                 // (component index bits, -1 = miss; a, b, c): triangle b0,b1,b2 / sphere refined local hit point
    // next-event-estimation state between k_shade, k_trace and k_resolve
    float4* sh;                  // 2 per slot: shadow ray (origin xyz, tmax) (direction xyz, -)
    float4* mis;                 // 2 per slot: BSDF-sampled light ray (origin xyz, -) (direction xyz = wi, light component id bits)
    float4* nee;                 // 4 per slot: (light-sampling term rgb, light choice pdf) (BSDF-sampling term rgb, flags bits)
                                 //             (throughput before this bounce's BSDF sample, -) (unused)
    uint32_t* occluded;          // 1 = the shadow ray was blocked (written by k_trace)
    uint32_t* mis_ok;            // 1 = the BSDF-sampled ray reached the chosen light and saw its emission
    float4* diff;                // textured scenes only, 4 per slot: the ray differential's offset rays (rx origin)(rx dir)(ry origin)(ry dir)
};
#define NEE_DONE 1u
#define NEE_SHADOW 2u
#define NEE_MIS 4u

#define ARN_NCLS 5                // shading classes: Lambert, Oren-Nayar, Plastic, Glass, Translucent
struct Queues {
    uint32_t* active[2];         // path ids for the current / next bounce
    uint32_t* connect;           // path ids with a pending direct-light term
    uint32_t* shadow;            // path ids with a shadow ray to trace
    uint32_t* mis;               // path ids with a BSDF-sampled light ray to trace
    uint32_t* cls[ARN_NCLS];     // hits of the current bounce, sorted by shading class (material sort)
    uint32_t* counts;            // queue sizes, double-buffered by the parity of the trace launch that consumes / fills them (cnt_* below)
    unsigned long long* stats;   // [0] extend rays [1] shadow rays [2] mis rays [3] invalid samples [4] extend rays of bounces>=1 [5..7] nodes/tris/spheres tested by extend (COUNT builds)
};

// Queue counters.  Trace launch j of a wave (j = 0: camera rays; j = b + 1: after shade(b)) has parity p = j & 1:
//   active(p)   path rays trace(j) follows      (filled by k_generate for j = 0, by shade(j - 1) otherwise)
//   cls(p, c)   hits of trace(j) by shading class (filled by trace(j), consumed by shade(j))
//   nee(p, k)   k = 0 connect, 1 shadow rays, 2 light rays of bounce j - 1 (filled by shade(j - 1); trace(j) reads 1 and 2,
//               resolve(j - 1) reads 0)
// The set with the other parity is idle while trace(j) runs — its last reader finished before trace(j) was launched and its
// next writer starts after it — so one thread of trace(j) zeroes it: no separate reset launches between the stages.
#define ARN_NCOUNTS 32
ARN_DEV uint32_t* cnt_active(const uint32_t* counts, uint32_t p) { return const_cast<uint32_t*>(counts) + p; }
ARN_DEV uint32_t* cnt_cls(const uint32_t* counts, uint32_t p, uint32_t c) { return const_cast<uint32_t*>(counts) + 2u + 5u * p + c; }
ARN_DEV uint32_t* cnt_nee(const uint32_t* counts, uint32_t p, uint32_t k) { return const_cast<uint32_t*>(counts) + 12u + 3u * p + k; }

struct WaveParams {
    float raster_view[16], view_parent[16];
    uint32_t has_lens; float lens_radius, focal_distance;
    uint32_t ortho;                      // OrthoCam (filming/ortho.rs)
    uint32_t filt_kind; float filt_a, filt_b;   // film filter (see filter1)
    int crop_x0, crop_y0, crop_w, crop_h;
    float fr_x, fr_y;
    int tile_dx, tile_dy, tile_lastx, tile_lasty, tiles_nx, tiles_ny, tile_rx, tile_ry;   // spawn_tiles grid: tile size, size of the last tile, count, sink growth (film.rs:104-135)
    uint32_t seed;
    uint32_t max_depth, min_depth; float rr_threshold;
    uint32_t n_tiles;
    const int4* tile_rect;                  // x0, y0, w, h of each of this rank's tiles
    const unsigned long long* tile_prefix;  // pixels before tile i (n_tiles + 1 entries)
    uint32_t spp_begin, spp_count;
    uint32_t textured, spp_total;            // image textures present: carry ray differentials, scaled by 1 / spp_total (pt.rs:141-142)
    uint32_t strat_ndim, sampledx, sampledy; // ARN_SAMPLER_STRATIFIED: the first strat_ndim 1-D / 2-D draws of a sample are stratified (0 = parity sampler)
};
// the draw `u` the sampler just returned was its (n - 1)-th of this kind: stratify it if the sampler is in that mode (warp-uniform test)
#define ARN_STRAT_1D(u, n) do { if (p.strat_ndim && (n) - 1u < p.strat_ndim) (u) = strat_1d(p.seed, p.sampledx, p.sampledy, spix, ssmp, (n) - 1u, (u)); } while (0)
#define ARN_STRAT_2D(u, n) do { if (p.strat_ndim && (n) - 1u < p.strat_ndim) (u) = strat_2d(p.seed, p.sampledx, p.sampledy, spix, ssmp, (n) - 1u, (u)); } while (0)

// ---- queue append: warp ballots + per-warp shared-memory staging -------------------
// Each warp compacts the ids it keeps into its own 64-entry staging row in shared memory
// (ballot + popc give the slots); whenever a row holds >= 32 ids the warp reserves 32 queue
// slots with ONE atomic and writes them as one aligned 128-byte store.  No block barrier:
// warps in a shading kernel finish at very different times.
struct WarpStage {
    uint32_t* row;       // 64 entries of this warp
    uint32_t fill;       // warp-uniform
};
ARN_DEV void stage_push(WarpStage& st, bool keep, uint32_t pid, uint32_t* __restrict__ queue, uint32_t* __restrict__ count) {
    const unsigned lane = threadIdx.x & 31u;
    unsigned ballot = __ballot_sync(0xffffffffu, keep);
    if (keep) st.row[st.fill + __popc(ballot & ((1u << lane) - 1u))] = pid;
    st.fill += __popc(ballot);
    __syncwarp();
    if (st.fill >= 32u) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(count, 32u);
        base = __shfl_sync(0xffffffffu, base, 0);
        queue[base + lane] = st.row[lane];
        __syncwarp();
        uint32_t rest = st.fill - 32u;
        uint32_t v = lane < rest ? st.row[32u + lane] : 0u;
        __syncwarp();
        if (lane < rest) st.row[lane] = v;
        st.fill = rest;
        __syncwarp();
    }
}
ARN_DEV void stage_flush(WarpStage& st, uint32_t* __restrict__ queue, uint32_t* __restrict__ count) {
    const unsigned lane = threadIdx.x & 31u;
    if (st.fill) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(count, st.fill);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (lane < st.fill) queue[base + lane] = st.row[lane];
        st.fill = 0;
    }
}

// Programmatic dependent launch (ARN_OPT_PDL, off by default): with the launch attribute set, the blocks of the NEXT kernel of a pipeline
// may become resident as this kernel's blocks retire and park at `wait` until this grid has completed and flushed; without the attribute
// both instructions do nothing.
#ifndef ARN_PDL_EARLY_TRIGGER
#define ARN_PDL_EARLY_TRIGGER 1        /* 0: no explicit trigger — the next kernel is released when this one's blocks exit (measured: no better) */
#endif
ARN_DEV void pdl_prologue() {
#if ARN_PDL_EARLY_TRIGGER
    asm volatile("griddepcontrol.launch_dependents;");
#endif
    asm volatile("griddepcontrol.wait;" ::: "memory");
}

// ---- K1 generate ------------------------------------------------------------------------
__global__ void __launch_bounds__(ARN_BLOCK) k_generate(const __grid_constant__ WaveParams p, PathBuf pb, Queues q,
                                                         unsigned long long wave_base, uint32_t n) {
    pdl_prologue();
    if (blockIdx.x == 0 && threadIdx.x < ARN_NCOUNTS) q.counts[threadIdx.x] = threadIdx.x == 0 ? n : 0u;     // begin the wave: every queue empty, n camera rays
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        unsigned long long g = wave_base + i;
        unsigned long long pl = g / p.spp_count;
        uint32_t s = p.spp_begin + (uint32_t)(g % p.spp_count);
        // tile lookup: largest t with prefix[t] <= pl
        uint32_t lo = 0, hi = p.n_tiles;
        while (hi - lo > 1) { uint32_t mid = (lo + hi) >> 1; if (p.tile_prefix[mid] <= pl) lo = mid; else hi = mid; }
        int4 tr = p.tile_rect[lo];
        uint32_t off = (uint32_t)(pl - p.tile_prefix[lo]);
        uint32_t px = (uint32_t)tr.x + off % (uint32_t)tr.z, py = (uint32_t)tr.y + off / (uint32_t)tr.z;   // row-major in the tile
        Sampler sm; sm.init(p.seed, px, py, s, 0, 0);
        const uint32_t spix = px | (py << 16), ssmp = s;
        float2 j = sm.next_2d(); ARN_STRAT_2D(j, sm.i2d);
        float2 pfilm = f2(j.x + (float)px, j.y + (float)py);
        float2 plens = sm.next_2d(); ARN_STRAT_2D(plens, sm.i2d);
        // PerspecCam::generate_path_differential, main ray
        float3 pview = xform_point(p.raster_view, f3(pfilm.x, pfilm.y, 0.f));
        float3 o = f3(0.f, 0.f, 0.f), d = f3(0.f, 0.f, 1.f);
        if (p.ortho) o = pview; else d = normalize(pview);       // OrthoCam::generate_path (ortho.rs:180-198)
        if (p.has_lens) {
            float2 dl = sample_concentric_disk(plens);
            float2 pln = f2(p.lens_radius * dl.x, p.lens_radius * dl.y);
            float ft = p.focal_distance / d.z;
            float3 pfocus = o + d * ft;
            o = f3(pln.x, pln.y, 0.f);
            d = normalize(pfocus - o);
        }
        if (p.textured) {       // generate_path_differential's offset rays (perspective.rs:312-319, ortho.rs:220-227), scale_differentials (ray.rs:282-291)
            float3 rxo, rxd, ryo, ryd;
            if (p.ortho) {
                float3 dx = xform_vector(p.raster_view, f3(1.f, 0.f, 0.f)), dy = xform_vector(p.raster_view, f3(0.f, 1.f, 0.f));
                rxo = o + dx; ryo = o + dy; rxd = d; ryd = d;
            } else {
                float3 or2v = xform_point(p.raster_view, f3(1.f, 0.f, 0.f));                   // sic: dx = 0 (perspective.rs:68-73)
                float3 dx = xform_point(p.raster_view, f3(1.f, 0.f, 0.f)) - or2v, dy = xform_point(p.raster_view, f3(0.f, 1.f, 0.f)) - or2v;
                rxo = o; ryo = o; rxd = normalize(pview + dx); ryd = normalize(pview + dy);
            }
            float3 wo = xform_point(p.view_parent, o), wd = xform_vector(p.view_parent, d);
            rxo = xform_point(p.view_parent, rxo); rxd = xform_vector(p.view_parent, rxd);
            ryo = xform_point(p.view_parent, ryo); ryd = xform_vector(p.view_parent, ryd);
            const float sc = 1.f / (float)p.spp_total;
            rxo = wo + (rxo - wo) * sc; ryo = wo + (ryo - wo) * sc; rxd = wd + (rxd - wd) * sc; ryd = wd + (ryd - wd) * sc;
            pb.diff[4 * i] = make_float4(rxo.x, rxo.y, rxo.z, 0.f); pb.diff[4 * i + 1] = make_float4(rxd.x, rxd.y, rxd.z, 0.f);
            pb.diff[4 * i + 2] = make_float4(ryo.x, ryo.y, ryo.z, 0.f); pb.diff[4 * i + 3] = make_float4(ryd.x, ryd.y, ryd.z, 0.f);
        }
        o = xform_point(p.view_parent, o); d = xform_vector(p.view_parent, d);
        pb.ray[2 * i] = make_float4(o.x, o.y, o.z, __uint_as_float(sm.key));
        pb.ray[2 * i + 1] = make_float4(d.x, d.y, d.z, __uint_as_float((0u) | (0u << 8) | (0u << 16) | (2u << 24)));
        pb.beta[i] = make_float4(1.f, 1.f, 1.f, 0.f);
        pb.L[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        pb.pfilm[i] = pfilm;
        pb.pix[i] = px | (py << 16);
        pb.smp[i] = s;
        q.active[0][i] = i;
    }
}

// ---- K2 extend: closest hit for every active path, hits sorted into per-material-class queues ----
ARN_DEV int shading_class(const arn_material& m) {
    switch (m.type) {
    case ARN_MAT_MATTE: return clampf(m.sigma, 0.f, 90.f) == 0.f ? 0 : 1;
    case ARN_MAT_PLASTIC: return 2;
    case ARN_MAT_GLASS: return 3;
    default: return 4;
    }
}
// K2/K4 trace: one launch over [path rays of queue `cur`] ++ [shadow rays] ++ [BSDF-sampled light rays],
// each segment padded to a multiple of 32 so that a warp runs one kind of query.
//   path ray   : closest hit -> hit record; hits sorted into the per-material-class queues
//   shadow ray : any hit (LightSample::occluded, lighting/mod.rs:125-133)      -> occluded[pid]
//   light ray  : closest hit, `ptr::eq(light, hit)` and lsi.le(-wi) (scene.rs:146-155) -> mis_ok[pid]
// ARN_TRAV_BINARY_SMEM: a block stages the pair records of a small tree in dynamic shared memory (k_trace and the batched queries)
ARN_DEV void stage_pairs(const DevScene& sc, bool block_has_work) {
    if (block_has_work) {                       // block-uniform
        for (uint32_t i = threadIdx.x; i < (ARN_PAIR_BYTES / 16u) * sc.n_pairs; i += blockDim.x) arn_spairs[i] = __ldg(&sc.pairs[i]);
        __syncthreads();
    }
}
template <int MODE>     // ARN_TRAV_BINARY / _COUNTED / _WIDE (traverse.cuh)
// the 4-wide instance serves trees that miss the caches: it trades a few spills for a fourth resident block per SM
// (64 registers; C4 k_trace 36.3 -> 35.9 ms, whole frame +4 %); the binary instance stays at three (80 registers)
#ifndef ARN_TRAV_MINB_WIDE
#define ARN_TRAV_MINB_WIDE (ARN_TRAV_MINB + 1)
#endif
// ARN_TRAV_BINARY_SMEM (trees whose pair records fit ARN_SMEM_NODE_BYTES, traverse2p): ONE block of ARN_BLOCK_SMEM threads per SM stages the
// records in dynamic shared memory once per launch (the same number of resident warps as three 256-thread blocks, one copy instead of three)
#ifndef ARN_BLOCK_SMEM
#define ARN_BLOCK_SMEM 768
#endif
#define ARN_TRACE_BLOCK(MODE) ((MODE) == ARN_TRAV_BINARY_SMEM ? ARN_BLOCK_SMEM : ARN_BLOCK)
__global__ void __launch_bounds__(ARN_TRACE_BLOCK(MODE), MODE == ARN_TRAV_BINARY_SMEM ? 1 : ((MODE == ARN_TRAV_WIDE || MODE == ARN_TRAV_CW8) ? ARN_TRAV_MINB_WIDE : ARN_TRAV_MINB)) k_trace(const __grid_constant__ DevScene sc, PathBuf pb, Queues q, int j) {
    constexpr bool COUNT = MODE == ARN_TRAV_COUNTED;
    constexpr int BLOCK = ARN_TRACE_BLOCK(MODE);
    pdl_prologue();
    uint32_t ctr[3] = {0, 0, 0};
    const uint32_t par = (uint32_t)j & 1u; const int cur = (int)par; const int first = j == 0;
    const uint32_t n_ext = *cnt_active(q.counts, par), n_sh = *cnt_nee(q.counts, par, 1), n_mis = *cnt_nee(q.counts, par, 2);
    if (blockIdx.x == 0 && threadIdx.x < 9) {          // empty the idle counter set (see cnt_* above)
        uint32_t* z = threadIdx.x == 0 ? cnt_active(q.counts, par ^ 1u) : (threadIdx.x < 6 ? cnt_cls(q.counts, par ^ 1u, threadIdx.x - 1) : cnt_nee(q.counts, par ^ 1u, threadIdx.x - 6));
        *z = 0u;
    }
    const uint32_t s1 = (n_ext + 31u) & ~31u, s2 = s1 + ((n_sh + 31u) & ~31u), s3 = s2 + ((n_mis + 31u) & ~31u);
    const uint32_t* __restrict__ ids = q.active[cur];
    // staging state lives in shared memory, not in registers: the traversal below is register-bound (occupancy) and
    // would otherwise carry five row pointers and five fill counters through every walk
    __shared__ uint32_t stage_rows[ARN_NCLS][BLOCK / 32][64];
    __shared__ uint32_t stage_fill[ARN_NCLS][BLOCK / 32];
    if ((threadIdx.x & 31u) == 0) {
#pragma unroll
        for (int c = 0; c < ARN_NCLS; c++) stage_fill[c][threadIdx.x >> 5] = 0;
    }
    __syncwarp();
    if (MODE == ARN_TRAV_BINARY_SMEM) stage_pairs(sc, blockIdx.x * blockDim.x < s3);      // a block without rays skips the copy
    for (uint32_t gi = blockIdx.x * blockDim.x + threadIdx.x; gi < s3; gi += gridDim.x * blockDim.x) {
        // the three kinds of query share ONE inlined walk (instruction-cache footprint): kind and `any` are warp-uniform
        const uint32_t kind = gi < s1 ? 0u : (gi < s2 ? 1u : 2u);                 // 0 path ray, 1 shadow ray, 2 BSDF-sampled light ray
        const uint32_t j = gi - (kind == 0u ? 0u : (kind == 1u ? s1 : s2));
        const uint32_t nk = kind == 0u ? n_ext : (kind == 1u ? n_sh : n_mis);
        const uint32_t* __restrict__ qk = kind == 0u ? ids : (kind == 1u ? q.shadow : q.mis);
        const float4* __restrict__ rk = kind == 0u ? pb.ray : (kind == 1u ? pb.sh : pb.mis);
        uint32_t pid = 0; int cls = -1;
        if (j < nk) {
            pid = __ldcg(&qk[j]);
            const float4 o = __ldcg(&rk[2 * pid]), d = __ldcg(&rk[2 * pid + 1]);   // one 32-byte record; streaming: keep L1 for nodes, slots and the stacks
            TravRay r; trav_init(r, f3(o.x, o.y, o.z), f3(d.x, d.y, d.z), kind == 1u ? o.w : ARN_INF);
            HitRec h;
            // shadow rays: any hit (LightSample::occluded, lighting/mod.rs:125-133).  The counted instance runs the reference's full
            // closest-hit query there (component/mod.rs:35-38) so that its counters are the reference traversal's: same boolean
            trace_ray<MODE>(sc, r, h, ctr, kind == 1u && !COUNT);
            if (kind == 0u) {
                __stcg(&pb.hit[pid], make_float4(__int_as_float(h.prim), h.a, h.b, h.c));
                if (h.prim >= 0) {
                    uint32_t ref = sc.prims[h.prim], mat;
                    if (ref & ARN_PRIM_SPHERE) mat = sc.spheres[ref & ~ARN_PRIM_SPHERE].material;
                    else mat = sc.meshes[sc.tri_mesh[ref]].material;
                    // `*ray = iray` (bvh.rs:108-113): every accepted hit on a transformed sphere leaves the traversal ray round-tripped
                    // through that sphere's frame — also when a nearer primitive is accepted afterwards.  Shading's wo is the ray that
                    // left the traversal, so a changed direction goes back into the ray record.
                    if (r.d.x != d.x || r.d.y != d.y || r.d.z != d.z) pb.ray[2 * pid + 1] = make_float4(r.d.x, r.d.y, r.d.z, d.w);
                    cls = shading_class(sc.materials[mat]);
                }
            } else if (kind == 1u) {
                __stcg(&pb.occluded[pid], h.prim >= 0 ? 1u : 0u);
            } else {
                // `ptr::eq(light, hit)` and lsi.le(-wi) (scene.rs:146-155)
                const float3 wi = f3(d.x, d.y, d.z);
                const uint32_t lcomp = __float_as_uint(d.w);
                uint32_t okl = 0;
                if (h.prim >= 0 && (uint32_t)h.prim == lcomp) {            // ptr::eq(light, hit.as_light()) (scene.rs:149)
                    const DevSphere& sp = sc.spheres[sc.prims[lcomp] & ~ARN_PRIM_SPHERE];
                    float3 pos = f3(h.a, h.b, h.c);
                    if (sp.has_transform) pos = xform_point(sp.local_parent, pos);
                    okl = is_black(light_le(sp, pos, -wi)) ? 0u : 1u;       // lsi.le(-wi)
                }
                __stcg(&pb.mis_ok[pid], okl);
            }
        }
        if (kind == 0u) {
#pragma unroll
            for (int c = 0; c < ARN_NCLS; c++) {
                WarpStage t; t.row = stage_rows[c][threadIdx.x >> 5]; t.fill = stage_fill[c][threadIdx.x >> 5];
                __syncwarp();
                stage_push(t, cls == c, pid, q.cls[c], cnt_cls(q.counts, par, c));
                if ((threadIdx.x & 31u) == 0) stage_fill[c][threadIdx.x >> 5] = t.fill;
                __syncwarp();
            }
        }
    }
#pragma unroll
    for (int c = 0; c < ARN_NCLS; c++) {
        WarpStage t; t.row = stage_rows[c][threadIdx.x >> 5]; t.fill = stage_fill[c][threadIdx.x >> 5];
        stage_flush(t, q.cls[c], cnt_cls(q.counts, par, c));
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        atomicAdd(&q.stats[0], (unsigned long long)n_ext);
        atomicAdd(&q.stats[1], (unsigned long long)n_sh);
        atomicAdd(&q.stats[2], (unsigned long long)n_mis);
        if (!first) atomicAdd(&q.stats[4], (unsigned long long)(n_ext + n_sh + n_mis));
    }
    if (COUNT) {
        unsigned long long a = ctr[0], b = ctr[1], c = ctr[2];
        for (int off = 16; off > 0; off >>= 1) { a += __shfl_down_sync(0xffffffffu, a, off); b += __shfl_down_sync(0xffffffffu, b, off); c += __shfl_down_sync(0xffffffffu, c, off); }
        if ((threadIdx.x & 31) == 0) { atomicAdd(&q.stats[5], a); atomicAdd(&q.stats[6], b); atomicAdd(&q.stats[7], c); }
    }
}

// ---- K3 shade ------------------------------------------------------------------------------------
// One instance per group of material classes (hits arrive sorted by class, so a launch only meets its own lobes):
//   SHADE_DIFFUSE  classes 1, 0  matte: one Lambert / Oren–Nayar lobe, inlined fast path, fewer registers
//   SHADE_GLASS    class 3       Fresnel + Torrance–Sparrow reflection / transmission lobes
//   SHADE_PLASTIC  class 2       one Ashikhmin–Shirley (Beckmann) lobe
//   SHADE_GENERIC  class 4       translucent, and anything else: every lobe kind
#define SHADE_GENERIC 0
#define SHADE_DIFFUSE 1
#define SHADE_PLASTIC 2
#define SHADE_GLASS 3
// TEX (with KIND = SHADE_GENERIC only): the instance for scenes with image textures — every class in one launch, image-plane
// differentials, textured material parameters and bump mapping (kernels/shade_tex.cuh)
template <int KIND, bool TEX = false>
__global__ void __launch_bounds__(ARN_BLOCK, KIND == SHADE_DIFFUSE ? ARN_SHADE_MINB_DIFFUSE : ARN_SHADE_MINB) k_shade(const __grid_constant__ DevScene sc, const __grid_constant__ WaveParams p, PathBuf pb, Queues q, int j) {
    constexpr bool DIFFUSE = KIND == SHADE_DIFFUSE;
    pdl_prologue();
    const uint32_t par = (uint32_t)j & 1u; const int cur = (int)par;      // shade(j) consumes the hits of trace(j)
    constexpr uint32_t LOBES = KIND == SHADE_PLASTIC ? LOBES_PLASTIC : (KIND == SHADE_GLASS ? LOBES_GLASS : LOBES_ALL);
    // One launch over the concatenation of this instance's class queues, each padded to a multiple of 32 so that
    // every warp shades ONE material class.
    const uint32_t ORDER = TEX ? 0x43210u : KIND == SHADE_DIFFUSE ? 0x55501u : (KIND == SHADE_GLASS ? 0x55553u : (KIND == SHADE_PLASTIC ? 0x55552u : 0x55554u));   // nibble k = class shaded k-th (5 = none)
    uint32_t seg_start[ARN_NCLS + 1];
    seg_start[0] = 0;
#pragma unroll
    for (int k = 0; k < ARN_NCLS; k++) { uint32_t c = (ORDER >> (4 * k)) & 0xFu; seg_start[k + 1] = seg_start[k] + (c < ARN_NCLS ? ((*cnt_cls(q.counts, par, c) + 31u) & ~31u) : 0u); }
    uint32_t* next = q.active[cur ^ 1];
    __shared__ uint32_t stage_rows[4][ARN_BLOCK / 32][64];
    WarpStage st_next, st_conn, st_sh, st_mis;
    st_next.row = stage_rows[0][threadIdx.x >> 5]; st_next.fill = 0;
    st_conn.row = stage_rows[1][threadIdx.x >> 5]; st_conn.fill = 0;
    st_sh.row = stage_rows[2][threadIdx.x >> 5]; st_sh.fill = 0;
    st_mis.row = stage_rows[3][threadIdx.x >> 5]; st_mis.fill = 0;
    const uint32_t n_round = seg_start[ARN_NCLS];       // multiple of 32: warp-uniform trip count
    for (uint32_t gi = blockIdx.x * blockDim.x + threadIdx.x; gi < n_round; gi += gridDim.x * blockDim.x) {
        uint32_t k = 0, base = 0;
#pragma unroll
        for (int j = 1; j < ARN_NCLS; j++) if (gi >= seg_start[j]) { k = (uint32_t)j; base = seg_start[j]; }
        const uint32_t cls = (ORDER >> (4 * k)) & 0xFu;
        const uint32_t i = gi - base;
        const uint32_t n = *cnt_cls(q.counts, par, cls);
        const uint32_t* __restrict__ ids = q.cls[cls];
        bool alive = false, nee = false, has_sh = false, has_mis = false;
        uint32_t pid = 0;
        if (i < n) {
            pid = ids[i];
            const float4 hr = pb.hit[pid];
            const int prim = __float_as_int(hr.x);
            if (prim >= 0) {
                const float4 o4 = pb.ray[2 * pid], d4 = pb.ray[2 * pid + 1];     // one 32-byte record: (-, sampler key) (direction, st)
                float3 raydir = f3(d4.x, d4.y, d4.z);
                uint32_t st = __float_as_uint(d4.w);
                uint32_t bounces = st & 0xffu; bool spec = ((st >> 8) & 1u) != 0;
                Sampler sm; sm.init_key(__float_as_uint(o4.w), (st >> 16) & 0xffu, st >> 24);
                uint32_t spix = 0, ssmp = 0;
                if (p.strat_ndim && (sm.i1d < p.strat_ndim || sm.i2d < p.strat_ndim)) { spix = pb.pix[pid]; ssmp = pb.smp[pid]; }     // stratified draws are keyed by (pixel, sample)
                float4 b4 = pb.beta[pid]; float3 beta = f3(b4.x, b4.y, b4.z);
                uint32_t ref = sc.prims[prim];
                Surf s; uint32_t mat; SurfTex sx;
                if (ref & ARN_PRIM_SPHERE) {
                    const DevSphere& sp = sc.spheres[ref & ~ARN_PRIM_SPHERE];
                    surf_sphere(sp, f3(hr.y, hr.z, hr.w), raydir, s, TEX ? &sx : nullptr);
                    mat = sp.material;
                    if ((bounces == 0 || spec) && sp.emissive) {                                              // pt.rs:72-78
                        // the only place a shade launch changes L: touch the radiance stream for these hits only
                        float4 l4 = pb.L[pid];
                        float3 L = f3(l4.x, l4.y, l4.z) + beta * light_le(sp, s.pos, -raydir);
                        pb.L[pid] = make_float4(L.x, L.y, L.z, 0.f);
                    }
                } else {
                    surf_triangle(sc, ref, hr.y, hr.z, hr.w, raydir, s, TEX ? &sx : nullptr);
                    mat = sc.meshes[sc.tri_mesh[ref]].material;
                }
                Bsdf bsdf;
                DxyInfo dxy;
                if (TEX) {          // pt.rs:80-83: dxy from the offset rays, then the material evaluates its textures (and bumps the frame)
                    const float4 r0 = pb.diff[4 * pid], r1 = pb.diff[4 * pid + 1], r2 = pb.diff[4 * pid + 2], r3 = pb.diff[4 * pid + 3];
                    dxy = compute_dxy(s.pos, s.ng, sx.duv_dpdu, sx.duv_dpdv, f3(r0.x, r0.y, r0.z), f3(r1.x, r1.y, r1.z), f3(r2.x, r2.y, r2.z), f3(r3.x, r3.y, r3.z));
                    const arn_material m = textured_material(sc, mat, s, sx, dxy);
                    bsdf_build(m, s, bsdf);
                } else bsdf_build(sc.materials[mat], s, bsdf);
                if (!DIFFUSE) bsdf_prepare<LOBES>(bsdf, s.wo);
                uint32_t flags = 0;
                if (bsdf.n > 0) {                                           // pt.rs:85-91 (have_n(ALL-SPECULAR) > 0 <=> any lobe)
                    // Scene::uniform_sample_one_light (scene.rs:58-66)
                    float u1 = sm.next(); ARN_STRAT_1D(u1, sm.i1d);
                    float2 ulight = sm.next_2d(); ARN_STRAT_2D(ulight, sm.i2d);
                    float2 uscatter = sm.next_2d(); ARN_STRAT_2D(uscatter, sm.i2d);
                    if (u1 == 0.f) u1 += ARN_EPS;                            // Distribution1D::search_offset
                    uint32_t lo = 0, hi = sc.n_lights + 1;
                    while (lo < hi) { uint32_t mid = lo + (hi - lo) / 2; if (sc.light_cdf[mid] < u1) lo = mid + 1; else hi = mid; }
                    uint32_t lidx = lo - 1;
                    float lightpdf = sc.light_integral > 0.f ? sc.light_func[lidx] / sc.light_integral : 0.f;
                    uint32_t lcomp = sc.light_prims[lidx];
                    // a Point / Spot / Distant light (pointlights.rs, distantlight.rs) or an emissive sphere
                    const bool analytic = (lcomp & ARN_LIGHT_ANALYTIC) != 0;
                    const bool delta = analytic && sc.analytic[lcomp & ~ARN_LIGHT_ANALYTIC].type != ARN_LIGHT_DISTANT;
                    const DevSphere& light = sc.spheres[analytic ? 0u : (sc.prims[lcomp] & ~ARN_PRIM_SPHERE)];   // not read when analytic
                    flags = NEE_DONE;
                    // Scene::evaluate_direct (scene.rs:83-167): light sampling half
                    LightSample ls = analytic ? analytic_sample(sc.analytic[lcomp & ~ARN_LIGHT_ANALYTIC], s.pos) : light_sample(light, s.pos, ulight);
                    float3 wi = normalize(ls.pfrom - ls.pto);
                    float3 A1 = grey(0.f);
                    if (!(ls.pdf == 0.f || is_black(ls.radiance))) {
                        float3 f; float spdf;
                        if (DIFFUSE) { f = bsdf_eval_k<true>(bsdf, s.wo, wi) * fabsf(dot(wi, s.ns)); spdf = bsdf_pdf_k<true>(bsdf, s.wo, wi); }
                        else { bsdf_eval_pdf<LOBES>(bsdf, s.wo, wi, f, spdf); f = f * fabsf(dot(wi, s.ns)); }
                        if (spdf == 0.f) f = grey(0.f);
                        float weight = delta ? 1.f : power_heuristic(ls.pdf, spdf);    // is_delta: no MIS weight (scene.rs:107-115); x * 1 is exact
                        A1 = ls.radiance * f * weight / ls.pdf;
                        if (!is_black(f)) {
                            // LightSample::occluded (lighting/mod.rs:125-133) + RawRay::spawn (ray.rs:93-98)
                            const float eps = ARN_EPS * 2.0f;
                            float3 dir = ls.pto - ls.pfrom;
                            float3 a = ls.pfrom + dir * eps;
                            float3 b = ls.pto + (-dir * eps);
                            float3 v = b - a;
                            float len = length(v);
                            float3 vd = v / len;
                            pb.sh[2 * pid] = make_float4(a.x, a.y, a.z, len);
                            pb.sh[2 * pid + 1] = make_float4(vd.x, vd.y, vd.z, 0.f);
                            flags |= NEE_SHADOW; has_sh = true;
                        }
                    }
                    // BSDF sampling half (scene.rs:128-165), skipped for delta lights.  A Distant light has no
                    // Light::pdf (default 0, lighting/mod.rs:64-66) and is nobody's `as_light()`: non-specular
                    // samples stop at lpdf == 0, specular ones trace a ray that can only add black.
                    float3 A2 = grey(0.f);
                    Sampled bs; bs.f = grey(0.f); bs.wi = f3(0.f, 1.f, 0.f); bs.pdf = 0.f; bs.type = 0;
                    if (!delta) bs = bsdf_sample_k<DIFFUSE, LOBES>(bsdf, s.wo, uscatter);
                    float3 f2v = bs.f * fabsf(dot(bs.wi, s.ns));
                    if (!is_black(f2v) && bs.pdf > 0.f) {
                        float weight = 1.f; bool skip = false;
                        if (!(bs.type & BXDF_SPECULAR)) {
                            float lpdf = analytic ? 0.f : light_pdf(light, s.pos, bs.wi);
                            if (lpdf == 0.f) skip = true; else weight = power_heuristic(bs.pdf, lpdf);
                        }
#ifdef ARN_DBG_PIX
                        if (pb.pix[pid] == (uint32_t)(ARN_DBG_PIX) && pb.smp[pid] == (uint32_t)(ARN_DBG_SMP))
                            printf("[dbg] pos %.9g %.9g %.9g ns %.9g %.9g %.9g wo %.9g %.9g %.9g uscatter %.9g %.9g\n      bs.f %.9g %.9g %.9g wi %.9g %.9g %.9g pdf %.9g type %u f2v %.9g %.9g %.9g lpdf %.9g weight %.9g skip %d\n",
                                   s.pos.x, s.pos.y, s.pos.z, s.ns.x, s.ns.y, s.ns.z, s.wo.x, s.wo.y, s.wo.z, uscatter.x, uscatter.y, bs.f.x, bs.f.y, bs.f.z, bs.wi.x, bs.wi.y, bs.wi.z, bs.pdf, bs.type, f2v.x, f2v.y, f2v.z,
                                   (bs.type & BXDF_SPECULAR) ? -1.f : light_pdf(light, s.pos, bs.wi), weight, (int)skip);
#endif
                        if (!skip) {
                            float3 mo = offset_towards(s, bs.wi);
                            if (!analytic) A2 = f2v * sphere_emission(light) * weight / bs.pdf;
                            pb.mis[2 * pid] = make_float4(mo.x, mo.y, mo.z, 0.f);
                            pb.mis[2 * pid + 1] = make_float4(bs.wi.x, bs.wi.y, bs.wi.z, __uint_as_float(lcomp));
                            flags |= NEE_MIS; has_mis = true;
                        }
                    }
                    // A direct-light term that is exactly zero adds beta * (0 / p_light) = +0 to L: it needs no record and no
                    // resolve, provided p_light is a positive finite number (0 / 0 and 0 / NaN must still poison the sample as
                    // they do in the reference, scene.rs:65 + pt.rs:89,152-156).  NaN terms are not "black" and keep their record.
                    nee = has_mis || !is_black(A1) || !(lightpdf > 0.f && lightpdf < ARN_INF) || !(beta.x < ARN_INF && beta.y < ARN_INF && beta.z < ARN_INF);
                    if (nee) {
                        pb.nee[4 * pid] = make_float4(A1.x, A1.y, A1.z, lightpdf);
                        pb.nee[4 * pid + 1] = make_float4(A2.x, A2.y, A2.z, __uint_as_float(flags));
                        pb.nee[4 * pid + 2] = make_float4(beta.x, beta.y, beta.z, 0.f);
                    }
                }
                // sample the BSDF for the next direction (pt.rs:92-107)
                float3 next_o = f3(0.f, 0.f, 0.f), next_d = next_o;
                float3 wo = -raydir;
                float2 ubsdf = sm.next_2d(); ARN_STRAT_2D(ubsdf, sm.i2d);
                Sampled bs = bsdf_sample_k<DIFFUSE, LOBES>(bsdf, wo, ubsdf);
                spec = (bs.type & BXDF_SPECULAR) != 0;
                alive = !(is_black(bs.f) || bs.pdf == 0.f);
                if (alive) {
                    beta = beta * (bs.f * (fabsf(dot(bs.wi, s.ns)) / bs.pdf));
                    bool valid = !(isnan(beta.x) || isnan(beta.y) || isnan(beta.z)) && !(isinf(beta.x) || isinf(beta.y) || isinf(beta.z))
                                 && beta.x >= 0.f && beta.y >= 0.f && beta.z >= 0.f;
                    if (!valid) alive = false;
                }
                if (alive) {
                    float3 no = offset_towards(s, bs.wi);
                    next_o = no; next_d = bs.wi;
                    bounces += 1;
                    if (bounces >= p.max_depth) alive = false;
                }
                if (alive) {                                               // Russian roulette (pt.rs:117-122)
                    float y = 0.212671f * beta.x + 0.715160f * beta.y + 0.072169f * beta.z;
                    if (y < p.rr_threshold && bounces >= p.min_depth) {
                        float qq = fmaxf(p.rr_threshold, 0.05f);
                        float urr = sm.next(); ARN_STRAT_1D(urr, sm.i1d);
                        if (urr < qq) alive = false;
                        else beta = beta / (1.f - qq);
                    }
                }
                if (alive) {         // the continuation ray: one 32-byte record, the sampler key travels with it
                    pb.beta[pid] = make_float4(beta.x, beta.y, beta.z, 0.f);
                    if (TEX) {      // spawn_ray_differential(wi, Some(&dxy)) (interaction.rs:236-251)
                        const float3 xo = next_o + dxy.dpdx, yo = next_o + dxy.dpdy;
                        pb.diff[4 * pid] = make_float4(xo.x, xo.y, xo.z, 0.f); pb.diff[4 * pid + 1] = make_float4(next_d.x, next_d.y, next_d.z, 0.f);
                        pb.diff[4 * pid + 2] = make_float4(yo.x, yo.y, yo.z, 0.f); pb.diff[4 * pid + 3] = make_float4(next_d.x, next_d.y, next_d.z, 0.f);
                    }
                    pb.ray[2 * pid] = make_float4(next_o.x, next_o.y, next_o.z, o4.w);
                    pb.ray[2 * pid + 1] = make_float4(next_d.x, next_d.y, next_d.z, __uint_as_float((bounces & 0xffu) | ((spec ? 1u : 0u) << 8) | ((sm.i1d & 0xffu) << 16) | (sm.i2d << 24)));
                }
            }
        }
        stage_push(st_next, alive, pid, next, cnt_active(q.counts, par ^ 1u));
        stage_push(st_conn, nee, pid, q.connect, cnt_nee(q.counts, par ^ 1u, 0));
        stage_push(st_sh, has_sh, pid, q.shadow, cnt_nee(q.counts, par ^ 1u, 1));
        stage_push(st_mis, has_mis, pid, q.mis, cnt_nee(q.counts, par ^ 1u, 2));
    }
    stage_flush(st_next, next, cnt_active(q.counts, par ^ 1u));
    stage_flush(st_conn, q.connect, cnt_nee(q.counts, par ^ 1u, 0));
    stage_flush(st_sh, q.shadow, cnt_nee(q.counts, par ^ 1u, 1));
    stage_flush(st_mis, q.mis, cnt_nee(q.counts, par ^ 1u, 2));
}

// ---- resolve: L += beta * (light term + BSDF term) / p_light  (evaluate_direct's sum, pt.rs:89) -----
__global__ void __launch_bounds__(ARN_BLOCK) k_resolve(PathBuf pb, Queues q, int j) {      // resolve(j): the direct-light terms shade(j) queued, after trace(j + 1) traced their rays
    pdl_prologue();
    const uint32_t n = *cnt_nee(q.counts, ((uint32_t)j + 1u) & 1u, 0);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        uint32_t pid = q.connect[i];
        const float4 a1 = pb.nee[4 * pid], a2 = pb.nee[4 * pid + 1], bo = pb.nee[4 * pid + 2];     // 48 bytes of one 64-byte record
        uint32_t flags = __float_as_uint(a2.w);
        float3 ret = f3(a1.x, a1.y, a1.z);
        if ((flags & NEE_SHADOW) && pb.occluded[pid]) ret = grey(0.f);
        if ((flags & NEE_MIS) && pb.mis_ok[pid]) ret = ret + f3(a2.x, a2.y, a2.z);
        if (is_black(ret) && a1.w > 0.f && a1.w < ARN_INF && bo.x < ARN_INF && bo.y < ARN_INF && bo.z < ARN_INF) continue;   // L + beta * (+0) = L: leave the radiance stream alone
        float3 term = ret / a1.w;                                        // evaluate_direct(..) / lightpdf
        float4 l4 = pb.L[pid];
        float3 L = f3(l4.x, l4.y, l4.z) + f3(bo.x, bo.y, bo.z) * term;   // ret += beta * term
        pb.L[pid] = make_float4(L.x, L.y, L.z, 0.f);
    }
}

// ---- K6 accumulate: filtered film splat of every sample of the wave (film.rs:297-319) ---------
// film filter weights feed sums only (no discrete decision): libdevice f32 sinf (<= 2 ulp) instead of the
// correctly-rounded f64 route; covered by the film-sum tolerance of tests/test_gpu_parity.py
ARN_DEV float sinc1(float x) { if (x < 1.0e-5f) return 1.f; float xpi = x * ARN_PI; return sinf(xpi) / xpi; }
ARN_DEV float mitchell1(float x, float b, float c) {                                 // MitchellFilter::mitchell_1d (filters.rs:154-169)
    const float INV_SIX = 1.0f / 6.0f;
    if (x > 1.0f) return (-b - 6.0f * c) * x * x * x + (6.0f * b + 30.0f * c) * x * x - (12.0f * b + 48.0f * c) * x + (8.0f * b + 24.0f * c) * INV_SIX;
    return (12.0f - 9.0f * b - 6.0f * c) * x * x * x + (-18.0f - 12.0f * b + 6.0f * c) * x * x + (6.0f - 2.0f * b) * INV_SIX;
}
// One axis of the film filter at signed offset x (every filter of sample/filters.rs is a product of two such
// factors); r = that axis' radius.  Lanczos: p.filt_a = 1 / tau (film.rs:47-51: tau = 3 after deserialisation).
ARN_DEV float filter1(const WaveParams& p, float x, float r) {
    switch (p.filt_kind) {
    case ARN_FILTER_BOX: return 1.f;
    case ARN_FILTER_TRIANGLE: return r - fabsf(x);
    case ARN_FILTER_GAUSSIAN: return expf(p.filt_a * x * x) - p.filt_a * r * r;       // filt_a = -alpha; sic (filters.rs:104-107,121-125)
    case ARN_FILTER_MITCHELL: return mitchell1(fabsf(2.0f * ((1.0f / r) * x)), p.filt_a, p.filt_b);
    default: return sinc1(x * p.filt_a) * sinc1(x);
    }
}

// FilmTile sink of the tile that owns tile pixel (px, py): the tile grown by the (truncated) filter radius; the crop
// window clip is applied by the callers.  Only differs from the crop clip for fractional radii.
ARN_DEV void tile_sink(const WaveParams& p, int px, int py, int& sx0, int& sy0, int& sx1, int& sy1) {
    int ix = min(px / p.tile_dx, p.tiles_nx - 1), iy = min(py / p.tile_dy, p.tiles_ny - 1);
    sx0 = ix * p.tile_dx - p.tile_rx; sy0 = iy * p.tile_dy - p.tile_ry;
    sx1 = ix * p.tile_dx + (ix == p.tiles_nx - 1 ? p.tile_lastx : p.tile_dx) + p.tile_rx;
    sy1 = iy * p.tile_dy + (iy == p.tiles_ny - 1 ? p.tile_lasty : p.tile_dy) + p.tile_ry;
}

__global__ void __launch_bounds__(ARN_BLOCK) k_accumulate(const __grid_constant__ WaveParams p, PathBuf pb, Queues q, float4* __restrict__ film, uint32_t n) {
    pdl_prologue();
    unsigned long long invalid = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float4 l4 = pb.L[i];
        float3 L = f3(l4.x, l4.y, l4.z);
        bool valid = !(isnan(L.x) || isnan(L.y) || isnan(L.z)) && !(isinf(L.x) || isinf(L.y) || isinf(L.z)) && L.x >= 0.f && L.y >= 0.f && L.z >= 0.f;
        if (!valid) { L = grey(0.f); invalid++; }                       // pt.rs:152-156
        float2 pos = pb.pfilm[i];
        float cx = pos.x - p.fr_x + 0.5f, cy = pos.y - p.fr_y + 0.5f;
        float fx = pos.x + p.fr_x - 0.5f, fy = pos.y + p.fr_y - 0.5f;
        int x0 = (int)cx, y0 = (int)cy, x1 = (int)fx + 1, y1 = (int)fy + 1;   // truncation toward zero
        if (x0 > x1) { int t = x0; x0 = x1; x1 = t; }
        if (y0 > y1) { int t = y0; y0 = y1; y1 = t; }
        // the tile sink (tile grown by the radius, clipped to the crop window) never clips more
        // than the crop window does for samples inside the tile (DESIGN.md "Film")
        x0 = max(x0, p.crop_x0); y0 = max(y0, p.crop_y0); x1 = min(x1, p.crop_x0 + p.crop_w); y1 = min(y1, p.crop_y0 + p.crop_h);
        {
            uint32_t pix = pb.pix[i]; int sx0, sy0, sx1, sy1;
            tile_sink(p, (int)(pix & 0xffffu), (int)(pix >> 16), sx0, sy0, sx1, sy1);
            x0 = max(x0, sx0); y0 = max(y0, sy0); x1 = min(x1, sx1); y1 = min(y1, sy1);
        }
        for (int y = y0; y < y1; y++) {
            float wy = filter1(p, ((float)y + 0.5f) - pos.y, p.fr_y);
            for (int x = x0; x < x1; x++) {
                float wx = filter1(p, ((float)x + 0.5f) - pos.x, p.fr_x);
                float w = wx * wy;
                float3 c = L * w;
                atomicAdd(&film[(size_t)(y - p.crop_y0) * (size_t)p.crop_w + (size_t)(x - p.crop_x0)], make_float4(c.x, c.y, c.z, w));
            }
        }
    }
    for (int off = 16; off > 0; off >>= 1) invalid += __shfl_down_sync(0xffffffffu, invalid, off);
    if ((threadIdx.x & 31) == 0 && invalid) atomicAdd(&q.stats[3], invalid);
}

// ---- K6 accumulate, warp-per-pixel form (filter radius <= 4): a warp owns ONE pixel of the wave (its samples are
// consecutive path slots) and keeps the 9x9 window of affected film pixels in registers (3 targets per lane): one vector
// atomic per target per pixel instead of 64 per sample.  Two phases per chunk of 32 samples: (1) lane k validates sample k,
// clips its splat box and evaluates its 9 + 9 separable filter factors (zero outside the box: adding L * 0 and 0 is exact)
// into shared memory; (2) the warp walks the chunk in sample order, every lane adding w = wx * wy and L * w to its three
// targets.  Every (L*w, w) term is the scatter kernel's; only the order of the additions across pixels differs.
#define ARN_ACC_STRIDE 19            // 18 factors per sample, odd stride: conflict-free stores
__global__ void __launch_bounds__(ARN_BLOCK) k_accumulate_px(const __grid_constant__ WaveParams p, PathBuf pb, Queues q, float4* __restrict__ film,
                                                              unsigned long long wave_base, uint32_t n) {
    __shared__ float s_w[ARN_BLOCK / 32][32 * ARN_ACC_STRIDE];
    __shared__ float4 s_L[ARN_BLOCK / 32][32];
    pdl_prologue();
    const unsigned lane = threadIdx.x & 31u, wib = threadIdx.x >> 5;
    float* __restrict__ wrow = s_w[wib];
    float4* __restrict__ lrow = s_L[wib];
    const unsigned long long warp = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const unsigned long long nwarps = ((unsigned long long)gridDim.x * blockDim.x) >> 5;
    const unsigned long long first_pl = wave_base / p.spp_count, last_pl = (wave_base + n - 1) / p.spp_count;
    unsigned long long invalid = 0;
    int tx[3], ty[3];
#pragma unroll
    for (int j = 0; j < 3; j++) { const int t = min((int)lane + 32 * j, 80); tx[j] = t % 9; ty[j] = 9 + t / 9; }
    for (unsigned long long pl = first_pl + warp; pl <= last_pl; pl += nwarps) {
        unsigned long long g0 = pl * p.spp_count, g1 = g0 + p.spp_count;
        uint32_t s_begin = (uint32_t)((g0 > wave_base ? g0 : wave_base) - wave_base);
        uint32_t s_end = (uint32_t)((g1 < wave_base + n ? g1 : wave_base + n) - wave_base);
        uint32_t pix = pb.pix[s_begin];
        int px = (int)(pix & 0xffffu), py = (int)(pix >> 16);
        int sx0, sy0, sx1, sy1;
        tile_sink(p, px, py, sx0, sy0, sx1, sy1);
        float4 acc[3];
#pragma unroll
        for (int j = 0; j < 3; j++) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (uint32_t base = s_begin; base < s_end; base += 32u) {
            const uint32_t cnt = min(32u, s_end - base);
            if (lane < cnt) {
                const uint32_t s = base + lane;
                float4 l4 = pb.L[s];
                bool valid = !(isnan(l4.x) || isnan(l4.y) || isnan(l4.z)) && !(isinf(l4.x) || isinf(l4.y) || isinf(l4.z)) && l4.x >= 0.f && l4.y >= 0.f && l4.z >= 0.f;
                if (!valid) { l4 = make_float4(0.f, 0.f, 0.f, 0.f); invalid++; }     // pt.rs:152-156
                lrow[lane] = l4;
                float2 pos = pb.pfilm[s];
                float cx = pos.x - p.fr_x + 0.5f, cy = pos.y - p.fr_y + 0.5f;
                float fx = pos.x + p.fr_x - 0.5f, fy = pos.y + p.fr_y - 0.5f;
                int x0 = (int)cx, y0 = (int)cy, x1 = (int)fx + 1, y1 = (int)fy + 1;   // truncation toward zero
                if (x0 > x1) { int t = x0; x0 = x1; x1 = t; }
                if (y0 > y1) { int t = y0; y0 = y1; y1 = t; }
                x0 = max(x0, sx0); y0 = max(y0, sy0); x1 = min(x1, sx1); y1 = min(y1, sy1);
                float* __restrict__ w = wrow + lane * ARN_ACC_STRIDE;
#pragma unroll 1
                for (int i = 0; i < 9; i++) {
                    const int X = px - 4 + i, Y = py - 4 + i;
                    const float wx = filter1(p, ((float)X + 0.5f) - pos.x, p.fr_x), wy = filter1(p, ((float)Y + 0.5f) - pos.y, p.fr_y);
                    w[i] = (X >= x0 && X < x1) ? wx : 0.f;
                    w[9 + i] = (Y >= y0 && Y < y1) ? wy : 0.f;
                }
            }
            __syncwarp();
            for (uint32_t k = 0; k < cnt; k++) {
                const float4 l4 = lrow[k];
                const float* __restrict__ w = wrow + k * ARN_ACC_STRIDE;
#pragma unroll
                for (int j = 0; j < 3; j++) {
                    const float wv = w[tx[j]] * w[ty[j]];
                    acc[j].x += l4.x * wv; acc[j].y += l4.y * wv; acc[j].z += l4.z * wv; acc[j].w += wv;
                }
            }
            __syncwarp();
        }
#pragma unroll
        for (int j = 0; j < 3; j++) {
            int t = (int)lane + 32 * j;
            int X = px - 4 + t % 9, Y = py - 4 + t / 9;
            if (t < 81 && X >= p.crop_x0 && X < p.crop_x0 + p.crop_w && Y >= p.crop_y0 && Y < p.crop_y0 + p.crop_h
                && (acc[j].x != 0.f || acc[j].y != 0.f || acc[j].z != 0.f || acc[j].w != 0.f))
                atomicAdd(&film[(size_t)(Y - p.crop_y0) * (size_t)p.crop_w + (size_t)(X - p.crop_x0)], acc[j]);
        }
    }
    for (int off = 16; off > 0; off >>= 1) invalid += __shfl_down_sync(0xffffffffu, invalid, off);
    if (lane == 0 && invalid) atomicAdd(&q.stats[3], invalid);
}

// diagnostic: per-sample radiance, indexed ((y*crop_w + x)*spp_count + (s - spp_begin)) by the TILE pixel (x, y),
// which runs over [0, crop_w) x [0, crop_h) whatever crop.pmin is (film.rs:118-121) (parity tests)
__global__ void __launch_bounds__(ARN_BLOCK) k_store_radiance(const __grid_constant__ WaveParams p, PathBuf pb, float4* __restrict__ out, uint32_t n) {
    pdl_prologue();
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        uint32_t pix = pb.pix[i]; uint32_t x = pix & 0xffffu, y = pix >> 16;
        size_t idx = ((size_t)y * p.crop_w + x) * p.spp_count + (pb.smp[i] - p.spp_begin);
        out[idx] = pb.L[i];
    }
}

// ---- standalone batched queries (arn_intersect_closest / arn_intersect_any) ----------------------
template <int MODE>
__global__ void __launch_bounds__(ARN_TRACE_BLOCK(MODE), MODE == ARN_TRAV_BINARY_SMEM ? 1 : ARN_TRAV_MINB) k_closest_batch(const __grid_constant__ DevScene sc, const arn_ray* __restrict__ rays, size_t n, arn_hit* __restrict__ hits,
                                                             unsigned long long* ctr_out) {
    constexpr bool COUNT = MODE == ARN_TRAV_COUNTED;
    uint32_t ctr[3] = {0, 0, 0};
    if (MODE == ARN_TRAV_BINARY_SMEM) stage_pairs(sc, (size_t)blockIdx.x * blockDim.x < n);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        arn_ray ry = rays[i];
        TravRay r; trav_init(r, f3(ry.o[0], ry.o[1], ry.o[2]), f3(ry.d[0], ry.d[1], ry.d[2]), ry.tmax);
        uint32_t before = ctr[0];
        HitRec h; trace_ray<MODE>(sc, r, h, ctr, false);
        arn_hit o; o.prim_id = h.prim; o.t = h.prim >= 0 ? r.tmax : ARN_INF;
        hits[i] = o;
        if (COUNT) {   // lane-utilisation probe: sum over warps of the longest ray's node count
            uint32_t steps = ctr[0] - before, mx = steps;
            for (int off = 16; off > 0; off >>= 1) mx = max(mx, __shfl_xor_sync(__activemask(), mx, off));
            if ((threadIdx.x & 31) == 0) atomicAdd(&ctr_out[3], (unsigned long long)mx);
        }
    }
    if (COUNT) {
        unsigned long long a = ctr[0], b = ctr[1], c = ctr[2];
        for (int off = 16; off > 0; off >>= 1) { a += __shfl_down_sync(0xffffffffu, a, off); b += __shfl_down_sync(0xffffffffu, b, off); c += __shfl_down_sync(0xffffffffu, c, off); }
        if ((threadIdx.x & 31) == 0) { atomicAdd(&ctr_out[0], a); atomicAdd(&ctr_out[1], b); atomicAdd(&ctr_out[2], c); }
    }
}
template <int MODE>
__global__ void __launch_bounds__(ARN_TRACE_BLOCK(MODE), MODE == ARN_TRAV_BINARY_SMEM ? 1 : ARN_TRAV_MINB) k_any_batch(const __grid_constant__ DevScene sc, const arn_ray* __restrict__ rays, size_t n, uint8_t* __restrict__ out) {
    if (MODE == ARN_TRAV_BINARY_SMEM) stage_pairs(sc, (size_t)blockIdx.x * blockDim.x < n);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        arn_ray ry = rays[i];
        TravRay r; trav_init(r, f3(ry.o[0], ry.o[1], ry.o[2]), f3(ry.d[0], ry.d[1], ry.d[2]), ry.tmax);
        HitRec h; trace_ray<MODE>(sc, r, h, nullptr, true);
        out[i] = h.prim >= 0 ? 1 : 0;
    }
}

}  // namespace arn
