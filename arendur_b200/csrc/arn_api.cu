// C-ABI of the B200 path-tracing core (include/arn.h): contexts, scene upload, batched
// intersection queries and the wavefront path tracer.  Compiled for sm_100a only, with
// --fmad=false (see kernels/dev_math.cuh).  There is no CPU fallback: every entry point that
// computes anything needs a CUDA device and fails with ARN_E_CUDA otherwise.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>
#include "../../include/arn.h"
#include "kernels/wavefront.cuh"
#include "kernels/trace_refill.cuh"
#include "kernels/lbvh.cuh"
#include "kernels/wide_build.cuh"
#include "kernels/cw8_build.cuh"
#include <cub/device/device_scan.cuh>
#include <cub/device/device_radix_sort.cuh>

using namespace arn;

#define ARN_MAX_PIPES 8

namespace {
std::string g_last_error;
std::mutex g_err_mutex;
}

struct arn_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    int sm_count = 0;
    std::string err;
    std::recursive_mutex mu;     // calls on one context are serialised (SURVEY.md §8(b) Threading); recursive: host-buffer entry points call the device ones
    // wave pipelines: each owns a stream and a set of wavefront buffers (allocated lazily, sized to the wave
    // capacity).  Consecutive waves of a render go to consecutive pipelines and run concurrently, so the thin
    // late-bounce launches of one wave overlap the wide early launches of the next.  pipes[0] runs on `stream`.
    struct Pipe { cudaStream_t stream = nullptr; size_t wave_cap = 0; PathBuf pb{}; Queues q{}; void* pool = nullptr; cudaEvent_t done = nullptr;
                  TraceBuf tb{}; void* tb_pool = nullptr; size_t tb_cap = 0;
                  void* diff_pool = nullptr; size_t diff_cap = 0; };                // ray-differential stream (textured scenes)     // tb: ray-stream buffers of the lane-refilling trace (allocated on first use)
    Pipe pipes[ARN_MAX_PIPES];
    int opt_pipes = 0;           // ARN_OPT_PIPELINES: 0 = auto (4, or 8 for trees large enough for the 4-wide walk)
    // tile tables
    int4* d_tile_rect = nullptr; unsigned long long* d_tile_prefix = nullptr; size_t tile_cap = 0;
    // event pool for per-kernel timing
    std::vector<cudaEvent_t> events;
    // launch geometry (blocks per kernel, persistent grid-stride)
    int g_generate = 0, g_trace = 0, g_shade = 0, g_shade_d = 0, g_resolve = 0, g_accum = 0, g_accum_px = 0, g_closest = 0, g_any = 0;
    // scratch for batched queries through host buffers
    void* d_rays = nullptr; void* d_hits = nullptr; size_t rays_cap = 0;
    unsigned long long* d_ctr = nullptr;
    void* d_film = nullptr; size_t film_cap = 0;     // arn_render_pt's device film
    bool opt_count = false;
    int opt_width = 0;           // ARN_OPT_BVH_WIDTH: 0 auto, 2 binary, 4 wide
    int g_trace_w = 0, g_closest_w = 0, g_any_w = 0, g_shade_p = 0, g_shade_g = 0;
    int g_trace_8 = 0, g_closest_8 = 0, g_any_8 = 0;      // compressed 8-wide walk
    size_t opt_wave = 0;
    int opt_pdl = 0;             // ARN_OPT_PDL: programmatic dependent launch along a pipeline's kernel chain
    int opt_smem_off = 0;        // ARN_OPT_SMEM_NODES = 1: never walk small trees from shared memory
    int opt_refill = 0;          // ARN_OPT_TRACE_REFILL: lane-refilling trace (kernels/trace_refill.cuh) for trees walked with the binary nodes
    int g_setup = 0, g_refill = 0, g_classify = 0, g_shade_tex = 0;
};

struct arn_scene {
    arn_ctx* ctx = nullptr;
    DevScene dev{};
    std::vector<void*> allocs;
    char* pool = nullptr; size_t pool_size = 0, pool_off = 0;   // one device allocation per scene, carved in 256-byte steps
    uint32_t max_depth = 0;
    uint32_t class_mask = 0;     // shading classes present in the material table
    uint64_t bytes = 0;
};

namespace {

int set_err(arn_ctx* c, int code, const std::string& msg) {
    { std::lock_guard<std::mutex> g(g_err_mutex); g_last_error = msg; }
    if (c) c->err = msg;
    return code;
}
#define CUDA_TRY(ctx, call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
    return set_err(ctx, e_ == cudaErrorMemoryAllocation ? ARN_E_OOM : ARN_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); } } while (0)

inline size_t align256(size_t b) { return (b + 255) & ~(size_t)255; }
void* pool_take(arn_scene* s, size_t bytes) {
    size_t o = s->pool_off;
    if (!s->pool || o + align256(bytes) > s->pool_size) return nullptr;
    s->pool_off = o + align256(bytes);
    return s->pool + o;
}
template <typename T, typename P> int dev_upload(arn_scene* s, const T* host, size_t n, P* out) {
    *out = nullptr;
    if (n == 0 || !host) return ARN_OK;
    void* d = pool_take(s, n * sizeof(T));
    if (!d) return set_err(s->ctx, ARN_E_OOM, "scene pool exhausted (internal sizing error)");
    CUDA_TRY(s->ctx, cudaMemcpyAsync(d, host, n * sizeof(T), cudaMemcpyHostToDevice, s->ctx->stream));
    s->bytes += n * sizeof(T);
    *out = (P)d;
    return ARN_OK;
}

int grid_for(arn_ctx* c, const void* kernel) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, ARN_BLOCK, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
    return c->sm_count * per_sm;
}

// One launch of the render chain; `pdl`: with the programmatic-stream-serialization attribute (kernels/wavefront.cuh, pdl_prologue)
template <typename... KArgs, typename... Args>
void launch_chain(bool pdl, void (*kernel)(KArgs...), int grid, int block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)block); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl ? 1u : 0u;
    cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// Pair records of a small tree (kernels/traverse.cuh, traverse2p): one 128-byte record per interior node, in node order.  Returns the number
// of records, 0 when the tree is a single leaf, not a full binary tree, or too large for the shared-memory budget (`rec` is then left empty).
uint32_t build_pair_records(const arn_node* nodes, uint32_t n_nodes, std::vector<float>* rec) {
    rec->clear();
    if (n_nodes < 3 || (nodes[0].len_axis >> 2) != 0 || ((size_t)n_nodes - 1) / 2 * ARN_PAIR_BYTES > ARN_SMEM_NODE_BYTES) return 0;
    const uint32_t n_int = (n_nodes - 1) / 2;
    std::vector<uint32_t> pair_of(n_nodes, 0u);
    uint32_t k = 0;
    for (uint32_t i = 0; i < n_nodes; i++) if ((nodes[i].len_axis >> 2) == 0) pair_of[i] = k++;
    if (k != n_int) return 0;
    rec->assign((size_t)n_int * (ARN_PAIR_BYTES / 4), 0.f);
    for (uint32_t i = 0; i < n_nodes; i++) {
        const arn_node& nd = nodes[i];
        if ((nd.len_axis >> 2) != 0) continue;
        float* r = &(*rec)[(size_t)pair_of[i] * (ARN_PAIR_BYTES / 4)];
        const uint32_t ch[2] = {i + 1, i + nd.offset};
        for (int cidx = 0; cidx < 2; cidx++) {
            if (ch[cidx] >= n_nodes) { rec->clear(); return 0; }
            const arn_node& cn = nodes[ch[cidx]];
            for (int a = 0; a < 3; a++) {           // axis block: (A.min, A.max, B.min, B.max) then the same with min / max swapped
                float* q = r + a * 8 + cidx * 2;
                q[0] = cn.bmin[a]; q[1] = cn.bmax[a]; q[4] = cn.bmax[a]; q[5] = cn.bmin[a];
            }
            const bool leaf = (cn.len_axis >> 2) != 0;
            const uint32_t w0 = leaf ? cn.offset : pair_of[ch[cidx]] * ARN_PAIR_BYTES, w1 = cn.len_axis;
            for (int rep = 0; rep < 2; rep++) { std::memcpy(r + 24 + rep * 4 + cidx * 2, &w0, 4); std::memcpy(r + 25 + rep * 4 + cidx * 2, &w1, 4); }
        }
    }
    return n_int;
}

cudaEvent_t get_event(arn_ctx* c, size_t i) {
    while (c->events.size() <= i) { cudaEvent_t e; cudaEventCreate(&e); c->events.push_back(e); }
    return c->events[i];
}

int ensure_wave(arn_ctx* ctx, arn_ctx::Pipe* c, size_t cap) {
    if (c->wave_cap >= cap) return ARN_OK;
    if (c->pool) { cudaFree(c->pool); c->pool = nullptr; c->wave_cap = 0; }
    // one pool, carved into 256-byte aligned SoA streams
    size_t off = 0;
    auto carve = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    size_t o_ray = carve(cap * 32), o_beta = carve(cap * 16), o_L = carve(cap * 16);
    size_t o_pfilm = carve(cap * 8), o_pix = carve(cap * 4), o_smp = carve(cap * 4);
    size_t o_hit = carve(cap * 16);
    size_t o_sh = carve(cap * 32), o_mis = carve(cap * 32), o_nee = carve(cap * 64), o_occ = carve(cap * 4), o_mok = carve(cap * 4);
    size_t o_q0 = carve(cap * 4), o_q1 = carve(cap * 4), o_qc = carve(cap * 4), o_qs = carve(cap * 4), o_qm = carve(cap * 4);
    size_t o_cls[ARN_NCLS]; for (int k = 0; k < ARN_NCLS; k++) o_cls[k] = carve(cap * 4);
    size_t o_counts = carve(ARN_NCOUNTS * 4), o_stats = carve(64);
    CUDA_TRY(ctx, cudaMalloc(&c->pool, off));
    char* b = (char*)c->pool;
    c->pb.ray = (float4*)(b + o_ray); c->pb.beta = (float4*)(b + o_beta); c->pb.L = (float4*)(b + o_L);
    c->pb.pfilm = (float2*)(b + o_pfilm); c->pb.pix = (uint32_t*)(b + o_pix); c->pb.smp = (uint32_t*)(b + o_smp);
    c->pb.hit = (float4*)(b + o_hit);
    c->pb.sh = (float4*)(b + o_sh); c->pb.mis = (float4*)(b + o_mis); c->pb.nee = (float4*)(b + o_nee);
    c->pb.occluded = (uint32_t*)(b + o_occ); c->pb.mis_ok = (uint32_t*)(b + o_mok);
    c->q.active[0] = (uint32_t*)(b + o_q0); c->q.active[1] = (uint32_t*)(b + o_q1); c->q.connect = (uint32_t*)(b + o_qc); c->q.shadow = (uint32_t*)(b + o_qs); c->q.mis = (uint32_t*)(b + o_qm);
    for (int k = 0; k < ARN_NCLS; k++) c->q.cls[k] = (uint32_t*)(b + o_cls[k]);
    c->q.counts = (uint32_t*)(b + o_counts); c->q.stats = (unsigned long long*)(b + o_stats);
    c->wave_cap = cap;
    return ARN_OK;
}

int ensure_diff_buf(arn_ctx* ctx, arn_ctx::Pipe* c) {
    if (c->diff_cap < c->wave_cap) {
        if (c->diff_pool) { cudaFree(c->diff_pool); c->diff_pool = nullptr; c->diff_cap = 0; }
        CUDA_TRY(ctx, cudaMalloc(&c->diff_pool, c->wave_cap * 4 * sizeof(float4)));
        c->diff_cap = c->wave_cap;
    }
    c->pb.diff = (float4*)c->diff_pool;
    return ARN_OK;
}

int ensure_trace_buf(arn_ctx* ctx, arn_ctx::Pipe* c, size_t wave_cap) {
    const size_t cap = 3 * wave_cap;                      // path + shadow + light rays of one bounce
    if (c->tb_cap >= cap) return ARN_OK;
    if (c->tb_pool) { cudaFree(c->tb_pool); c->tb_pool = nullptr; c->tb_cap = 0; }
    CUDA_TRY(ctx, cudaMalloc(&c->tb_pool, cap * 6 * sizeof(float4)));
    c->tb.rs = (float4*)c->tb_pool; c->tb.res = c->tb.rs + 5 * cap; c->tb.cap = (uint32_t)cap;
    c->tb_cap = cap;
    return ARN_OK;
}

size_t wave_capacity_default(int pipes, bool big_scene, bool smem_walk, unsigned long long total) {
    const char* e = std::getenv("ARN_WAVE");
    if (e) { long v = std::atol(e); if (v >= 1024) return (size_t)v; }
    if (big_scene && pipes >= 6) return (size_t)1 << 17;
    // the shared-memory walk stages the pair records once per launch and block: larger launches amortise it (C3, 64 spp, 4 pipelines:
    // 2^19 253.2 ms, 2^20 244.2, 2^21 247.9; tools/wave_sweep.py) — as long as every pipeline still gets two waves
    if (smem_walk && total >= ((unsigned long long)pipes << 21)) return (size_t)1 << 20;
    return pipes >= 3 ? (size_t)1 << 19 : (size_t)1 << 20;
}

}  // namespace

extern "C" {

const char* arn_version(void) { return "arendur_b200 0.1 sm_100a"; }

const char* arn_last_error(const arn_ctx* ctx) {
    if (ctx) return ctx->err.c_str();
    return g_last_error.c_str();
}

int arn_ctx_create(int device, arn_ctx** out) {
    if (!out) return set_err(nullptr, ARN_E_INVALID, "arn_ctx_create: out is NULL");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return set_err(nullptr, ARN_E_CUDA, std::string("no CUDA device available (this library has no CPU fallback): ") + cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return set_err(nullptr, ARN_E_INVALID, "arn_ctx_create: device index out of range");
    arn_ctx* c = new arn_ctx;
    c->device = device;
    // any failure below releases what was created so far (arn_ctx_destroy tolerates a half-built context)
#define CTX_TRY(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { arn_ctx_destroy(c); \
    return set_err(nullptr, e_ == cudaErrorMemoryAllocation ? ARN_E_OOM : ARN_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); } } while (0)
    CTX_TRY( cudaSetDevice(device));
    cudaDeviceProp prop;
    CTX_TRY( cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) { arn_ctx_destroy(c); return set_err(nullptr, ARN_E_UNSUPPORTED, "this build targets sm_100a (B200) only; found compute capability " + std::to_string(prop.major) + "." + std::to_string(prop.minor)); }
    c->sm_count = prop.multiProcessorCount;
    CTX_TRY( cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    c->pipes[0].stream = c->stream;
    for (int i = 1; i < ARN_MAX_PIPES; i++) CTX_TRY( cudaStreamCreateWithFlags(&c->pipes[i].stream, cudaStreamNonBlocking));
    for (int i = 0; i < ARN_MAX_PIPES; i++) CTX_TRY( cudaEventCreateWithFlags(&c->pipes[i].done, cudaEventDisableTiming));
    { const char* e = std::getenv("ARN_PIPES"); if (e) { int v = std::atoi(e); if (v >= 1 && v <= ARN_MAX_PIPES) c->opt_pipes = v; } }
    c->g_generate = grid_for(c, (const void*)k_generate);
    c->g_trace = grid_for(c, (const void*)k_trace<ARN_TRAV_BINARY>);
    c->g_trace_w = grid_for(c, (const void*)k_trace<ARN_TRAV_WIDE>);
    CTX_TRY( cudaFuncSetAttribute((const void*)k_trace<ARN_TRAV_BINARY_SMEM>, cudaFuncAttributeMaxDynamicSharedMemorySize, ARN_SMEM_NODE_BYTES));
    CTX_TRY( cudaFuncSetAttribute((const void*)k_closest_batch<ARN_TRAV_BINARY_SMEM>, cudaFuncAttributeMaxDynamicSharedMemorySize, ARN_SMEM_NODE_BYTES));
    CTX_TRY( cudaFuncSetAttribute((const void*)k_any_batch<ARN_TRAV_BINARY_SMEM>, cudaFuncAttributeMaxDynamicSharedMemorySize, ARN_SMEM_NODE_BYTES));
    { const char* e = std::getenv("ARN_SMEM_NODES"); if (e) c->opt_smem_off = std::atoi(e) == 0; }
    { const char* e = std::getenv("ARN_PDL"); if (e) c->opt_pdl = std::atoi(e) != 0; }
    c->g_trace_8 = grid_for(c, (const void*)k_trace<ARN_TRAV_CW8>);
    c->g_closest_8 = grid_for(c, (const void*)k_closest_batch<ARN_TRAV_CW8>);
    c->g_any_8 = grid_for(c, (const void*)k_any_batch<ARN_TRAV_CW8>);
    c->g_shade = grid_for(c, (const void*)k_shade<SHADE_GENERIC>);
    c->g_shade_tex = grid_for(c, (const void*)k_shade<SHADE_GENERIC, true>);
    c->g_shade_p = grid_for(c, (const void*)k_shade<SHADE_PLASTIC>);
    c->g_shade_g = grid_for(c, (const void*)k_shade<SHADE_GLASS>);
    c->g_shade_d = grid_for(c, (const void*)k_shade<SHADE_DIFFUSE>);
    c->g_resolve = grid_for(c, (const void*)k_resolve);
    c->g_setup = grid_for(c, (const void*)k_ray_setup); c->g_refill = grid_for(c, (const void*)k_trace_refill); c->g_classify = grid_for(c, (const void*)k_classify);
    { const char* e = std::getenv("ARN_REFILL"); if (e) c->opt_refill = std::atoi(e) != 0; }
    c->g_accum = grid_for(c, (const void*)k_accumulate);
    c->g_accum_px = grid_for(c, (const void*)k_accumulate_px);
    c->g_closest = grid_for(c, (const void*)k_closest_batch<ARN_TRAV_BINARY>);
    c->g_closest_w = grid_for(c, (const void*)k_closest_batch<ARN_TRAV_WIDE>);
    c->g_any = grid_for(c, (const void*)k_any_batch<ARN_TRAV_BINARY>);
    c->g_any_w = grid_for(c, (const void*)k_any_batch<ARN_TRAV_WIDE>);
    CTX_TRY( cudaMalloc(&c->d_ctr, 64));
#undef CTX_TRY
    *out = c;
    return ARN_OK;
}

void arn_ctx_destroy(arn_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    for (int i = 0; i < ARN_MAX_PIPES; i++) {
        if (c->pipes[i].stream) cudaStreamSynchronize(c->pipes[i].stream);
        if (c->pipes[i].pool) cudaFree(c->pipes[i].pool);
        if (c->pipes[i].tb_pool) cudaFree(c->pipes[i].tb_pool);
        if (c->pipes[i].diff_pool) cudaFree(c->pipes[i].diff_pool);
        if (c->pipes[i].done) cudaEventDestroy(c->pipes[i].done);
        if (i > 0 && c->pipes[i].stream) cudaStreamDestroy(c->pipes[i].stream);
    }
    if (c->d_tile_rect) cudaFree(c->d_tile_rect);
    if (c->d_tile_prefix) cudaFree(c->d_tile_prefix);
    if (c->d_rays) cudaFree(c->d_rays);
    if (c->d_hits) cudaFree(c->d_hits);
    if (c->d_ctr) cudaFree(c->d_ctr);
    if (c->d_film) cudaFree(c->d_film);
    for (cudaEvent_t e : c->events) cudaEventDestroy(e);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

int arn_ctx_set_option(arn_ctx* c, int option, long long value) {
    if (!c) return ARN_E_INVALID;
    switch (option) {
    case ARN_OPT_COUNT_TRAVERSAL: c->opt_count = value != 0; return ARN_OK;
    case ARN_OPT_BVH_WIDTH: if (value != 0 && value != 2 && value != 4 && value != 8) return set_err(c, ARN_E_INVALID, "BVH width must be 0 (auto), 2, 4 or 8"); c->opt_width = (int)value; return ARN_OK;
    case ARN_OPT_PIPELINES: if (value < 0 || value > ARN_MAX_PIPES) return set_err(c, ARN_E_INVALID, "pipelines must be in 0 (auto) ..8"); c->opt_pipes = (int)value; return ARN_OK;
    case ARN_OPT_TRACE_REFILL: c->opt_refill = value != 0; return ARN_OK;
    case ARN_OPT_SMEM_NODES: c->opt_smem_off = value != 0; return ARN_OK;
    case ARN_OPT_PDL: c->opt_pdl = value != 0; return ARN_OK;
    case ARN_OPT_WAVE_CAPACITY: if (value != 0 && value < 1024) return set_err(c, ARN_E_INVALID, "wave capacity must be >= 1024"); c->opt_wave = (size_t)value; return ARN_OK;
    default: return set_err(c, ARN_E_INVALID, "unknown option");
    }
}

int arn_ctx_synchronize(arn_ctx* c) { if (!c) return ARN_E_INVALID; cudaSetDevice(c->device); CUDA_TRY(c, cudaStreamSynchronize(c->stream)); return ARN_OK; }
void* arn_ctx_stream(arn_ctx* c) { return c ? (void*)c->stream : nullptr; }

void arn_scene_destroy(arn_scene* s) {
    if (!s) return;
    cudaSetDevice(s->ctx->device);
    cudaStreamSynchronize(s->ctx->stream);
    for (void* p : s->allocs) cudaFree(p);
    delete s;
}

int arn_scene_upload(arn_ctx* c, const arn_scene_desc* d, arn_scene** out) {
    if (!c || !d || !out) return set_err(c, ARN_E_INVALID, "arn_scene_upload: NULL argument");
    const bool verbose = std::getenv("ARN_VERBOSE") != nullptr;
    auto t_start = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (!verbose) return;
        auto now = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[arn upload] %-28s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(now - t_start).count());
        t_start = now;
    };
    if (!d->n_prims || !d->prims || !d->n_nodes || !d->nodes || !d->order) return set_err(c, ARN_E_INVALID, "arn_scene_upload: scene has no primitives or no BVH");
    if (d->n_triangles && (!d->positions || !d->indices || !d->tri_mesh || !d->meshes)) return set_err(c, ARN_E_INVALID, "arn_scene_upload: triangle arrays missing");
    if (d->n_prims >= 0x80000000u) return set_err(c, ARN_E_INVALID, "arn_scene_upload: too many primitives");
    if (d->n_spheres && !d->spheres) return set_err(c, ARN_E_INVALID, "arn_scene_upload: n_spheres > 0 but spheres is NULL");
    if (d->n_materials && !d->materials) return set_err(c, ARN_E_INVALID, "arn_scene_upload: n_materials > 0 but materials is NULL");
    if (d->n_lights && (!d->light_prims || !d->light_func || !d->light_cdf)) return set_err(c, ARN_E_INVALID, "arn_scene_upload: n_lights > 0 but light_prims / light_func / light_cdf is NULL");
    if (d->n_analytic_lights && !d->analytic_lights) return set_err(c, ARN_E_INVALID, "arn_scene_upload: n_analytic_lights > 0 but analytic_lights is NULL");
    // validate references
    for (uint32_t i = 0; i < d->n_prims; i++) {
        uint32_t r = d->prims[i];
        if (r & ARN_PRIM_SPHERE) { if ((r & ~ARN_PRIM_SPHERE) >= d->n_spheres) return set_err(c, ARN_E_INVALID, "component references a missing sphere"); }
        else if (r >= d->n_triangles) return set_err(c, ARN_E_INVALID, "component references a missing triangle");
        if (d->order[i] >= d->n_prims) return set_err(c, ARN_E_INVALID, "BVH order entry out of range");
    }
    for (size_t i = 0; i < (size_t)d->n_triangles * 3; i++) if (d->indices[i] >= d->n_vertices) return set_err(c, ARN_E_INVALID, "triangle index out of range");
    for (uint32_t i = 0; i < d->n_triangles; i++) if (d->tri_mesh[i] >= d->n_meshes) return set_err(c, ARN_E_INVALID, "triangle mesh id out of range");
    for (uint32_t i = 0; i < d->n_meshes; i++) {
        if (d->meshes[i].material >= d->n_materials) return set_err(c, ARN_E_INVALID, "mesh material out of range");
        if (d->meshes[i].has_normals && !d->normals) return set_err(c, ARN_E_INVALID, "mesh claims normals but none were passed");
        if (d->meshes[i].has_uvs && !d->uvs) return set_err(c, ARN_E_INVALID, "mesh claims uvs but none were passed");
    }
    for (uint32_t i = 0; i < d->n_spheres; i++) if (d->spheres[i].material >= d->n_materials) return set_err(c, ARN_E_INVALID, "sphere material out of range");
    for (uint32_t i = 0; i < d->n_lights; i++) {
        if (d->light_prims[i] & ARN_LIGHT_ANALYTIC) {
            uint32_t k = d->light_prims[i] & ~ARN_LIGHT_ANALYTIC;
            if (k >= d->n_analytic_lights || !d->analytic_lights || d->analytic_lights[k].type > ARN_LIGHT_DISTANT) return set_err(c, ARN_E_INVALID, "light references a missing Point/Spot/Distant light");
            continue;
        }
        if (d->light_prims[i] >= d->n_prims || !(d->prims[d->light_prims[i]] & ARN_PRIM_SPHERE))
            return set_err(c, ARN_E_UNSUPPORTED, "lights must be emissive sphere primitives (triangle emitters do not work in arendur: surface_area() == 0, SURVEY.md Appendix A-2)");
    }
    if (d->n_textures && (!d->textures || !d->texels)) return set_err(c, ARN_E_INVALID, "arn_scene_upload: n_textures > 0 but textures / texels is NULL");
    for (uint32_t i = 0; i < d->n_textures; i++) {
        const arn_texture& t = d->textures[i];
        if ((t.channels != 1 && t.channels != 3) || t.n_levels < 1 || t.n_levels > ARN_TEX_MAX_LEVELS || t.wrapping > ARN_WRAP_CLAMP) return set_err(c, ARN_E_INVALID, "texture: channels must be 1 or 3, 1..16 levels, a valid wrap mode");
        for (uint32_t l = 0; l < t.n_levels; l++)
            if (!t.level_w[l] || !t.level_h[l] || (uint64_t)t.level_offset[l] + (uint64_t)t.level_w[l] * t.level_h[l] * t.channels > d->n_texel_floats) return set_err(c, ARN_E_INVALID, "texture level outside the texel array");
    }
    for (uint32_t i = 0; i < d->n_materials; i++) {
        const arn_material& m = d->materials[i];
        const uint32_t ids[4] = {m.kd_tex, m.ks_tex, m.aux_tex, m.bump_tex};
        for (int k = 0; k < 4; k++) {
            if (ids[k] > d->n_textures) return set_err(c, ARN_E_INVALID, "material references a missing texture");
            if (ids[k] && d->textures[ids[k] - 1].channels != (k < 2 ? 3u : 1u)) return set_err(c, ARN_E_INVALID, "kd / ks take RGB textures, sigma / roughness / bump Luma textures");
        }
    }
    lap("reference validation");
    // tree walk: bounds of child offsets, leaf ranges, maximum stack depth
    uint32_t max_depth = 0;
    {
        std::vector<std::pair<uint32_t, uint32_t>> st; st.push_back({0u, 1u});
        size_t visited = 0;
        while (!st.empty()) {
            auto [idx, depth] = st.back(); st.pop_back();
            if (idx >= d->n_nodes) return set_err(c, ARN_E_INVALID, "BVH child index out of range");
            if (++visited > d->n_nodes) return set_err(c, ARN_E_INVALID, "BVH is not a tree");
            max_depth = std::max(max_depth, depth);
            const arn_node& nd = d->nodes[idx];
            uint32_t len = nd.len_axis >> 2;
            if (len == 0) { if (nd.offset < 2) return set_err(c, ARN_E_INVALID, "BVH interior offset invalid"); st.push_back({idx + nd.offset, depth + 1}); st.push_back({idx + 1, depth + 1}); }
            else if ((uint64_t)nd.offset + len > d->n_prims) return set_err(c, ARN_E_INVALID, "BVH leaf range out of bounds");
        }
    }
    if (max_depth > ARN_STACK) return set_err(c, ARN_E_UNSUPPORTED, "BVH deeper than the traversal stack (" + std::to_string(max_depth) + " > " + std::to_string(ARN_STACK) + ")");

    lap("tree walk");
    cudaSetDevice(c->device);
    arn_scene* s = new arn_scene; s->ctx = c; s->max_depth = max_depth;
    std::vector<void*> scratch;          // device scratch of this upload, freed after its final sync
    auto fail = [&](int rc) { cudaStreamSynchronize(c->stream); for (void* p : scratch) cudaFree(p); arn_scene_destroy(s); return rc; };
    int rc;
    // one device allocation for the whole scene (and one for the upload's scratch): 2 cudaMalloc instead of 20
    const size_t n_interior_all = ((size_t)d->n_nodes - 1) / 2;
    {
        size_t total = align256((size_t)d->n_nodes * sizeof(arn_node)) + align256(n_interior_all * 4 * sizeof(arn_node)) + align256((size_t)d->n_prims * 48)
                     + align256((size_t)d->n_spheres * sizeof(arn_sphere)) + align256((size_t)d->n_triangles * 12) + align256((size_t)d->n_vertices * 12) * 2
                     + align256((size_t)d->n_vertices * 8) + align256((size_t)d->n_triangles * 4) + align256((size_t)d->n_meshes * sizeof(arn_mesh))
                     + align256((size_t)d->n_materials * sizeof(arn_material)) + align256((size_t)d->n_prims * 4) + align256((size_t)d->n_lights * 4) * 2
                     + align256((size_t)d->n_analytic_lights * sizeof(arn_analytic_light)) + align256(((size_t)d->n_lights + 1) * 4) + 4096 + align256(ARN_SMEM_NODE_BYTES)
                     + align256((size_t)d->n_textures * sizeof(arn_texture)) + align256((size_t)d->n_texel_floats * 4);
        void* pool = nullptr;
        cudaError_t ce = cudaMalloc(&pool, total);
        if (ce != cudaSuccess) { set_err(c, ARN_E_OOM, std::string("scene memory: ") + cudaGetErrorString(ce)); return fail(ARN_E_OOM); }
        s->allocs.push_back(pool); s->pool = (char*)pool; s->pool_size = total;
        size_t sbytes = align256(n_interior_all * sizeof(uint2)) * 2 + align256((ARN_STACK + 4) * sizeof(uint32_t)) + align256((size_t)d->n_prims * 4);
        void* sp = nullptr;
        if (cudaMalloc(&sp, sbytes) != cudaSuccess) { set_err(c, ARN_E_OOM, "upload scratch: out of device memory"); return fail(ARN_E_OOM); }
        scratch.push_back(sp);
    }
    char* scratch_p = (char*)scratch[0];
    auto scratch_take = [&](size_t bytes) { char* r = scratch_p; scratch_p += align256(bytes); return (void*)r; };
    // nodes: same 32-byte records, read on the device as float4 pairs
    const arn_node* dn = nullptr;
    if ((rc = dev_upload(s, d->nodes, d->n_nodes, &dn)) != ARN_OK) return fail(rc);
    s->dev.nodes = (const float4*)dn;
    lap("node upload");
    s->dev.pairs = nullptr; s->dev.n_pairs = 0; s->dev.root_axis = 0;
    std::vector<float> rec;                             // staging of the pair records: lives until the upload's final sync
    {
        const uint32_t n_int = build_pair_records(d->nodes, d->n_nodes, &rec);
        if (n_int) {
            const float4* dp = nullptr;
            if ((rc = dev_upload(s, (const float4*)rec.data(), rec.size() / 4, &dp)) != ARN_OK) return fail(rc);
            s->dev.pairs = dp; s->dev.n_pairs = n_int; s->dev.root_axis = d->nodes[0].len_axis & 3u;
        }
    }
    {   // 4-wide collapse on the device (kernels/wide_build.cuh; layout: kernels/traverse.cuh, traverse4)
        auto is_leaf = [&](uint32_t i) { return (d->nodes[i].len_axis >> 2) != 0; };
        arn_node root = d->nodes[0];
        if (is_leaf(0)) root.len_axis = ((root.len_axis >> 2) << 8) | ARN_W_LEAF;
        else {
            uint32_t a = 1, b = d->nodes[0].offset;
            uint32_t axes = (d->nodes[0].len_axis & 3u) | ((is_leaf(a) ? 0u : (d->nodes[a].len_axis & 3u)) << 2) | ((is_leaf(b) ? 0u : (d->nodes[b].len_axis & 3u)) << 4);
            root.offset = 0; root.len_axis = (axes << 2) | ARN_W_INNER;
        }
        std::memcpy(&s->dev.root0, &root, 16); std::memcpy(&s->dev.root1, (const char*)&root + 16, 16);
        auto amax = [](float a, float b) { a = std::fabs(a); b = std::fabs(b); return a > b ? a : b; };
        s->dev.absmax = make_float3(amax(root.bmin[0], root.bmax[0]), amax(root.bmin[1], root.bmax[1]), amax(root.bmin[2], root.bmax[2]));
        const size_t n_interior = ((size_t)d->n_nodes - 1) / 2;        // full binary tree; every wide node is rooted at an interior node
        if (n_interior > 0) {
            arn_node* d_wide = (arn_node*)pool_take(s, n_interior * 4 * sizeof(arn_node));
            if (!d_wide) { set_err(c, ARN_E_OOM, "scene pool exhausted (wide nodes)"); return fail(ARN_E_OOM); }
            uint2* f0 = (uint2*)scratch_take(n_interior * sizeof(uint2)); uint2* f1 = (uint2*)scratch_take(n_interior * sizeof(uint2));
            uint32_t* d_counts = (uint32_t*)scratch_take((ARN_STACK + 4) * sizeof(uint32_t));
            uint32_t* d_wide_count = d_counts + ARN_STACK + 2;
            k_wide_begin<<<1, 1, 0, c->stream>>>(f0, d_counts, d_wide_count, 1);
            int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device);
            const int grid = (int)std::min<size_t>((size_t)sms * 8, (n_interior + 255) / 256);
            const int levels = (int)(max_depth + 1) / 2 + 1;
            for (int level = 0; level < levels; level++)
                k_wide_level<<<grid, 256, 0, c->stream>>>(dn, d_wide, (level & 1) ? f1 : f0, (level & 1) ? f0 : f1, d_counts, level, d_wide_count);
            s->dev.wide = (const float4*)d_wide;
        }
    }
    lap("wide collapse + upload");
    if ((rc = dev_upload(s, d->spheres, d->n_spheres, &s->dev.spheres)) != ARN_OK) return fail(rc);
    if ((rc = dev_upload(s, d->indices, (size_t)d->n_triangles * 3, &s->dev.indices)) != ARN_OK) return fail(rc);
    if ((rc = dev_upload(s, d->positions, (size_t)d->n_vertices * 3, &s->dev.positions)) != ARN_OK) return fail(rc);
    if ((rc = dev_upload(s, d->prims, d->n_prims, &s->dev.prims)) != ARN_OK) return fail(rc);
    {   // ordered 48-byte primitive slots, gathered on the device from the arrays just uploaded
        uint32_t* d_order = (uint32_t*)scratch_take((size_t)d->n_prims * 4);
        float4* d_slots = (float4*)pool_take(s, (size_t)d->n_prims * 48);
        if (!d_slots) { set_err(c, ARN_E_OOM, "scene pool exhausted (primitive slots)"); return fail(ARN_E_OOM); }
        cudaMemcpyAsync(d_order, d->order, (size_t)d->n_prims * 4, cudaMemcpyHostToDevice, c->stream);
        int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device);
        const int grid = (int)std::min<size_t>((size_t)sms * 8, ((size_t)d->n_prims + 255) / 256);
        k_build_slots<<<grid, 256, 0, c->stream>>>(d_order, s->dev.prims, s->dev.indices, s->dev.positions, d_slots, d->n_prims);
        s->dev.tris = d_slots;
    }
    lap("slots (device gather)");
    // compressed 8-wide nodes + leaf blob (kernels/cw8_build.cuh) for trees the 4-wide records no longer keep in the caches
    // — only on request (ARN_OPT_BVH_WIDTH = 8 before the upload): on C4 the 8-wide walk moves 21 % fewer DRAM bytes but issues 63 % more
    // instructions and is 20 % slower than the 4-wide walk (profiles/r02_c4_cw8.txt), so `auto` never picks it
    if (d->n_nodes >= 3 && (d->nodes[0].len_axis >> 2) == 0 && c->opt_width == 8) {
        const size_t n_nodes = d->n_nodes, n_leaves = (n_nodes + 1) / 2;
        uint32_t* d_sizes = nullptr; uint32_t* d_off = nullptr; void* d_tmp = nullptr; uint2 *f0 = nullptr, *f1 = nullptr; uint32_t* d_cnt = nullptr;
        size_t tmp_bytes = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_sizes, d_off, (int)n_nodes, c->stream);
        auto salloc = [&](size_t bytes) -> void* { void* p = nullptr; if (cudaMalloc(&p, bytes ? bytes : 1) != cudaSuccess) return nullptr; scratch.push_back(p); return p; };
        d_sizes = (uint32_t*)salloc(n_nodes * 4); d_off = (uint32_t*)salloc(n_nodes * 4); d_tmp = salloc(tmp_bytes);
        f0 = (uint2*)salloc(n_interior_all * sizeof(uint2)); f1 = (uint2*)salloc(n_interior_all * sizeof(uint2)); d_cnt = (uint32_t*)salloc((ARN_STACK + 4) * 4);
        if (!d_sizes || !d_off || !d_tmp || !f0 || !f1 || !d_cnt) { cudaGetLastError(); set_err(c, ARN_E_OOM, "8-wide collapse scratch: out of device memory"); return fail(ARN_E_OOM); }
        int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device);
        const int grid = (int)std::min<size_t>((size_t)sms * 8, (n_nodes + 255) / 256);
        k_cw8_leaf_sizes<<<grid, 256, 0, c->stream>>>(dn, (uint32_t)n_nodes, d_sizes);
        cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, d_sizes, d_off, (int)n_nodes, c->stream);
        const size_t blob_f4 = 2 * n_leaves + 3 * (size_t)d->n_prims;
        const int levels = (int)(max_depth + 2) / 3 + 2;
        uint32_t* d_node_count = d_cnt + ARN_STACK + 2;
        // counting pass, then the exact allocation
        k_cw8_begin<<<1, 1, 0, c->stream>>>(f0, d_cnt, d_node_count);
        for (int level = 0; level < levels; level++)
            k_cw8_level<<<grid, 256, 0, c->stream>>>(dn, d_off, nullptr, (level & 1) ? f1 : f0, (level & 1) ? f0 : f1, d_cnt, level, d_node_count);
        uint32_t n_cw8 = 0;
        cudaMemcpyAsync(&n_cw8, d_node_count, 4, cudaMemcpyDeviceToHost, c->stream);
        if (cudaStreamSynchronize(c->stream) != cudaSuccess) { set_err(c, ARN_E_CUDA, "8-wide collapse (count) failed"); return fail(ARN_E_CUDA); }
        if (n_cw8 < (1u << 24) && blob_f4 < 0xffffffffull) {          // node index + leaf mask share a word; leaf references are 32-bit float4 indices
            void* extra = nullptr;
            if (cudaMalloc(&extra, (size_t)n_cw8 * 128 + blob_f4 * 16) != cudaSuccess) { cudaGetLastError(); set_err(c, ARN_E_OOM, "8-wide nodes: out of device memory"); return fail(ARN_E_OOM); }
            s->allocs.push_back(extra);
            uint4* d_cw8 = (uint4*)extra; float4* d_blob = (float4*)((char*)extra + (size_t)n_cw8 * 128);
            k_cw8_leaf_blob<<<grid, 256, 0, c->stream>>>(dn, (uint32_t)n_nodes, d_off, s->dev.tris, d_blob);
            k_cw8_begin<<<1, 1, 0, c->stream>>>(f0, d_cnt, d_node_count);
            for (int level = 0; level < levels; level++)
                k_cw8_level<<<grid, 256, 0, c->stream>>>(dn, d_off, d_cw8, (level & 1) ? f1 : f0, (level & 1) ? f0 : f1, d_cnt, level, d_node_count);
            s->dev.cw8 = d_cw8; s->dev.blob = d_blob;
            s->bytes += (size_t)n_cw8 * 128 + blob_f4 * 16;
        }
        lap("8-wide collapse + leaf blob");
    }
    if ((rc = dev_upload(s, d->normals, d->normals ? (size_t)d->n_vertices * 3 : 0, &s->dev.normals)) != ARN_OK) return fail(rc);
    if ((rc = dev_upload(s, d->uvs, d->uvs ? (size_t)d->n_vertices * 2 : 0, &s->dev.uvs)) != ARN_OK) return fail(rc);
    if ((rc = dev_upload(s, d->tri_mesh, d->n_triangles, &s->dev.tri_mesh)) != ARN_OK) return fail(rc);
    if ((rc = dev_upload(s, d->meshes, d->n_meshes, &s->dev.meshes)) != ARN_OK) return fail(rc);
    if ((rc = dev_upload(s, d->materials, d->n_materials, &s->dev.materials)) != ARN_OK) return fail(rc);
    if ((rc = dev_upload(s, d->light_prims, d->n_lights, &s->dev.light_prims)) != ARN_OK) return fail(rc);
    if ((rc = dev_upload(s, d->analytic_lights, d->n_analytic_lights, &s->dev.analytic)) != ARN_OK) return fail(rc);
    if ((rc = dev_upload(s, d->light_func, d->n_lights, &s->dev.light_func)) != ARN_OK) return fail(rc);
    if ((rc = dev_upload(s, d->light_cdf, d->n_lights ? d->n_lights + 1 : 0, &s->dev.light_cdf)) != ARN_OK) return fail(rc);
    if (d->n_textures) {
        if ((rc = dev_upload(s, d->textures, d->n_textures, &s->dev.textures)) != ARN_OK) return fail(rc);
        if ((rc = dev_upload(s, d->texels, (size_t)d->n_texel_floats, &s->dev.texels)) != ARN_OK) return fail(rc);
        s->dev.n_textures = d->n_textures;
    }
    for (uint32_t i = 0; i < d->n_materials; i++) {
        const arn_material& m = d->materials[i];
        int cls = m.type == ARN_MAT_MATTE ? (!(m.sigma >= 0.f && m.sigma != 0.f) && !(m.sigma != m.sigma) ? 0 : 1) : (m.type == ARN_MAT_PLASTIC ? 2 : (m.type == ARN_MAT_GLASS ? 3 : 4));
        s->class_mask |= 1u << cls;
    }
    s->dev.light_integral = d->light_func_integral;
    s->dev.n_lights = d->n_lights; s->dev.n_nodes = d->n_nodes; s->dev.n_prims = d->n_prims; s->dev.n_spheres = d->n_spheres;
    lap("shading arrays upload");
    cudaError_t e = cudaStreamSynchronize(c->stream);     // the staging vector `slots` dies here
    if (e == cudaSuccess) e = cudaGetLastError();
    for (void* p : scratch) cudaFree(p);
    scratch.clear();
    lap("sync");
    if (e != cudaSuccess) { set_err(c, ARN_E_CUDA, std::string("scene upload: ") + cudaGetErrorString(e)); return fail(ARN_E_CUDA); }
    *out = s;
    return ARN_OK;
}

// Which tree the product kernels walk.  Both give the same bits (traverse.cuh); the 4-wide walk halves
// the chain of dependent node fetches and wins once the tree no longer sits in L1 (C4: +10 %), the
// binary walk issues fewer instructions per node and wins on cache-resident trees (Cornell: +15 %).
#define ARN_WIDE_MIN_NODES (1u << 16)
#define ARN_SMEM_BATCH_MIN ((size_t)1 << 15)      /* rays: below this a batched query does not amortise the copy of the pair records */
static bool use_cw8(const arn_scene* s) {        // compressed 8-wide walk: trees far larger than the caches (needs the 8-wide nodes, built at upload)
    int w = s->ctx->opt_width;
    return s->dev.cw8 != nullptr && w == 8;
}
static bool use_wide(const arn_scene* s) {
    int w = s->ctx->opt_width;
    if (use_cw8(s)) return false;
    return w == 4 || ((w == 0 || w == 8) && s->dev.n_nodes >= ARN_WIDE_MIN_NODES);
}

// ---------------------------------------------------------------- batched queries
int arn_intersect_closest_dev(arn_scene* s, const void* rays_dev, size_t n, void* hits_dev, arn_stats* stats) {
    if (!s || (n && (!rays_dev || !hits_dev))) return set_err(s ? s->ctx : nullptr, ARN_E_INVALID, "arn_intersect_closest_dev: NULL argument");
    arn_ctx* c = s->ctx; std::lock_guard<std::recursive_mutex> g(c->mu); cudaSetDevice(c->device);
    if (n == 0) return ARN_OK;
    const bool wide = use_wide(s), cw8 = use_cw8(s);
    int grid = (int)std::min<size_t>((size_t)(cw8 ? c->g_closest_8 : wide ? c->g_closest_w : c->g_closest), (n + ARN_BLOCK - 1) / ARN_BLOCK);
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (stats) { e0 = get_event(c, 0); e1 = get_event(c, 1); cudaEventRecord(e0, c->stream); }
    // batches large enough to pay for staging the pair records of a small tree take the shared-memory walk, like k_trace
    const bool smem = !c->opt_smem_off && !wide && !cw8 && s->dev.pairs != nullptr && n >= ARN_SMEM_BATCH_MIN;
    if (smem) k_closest_batch<ARN_TRAV_BINARY_SMEM><<<(int)std::min<size_t>((size_t)c->sm_count, (n + ARN_BLOCK_SMEM - 1) / ARN_BLOCK_SMEM), ARN_BLOCK_SMEM, (size_t)s->dev.n_pairs * ARN_PAIR_BYTES, c->stream>>>(s->dev, (const arn_ray*)rays_dev, n, (arn_hit*)hits_dev, nullptr);
    else if (cw8) k_closest_batch<ARN_TRAV_CW8><<<grid, ARN_BLOCK, 0, c->stream>>>(s->dev, (const arn_ray*)rays_dev, n, (arn_hit*)hits_dev, nullptr);
    else if (wide) k_closest_batch<ARN_TRAV_WIDE><<<grid, ARN_BLOCK, 0, c->stream>>>(s->dev, (const arn_ray*)rays_dev, n, (arn_hit*)hits_dev, nullptr);
    else k_closest_batch<ARN_TRAV_BINARY><<<grid, ARN_BLOCK, 0, c->stream>>>(s->dev, (const arn_ray*)rays_dev, n, (arn_hit*)hits_dev, nullptr);
    CUDA_TRY(c, cudaGetLastError());
    if (stats) {
        cudaEventRecord(e1, c->stream);
        CUDA_TRY(c, cudaEventSynchronize(e1));
        float ms = 0.f; cudaEventElapsedTime(&ms, e0, e1);
        std::memset(stats, 0, sizeof *stats);
        stats->extend_rays = n; stats->kernel_launches = 1; stats->gpu_ms = ms; stats->extend_ms = ms;
    }
    return ARN_OK;
}
int arn_intersect_any_dev(arn_scene* s, const void* rays_dev, size_t n, void* out_dev, arn_stats* stats) {
    if (!s || (n && (!rays_dev || !out_dev))) return set_err(s ? s->ctx : nullptr, ARN_E_INVALID, "arn_intersect_any_dev: NULL argument");
    arn_ctx* c = s->ctx; std::lock_guard<std::recursive_mutex> g(c->mu); cudaSetDevice(c->device);
    if (n == 0) return ARN_OK;
    const bool wide = use_wide(s), cw8 = use_cw8(s);
    int grid = (int)std::min<size_t>((size_t)(cw8 ? c->g_any_8 : wide ? c->g_any_w : c->g_any), (n + ARN_BLOCK - 1) / ARN_BLOCK);
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (stats) { e0 = get_event(c, 0); e1 = get_event(c, 1); cudaEventRecord(e0, c->stream); }
    const bool smem = !c->opt_smem_off && !wide && !cw8 && s->dev.pairs != nullptr && n >= ARN_SMEM_BATCH_MIN;
    if (smem) k_any_batch<ARN_TRAV_BINARY_SMEM><<<(int)std::min<size_t>((size_t)c->sm_count, (n + ARN_BLOCK_SMEM - 1) / ARN_BLOCK_SMEM), ARN_BLOCK_SMEM, (size_t)s->dev.n_pairs * ARN_PAIR_BYTES, c->stream>>>(s->dev, (const arn_ray*)rays_dev, n, (uint8_t*)out_dev);
    else if (cw8) k_any_batch<ARN_TRAV_CW8><<<grid, ARN_BLOCK, 0, c->stream>>>(s->dev, (const arn_ray*)rays_dev, n, (uint8_t*)out_dev);
    else if (wide) k_any_batch<ARN_TRAV_WIDE><<<grid, ARN_BLOCK, 0, c->stream>>>(s->dev, (const arn_ray*)rays_dev, n, (uint8_t*)out_dev);
    else k_any_batch<ARN_TRAV_BINARY><<<grid, ARN_BLOCK, 0, c->stream>>>(s->dev, (const arn_ray*)rays_dev, n, (uint8_t*)out_dev);
    CUDA_TRY(c, cudaGetLastError());
    if (stats) {
        cudaEventRecord(e1, c->stream);
        CUDA_TRY(c, cudaEventSynchronize(e1));
        float ms = 0.f; cudaEventElapsedTime(&ms, e0, e1);
        std::memset(stats, 0, sizeof *stats);
        stats->shadow_rays = n; stats->kernel_launches = 1; stats->gpu_ms = ms;
    }
    return ARN_OK;
}

static int ensure_ray_scratch(arn_ctx* c, size_t n) {
    if (c->rays_cap >= n) return ARN_OK;
    if (c->d_rays) cudaFree(c->d_rays);
    if (c->d_hits) cudaFree(c->d_hits);
    c->d_rays = c->d_hits = nullptr; c->rays_cap = 0;
    CUDA_TRY(c, cudaMalloc(&c->d_rays, n * sizeof(arn_ray)));
    CUDA_TRY(c, cudaMalloc(&c->d_hits, n * sizeof(arn_hit)));
    c->rays_cap = n;
    return ARN_OK;
}
int arn_intersect_closest(arn_scene* s, const arn_ray* rays, size_t n, arn_hit* hits) {
    if (!s || (n && (!rays || !hits))) return set_err(s ? s->ctx : nullptr, ARN_E_INVALID, "arn_intersect_closest: NULL argument");
    arn_ctx* c = s->ctx; std::lock_guard<std::recursive_mutex> g(c->mu); cudaSetDevice(c->device);
    if (n == 0) return ARN_OK;
    int rc = ensure_ray_scratch(c, n); if (rc != ARN_OK) return rc;
    CUDA_TRY(c, cudaMemcpyAsync(c->d_rays, rays, n * sizeof(arn_ray), cudaMemcpyHostToDevice, c->stream));
    rc = arn_intersect_closest_dev(s, c->d_rays, n, c->d_hits, nullptr); if (rc != ARN_OK) return rc;
    CUDA_TRY(c, cudaMemcpyAsync(hits, c->d_hits, n * sizeof(arn_hit), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return ARN_OK;
}
int arn_intersect_any(arn_scene* s, const arn_ray* rays, size_t n, uint8_t* out) {
    if (!s || (n && (!rays || !out))) return set_err(s ? s->ctx : nullptr, ARN_E_INVALID, "arn_intersect_any: NULL argument");
    arn_ctx* c = s->ctx; std::lock_guard<std::recursive_mutex> g(c->mu); cudaSetDevice(c->device);
    if (n == 0) return ARN_OK;
    int rc = ensure_ray_scratch(c, n); if (rc != ARN_OK) return rc;
    CUDA_TRY(c, cudaMemcpyAsync(c->d_rays, rays, n * sizeof(arn_ray), cudaMemcpyHostToDevice, c->stream));
    rc = arn_intersect_any_dev(s, c->d_rays, n, c->d_hits, nullptr); if (rc != ARN_OK) return rc;
    CUDA_TRY(c, cudaMemcpyAsync(out, c->d_hits, n, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return ARN_OK;
}

// Instrumented closest-hit pass: counters_out[0..2] = nodes tested, triangles tested, spheres
// tested, summed over the batch (the algorithmic-bytes figure of SURVEY.md §8(d)).  DEVICE buffers.
int arn_intersect_closest_counted_dev(arn_scene* s, const void* rays_dev, size_t n, void* hits_dev, uint64_t* counters_out) {
    if (!s || !rays_dev || !hits_dev || !counters_out) return set_err(s ? s->ctx : nullptr, ARN_E_INVALID, "arn_intersect_closest_counted_dev: NULL argument");
    arn_ctx* c = s->ctx; std::lock_guard<std::recursive_mutex> g(c->mu); cudaSetDevice(c->device);
    CUDA_TRY(c, cudaMemsetAsync(c->d_ctr, 0, 64, c->stream));
    int grid = (int)std::min<size_t>((size_t)c->g_closest, (n + ARN_BLOCK - 1) / ARN_BLOCK);
    if (grid < 1) grid = 1;
    k_closest_batch<ARN_TRAV_COUNTED><<<grid, ARN_BLOCK, 0, c->stream>>>(s->dev, (const arn_ray*)rays_dev, n, (arn_hit*)hits_dev, c->d_ctr);
    CUDA_TRY(c, cudaGetLastError());
    unsigned long long h[4];
    CUDA_TRY(c, cudaMemcpyAsync(h, c->d_ctr, 32, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    counters_out[0] = h[0]; counters_out[1] = h[1]; counters_out[2] = h[2];
    if (std::getenv("ARN_PROBE")) std::fprintf(stderr, "[arn probe] nodes %llu, sum of per-warp max %llu -> lane utilisation %.3f\n", h[0], h[3], (double)h[0] / (32.0 * (double)h[3]));
    return ARN_OK;
}

// ---------------------------------------------------------------- device BVH build (kernels/lbvh.cuh)
int arn_bvh_build_gpu(arn_ctx* c, uint32_t n, const float* bounds6, arn_node* nodes_out, uint32_t* order_out, uint32_t* n_nodes_out, float* build_ms_out) {
    if (!c || !bounds6 || !nodes_out || !order_out || !n_nodes_out) return set_err(c, ARN_E_INVALID, "arn_bvh_build_gpu: NULL argument");
    if (n == 0) return set_err(c, ARN_E_INVALID, "arn_bvh_build_gpu: no components (recursive_build asserts len != 0)");
    if (n >= 0x40000000u) return set_err(c, ARN_E_INVALID, "arn_bvh_build_gpu: too many components");
    std::lock_guard<std::recursive_mutex> g(c->mu); cudaSetDevice(c->device);
    const size_t total = 2 * (size_t)n - 1;
    std::vector<void*> allocs;
    auto cleanup = [&]() { for (void* p : allocs) cudaFree(p); };
    auto dalloc = [&](size_t bytes) -> void* { void* p = nullptr; if (cudaMalloc(&p, bytes ? bytes : 1) != cudaSuccess) return nullptr; allocs.push_back(p); return p; };
    float* d_bounds = (float*)dalloc((size_t)n * 24);
    unsigned long long* d_keys = (unsigned long long*)dalloc((size_t)n * 8); unsigned long long* d_keys2 = (unsigned long long*)dalloc((size_t)n * 8);
    uint32_t* d_vals = (uint32_t*)dalloc((size_t)n * 4); uint32_t* d_vals2 = (uint32_t*)dalloc((size_t)n * 4);
    int* d_cb = (int*)dalloc(32);
    uint32_t* d_left = (uint32_t*)dalloc((size_t)n * 4); uint32_t* d_right = (uint32_t*)dalloc((size_t)n * 4);
    uint32_t* d_axis = (uint32_t*)dalloc((size_t)n * 4); uint32_t* d_flag = (uint32_t*)dalloc((size_t)n * 4);
    uint32_t* d_parent = (uint32_t*)dalloc(total * 4); uint32_t* d_size = (uint32_t*)dalloc(total * 4);
    float* d_bb = (float*)dalloc(total * 24);
    arn_node* d_nodes = (arn_node*)dalloc(total * sizeof(arn_node));
    uint32_t* d_order = (uint32_t*)dalloc((size_t)n * 4);
    size_t temp_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, d_keys, d_keys2, d_vals, d_vals2, (int)n, 0, 63, c->stream);
    void* d_temp = dalloc(temp_bytes);
    if (!d_bounds || !d_keys || !d_keys2 || !d_vals || !d_vals2 || !d_cb || !d_left || !d_right || !d_axis || !d_flag || !d_parent || !d_size || !d_bb || !d_nodes || !d_order || !d_temp) {
        cleanup(); cudaGetLastError(); return set_err(c, ARN_E_CUDA, "arn_bvh_build_gpu: out of device memory");
    }
    auto fail = [&](cudaError_t e, const char* what) { cleanup(); return set_err(c, ARN_E_CUDA, std::string("arn_bvh_build_gpu: ") + what + ": " + cudaGetErrorString(e)); };
    cudaError_t e;
    if ((e = cudaMemcpyAsync(d_bounds, bounds6, (size_t)n * 24, cudaMemcpyHostToDevice, c->stream)) != cudaSuccess) return fail(e, "bounds upload");
    cudaEvent_t e0 = get_event(c, 0), e1 = get_event(c, 1);
    cudaEventRecord(e0, c->stream);
    LbvhBuf b; b.bounds = d_bounds; b.keys = d_keys; b.vals = d_vals; b.cbounds = d_cb; b.left = d_left; b.right = d_right; b.parent = d_parent;
    b.size = d_size; b.axis = d_axis; b.flag = d_flag; b.bb = d_bb; b.n = n;
    int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device);
    const int grid = (int)std::min<size_t>((size_t)sms * 8, (total + 255) / 256);
    k_lbvh_init<<<1, 32, 0, c->stream>>>(b);
    k_lbvh_bounds<<<grid, 256, 0, c->stream>>>(b);
    k_lbvh_keys<<<grid, 256, 0, c->stream>>>(b);
    if ((e = cub::DeviceRadixSort::SortPairs(d_temp, temp_bytes, d_keys, d_keys2, d_vals, d_vals2, (int)n, 0, 63, c->stream)) != cudaSuccess) return fail(e, "radix sort");
    b.keys = d_keys2; b.vals = d_vals2;
    k_lbvh_hierarchy<<<grid, 256, 0, c->stream>>>(b, d_keys2);
    k_lbvh_refit<<<grid, 256, 0, c->stream>>>(b);
    k_lbvh_emit<<<grid, 256, 0, c->stream>>>(b, d_nodes, d_order);
    cudaEventRecord(e1, c->stream);
    if ((e = cudaGetLastError()) != cudaSuccess) return fail(e, "kernel launch");
    if ((e = cudaMemcpyAsync(nodes_out, d_nodes, total * sizeof(arn_node), cudaMemcpyDeviceToHost, c->stream)) != cudaSuccess) return fail(e, "node download");
    if ((e = cudaMemcpyAsync(order_out, d_order, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream)) != cudaSuccess) return fail(e, "order download");
    if ((e = cudaStreamSynchronize(c->stream)) != cudaSuccess) return fail(e, "build");
    if (build_ms_out) { float ms = 0.f; cudaEventElapsedTime(&ms, e0, e1); *build_ms_out = ms; }
    *n_nodes_out = (uint32_t)total;
    cleanup();
    return ARN_OK;
}

// ---------------------------------------------------------------- path tracer
static int render_pt_impl(arn_scene* s, const arn_camera* cam, const arn_film* film, const arn_sampler* smp,
                          const arn_pt_params* prm, void* film_dev, arn_stats* stats, float4* radiance_dev) {
    if (!s || !cam || !film || !smp || !prm || !film_dev) return set_err(s ? s->ctx : nullptr, ARN_E_INVALID, "arn_render_pt: NULL argument");
    arn_ctx* c = s->ctx; cudaSetDevice(c->device);
    if (s->dev.n_lights == 0) return set_err(c, ARN_E_INVALID, "arn_render_pt: the scene has no lights (the reference indexes lights[0] and panics, renderer/scene.rs:53-55)");
    int cw = film->crop_max_x - film->crop_min_x, chh = film->crop_max_y - film->crop_min_y;
    if (cw <= 0 || chh <= 0) return set_err(c, ARN_E_INVALID, "arn_render_pt: empty crop window");
    if (film->crop_max_x > 65535 || film->crop_max_y > 65535 || film->crop_min_x < 0 || film->crop_min_y < 0) return set_err(c, ARN_E_UNSUPPORTED, "arn_render_pt: crop window must lie in [0, 65535]");
    uint32_t spp = smp->sampledx * smp->sampledy;
    uint32_t s0 = prm->spp_begin, s1 = prm->spp_end ? prm->spp_end : spp;
    if (s1 <= s0) return set_err(c, ARN_E_INVALID, "arn_render_pt: empty sample range");
    if (smp->mode > ARN_SAMPLER_STRATIFIED) return set_err(c, ARN_E_INVALID, "arn_render_pt: unknown sampler mode");
    if (smp->mode == ARN_SAMPLER_STRATIFIED && s1 > spp) return set_err(c, ARN_E_INVALID, "arn_render_pt: the stratified sampler has sampledx * sampledy samples per pixel");
    if (prm->max_depth == 0 || prm->max_depth > 80) return set_err(c, ARN_E_INVALID, "arn_render_pt: max_depth must be in 1..80");
    uint32_t world = prm->world_size ? prm->world_size : 1;
    if (prm->rank >= world) return set_err(c, ARN_E_INVALID, "arn_render_pt: rank >= world_size");
    // Film::spawn_tiles(nx, ny) (filming/film.rs:104-135), ix-major order; this rank takes the tiles with (ix + iy) % world == rank
    // (diagonal interleave: every rank gets a share of every row and column of the picture, so the costly regions are spread)
    long nx = prm->tiles_x ? prm->tiles_x : 16, ny = prm->tiles_y ? prm->tiles_y : 16;
    long dx = cw / nx, dy = chh / ny;
    if (dx <= 0 || dy <= 0) return set_err(c, ARN_E_INVALID, "arn_render_pt: crop window smaller than the tile grid (spawn_tiles divides by zero)");
    long lastx = dx + cw % dx, lasty = dy + chh % dy;
    std::vector<int4> rects; std::vector<unsigned long long> prefix; prefix.push_back(0);
    for (long ix = 0, t = 0; ix < nx; ix++) for (long iy = 0; iy < ny; iy++, t++) {
        long cdx = ix == nx - 1 ? lastx : dx, cdy = iy == ny - 1 ? lasty : dy;
        // sic: tiles are laid out from (0, 0), not from crop.pmin (film.rs:118-121); a tile that, grown by the filter
        // radius, does not meet the crop window makes the reference panic (`.intersect(..).unwrap()`, film.rs:129)
        {
            long rx = (long)film->filter_radius_x, ry = (long)film->filter_radius_y;
            long gx0 = std::max(ix * dx - rx, (long)film->crop_min_x), gx1 = std::min(ix * dx + cdx + rx, (long)film->crop_max_x);
            long gy0 = std::max(iy * dy - ry, (long)film->crop_min_y), gy1 = std::min(iy * dy + cdy + ry, (long)film->crop_max_y);
            if (gx0 > gx1 || gy0 > gy1) return set_err(c, ARN_E_INVALID, "arn_render_pt: a film tile does not meet the crop window (Film::spawn_tiles lays tiles out from (0,0) and panics on this, film.rs:118-129)");
        }
        // ranks own whole tiles, or (partition_subdiv = k > 1) the k x k cells of every tile: finer interleave, same film
        const long sub = prm->partition_subdiv > 1 ? (long)prm->partition_subdiv : 1;
        const long ncx = std::min(sub, cdx), ncy = std::min(sub, cdy), sdx = cdx / ncx, sdy = cdy / ncy;
        for (long jx = 0; jx < ncx; jx++) for (long jy = 0; jy < ncy; jy++) {
            if ((uint32_t)((ix * sub + jx + iy * sub + jy) % (long)world) != prm->rank) continue;
            long w = jx == ncx - 1 ? cdx - jx * sdx : sdx, h = jy == ncy - 1 ? cdy - jy * sdy : sdy;
            rects.push_back(make_int4((int)(ix * dx + jx * sdx), (int)(iy * dy + jy * sdy), (int)w, (int)h));
            prefix.push_back(prefix.back() + (unsigned long long)w * (unsigned long long)h);
        }
    }
    std::lock_guard<std::recursive_mutex> guard(c->mu);
    if (rects.empty()) { if (stats) std::memset(stats, 0, sizeof *stats); return ARN_OK; }
    if (c->tile_cap < rects.size()) {
        if (c->d_tile_rect) cudaFree(c->d_tile_rect);
        if (c->d_tile_prefix) cudaFree(c->d_tile_prefix);
        c->d_tile_rect = nullptr; c->d_tile_prefix = nullptr; c->tile_cap = 0;
        CUDA_TRY(c, cudaMalloc(&c->d_tile_rect, rects.size() * sizeof(int4)));
        CUDA_TRY(c, cudaMalloc(&c->d_tile_prefix, (rects.size() + 1) * sizeof(unsigned long long)));
        c->tile_cap = rects.size();
    }
    CUDA_TRY(c, cudaMemcpyAsync(c->d_tile_rect, rects.data(), rects.size() * sizeof(int4), cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(c, cudaMemcpyAsync(c->d_tile_prefix, prefix.data(), prefix.size() * sizeof(unsigned long long), cudaMemcpyHostToDevice, c->stream));

    unsigned long long total = prefix.back() * (unsigned long long)(s1 - s0);
    // measured on C3 (tools/prof_cornell.py): 1 pipeline x 2^20 samples 400.8 ms, 2 x 2^20 350.5, 4 x 2^20 336.2, 4 x 2^19 330.8 (= 8 x 2^19)
    // large scenes (DRAM-latency-bound trace) prefer more and smaller waves: C4 4 x 2^19 191.7 ms, 8 x 2^18 188.8, 8 x 2^17 185.9, 8 x 2^16 190.4
    const bool big = s->dev.n_nodes >= ARN_WIDE_MIN_NODES;
    const int pipes_wanted = c->opt_count ? 1 : (c->opt_pipes ? c->opt_pipes : (big ? 8 : 4));
    const bool smem_walk = !c->opt_smem_off && !use_wide(s) && !use_cw8(s) && !c->opt_refill && !c->opt_count && s->dev.pairs != nullptr;
    size_t cap = c->opt_wave ? c->opt_wave : wave_capacity_default(pipes_wanted, big, smem_walk, total);
    if ((unsigned long long)cap > total) cap = (size_t)((total + ARN_BLOCK - 1) / ARN_BLOCK * ARN_BLOCK);
    const unsigned long long n_waves = (total + cap - 1) / cap;
    // per-kernel event timing needs serial launches: it is only taken with one pipeline (ARN_OPT_PIPELINES = 1)
    const int np = (int)std::min<unsigned long long>((unsigned long long)pipes_wanted, n_waves);
    const bool wide = use_wide(s), cw8 = use_cw8(s);
    const bool refill = c->opt_refill && !wide && !cw8 && !c->opt_count;
    const size_t node_bytes = (size_t)s->dev.n_pairs * ARN_PAIR_BYTES;
    const bool smem_nodes = smem_walk;
    const bool textured = s->dev.n_textures != 0;
    for (int i = 0; i < np; i++) {
        int rc = ensure_wave(c, &c->pipes[i], cap); if (rc != ARN_OK) return rc;
        if (refill) { rc = ensure_trace_buf(c, &c->pipes[i], c->pipes[i].wave_cap); if (rc != ARN_OK) return rc; }
        if (textured) { rc = ensure_diff_buf(c, &c->pipes[i]); if (rc != ARN_OK) return rc; }
    }

    WaveParams wp;
    std::memcpy(wp.raster_view, cam->raster_view, 64); std::memcpy(wp.view_parent, cam->view_parent, 64);
    wp.has_lens = cam->has_lens; wp.lens_radius = cam->lens_radius; wp.focal_distance = cam->focal_distance; wp.ortho = cam->ortho;
    // the film's filter (sample/filters.rs): constructor asserts, then the per-axis parameters filter1 reads
    if (film->filter_kind > ARN_FILTER_MITCHELL) return set_err(c, ARN_E_INVALID, "arn_render_pt: unknown film filter");
    if (!(film->filter_radius_x > 0.f) || !(film->filter_radius_y > 0.f)) return set_err(c, ARN_E_INVALID, "arn_render_pt: filter radius must be positive (assert! in every Filter::new)");
    wp.filt_kind = film->filter_kind; wp.filt_a = film->filter_a; wp.filt_b = film->filter_b;
    if (film->filter_kind == ARN_FILTER_LANCZOS) wp.filt_a = 1.0f / (film->filter_a > 0.f ? film->filter_a : 3.0f);    // inv_tau (filters.rs:203-206)
    if (film->filter_kind == ARN_FILTER_GAUSSIAN) wp.filt_a = -film->filter_a;                                           // neg_alpha (:104)
    wp.crop_x0 = film->crop_min_x; wp.crop_y0 = film->crop_min_y; wp.crop_w = cw; wp.crop_h = chh;
    wp.fr_x = film->filter_radius_x; wp.fr_y = film->filter_radius_y;
    wp.tile_dx = (int)dx; wp.tile_dy = (int)dy; wp.tile_lastx = (int)lastx; wp.tile_lasty = (int)lasty; wp.tiles_nx = (int)nx; wp.tiles_ny = (int)ny;
    wp.tile_rx = (int)(long)film->filter_radius_x; wp.tile_ry = (int)(long)film->filter_radius_y;       // filter_radius.cast()
    wp.seed = smp->seed; wp.max_depth = prm->max_depth; wp.min_depth = prm->min_depth; wp.rr_threshold = prm->rr_threshold;
    wp.n_tiles = (uint32_t)rects.size(); wp.tile_rect = c->d_tile_rect; wp.tile_prefix = c->d_tile_prefix;
    wp.spp_begin = s0; wp.spp_count = s1 - s0;
    wp.textured = textured ? 1u : 0u; wp.spp_total = spp;
    wp.strat_ndim = smp->mode == ARN_SAMPLER_STRATIFIED ? smp->ndim : 0u; wp.sampledx = smp->sampledx; wp.sampledy = smp->sampledy;

    size_t ev = 0;
    cudaEvent_t e_begin = get_event(c, ev++), e_end = get_event(c, ev++);
    const bool time_kernels = stats != nullptr && np == 1;
    std::vector<std::pair<size_t, int>> ext_events;      // (event index, bounce)
    uint64_t launches = 0;
    cudaEventRecord(e_begin, c->stream);
    for (int i = 0; i < np; i++) {
        if (i > 0) cudaStreamWaitEvent(c->pipes[i].stream, e_begin, 0);       // after the tile tables / whatever the caller queued
        CUDA_TRY(c, cudaMemsetAsync(c->pipes[i].q.stats, 0, 64, c->pipes[i].stream));
    }
    // ARN_OPT_PDL: every kernel of the chain below starts with pdl_prologue(); the variants without it (counted, refill, textured) launch plainly
    const bool pdl = c->opt_pdl && !time_kernels && !refill && !c->opt_count && !textured;
#define ARN_LAUNCH(kern, grid, block, smem, ...) do { if (pdl) launch_chain(true, kern, grid, block, smem, st, __VA_ARGS__); else kern<<<grid, block, smem, st>>>(__VA_ARGS__); } while (0)
    unsigned long long wave = 0;
    for (unsigned long long base = 0; base < total; base += cap, wave++) {
        arn_ctx::Pipe& P = c->pipes[wave % (unsigned long long)np];
        cudaStream_t st = P.stream;
        uint32_t n = (uint32_t)std::min<unsigned long long>(cap, total - base);
        // k_generate empties the wave's queue counters; every k_trace empties the counter set of the other parity (wavefront.cuh, cnt_*)
        ARN_LAUNCH(k_generate, std::min(c->g_generate, (int)((n + ARN_BLOCK - 1) / ARN_BLOCK)), ARN_BLOCK, 0, wp, P.pb, P.q, base, n);
        launches += 1;
        auto trace = [&](int j) {
            if (time_kernels) { size_t i0 = ev; cudaEventRecord(get_event(c, ev++), st); ext_events.push_back({i0, j}); }
            if (refill) {
                k_ray_setup<<<c->g_setup, ARN_BLOCK, 0, st>>>(s->dev, P.pb, P.q, P.tb, j);
                k_trace_refill<<<c->g_refill, ARN_BLOCK, 0, st>>>(s->dev, P.q, P.tb, j);
                k_classify<<<c->g_classify, ARN_BLOCK, 0, st>>>(s->dev, P.pb, P.q, P.tb, j);
                launches += 2;
            }
            else if (c->opt_count) k_trace<ARN_TRAV_COUNTED><<<c->g_trace, ARN_BLOCK, 0, st>>>(s->dev, P.pb, P.q, j);
            else if (cw8) ARN_LAUNCH(k_trace<ARN_TRAV_CW8>, c->g_trace_8, ARN_BLOCK, 0, s->dev, P.pb, P.q, j);
            else if (wide) ARN_LAUNCH(k_trace<ARN_TRAV_WIDE>, c->g_trace_w, ARN_BLOCK, 0, s->dev, P.pb, P.q, j);
            else if (smem_nodes) ARN_LAUNCH(k_trace<ARN_TRAV_BINARY_SMEM>, c->sm_count, ARN_BLOCK_SMEM, node_bytes, s->dev, P.pb, P.q, j);
            else ARN_LAUNCH(k_trace<ARN_TRAV_BINARY>, c->g_trace, ARN_BLOCK, 0, s->dev, P.pb, P.q, j);
            if (time_kernels) cudaEventRecord(get_event(c, ev++), st);
        };
        trace(0);                                                  // camera rays
        launches += 1;
        for (uint32_t b = 0; b < prm->max_depth; b++) {
            // shade(b): consumes the class queues of trace(b), fills the next active queue + connect / shadow / light-ray queues
            // heavy classes first: the tail of the bounce is cheap Lambert work
            if (textured) { k_shade<SHADE_GENERIC, true><<<c->g_shade_tex, ARN_BLOCK, 0, st>>>(s->dev, wp, P.pb, P.q, (int)b); launches++; }
            else {
            if (s->class_mask & 0x08u) { ARN_LAUNCH(k_shade<SHADE_GLASS>, c->g_shade_g, ARN_BLOCK, 0, s->dev, wp, P.pb, P.q, (int)b); launches++; }
            if (s->class_mask & 0x04u) { ARN_LAUNCH(k_shade<SHADE_PLASTIC>, c->g_shade_p, ARN_BLOCK, 0, s->dev, wp, P.pb, P.q, (int)b); launches++; }
            if (s->class_mask & 0x10u) { ARN_LAUNCH(k_shade<SHADE_GENERIC>, c->g_shade, ARN_BLOCK, 0, s->dev, wp, P.pb, P.q, (int)b); launches++; }
            if (s->class_mask & 0x03u) { ARN_LAUNCH(k_shade<SHADE_DIFFUSE>, c->g_shade_d, ARN_BLOCK, 0, s->dev, wp, P.pb, P.q, (int)b); launches++; }
            }
            trace((int)b + 1);                                     // path rays of bounce b+1, shadow + light rays of bounce b
            ARN_LAUNCH(k_resolve, c->g_resolve, ARN_BLOCK, 0, P.pb, P.q, (int)b);
            launches += 2;
        }
        if (film->filter_radius_x <= 4.f && film->filter_radius_y <= 4.f && film->filter_radius_x >= 0.5f && film->filter_radius_y >= 0.5f) {
            unsigned long long npix = (base + n - 1) / wp.spp_count - base / wp.spp_count + 1;
            int blocks = (int)std::min<unsigned long long>((unsigned long long)c->g_accum_px, (npix * 32 + ARN_BLOCK - 1) / ARN_BLOCK);
            ARN_LAUNCH(k_accumulate_px, std::max(blocks, 1), ARN_BLOCK, 0, wp, P.pb, P.q, (float4*)film_dev, base, n);
        } else
        ARN_LAUNCH(k_accumulate, std::min(c->g_accum, (int)((n + ARN_BLOCK - 1) / ARN_BLOCK)), ARN_BLOCK, 0, wp, P.pb, P.q, (float4*)film_dev, n);
        launches += 1;
        if (radiance_dev) { ARN_LAUNCH(k_store_radiance, std::min(c->g_accum, (int)((n + ARN_BLOCK - 1) / ARN_BLOCK)), ARN_BLOCK, 0, wp, P.pb, radiance_dev, n); launches += 1; }
        CUDA_TRY(c, cudaGetLastError());
    }
#undef ARN_LAUNCH
    // join: everything the pipelines did is ordered before whatever follows on the context's stream
    for (int i = 1; i < np; i++) { cudaEventRecord(c->pipes[i].done, c->pipes[i].stream); cudaStreamWaitEvent(c->stream, c->pipes[i].done, 0); }
    cudaEventRecord(e_end, c->stream);
    if (stats) {
        unsigned long long hs[8] = {0, 0, 0, 0, 0, 0, 0, 0}, one[8];
        for (int i = 0; i < np; i++) {
            CUDA_TRY(c, cudaMemcpyAsync(one, c->pipes[i].q.stats, 64, cudaMemcpyDeviceToHost, c->stream));
            CUDA_TRY(c, cudaStreamSynchronize(c->stream));
            for (int k = 0; k < 8; k++) hs[k] += one[k];
        }
        std::memset(stats, 0, sizeof *stats);
        stats->camera_rays = total; stats->extend_rays = hs[0]; stats->shadow_rays = hs[1]; stats->mis_rays = hs[2];
        stats->invalid_samples = hs[3]; stats->extend_bounce_rays = hs[4]; stats->kernel_launches = launches;
        stats->extend_nodes = hs[5]; stats->extend_tris = hs[6]; stats->extend_spheres = hs[7];
        float ms = 0.f; cudaEventElapsedTime(&ms, e_begin, e_end); stats->gpu_ms = ms;
        double ext = 0.0, extb = 0.0;
        for (auto& pr : ext_events) { float m = 0.f; cudaEventElapsedTime(&m, c->events[pr.first], c->events[pr.first + 1]); ext += m; if (pr.second > 0) extb += m; }
        stats->extend_ms = ext; stats->extend_bounce_ms = extb;
    }
    return ARN_OK;
}

int arn_render_pt_dev(arn_scene* s, const arn_camera* cam, const arn_film* film, const arn_sampler* smp,
                      const arn_pt_params* prm, void* film_dev, arn_stats* stats) {
    return render_pt_impl(s, cam, film, smp, prm, film_dev, stats, nullptr);
}

// Diagnostic twin of arn_render_pt: additionally returns the radiance of every camera sample,
// radiance_out[((y*crop_w + x)*n_spp + s)*4 .. +3] (HOST, pixels of other ranks' tiles stay 0).
int arn_render_pt_samples(arn_scene* s, const arn_camera* cam, const arn_film* film, const arn_sampler* smp,
                          const arn_pt_params* prm, float* film_out, float* radiance_out, arn_stats* stats) {
    if (!s || !film || !smp || !prm || !film_out || !radiance_out) return set_err(s ? s->ctx : nullptr, ARN_E_INVALID, "arn_render_pt_samples: NULL argument");
    arn_ctx* c = s->ctx; std::lock_guard<std::recursive_mutex> g(c->mu); cudaSetDevice(c->device);
    long cw = film->crop_max_x - film->crop_min_x, chh = film->crop_max_y - film->crop_min_y;
    uint32_t spp = smp->sampledx * smp->sampledy; uint32_t s0 = prm->spp_begin, s1 = prm->spp_end ? prm->spp_end : spp;
    if (cw <= 0 || chh <= 0 || s1 <= s0) return set_err(c, ARN_E_INVALID, "arn_render_pt_samples: empty crop window or sample range");
    size_t fbytes = (size_t)cw * chh * 16, rbytes = fbytes * (s1 - s0);
    void *d_film = nullptr, *d_rad = nullptr;
    CUDA_TRY(c, cudaMalloc(&d_film, fbytes));
    if (cudaMalloc(&d_rad, rbytes) != cudaSuccess) { cudaFree(d_film); return set_err(c, ARN_E_OOM, "arn_render_pt_samples: radiance buffer"); }
    cudaMemsetAsync(d_film, 0, fbytes, c->stream); cudaMemsetAsync(d_rad, 0, rbytes, c->stream);
    arn_stats local;
    int rc = render_pt_impl(s, cam, film, smp, prm, d_film, stats ? stats : &local, (float4*)d_rad);
    if (rc == ARN_OK) {
        cudaMemcpyAsync(film_out, d_film, fbytes, cudaMemcpyDeviceToHost, c->stream);
        cudaMemcpyAsync(radiance_out, d_rad, rbytes, cudaMemcpyDeviceToHost, c->stream);
        cudaError_t e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) rc = set_err(c, ARN_E_CUDA, std::string("download: ") + cudaGetErrorString(e));
    }
    cudaFree(d_film); cudaFree(d_rad);
    return rc;
}

int arn_render_pt(arn_scene* s, const arn_camera* cam, const arn_film* film, const arn_sampler* smp,
                  const arn_pt_params* prm, float* film_out, arn_stats* stats) {
    if (!s || !film || !film_out) return set_err(s ? s->ctx : nullptr, ARN_E_INVALID, "arn_render_pt: NULL argument");
    arn_ctx* c = s->ctx; std::lock_guard<std::recursive_mutex> g(c->mu); cudaSetDevice(c->device);
    long cw = film->crop_max_x - film->crop_min_x, chh = film->crop_max_y - film->crop_min_y;
    if (cw <= 0 || chh <= 0) return set_err(c, ARN_E_INVALID, "arn_render_pt: empty crop window");
    size_t bytes = (size_t)cw * (size_t)chh * 16;
    // the device film lives with the context and only grows: no allocation on the steady-state path
    if (c->film_cap < bytes) {
        if (c->d_film) { cudaStreamSynchronize(c->stream); cudaFree(c->d_film); c->d_film = nullptr; c->film_cap = 0; }
        CUDA_TRY(c, cudaMalloc(&c->d_film, bytes));
        c->film_cap = bytes;
    }
    void* d_film = c->d_film;
    cudaMemsetAsync(d_film, 0, bytes, c->stream);
    arn_stats local;
    int rc = arn_render_pt_dev(s, cam, film, smp, prm, d_film, stats ? stats : &local);
    if (rc == ARN_OK) {
        cudaError_t e = cudaMemcpyAsync(film_out, d_film, bytes, cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) rc = set_err(c, ARN_E_CUDA, std::string("film download: ") + cudaGetErrorString(e));
    }
    return rc;
}

}  // extern "C"

// ---------------------------------------------------------------- self test of the fast transcendentals (kernels/cr_math.cuh)
namespace {
// every f32 bit pattern: wrapper (short f64 kernel, library fallback) vs the library value rounded once — must be identical
__global__ void __launch_bounds__(256) k_selftest_math(unsigned long long* out, uint32_t first, uint32_t count_log2) {
    unsigned long long bad[5] = {0, 0, 0, 0, 0};
    const unsigned long long n = 1ull << count_log2;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x) {
        const uint32_t bits = first + (uint32_t)i;
        const float x = __uint_as_float(bits);
        auto same = [](float a, float b) { return __float_as_uint(a) == __float_as_uint(b) || (a != a && b != b); };
        float s, c; cr_sincosf(x, s, c);
        double ds, dc; sincos((double)x, &ds, &dc);
        bad[0] += !same(s, (float)ds) || !same(cr_sinf(x), (float)sin((double)x));
        bad[1] += !same(c, (float)dc) || !same(cr_cosf(x), (float)cos((double)x));
        bad[2] += !same(cr_expf(x), (float)exp((double)x));
        bad[3] += !same(cr_logf(x), (float)log((double)x));
        const uint32_t h = mix32(bits);                                         // pow: this base, a pseudo-random exponent in [-4, 4] or the Beckmann sampler's [0.4, 1.1]
        const float b = (h & 1u) ? -4.f + 8.f * u01(h) : 0.4f + 0.7f * u01(h);
        bad[4] += !same(cr_powf(x, b), (float)pow((double)x, (double)b));
    }
    for (int k = 0; k < 5; k++) {
        unsigned long long v = bad[k];
        for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(&out[k], v);
    }
}
}  // namespace

extern "C" int arn_selftest_pair_records(const arn_node* nodes, uint32_t n_nodes, float* records_out, uint32_t* n_records_out) {
    if (!nodes || !n_records_out) return ARN_E_INVALID;
    std::vector<float> rec;
    const uint32_t n = build_pair_records(nodes, n_nodes, &rec);
    *n_records_out = n;
    if (n && records_out) std::memcpy(records_out, rec.data(), rec.size() * sizeof(float));
    return ARN_OK;
}

int arn_selftest_math(arn_ctx* c, uint32_t first_bits, uint32_t count_log2, uint64_t* mismatches5) {
    if (!c || !mismatches5 || count_log2 > 32) return set_err(c, ARN_E_INVALID, "arn_selftest_math: bad argument");
    std::lock_guard<std::recursive_mutex> g(c->mu); cudaSetDevice(c->device);
    CUDA_TRY(c, cudaMemsetAsync(c->d_ctr, 0, 64, c->stream));
    k_selftest_math<<<c->sm_count * 8, 256, 0, c->stream>>>(c->d_ctr, first_bits, count_log2);
    CUDA_TRY(c, cudaGetLastError());
    unsigned long long h[5];
    CUDA_TRY(c, cudaMemcpyAsync(h, c->d_ctr, 40, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    for (int k = 0; k < 5; k++) mismatches5[k] = h[k];
    return ARN_OK;
}

namespace {
__global__ void __launch_bounds__(128) k_selftest_bsdf(const arn_material m, uint32_t n, const float* __restrict__ wo3, const float* __restrict__ u2,
                                                       const float* __restrict__ wi3, const float* __restrict__ fr, float* __restrict__ out12) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float3 wo = f3(wo3[3 * i], wo3[3 * i + 1], wo3[3 * i + 2]), wi = f3(wi3[3 * i], wi3[3 * i + 1], wi3[3 * i + 2]);
        Surf s; s.pos = f3(0.f, 0.f, 0.f); s.perr = s.pos; s.wo = wo; s.ng = f3(0.f, 0.f, 1.f); s.ns = s.ng; s.dpdu = f3(1.f, 0.f, 0.f);
        if (fr) { const float* q = fr + 9 * (size_t)i; s.dpdu = f3(q[0], q[1], q[2]); s.ns = f3(q[3], q[4], q[5]); s.ng = f3(q[6], q[7], q[8]); }
        Bsdf b; bsdf_build(m, s, b);
        bsdf_prepare<LOBES_ALL>(b, wo);
        const Sampled sm = bsdf_sample<LOBES_ALL>(b, wo, f2(u2[2 * i], u2[2 * i + 1]));
        float3 f; float pdf; bsdf_eval_pdf<LOBES_ALL>(b, wo, wi, f, pdf);
        float* o = out12 + 12 * (size_t)i;
        o[0] = sm.f.x; o[1] = sm.f.y; o[2] = sm.f.z; o[3] = sm.wi.x; o[4] = sm.wi.y; o[5] = sm.wi.z; o[6] = sm.pdf; o[7] = (float)sm.type;
        o[8] = f.x; o[9] = f.y; o[10] = f.z; o[11] = pdf;
    }
}
}  // namespace

extern "C" int arn_selftest_bsdf(arn_ctx* c, const arn_material* m, size_t n, const float* wo3, const float* u2, const float* wi3, const float* frame9, float* out12) {
    if (!c || !m || !wo3 || !u2 || !wi3 || !out12 || n == 0 || n > (1u << 24)) return set_err(c, ARN_E_INVALID, "arn_selftest_bsdf: bad argument");
    if (m->type > ARN_MAT_TRANSLUCENT || m->kd_tex || m->ks_tex || m->aux_tex || m->bump_tex) return set_err(c, ARN_E_INVALID, "arn_selftest_bsdf: constant materials only");
    std::lock_guard<std::recursive_mutex> g(c->mu); cudaSetDevice(c->device);
    float* d = nullptr;
    CUDA_TRY(c, cudaMalloc(&d, n * 29 * sizeof(float)));
    float *dwo = d, *du = d + 3 * n, *dwi = d + 5 * n, *dout = d + 8 * n, *dfr = d + 20 * n;
    cudaError_t e = cudaMemcpyAsync(dwo, wo3, n * 12, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(du, u2, n * 8, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dwi, wi3, n * 12, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess && frame9) e = cudaMemcpyAsync(dfr, frame9, n * 36, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) { k_selftest_bsdf<<<(unsigned)((n + 127) / 128), 128, 0, c->stream>>>(*m, (uint32_t)n, dwo, du, dwi, frame9 ? dfr : nullptr, dout); e = cudaGetLastError(); }
    if (e == cudaSuccess) e = cudaMemcpyAsync(out12, dout, n * 48, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(d);
    if (e != cudaSuccess) return set_err(c, ARN_E_CUDA, cudaGetErrorString(e));
    return ARN_OK;
}

// ---------------------------------------------------------------- multi-GPU film merge (Film::merge_into across GPUs)
// NCCL is resolved at first use with dlopen("libnccl.so.2"): a process that already loaded NCCL (torch, or the Rust
// host's own binding) gets that same copy, a stand-alone host gets the system library, and single-GPU users never
// touch it.  Only the handful of entry points below is used; their C signatures are NCCL's stable public ABI (nccl.h).
namespace {
struct NcclUniqueId { char internal[128]; };          // ncclUniqueId
struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(NcclUniqueId*) = nullptr;
    int (*CommInitRank)(void**, int, NcclUniqueId, int) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    int (*Reduce)(const void*, void*, size_t, int, int, int, void*, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    std::string why;
};
NcclApi* nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) { api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (api.lib) break; }
        if (!api.lib) { api.why = std::string("libnccl.so.2 could not be loaded: ") + (dlerror() ? dlerror() : "?"); return; }
        api.GetUniqueId = (int (*)(NcclUniqueId*))dlsym(api.lib, "ncclGetUniqueId");
        api.CommInitRank = (int (*)(void**, int, NcclUniqueId, int))dlsym(api.lib, "ncclCommInitRank");
        api.CommDestroy = (int (*)(void*))dlsym(api.lib, "ncclCommDestroy");
        api.Reduce = (int (*)(const void*, void*, size_t, int, int, int, void*, cudaStream_t))dlsym(api.lib, "ncclReduce");
        api.GetErrorString = (const char* (*)(int))dlsym(api.lib, "ncclGetErrorString");
        if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.Reduce) { api.why = "libnccl.so.2 lacks the expected entry points"; api.lib = nullptr; }
    });
    return &api;
}
int nccl_fail(arn_ctx* c, const char* what, int r) {
    NcclApi* a = nccl_api();
    return set_err(c, ARN_E_NCCL, std::string(what) + ": " + (a->GetErrorString ? a->GetErrorString(r) : "NCCL error") + " (" + std::to_string(r) + ")");
}
__global__ void __launch_bounds__(256) k_film_merge(float4* __restrict__ dst, const float4* __restrict__ src, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float4 a = dst[i]; const float4 b = src[i];
        a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;          // spectrum_sum += .., filter_weight_sum += .. (film.rs:97-98)
        dst[i] = a;
    }
}
}  // namespace

extern "C" {

int arn_nccl_unique_id(void* id128_out) {
    if (!id128_out) return set_err(nullptr, ARN_E_INVALID, "arn_nccl_unique_id: NULL argument");
    NcclApi* a = nccl_api();
    if (!a->lib) return set_err(nullptr, ARN_E_NCCL, a->why);
    NcclUniqueId id;
    int r = a->GetUniqueId(&id);
    if (r != 0) return nccl_fail(nullptr, "ncclGetUniqueId", r);
    std::memcpy(id128_out, &id, sizeof id);
    return ARN_OK;
}

int arn_nccl_comm_create(arn_ctx* c, const void* id128, int rank, int world_size, void** comm_out) {
    if (!c || !id128 || !comm_out || world_size < 1 || rank < 0 || rank >= world_size) return set_err(c, ARN_E_INVALID, "arn_nccl_comm_create: bad argument");
    NcclApi* a = nccl_api();
    if (!a->lib) return set_err(c, ARN_E_NCCL, a->why);
    cudaSetDevice(c->device);
    NcclUniqueId id; std::memcpy(&id, id128, sizeof id);
    void* comm = nullptr;
    int r = a->CommInitRank(&comm, world_size, id, rank);
    if (r != 0) return nccl_fail(c, "ncclCommInitRank", r);
    *comm_out = comm;
    return ARN_OK;
}

int arn_nccl_comm_destroy(void* comm) {
    if (!comm) return ARN_OK;
    NcclApi* a = nccl_api();
    if (!a->lib) return set_err(nullptr, ARN_E_NCCL, a->why);
    int r = a->CommDestroy(comm);
    return r == 0 ? ARN_OK : nccl_fail(nullptr, "ncclCommDestroy", r);
}

int arn_film_reduce(arn_ctx* c, void* comm, void* film_dev, size_t n_pixels, int root) {
    if (!c || !comm || (n_pixels && !film_dev) || root < 0) return set_err(c, ARN_E_INVALID, "arn_film_reduce: bad argument");
    if (n_pixels == 0) return ARN_OK;
    NcclApi* a = nccl_api();
    if (!a->lib) return set_err(c, ARN_E_NCCL, a->why);
    std::lock_guard<std::recursive_mutex> g(c->mu); cudaSetDevice(c->device);
    int r = a->Reduce(film_dev, film_dev, n_pixels * 4, /*ncclFloat32*/ 7, /*ncclSum*/ 0, root, comm, c->stream);
    return r == 0 ? ARN_OK : nccl_fail(c, "ncclReduce", r);
}

int arn_film_merge(arn_ctx* c, void* dst, const void* src, size_t n_pixels) {
    if (!c || (n_pixels && (!dst || !src))) return set_err(c, ARN_E_INVALID, "arn_film_merge: NULL argument");
    if (n_pixels == 0) return ARN_OK;
    std::lock_guard<std::recursive_mutex> g(c->mu); cudaSetDevice(c->device);
    int grid = (int)std::min<size_t>((size_t)c->sm_count * 8, (n_pixels + 255) / 256);
    k_film_merge<<<grid, 256, 0, c->stream>>>((float4*)dst, (const float4*)src, n_pixels);
    CUDA_TRY(c, cudaGetLastError());
    return ARN_OK;
}

}  // extern "C"
