/* arn.h — C-ABI of the B200-native path-tracing core for arendur.
 *
 * This is the drop-in boundary: plain pointers and sizes, no
 * C++/torch types.  A Rust shim implementing arendur's `Renderer` / `Composable`
 * traits binds exactly these entry points (see INTEGRATION.md for the `extern "C"`
 * block).  Every entry point cites the reference interface it replaces
 * (paths relative to the arendur source tree).
 *
 * Conventions
 *   - every function returns an `int` status: ARN_OK (0) or a negative ARN_E_* code;
 *     a human-readable message for the last failure on a context is available through
 *     arn_last_error().  Nothing ever unwinds across this boundary (the reference
 *     panics instead: `assert!/expect/unwrap`, e.g. component/bvh.rs:103,117).
 *   - all floating point is IEEE f32 (reference: geometry/foundamental.rs:15).
 *   - matrices are 16 floats, COLUMN-major, as cgmath's Matrix4 (x,y,z,w columns).
 *   - a miss is prim_id = -1, t = +inf.
 *   - handles are opaque; the caller owns every host buffer it passes in; uploads copy.
 */
#ifndef ARN_H_
#define ARN_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ARN_OK             0
#define ARN_E_INVALID     -1   /* bad argument / inconsistent description          */
#define ARN_E_CUDA        -2   /* CUDA runtime error (message has the cuda string) */
#define ARN_E_OOM         -3   /* host or device allocation failed                 */
#define ARN_E_IO          -4   /* file could not be read / parsed                  */
#define ARN_E_UNSUPPORTED -5   /* valid in arendur but outside this hot path       */
#define ARN_E_NCCL        -6   /* NCCL missing or a collective failed               */

/* ---------------------------------------------------------------- flattened scene */

/* One BVH node, 32 bytes, 32-byte aligned on the device.
 * Replaces `LinearNode` (component/bvh.rs:136-146; 48 B in Rust).
 *   interior: (len_axis >> 2) == 0, `offset` = distance to the second child
 *             (first child is the next node, pre-order), axis = len_axis & 3 in 0..2
 *   leaf    : len = len_axis >> 2 > 0, `offset` = first slot in the ordered primitive
 *             list, len_axis & 3 == 3 (the reference stores split_axis = 4). */
typedef struct arn_node {
    float    bmin[3];
    float    bmax[3];
    uint32_t offset;
    uint32_t len_axis;
} arn_node;

#define ARN_PRIM_SPHERE 0x80000000u   /* component reference: high bit = sphere table */

/* BVH build strategy — `BVHStrategy` (component/bvh.rs:40-47). */
#define ARN_BVH_SAH         0
#define ARN_BVH_MIDDLECOUNT 1
#define ARN_BVH_MIDPOINT    2

/* Material record — the four materials `load_obj`/arencli can produce
 * (component/mod.rs:139-164, material/{matte,plastic,glass,translucent}.rs), with
 * constant textures only (texturing/textures/mod.rs:16-33).
 * `alpha` = roughness_to_alpha(roughness) (bxdf/microfacet.rs:57-63), evaluated once
 * on the host. */
#define ARN_MAT_MATTE       0
#define ARN_MAT_PLASTIC     1
#define ARN_MAT_GLASS       2
#define ARN_MAT_TRANSLUCENT 3
typedef struct arn_material {
    uint32_t type;
    float    kd[3];       /* Matte kd / diffuse                                  */
    float    ks[3];       /* specular                                            */
    float    sigma;       /* Matte: Oren–Nayar sigma, already clamped to [0,90]  */
    float    roughness;
    float    alpha;
    float    eta;         /* Glass: optical density                              */
    float    dissolve;    /* Translucent                                         */
    /* image textures: 0 = the constant above, k + 1 = entry k of arn_scene_desc.textures.
     * kd_tex / ks_tex are RGB textures replacing kd / ks (`diffuse` / `specular` of the material sources), aux_tex a Luma texture
     * replacing sigma (Matte) or roughness (others), bump_tex the Luma displacement of add_bumping (material/mod.rs:42-86) */
    uint32_t kd_tex, ks_tex, aux_tex, bump_tex;
} arn_material;           /* 64 bytes */

/* `ImageTexture<Float, _, UVMapping>` (texturing/textures/image.rs:26-35, texturing/mappings.rs:14-31) with its `MipMap`
 * (image.rs:212-216) flattened.  The pyramid levels are built on the host side exactly as MipMap::new does (`image` crate
 * decode + Lanczos3 resize per level + convert_in, image.rs:218-262: third-party code outside the reference tree) and passed
 * as float texels, level 0 first, row-major, `channels` floats per texel. */
#define ARN_WRAP_REPEAT 0u   /* ImageWrapMode (image.rs:573-581) */
#define ARN_WRAP_BLACK  1u
#define ARN_WRAP_CLAMP  2u
#define ARN_TEX_MAX_LEVELS 16
typedef struct arn_texture {
    uint32_t channels;           /* 3: RGBImageTexture, 1: LumaImageTexture                                  */
    uint32_t n_levels;           /* pyramid.len(), 1..16                                                      */
    uint32_t trilinear;          /* ImageInfo.trilinear: trilinear look-up, else EWA                          */
    uint32_t wrapping;           /* ARN_WRAP_*                                                                */
    float    max_aniso;          /* ImageInfo.max_aniso                                                       */
    float    scale_u, scale_v, shift_u, shift_v;     /* UVMapping.scaling / .shifting                         */
    uint32_t level_w[ARN_TEX_MAX_LEVELS], level_h[ARN_TEX_MAX_LEVELS];
    uint32_t level_offset[ARN_TEX_MAX_LEVELS];       /* index of the level's first float in arn_scene_desc.texels */
} arn_texture;

/* Mesh record: one `TriangleMesh` (shape/triangle.rs:26-36).  Vertices are already
 * in world space (from_model_transformed, triangle.rs:120-160). */
typedef struct arn_mesh {
    uint32_t material;
    uint32_t has_normals;
    uint32_t has_uvs;
    uint32_t reserved;
} arn_mesh;

/* Sphere primitive: `ShapedPrimitive<Sphere, _>` (component/shape.rs:21-26) optionally
 * wrapped in `TransformedComposable` (component/transformed.rs:20-24). */
typedef struct arn_sphere {
    float    radius, zmin, zmax, phimax, thetamin, thetamax;  /* shape/sphere.rs:19-31 */
    uint32_t material;
    uint32_t has_transform;   /* 0: bare ShapedPrimitive (no ray round trip)           */
    uint32_t emissive;        /* lighting_profile.is_some()                            */
    float    emission[3];     /* constant emission texture value                       */
    float    local_parent[16];
    float    parent_local[16];
} arn_sphere;

/* `PointLight`, `SpotLight` (lighting/pointlights.rs:16-22,88-101) and `DistantLight`
 * (lighting/distantlight.rs:16-21): the lights arencli reads from the scene file's `lights` array
 * (examples/arencli.rs:95-98,487-509).  Fields are the reference's own (its serde form), so a
 * deserialised light is passed through unchanged. */
#define ARN_LIGHT_POINT   0u
#define ARN_LIGHT_SPOT    1u
#define ARN_LIGHT_DISTANT 2u
#define ARN_LIGHT_ANALYTIC 0x80000000u   /* light_prims[i] = ARN_LIGHT_ANALYTIC | k: entry k of analytic_lights */
typedef struct arn_analytic_light {
    uint32_t type;               /* ARN_LIGHT_*                                                    */
    float    pos[3];             /* Point / Spot: posw                                             */
    float    intensity[3];
    float    cost, cosf;         /* Spot: cos(total angle), cos(falloff start)                     */
    float    parent_local[16];   /* Spot: column-major; falloff reads its z row only               */
    float    dir[3];             /* Distant: direction the light travels in (normalised by ::new)  */
    float    world_radius;       /* Distant: bounding-sphere radius; power and pfrom scale with it */
} arn_analytic_light;

/* Everything `Scene::new(lights, BVH::new(components, SAH))` holds
 * (renderer/scene.rs:23-51, component/bvh.rs:49-79), flattened. */
typedef struct arn_scene_desc {
    /* triangle soup: TriangleMesh::{vertices, indices, normals, uvs}            */
    uint32_t        n_vertices;
    const float*    positions;     /* n_vertices * 3, world space                */
    const float*    normals;       /* n_vertices * 3 or NULL (per-mesh flag says if valid) */
    const float*    uvs;           /* n_vertices * 2 or NULL                     */
    uint32_t        n_triangles;
    const uint32_t* indices;       /* n_triangles * 3 into the vertex arrays     */
    const uint32_t* tri_mesh;      /* n_triangles: mesh id of each triangle      */
    uint32_t        n_meshes;
    const arn_mesh* meshes;
    uint32_t        n_spheres;
    const arn_sphere* spheres;
    uint32_t        n_materials;
    const arn_material* materials;
    /* component list in the order handed to BVH::new: entry = triangle index, or
     * ARN_PRIM_SPHERE | sphere index.  prim_id in every hit indexes THIS list.   */
    uint32_t        n_prims;
    const uint32_t* prims;
    /* flattened BVH (arn_bvh_build output, or the Rust side's own BVH flattened) */
    uint32_t        n_nodes;
    const arn_node* nodes;
    const uint32_t* order;         /* n_prims: ordered slot -> index into prims   */
    /* lights in `Scene.lights` order (arencli: the file's `lights` array, then the emissive
     * primitives), with the power distribution of Scene::new (renderer/scene.rs:31-51,
     * sample/distribution.rs:25-63) */
    uint32_t        n_lights;
    const uint32_t* light_prims;   /* n_lights: index into prims (an emissive sphere), or
                                      ARN_LIGHT_ANALYTIC | index into analytic_lights        */
    const float*    light_func;    /* n_lights: power().to_xyz().y                */
    const float*    light_cdf;     /* n_lights + 1                                */
    float           light_func_integral;
    uint32_t        n_analytic_lights;
    const arn_analytic_light* analytic_lights;
    /* image textures referenced by the materials (N4); a scene with n_textures > 0 is rendered with ray differentials
     * (geometry/interaction.rs:204-251, filming/perspective.rs:292-320) */
    uint32_t        n_textures;
    const arn_texture* textures;
    uint64_t        n_texel_floats;
    const float*    texels;
} arn_scene_desc;

/* `PerspecCam` (filming/perspective.rs:25-38) or `OrthoCam` (filming/ortho.rs:19-28) reduced to
 * what ray generation reads (perspective.rs:292-320, ortho.rs:180-198): raster_view =
 * inverse(view_screen) * raster_screen (filming/projective.rs:29-37) and view_parent. */
typedef struct arn_camera {
    float    raster_view[16];
    float    view_parent[16];
    uint32_t has_lens;
    float    lens_radius;
    float    focal_distance;
    uint32_t ortho;              /* 0 PerspecCam, 1 OrthoCam (rays leave the raster point along +z) */
} arn_camera;

/* `Film` (filming/film.rs:38-45).  A deserialised film always filters with Lanczos(radius (4,4),
 * tau 3) whatever the file says (film.rs:42,47-51) and `filter_radius` only sizes the splat box;
 * `Film::new(resolution, crop, filter)` accepts any `Filter` of sample/filters.rs, selected here by
 * filter_kind (0 = that default Lanczos, so a zero-initialised tail keeps the arencli behaviour). */
#define ARN_FILTER_LANCZOS  0u   /* LanczosSincFilter (filters.rs:189-240): filter_a = tau, 0 = 3             */
#define ARN_FILTER_BOX      1u   /* BoxFilter (:35-59)                                                        */
#define ARN_FILTER_TRIANGLE 2u   /* TriangleFilter (:61-85): (rx - |x|) (ry - |y|)                            */
#define ARN_FILTER_GAUSSIAN 3u   /* GaussianFilter (:87-127): filter_a = alpha; sic: subtracts -alpha r^2     */
#define ARN_FILTER_MITCHELL 4u   /* MitchellFilter (:129-187): filter_a = b, filter_b = c                     */
typedef struct arn_film {
    uint32_t res_x, res_y;
    int32_t  crop_min_x, crop_min_y, crop_max_x, crop_max_y;  /* pixels, max exclusive */
    float    filter_radius_x, filter_radius_y;   /* Film.filter_radius: the splat box AND the filter's own radius */
    uint32_t filter_kind;
    float    filter_a, filter_b;
} arn_film;

/* Sampler: `StrataSampler{sampledx, sampledy, ndim}` (sample/strata.rs:25-31) gives the
 * sample count; the draws themselves come from the counter-based ParitySampler
 * (DESIGN.md "Sampler"). */
#define ARN_SAMPLER_PARITY     0u   /* every draw = hash(seed, pixel, sample, stream, draw index): what the reference's sampler
                                      amounts to, draw for draw, through its never-reset dimension counter (SURVEY A-17)   */
#define ARN_SAMPLER_STRATIFIED 1u   /* the sampler AS INTENDED (not reference behaviour): the dimension counters
                                      restart with every sample; the first `ndim` 1-D draws of sample s are
                                      (perm_d(s) + u) / spp (the sum kept below perm_d(s) + 1), the first `ndim` 2-D draws fall into cell c = perm_d(s) of the
                                      sampledx x sampledy grid, x-major like strata.rs:67-80: ((c / sampledy + u) / sampledx,
                                      (c % sampledy + u') / sampledy); perm_d = a hash permutation of [0, spp) keyed by
                                      (seed, pixel, dimension); u, u' and every later draw are the parity sampler's        */
typedef struct arn_sampler {
    uint32_t sampledx, sampledy, ndim;
    uint32_t seed;
    uint32_t mode;                  /* ARN_SAMPLER_*                                                                        */
} arn_sampler;

/* `PTRenderer` constants (renderer/pt.rs:37-52): rr_threshold = 0.05,
 * min_depth = max_depth / 2. Tiles: film.spawn_tiles(16,16) (pt.rs:131). */
typedef struct arn_pt_params {
    uint32_t max_depth;
    uint32_t min_depth;
    float    rr_threshold;
    uint32_t tiles_x, tiles_y;   /* tile grid for multi-GPU partitioning (16,16)       */
    uint32_t rank, world_size;   /* this context renders the tiles (ix, iy) with (ix + iy) % world == rank */
    uint32_t spp_begin, spp_end; /* sample index range to render, [0, spp) for all      */
    uint32_t partition_subdiv;   /* multi-GPU load balance: 0 or 1 = ranks own whole tiles; k > 1 = every tile is cut into
                                    k x k cells (the last one absorbs the remainder, like spawn_tiles) and cell (jx, jy) of
                                    tile (ix, iy) goes to rank (ix*k + jx + iy*k + jy) % world_size.  Only the assignment of
                                    pixels to ranks changes: tile sinks stay those of the tiles_x x tiles_y grid */
} arn_pt_params;

typedef struct arn_ray {
    float o[3];
    float d[3];
    float tmax;
} arn_ray;                       /* 28 bytes: RawRay (geometry/ray.rs:64-69) minus the cache */

typedef struct arn_hit {
    int32_t prim_id;             /* index into arn_scene_desc.prims, -1 = miss */
    float   t;
} arn_hit;

typedef struct arn_stats {
    uint64_t camera_rays;        /* camera samples generated                               */
    uint64_t extend_rays;        /* closest-hit traversals on path rays (incl. camera rays) */
    uint64_t shadow_rays;        /* LightSample::occluded traversals                        */
    uint64_t mis_rays;           /* BSDF-sampled light rays (scene.rs:146)                  */
    uint64_t invalid_samples;    /* radiance replaced by black (pt.rs:152-156)              */
    uint64_t kernel_launches;    /* kernels of this library launched by the call            */
    double   gpu_ms;             /* device time of the call, CUDA events                    */
    double   extend_ms;          /* device time of the k_trace launches (path + shadow + light rays); needs serial
                                    launches: 0 unless ARN_OPT_PIPELINES is 1                              */
    double   extend_bounce_ms;   /* ... without each wave's first launch (camera rays): incoherent rays */
    uint64_t extend_bounce_rays; /* rays traced by those launches (path rays of bounces >= 1, shadow, light) */
    /* filled only when ARN_OPT_COUNT_TRAVERSAL is on (instrumented kernels, not for timing):
     * BVH nodes / triangles / spheres tested by ALL traversals of k_trace — the Nn, Nt
     * of the algorithmic-bytes figure, SURVEY.md §8(d) */
    uint64_t extend_nodes, extend_tris, extend_spheres;
} arn_stats;

typedef struct arn_ctx   arn_ctx;
typedef struct arn_scene arn_scene;

/* ---------------------------------------------------------------- host-only pieces */

/* Build the reference's BVH over `n` components given their parent-space bounds
 * (bounds6 = pmin.xyz, pmax.xyz per component) and `intersection_cost()` values.
 * Replaces BVH::new + recursive_build + BuildNode::flatten
 * (component/bvh.rs:58-79,246-465,219-243) INCLUDING its SAH quirks, so that the
 * tree, the primitive order and therefore every tie-break equal the reference's.
 * nodes_out must hold 2*n-1 nodes, order_out n entries. */
int arn_bvh_build(uint32_t n, const float* bounds6, const float* costs, int strategy,
                  arn_node* nodes_out, uint32_t* order_out, uint32_t* n_nodes_out);

/* Scene::new's light distribution: Distribution1D::new (sample/distribution.rs:25-63).
 * cdf_out has n+1 entries. */
int arn_light_distribution(uint32_t n, const float* func, float* cdf_out, float* integral_out);

/* TilePixel::finalize + ToNorm<u8> (filming/film.rs:338-344, spectrum/macros.rs:164-180):
 * film = n_pixels * float4 (sum r,g,b, weight sum).  rgb_out (n*3 floats) and/or
 * rgb8_out (n*3 bytes) may be NULL. */
int arn_film_finalize(const float* film, size_t n_pixels, float* rgb_out, uint8_t* rgb8_out);

/* ---------------------------------------------------------------- device contexts */

/* One context per GPU / rank; owns a stream set and the wavefront queues. */
int  arn_ctx_create(int device, arn_ctx** out);
void arn_ctx_destroy(arn_ctx* ctx);
const char* arn_last_error(const arn_ctx* ctx);   /* ctx may be NULL: last global error */

/* Device BVH build: a linear BVH (63-bit Morton order, Karras hierarchy, bottom-up
 * bounds) over the same component bounds, emitted in the same pre-order 32-byte node layout and ordered
 * component list as arn_bvh_build (`LinearNode`, component/bvh.rs:136-146,219-243), for scenes where the
 * host build of the reference's SAH tree (seconds at 10^7 triangles) dominates start-up.  The topology is
 * NOT the reference's: hits agree except where primitives tie in t or a transformed-sphere hit rewrote the
 * ray earlier in a different visiting order.  bounds6 / nodes_out / order_out are HOST buffers
 * (2n-1 nodes, n entries); build_ms_out (may be NULL) = device time of the build without the copies. */
int arn_bvh_build_gpu(arn_ctx* ctx, uint32_t n, const float* bounds6, arn_node* nodes_out,
                      uint32_t* order_out, uint32_t* n_nodes_out, float* build_ms_out);

/* Upload a flattened scene (copies; desc buffers may be freed afterwards).
 * The handle is immutable and may be used from any host thread. */
int  arn_scene_upload(arn_ctx* ctx, const arn_scene_desc* desc, arn_scene** out);
void arn_scene_destroy(arn_scene* scene);

/* Batched `Composable::intersect_ray` on the aggregate (component/bvh.rs:97-128):
 * closest hit for n rays.  HOST buffers; copies are part of the call. */
int arn_intersect_closest(arn_scene* scene, const arn_ray* rays, size_t n, arn_hit* hits_out);
/* Batched `Composable::can_intersect` (component/mod.rs:35-38): 1 = occluded. */
int arn_intersect_any(arn_scene* scene, const arn_ray* rays, size_t n, uint8_t* out);

/* Same with DEVICE buffers, asynchronous on the context's stream; `stats` may be NULL.
 * rays_dev: n * 28 B, hits_dev: n * 8 B. */
int arn_intersect_closest_dev(arn_scene* scene, const void* rays_dev, size_t n, void* hits_dev,
                              arn_stats* stats);
int arn_intersect_any_dev(arn_scene* scene, const void* rays_dev, size_t n, void* out_dev,
                          arn_stats* stats);

/* Instrumented closest-hit pass over DEVICE buffers: counters_out[0..2] = BVH nodes tested,
 * triangles tested, spheres tested, summed over the batch — the Nn / Nt of the
 * algorithmic-bytes figure.  Synchronous. */
int arn_intersect_closest_counted_dev(arn_scene* scene, const void* rays_dev, size_t n, void* hits_dev,
                                      uint64_t* counters_out);

/* `PTRenderer::render` (renderer/pt.rs:128-176) up to and including the tile merge
 * (`Film::collect_into`, film.rs:171-183) but before finalize: accumulates
 * (sum r, sum g, sum b, weight sum) per crop-window pixel, row-major, into `film_out`
 * (HOST, crop_w * crop_h * 4 floats, overwritten). */
int arn_render_pt(arn_scene* scene, const arn_camera* cam, const arn_film* film,
                  const arn_sampler* sampler, const arn_pt_params* params,
                  float* film_out, arn_stats* stats);
/* Device-resident variant: film_dev is a DEVICE buffer that is ACCUMULATED into
 * (caller zeroes it); used for multi-GPU, where the caller reduces the per-rank films
 * (Film::merge_into semantics, film.rs:82-101) with NCCL. */
int arn_render_pt_dev(arn_scene* scene, const arn_camera* cam, const arn_film* film,
                      const arn_sampler* sampler, const arn_pt_params* params,
                      void* film_dev, arn_stats* stats);

/* ---------------------------------------------------------------- multi-GPU film merge
 * One arn_ctx (and one process or thread) per GPU; every rank renders its tiles of the frame into a
 * full-frame device film (arn_render_pt_dev) and the films are summed on the root — `Film::merge_into`
 * / `Film::collect_into` (filming/film.rs:82-101,171-183) across GPUs.  The exchange is one
 * ncclReduce(sum) over NVLink/NVSwitch, enqueued on the context's stream.  libnccl.so.2 is resolved
 * at first use (dlopen), so single-GPU users do not need it. */

/* Sum `film_dev` (n_pixels * float4: sum r, g, b, weight) over the ranks of `nccl_comm` into rank
 * `root`'s buffer, in place, on the context's stream (asynchronous; other ranks' buffers are left
 * unchanged).  `nccl_comm` is an ncclComm_t the caller owns (ncclCommInitRank on the Rust side, or
 * arn_nccl_comm_create below). */
int arn_film_reduce(arn_ctx* ctx, void* nccl_comm, void* film_dev, size_t n_pixels, int root);
/* dst += src on the device, pixel by pixel (the per-pixel sum of `Film::merge_into`, film.rs:82-101): lets a root
 * keep a running film across several reduced slices of samples.  Asynchronous on the context's stream. */
int arn_film_merge(arn_ctx* ctx, void* dst_film_dev, const void* src_film_dev, size_t n_pixels);
/* Communicator plumbing for hosts that have no NCCL binding of their own: rank 0 obtains a 128-byte
 * ncclUniqueId, ships it to the other ranks by any means, and every rank calls arn_nccl_comm_create
 * on its own context (collective; the context's device becomes the communicator's device). */
int arn_nccl_unique_id(void* id128_out);
int arn_nccl_comm_create(arn_ctx* ctx, const void* id128, int rank, int world_size, void** comm_out);
int arn_nccl_comm_destroy(void* nccl_comm);

/* Diagnostic twin of arn_render_pt (parity tests): additionally returns the radiance
 * `calculate_lighting` produced for every camera sample (renderer/pt.rs:144-148),
 * radiance_out[((y*crop_w + x)*n_spp + s)*4 + {0,1,2}], HOST buffer of crop_w*crop_h*n_spp*4 floats; (x, y) is the
 * TILE pixel, which runs over [0, crop_w) x [0, crop_h) whatever crop.pmin is (Film::spawn_tiles, film.rs:118-121). */
int arn_render_pt_samples(arn_scene* scene, const arn_camera* cam, const arn_film* film,
                          const arn_sampler* sampler, const arn_pt_params* params,
                          float* film_out, float* radiance_out, arn_stats* stats);

/* Context options. */
#define ARN_OPT_COUNT_TRAVERSAL 1   /* value != 0: arn_render_pt* use instrumented extend kernels          */
#define ARN_OPT_WAVE_CAPACITY   2   /* camera samples per wave; 0 = auto: 1 << 19 (1 << 17 for large trees, 1 << 20 with < 3 pipelines or when small trees are walked from shared memory and every pipeline gets two waves); env ARN_WAVE       */
#define ARN_OPT_PIPELINES       4   /* 1..8 concurrent wave pipelines; 0 = auto (default): 4, or 8 for large trees (also env ARN_PIPES). Per-kernel
                                       timings in arn_stats (extend_ms, extend_bounce_ms) need serial launches and
                                       are only filled with 1                                                      */
#define ARN_OPT_TRACE_REFILL    5   /* value != 0: lane-refilling trace (ray stream + persistent warps, kernels/trace_refill.cuh) on trees walked
                                       with the binary nodes; same bits; also env ARN_REFILL */
#define ARN_OPT_BVH_WIDTH       3   /* 0 auto (default), 2 binary nodes, 4 the 4-wide collapse, 8 the compressed 8-wide collapse (set BEFORE the
                                       upload for trees below 2^20 nodes: the 8-wide nodes are built at upload); same bits */
#define ARN_OPT_SMEM_NODES      6   /* 0 auto (default): trees with up to ARN_SMEM_NODE_BYTES / 112 interior nodes are walked by k_trace from
                                       pair records staged in shared memory; 1: never (also env ARN_SMEM_NODES=0); same bits */
#define ARN_SMEM_NODE_BYTES (160 * 1024)
#define ARN_OPT_PDL             7   /* value != 0: programmatic dependent launch along each pipeline's kernel chain (measured; off by
                                       default; also env ARN_PDL) */
int arn_ctx_set_option(arn_ctx* ctx, int option, long long value);

int arn_ctx_synchronize(arn_ctx* ctx);
/* The context's CUDA stream as a cudaStream_t cast to void* (for event timing). */
void* arn_ctx_stream(arn_ctx* ctx);

/* Diagnostic: runs the fast correctly-rounded sin / cos / exp / log / pow (kernels/cr_math.cuh) over the 2^count_log2 f32 bit
 * patterns starting at first_bits and counts results that differ from the f64 library value rounded once — the definition
 * both this library and the oracle use.  mismatches5 = {sin, cos, exp, log, pow}; all must be 0. */
int arn_selftest_math(arn_ctx* ctx, uint32_t first_bits, uint32_t count_log2, uint64_t* mismatches5);

/* Diagnostic (host only, no device needed): the pair records arn_scene_upload builds for k_trace's shared-memory walk of a small tree
 * (kernels/traverse.cuh, traverse2p): 32 floats per interior node, in node order.  *n_records_out = 0 when the tree does not qualify
 * (single leaf, or more than ARN_SMEM_NODE_BYTES / 128 interior nodes).  records_out may be NULL to query the count. */
int arn_selftest_pair_records(const arn_node* nodes, uint32_t n_nodes, float* records_out, uint32_t* n_records_out);

/* Diagnostic: the BSDF of material `m` (Material::compute_scattering with constant textures, material/{matte,plastic,glass,translucent}.rs) at a fixed local
 * frame (dpdu = x, shading normal = geometric normal = z; or, with frame9 != NULL, dpdu / shading normal / geometric normal per probe),
 * for n host-side triples (wo[3], u[2], wi[3]):
 * out[12 i ..] = evaluate_sampled(wo, u, ALL) -> f(3), wi(3), pdf, type;  evaluate(wo, wi, ALL)(3);  pdf(wo, wi, ALL)
 * (material/bsdf.rs:82-145,205-222).  Kernel-level parity probe: tests compare it with the oracle's on random directions. */
int arn_selftest_bsdf(arn_ctx* ctx, const arn_material* m, size_t n, const float* wo3, const float* u2, const float* wi3, const float* frame9, float* out12);

/* Library identification: "arendur_b200 <version> sm_100a". */
const char* arn_version(void);

#ifdef __cplusplus
}
#endif
#endif /* ARN_H_ */
