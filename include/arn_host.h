/* arn_host.h — C entry points of the HOST layer of the B200 path-tracing core: scene
 * ingest and flattening (what the Rust side of arendur does before it would call arn.h).
 *
 * The reference is compiled Rust and no Rust toolchain exists in this environment, so the
 * host side above the C-ABI is C++ (arendur_b200/csrc/host/, mirror classes in
 * arendur_b200/csrc/host/arendur.hpp).  These C functions expose that layer to the Python
 * harness (tests/, bench.py) and to the `arencli` equivalent.  Each cites the reference
 * code path it mirrors.
 */
#ifndef ARN_HOST_H_
#define ARN_HOST_H_

#include "arn.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct arn_hscene arn_hscene;   /* scene under construction: components + lights */

int  arn_hscene_create(arn_hscene** out);
void arn_hscene_destroy(arn_hscene* h);
const char* arn_hscene_last_error(const arn_hscene* h);

/* Arc<Material> with constant textures (material/{matte,plastic,glass,translucent}.rs);
 * `alpha` is filled in (roughness_to_alpha). Returns the material id (>= 0) or an error. */
int arn_hscene_add_material(arn_hscene* h, const arn_material* m);

/* TriangleMesh::from_model / from_model_transformed (shape/triangle.rs:82-160) followed by
 * pushing every TriangleInstance to the component list (component/mod.rs:174-183).
 * positions: n_vertices*3, indices: n_indices (n_indices/3 triangles, trailing remainder
 * ignored as the TriangleInstance iterator does, triangle.rs:229-241), normals / uvs may be
 * NULL, transform16 may be NULL (identity = from_model).  Returns the mesh id or an error. */
int arn_hscene_add_mesh(arn_hscene* h, const float* positions, uint32_t n_vertices,
                        const uint32_t* indices, uint32_t n_indices,
                        const float* normals, const float* uvs,
                        const float* transform16, uint32_t material);

/* ShapedPrimitive::new(Sphere::new(radius,zmin,zmax,phimax), material, emission) optionally
 * wrapped in TransformedComposable (examples/arencli.rs:121-161).  emission3 NULL = not a
 * light; otherwise the primitive is also pushed to `lights`.  transform16 NULL = bare shape.
 * Returns the component index or an error. */
int arn_hscene_add_sphere(arn_hscene* h, float radius, float zmin, float zmax, float phimax,
                          uint32_t material, const float* emission3, const float* transform16);

/* `ImageTexture::new(info, UVMapping, ..)` (texturing/textures/image.rs:104-140) with the pyramid already built by the caller
 * (MipMap::new decodes and resizes with the `image` crate, :218-262 — outside the reference tree): `tex->level_offset[l]` indexes
 * `texels` (n_floats floats, `channels` per texel, row-major).  Add textures BEFORE the materials that use them.  Returns the
 * id to put into arn_material.{kd,ks,aux,bump}_tex (k + 1) or an error. */
int arn_hscene_add_texture(arn_hscene* h, const arn_texture* tex, const float* texels, uint64_t n_floats);

/* The same from a picture file: `MipMap::new` restated for PNG files (host/image_io.hpp) — one Lanczos3 resize of the original
 * per level (next-power-of-two sizes), `to_rgb()` (`params->channels` = 3) or `to_luma()` (1), `convert_in(gamma, scale)`.
 * `params` supplies trilinear / wrapping / max_aniso / UV scale and shift; its level fields are ignored.  PARITY UNPINNED for the
 * pyramid values: decoder and resampler belong to the `image` crate, which is not part of the reference tree.  `load_obj`'s
 * map_Kd / map_Ks / map_bump and the scene file's `Image` textures go through the same routine.  mean3_out (optional) receives
 * MipMap::mean.  Returns the texture id, ARN_E_IO if the file cannot be opened or decoded. */
int arn_hscene_add_texture_file(arn_hscene* h, const char* path, const arn_texture* params, int gamma, float scale, float* mean3_out);

/* `lights.push(light.to_arc())` for the scene file's Point / Spot / Distant lights
 * (examples/arencli.rs:95-98).  These come first in `Scene.lights`, before the emissive
 * primitives.  Returns the index into analytic_lights or an error. */
int arn_hscene_add_light(arn_hscene* h, const arn_analytic_light* light);

/* SpotLight::new(pos, towards, intensity, total_angle, start_falloff_angle)
 * (lighting/pointlights.rs:103-122): builds parent_local = rotation(towards -> +z) * translation(pos)
 * with cgmath's Quaternion::from_arc.  cgmath is a crates.io dependency that is not vendored in the
 * reference tree; its published from_arc / Matrix4::from(Quaternion) are restated (parity unpinned —
 * scene files carry the matrices themselves and do not go through this helper). */
int arn_spot_light_make(const float* pos3, const float* towards3, const float* intensity3,
                        float total_angle, float start_falloff_angle, arn_analytic_light* out);
/* PointLight::new (pointlights.rs:25-27) / DistantLight::new + set_world_bounds given the radius
 * (distantlight.rs:26-50). */
int arn_point_light_make(const float* pos3, const float* intensity3, arn_analytic_light* out);
int arn_distant_light_make(const float* intensity3, const float* dir3, float world_radius, arn_analytic_light* out);

/* component::load_obj (component/mod.rs:65-185): tobj-compatible OBJ + MTL ingest, material
 * choice per MTL, one mesh per model. Returns the number of triangles added or an error. */
int arn_hscene_load_obj(arn_hscene* h, const char* path, const float* transform16);

/* examples/arencli.rs parse_input (:70-204): reads the JSON scene description; fills the
 * camera, film, sampler and PT parameters.  Component order is fixed as: meshes in file
 * order, then shaped primitives in file order (the reference's HashMap order is random).
 * base_dir: directory against which relative mesh filenames are resolved (NULL = cwd). */
int arn_hscene_load_json(arn_hscene* h, const char* json_path, const char* base_dir,
                         arn_camera* cam, arn_film* film, arn_sampler* sampler, arn_pt_params* params,
                         char* outputfilename, size_t outputfilename_cap);

/* BVH::new(&components, strategy) + Scene::new(lights, bvh) (component/bvh.rs:58-79,
 * renderer/scene.rs:31-51).  After this the description is complete. */
int arn_hscene_build(arn_hscene* h, int strategy);
/* The same with the tree built on the device by arn_bvh_build_gpu (include/arn.h): start-up in
 * milliseconds instead of seconds at 10^7 triangles, at the price of the reference's topology
 * (tie-breaks may differ).  build_ms_out may be NULL. */
int arn_hscene_build_gpu(arn_hscene* h, arn_ctx* ctx, float* build_ms_out);

/* The flattened description; pointers stay valid until the hscene is modified or destroyed. */
const arn_scene_desc* arn_hscene_desc(const arn_hscene* h);

/* PerspecCam::new (filming/perspective.rs:42-90) + ProjCameraInfo::new (projective.rs:24-45).
 * parent_view16: the JSON "transform"; screen4 = pmin.x, pmin.y, pmax.x, pmax.y. */
int arn_camera_make(const float* parent_view16, const float* screen4, float znear, float zfar,
                    float fov, int has_lens, float lens_radius, float focal_distance,
                    float res_x, float res_y, arn_camera* out);

/* OrthoCam::new(view_parent, screen, znear, zfar, lens, film) (filming/ortho.rs:30-67).  Note the
 * reference's argument is view_parent here (PerspecCam::new takes parent_view). */
int arn_ortho_camera_make(const float* view_parent16, const float* screen4, float znear, float zfar,
                          int has_lens, float lens_radius, float focal_distance,
                          float res_x, float res_y, arn_camera* out);

/* Image::save (filming/film.rs:380-391): finalize + 8-bit RGB PNG, no gamma. */
int arn_save_png(const char* path, const float* film, uint32_t width, uint32_t height);

#ifdef __cplusplus
}
#endif
#endif
