// arendur.hpp — C++ mirror of the slice of arendur's Rust API that sits on the path-tracing hot
// path, implemented on top of the C-ABI (include/arn.h, include/arn_host.h).
//
// The reference is compiled Rust and no Rust toolchain exists in this environment, so this header is
// the host-side "plugin interface" a user of arendur would recognise: same type and method names, same
// argument meaning, same error behaviour (the reference panics via assert!/expect/unwrap — here every
// such condition throws arendur::Panic carrying the C-ABI's message).  Everything heavy happens behind
// the C-ABI: BVH::build == arn_bvh_build, PTRenderer::render == arn_render_pt on the GPU.
//
//   reference                                          here
//   component::load_obj (component/mod.rs:65)          arendur::load_obj
//   Sphere::new / ::full (shape/sphere.rs:133-163)     arendur::Sphere
//   MatteMaterial / Plastic / Glass / Translucent      arendur::Material::{matte,plastic,glass,translucent}
//   ShapedPrimitive::new + TransformedComposable::new  arendur::Components::push_shaped
//   BVH::new(&components, BVHStrategy::SAH)            arendur::BVH::build
//   Scene::new(lights, aggregate)                      arendur::Scene
//   PerspecCam::new, Film (filming/*.rs)               arendur::PerspecCam, arendur::Film
//   StrataSampler (sample/strata.rs)                   arendur::StrataSampler (draws: ParitySampler)
//   PTRenderer::new + Renderer::render (pt.rs)         arendur::PTRenderer
//   Composable::intersect_ray / can_intersect          arendur::Scene::intersect_ray / can_intersect (batched too)
#pragma once
#include <array>
#include <cmath>
#include <limits>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>
#include "../../../include/arn_host.h"

namespace arendur {

typedef float Float;                                      // geometry/foundamental.rs:15
typedef std::array<Float, 16> Matrix4f;                   // column-major, like cgmath
inline Matrix4f identity() { return Matrix4f{1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1}; }
inline Matrix4f from_translation(Float x, Float y, Float z) { Matrix4f m = identity(); m[12] = x; m[13] = y; m[14] = z; return m; }

struct Panic : std::runtime_error { int code; Panic(int c, const std::string& m) : std::runtime_error(m), code(c) {} };

enum class BVHStrategy { SAH = ARN_BVH_SAH, MiddleCount = ARN_BVH_MIDDLECOUNT, MidPoint = ARN_BVH_MIDPOINT };

struct RGBSpectrumf { Float r, g, b; };

struct Material {                                          // material/*.rs with ConstantTexture inputs
    arn_material m;
    static Material matte(RGBSpectrumf kd, Float sigma) { Material x{}; x.m.type = ARN_MAT_MATTE; set(x.m.kd, kd); x.m.sigma = sigma; return x; }
    static Material plastic(RGBSpectrumf kd, RGBSpectrumf ks, Float roughness) { Material x{}; x.m.type = ARN_MAT_PLASTIC; set(x.m.kd, kd); set(x.m.ks, ks); x.m.roughness = roughness; return x; }
    static Material glass(RGBSpectrumf kd, RGBSpectrumf ks, Float roughness, Float eta) { Material x = plastic(kd, ks, roughness); x.m.type = ARN_MAT_GLASS; x.m.eta = eta; return x; }
    static Material translucent(RGBSpectrumf kd, RGBSpectrumf ks, Float roughness, Float dissolve) { Material x = plastic(kd, ks, roughness); x.m.type = ARN_MAT_TRANSLUCENT; x.m.dissolve = dissolve; return x; }
private:
    static void set(float* d, RGBSpectrumf s) { d[0] = s.r; d[1] = s.g; d[2] = s.b; }
};

struct Sphere {                                            // shape/sphere.rs:19-31
    Float radius, zmin, zmax, phimax;
    static Sphere make(Float radius, Float zmin, Float zmax, Float phimax) { return Sphere{radius, zmin, zmax, phimax}; }   // Sphere::new
    static Sphere full(Float radius) { return Sphere{radius, -radius, radius, 2.0f * 3.14159265358979323846f}; }          // Sphere::full
};

struct RawRay {                                            // geometry/ray.rs:64-69
    Float origin[3], dir[3], tmax;
    static RawRay from_od(const Float o[3], const Float d[3]) { RawRay r; for (int i = 0; i < 3; i++) { r.origin[i] = o[i]; r.dir[i] = d[i]; } r.tmax = std::numeric_limits<Float>::infinity(); return r; }
};
struct Hit { int prim_id; Float t; bool is_some() const { return prim_id >= 0; } };

// lighting::pointlights::{PointLight, SpotLight}, lighting::distantlight::DistantLight — same constructors
struct Light {
    arn_analytic_light l;
    static Light point(const Float pos[3], const RGBSpectrumf& intensity) {                       // PointLight::new
        Light x{}; Float i[3] = {intensity.r, intensity.g, intensity.b};
        int rc = arn_point_light_make(pos, i, &x.l); if (rc != ARN_OK) throw Panic(rc, "PointLight::new"); return x;
    }
    static Light spot(const Float pos[3], const Float towards[3], const RGBSpectrumf& intensity, Float total_angle, Float start_falloff_angle) {   // SpotLight::new
        Light x{}; Float i[3] = {intensity.r, intensity.g, intensity.b};
        int rc = arn_spot_light_make(pos, towards, i, total_angle, start_falloff_angle, &x.l);
        if (rc != ARN_OK) throw Panic(rc, arn_hscene_last_error(nullptr));                         // the assert!s of SpotLight::new
        return x;
    }
    // DistantLight::new(intensity, dir) followed by set_world_bounds(): pass the bounding-sphere radius (the reference
    // leaves it infinite until set_world_bounds is called, which makes the light sample NaN)
    static Light distant(const RGBSpectrumf& intensity, const Float dir[3], Float world_radius) {
        Light x{}; Float i[3] = {intensity.r, intensity.g, intensity.b};
        int rc = arn_distant_light_make(i, dir, world_radius, &x.l); if (rc != ARN_OK) throw Panic(rc, "DistantLight::new"); return x;
    }
};

// The `Vec<ComponentPointer>` + `lights` that arencli assembles (examples/arencli.rs:88-194).
class Components {
public:
    Components() { check(arn_hscene_create(&h_)); }
    ~Components() { arn_hscene_destroy(h_); }
    Components(const Components&) = delete; Components& operator=(const Components&) = delete;
    int add_material(const Material& m) { return check(arn_hscene_add_material(h_, &m.m)); }
    // TriangleMesh::from_model_transformed + one TriangleInstance per face
    int push_mesh(const std::vector<Float>& positions, const std::vector<uint32_t>& indices, int material, const Matrix4f* transform = nullptr,
                  const std::vector<Float>* normals = nullptr, const std::vector<Float>* uvs = nullptr) {
        return check(arn_hscene_add_mesh(h_, positions.data(), (uint32_t)(positions.size() / 3), indices.data(), (uint32_t)indices.size(),
                                         normals ? normals->data() : nullptr, uvs ? uvs->data() : nullptr, transform ? transform->data() : nullptr, (uint32_t)material));
    }
    // ShapedPrimitive::new(sphere, material, emission) [wrapped in TransformedComposable::new(.., transform, inverse)]
    int push_shaped(const Sphere& s, int material, const RGBSpectrumf* emission = nullptr, const Matrix4f* transform = nullptr) {
        Float e[3]; if (emission) { e[0] = emission->r; e[1] = emission->g; e[2] = emission->b; }
        return check(arn_hscene_add_sphere(h_, s.radius, s.zmin, s.zmax, s.phimax, (uint32_t)material, emission ? e : nullptr, transform ? transform->data() : nullptr));
    }
    // `lights.push(light.to_arc())` (examples/arencli.rs:95-98): ahead of the emissive primitives in Scene.lights
    int push_light(const Light& light) { return check(arn_hscene_add_light(h_, &light.l)); }
    arn_hscene* raw() { return h_; }
    int check(int rc) const { if (rc < 0) throw Panic(rc, arn_hscene_last_error(h_)); return rc; }
private:
    arn_hscene* h_ = nullptr;
};

// component::load_obj(path, transform) -> Result<Vec<ComponentPointer>, LoadError>: pushes into `out`, returns #triangles
inline int load_obj(Components& out, const std::string& path, const Matrix4f& transform) {
    return out.check(arn_hscene_load_obj(out.raw(), path.c_str(), transform.data()));
}

// sample::filters::{BoxFilter, TriangleFilter, GaussianFilter, MitchellFilter, LanczosSincFilter}::new
struct Filter {
    uint32_t kind; Float rx, ry, a, b;
    static Filter box(Float rx, Float ry) { return check({ARN_FILTER_BOX, rx, ry, 0.f, 0.f}); }
    static Filter triangle(Float rx, Float ry) { return check({ARN_FILTER_TRIANGLE, rx, ry, 0.f, 0.f}); }
    static Filter gaussian(Float alpha, Float rx, Float ry) { return check({ARN_FILTER_GAUSSIAN, rx, ry, alpha, 0.f}); }
    static Filter mitchell(Float rx, Float ry, Float b, Float c) { return check({ARN_FILTER_MITCHELL, rx, ry, b, c}); }
    static Filter lanczos(Float rx, Float ry, Float tau) { if (!(tau > 0.f)) throw Panic(ARN_E_INVALID, "assertion failed: tau > 0.0"); return check({ARN_FILTER_LANCZOS, rx, ry, tau, 0.f}); }
private:
    static Filter check(Filter f) { if (!(f.rx > 0.f) || !(f.ry > 0.f)) throw Panic(ARN_E_INVALID, "assertion failed: radius > 0.0"); return f; }
};

struct Film {                                              // filming/film.rs:38-45
    arn_film f;
    // a deserialised film: Lanczos(tau 3) whatever the radius (film.rs:42,47-51)
    static Film make(uint32_t res_x, uint32_t res_y, Float filter_radius = 4.f) {
        Film x{}; x.f.res_x = res_x; x.f.res_y = res_y; x.f.crop_max_x = (int32_t)res_x; x.f.crop_max_y = (int32_t)res_y;
        x.f.filter_radius_x = x.f.filter_radius_y = filter_radius; return x;
    }
    // Film::new(resolution, crop_window = whole frame, filter) (film.rs:55-80)
    static Film make(uint32_t res_x, uint32_t res_y, const Filter& filter) {
        Film x = make(res_x, res_y); x.f.filter_radius_x = filter.rx; x.f.filter_radius_y = filter.ry;
        x.f.filter_kind = filter.kind; x.f.filter_a = filter.a; x.f.filter_b = filter.b; return x;
    }
};

struct PerspecCam {                                        // filming/perspective.rs:25-38
    arn_camera cam; Film film;
    // PerspecCam::new(parent_view, screen, znear, zfar, fov, lens, film)
    static PerspecCam make(const Matrix4f& parent_view, const Float screen[4], Float znear, Float zfar, Float fov, const Float* lens, const Film& film) {
        PerspecCam c{}; c.film = film;
        int rc = arn_camera_make(parent_view.data(), screen, znear, zfar, fov, lens ? 1 : 0, lens ? lens[0] : 0.f, lens ? lens[1] : 0.f, (Float)film.f.res_x, (Float)film.f.res_y, &c.cam);
        if (rc != ARN_OK) throw Panic(rc, arn_hscene_last_error(nullptr));     // "matrix inversion failure" / assert!(znear < zfar)
        return c;
    }
};

struct OrthoCam {                                          // filming/ortho.rs:19-28; the record PTRenderer takes is shared
    // OrthoCam::new(view_parent, screen, znear, zfar, lens, film) — view_parent, unlike PerspecCam::new
    static PerspecCam make(const Matrix4f& view_parent, const Float screen[4], Float znear, Float zfar, const Float* lens, const Film& film) {
        PerspecCam c{}; c.film = film;
        int rc = arn_ortho_camera_make(view_parent.data(), screen, znear, zfar, lens ? 1 : 0, lens ? lens[0] : 0.f, lens ? lens[1] : 0.f, (Float)film.f.res_x, (Float)film.f.res_y, &c.cam);
        if (rc != ARN_OK) throw Panic(rc, arn_hscene_last_error(nullptr));
        return c;
    }
};

struct StrataSampler { arn_sampler s; static StrataSampler make(uint32_t sampledx, uint32_t sampledy, uint32_t ndim, uint32_t seed = 0) { return StrataSampler{{sampledx, sampledy, ndim, seed, ARN_SAMPLER_PARITY}}; }
    StrataSampler as_intended() const { StrataSampler r = *this; r.s.mode = ARN_SAMPLER_STRATIFIED; return r; } };   // the stratification the reference meant to have (SURVEY A-17)

class Device {                                             // one GPU
public:
    explicit Device(int index = 0) { int rc = arn_ctx_create(index, &c_); if (rc != ARN_OK) throw Panic(rc, arn_last_error(nullptr)); }
    ~Device() { arn_ctx_destroy(c_); }
    Device(const Device&) = delete; Device& operator=(const Device&) = delete;
    arn_ctx* raw() { return c_; }
private:
    arn_ctx* c_ = nullptr;
};

// BVH::new(&components, strategy): builds the reference's tree over the component list.
struct BVH {
    static void build(Components& comps, BVHStrategy strategy = BVHStrategy::SAH) { comps.check(arn_hscene_build(comps.raw(), (int)strategy)); }
};

// Scene::new(lights, Arc::new(bvh)) resident on a device.
class Scene {
public:
    Scene(Device& dev, Components& comps) : dev_(dev) {
        const arn_scene_desc* d = arn_hscene_desc(comps.raw());
        if (!d) throw Panic(ARN_E_INVALID, "Scene::new: call BVH::build first");
        int rc = arn_scene_upload(dev.raw(), d, &s_); if (rc != ARN_OK) throw Panic(rc, arn_last_error(dev.raw()));
    }
    ~Scene() { arn_scene_destroy(s_); }
    Scene(const Scene&) = delete; Scene& operator=(const Scene&) = delete;
    // Composable::intersect_ray on the aggregate: updates ray.tmax on a hit (component/mod.rs:27-31)
    Hit intersect_ray(RawRay& ray) {
        arn_ray r; for (int i = 0; i < 3; i++) { r.o[i] = ray.origin[i]; r.d[i] = ray.dir[i]; } r.tmax = ray.tmax;
        arn_hit h; check(arn_intersect_closest(s_, &r, 1, &h));
        if (h.prim_id >= 0) ray.tmax = h.t;
        return Hit{h.prim_id, h.t};
    }
    bool can_intersect(const RawRay& ray) {
        arn_ray r; for (int i = 0; i < 3; i++) { r.o[i] = ray.origin[i]; r.d[i] = ray.dir[i]; } r.tmax = ray.tmax;
        uint8_t o = 0; check(arn_intersect_any(s_, &r, 1, &o)); return o != 0;
    }
    void intersect_rays(const std::vector<arn_ray>& rays, std::vector<arn_hit>& hits) { hits.resize(rays.size()); check(arn_intersect_closest(s_, rays.data(), rays.size(), hits.data())); }
    arn_scene* raw() { return s_; }
    void check(int rc) { if (rc != ARN_OK) throw Panic(rc, arn_last_error(dev_.raw())); }
private:
    Device& dev_; arn_scene* s_ = nullptr;
};

// PTRenderer::new(sampler, camera, filename, max_depth, multithreaded) + Renderer::render(&scene)
class PTRenderer {
public:
    PTRenderer(const StrataSampler& sampler, const PerspecCam& camera, const std::string& filename, size_t max_depth, bool /*multithreaded*/)
        : sampler_(sampler), camera_(camera), filename_(filename) {
        prm_ = arn_pt_params{}; prm_.max_depth = (uint32_t)max_depth; prm_.min_depth = (uint32_t)(max_depth / 2); prm_.rr_threshold = 0.05f;   // pt.rs:47-48
        prm_.tiles_x = prm_.tiles_y = 16; prm_.world_size = 1;                                                                                  // pt.rs:131
    }
    // renders, merges the tiles and saves the PNG (pt.rs:128-176); returns the accumulators (sum r,g,b, weight) per pixel
    const std::vector<Float>& render(Scene& scene, arn_stats* stats = nullptr) {
        const arn_film& f = camera_.film.f;
        size_t w = (size_t)(f.crop_max_x - f.crop_min_x), h = (size_t)(f.crop_max_y - f.crop_min_y);
        film_.assign(w * h * 4, 0.f);
        scene.check(arn_render_pt(scene.raw(), &camera_.cam, &f, &sampler_.s, &prm_, film_.data(), stats));
        if (!filename_.empty() && arn_save_png(filename_.c_str(), film_.data(), (uint32_t)w, (uint32_t)h) != ARN_OK)
            std::fprintf(stderr, "Path tracing result saving at %s failed\n", filename_.c_str());     // warn!, pt.rs:173
        return film_;
    }
private:
    StrataSampler sampler_; PerspecCam camera_; std::string filename_; arn_pt_params prm_; std::vector<Float> film_;
};

}  // namespace arendur
