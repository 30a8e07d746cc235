// Reads like the reference's own tests (src/shape/tests.rs, src/geometry/tests.rs style) but goes through
// the C++ mirror API (arendur_b200/csrc/host/arendur.hpp) and the GPU.  Built and run by tests/test_cpp_mirror.py.
#include <cassert>
#include <cstdio>
#include <cmath>
#include "../../arendur_b200/csrc/host/arendur.hpp"
using namespace arendur;

static void test_sy_intersect(Device& dev) {               // shape/tests.rs:21-49
    Components comps;
    int m = comps.add_material(Material::matte({0.5f, 0.5f, 0.5f}, 0.f));
    comps.push_shaped(Sphere::full(1.0f), m);
    BVH::build(comps);
    Scene scene(dev, comps);
    const Float o[3] = {0.f, 0.f, -10.f}, d[3] = {0.f, 0.f, 1.f};
    RawRay ray = RawRay::from_od(o, d);
    assert(scene.can_intersect(ray));
    Hit h = scene.intersect_ray(ray);
    assert(h.is_some() && std::fabs(h.t - 9.0f) < 1e-5f && ray.tmax == h.t);   // tmax updated on a hit
    const Float away[3] = {0.f, 0.f, -1.f};
    RawRay miss = RawRay::from_od(o, away);
    assert(!scene.can_intersect(miss) && !scene.intersect_ray(miss).is_some());
}

static void test_panics() {
    bool threw = false;
    try { Components c; int m = c.add_material(Material::matte({1, 1, 1}, 0)); c.push_shaped(Sphere::make(-1.f, -1.f, 1.f, 1.f), m); }
    catch (const Panic& p) { threw = true; assert(std::string(p.what()).find("radius should be positive") != std::string::npos); }   // assert! in Sphere::new
    assert(threw);
    threw = false;
    try { Components c; BVH::build(c); } catch (const Panic&) { threw = true; }                                                      // recursive_build asserts len != 0
    assert(threw);
}

static void test_render_small(Device& dev) {
    Components comps;
    int wall = comps.add_material(Material::matte({0.7f, 0.7f, 0.7f}, 0.f));
    int lm = comps.add_material(Material::matte({0.5f, 0.5f, 0.5f}, 3.0f));
    comps.push_mesh({-2, -2, 4, 2, -2, 4, 2, 2, 4, -2, 2, 4}, {0, 1, 2, 0, 2, 3}, wall);
    RGBSpectrumf e{15.5f, 10.5f, 5.5f}; Matrix4f t = from_translation(0.f, 0.f, -3.f);
    comps.push_shaped(Sphere::make(1.5f, -2.f, 2.f, 6.28f), lm, &e, &t);
    BVH::build(comps);
    Scene scene(dev, comps);
    const Float screen[4] = {-1.f, -1.f, 1.f, 1.f};
    PerspecCam cam = PerspecCam::make(identity(), screen, 0.1f, 1000.f, 1.2707964f, nullptr, Film::make(64, 64));
    PTRenderer r(StrataSampler::make(2, 2, 8), cam, "", 8, true);
    arn_stats st;
    const std::vector<Float>& film = r.render(scene, &st);
    assert(st.camera_rays == 64 * 64 * 4 && st.invalid_samples == 0 && st.kernel_launches > 0);
    double lum = 0; for (size_t i = 0; i < film.size(); i += 4) if (film[i + 3] != 0.f) lum += film[i] / film[i + 3];
    assert(lum > 0.0);                                      // the wall is lit by the sphere behind the camera
}

// a wall lit only by scene-file style lights: SpotLight::new + PointLight::new, no emissive primitive
static void test_render_analytic_lights(Device& dev) {
    Components comps;
    int wall = comps.add_material(Material::matte({0.7f, 0.7f, 0.7f}, 0.f));
    comps.push_mesh({-2, -2, 4, 2, -2, 4, 2, 2, 4, -2, 2, 4}, {0, 1, 2, 0, 2, 3}, wall);
    const Float pos[3] = {0.f, 0.f, 1.f}, towards[3] = {0.f, 0.f, 1.f}, ppos[3] = {1.f, 1.f, 2.f};
    comps.push_light(Light::spot(pos, towards, {8.f, 8.f, 8.f}, 0.6f, 0.3f));
    comps.push_light(Light::point(ppos, {1.f, 2.f, 3.f}));
    bool threw = false;
    try { Light::spot(pos, towards, {1, 1, 1}, 0.2f, 0.3f); } catch (const Panic&) { threw = true; }     // assert!(total_angle > start_falloff_angle)
    assert(threw);
    BVH::build(comps);
    Scene scene(dev, comps);
    const Float screen[4] = {-1.f, -1.f, 1.f, 1.f};
    PerspecCam cam = PerspecCam::make(identity(), screen, 0.1f, 1000.f, 1.2707964f, nullptr, Film::make(32, 32));
    PTRenderer r(StrataSampler::make(2, 2, 8), cam, "", 3, true);
    arn_stats st;
    const std::vector<Float>& film = r.render(scene, &st);
    assert(st.shadow_rays > 0 && st.mis_rays == 0 && st.invalid_samples == 0);
    // OrthoCam + Film::new with a Mitchell filter
    PerspecCam ocam = OrthoCam::make(identity(), screen, 0.1f, 1000.f, nullptr, Film::make(32, 32, Filter::mitchell(2.f, 2.f, 1.f / 3.f, 1.f / 3.f)));
    PTRenderer r2(StrataSampler::make(1, 1, 8), ocam, "", 2, true);
    const std::vector<Float>& film2 = r2.render(scene, &st);
    assert(st.camera_rays == 32 * 32 && film2.size() == 32 * 32 * 4);
    threw = false;
    try { Filter::box(0.f, 1.f); } catch (const Panic&) { threw = true; }                                   // assert!(radius.x > 0.0)
    assert(threw);
    double lum = 0; for (size_t i = 0; i < film.size(); i += 4) if (film[i + 3] != 0.f) lum += film[i] / film[i + 3];
    assert(lum > 0.0);
}

int main() {
    test_panics();
    Device dev(0);
    test_sy_intersect(dev);
    test_render_small(dev);
    test_render_analytic_lights(dev);
    std::printf("mirror API tests: OK\n");
    return 0;
}
