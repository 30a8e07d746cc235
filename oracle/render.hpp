// ORACLE — test infrastructure only (see geom.hpp header).
// CPU restatement of arendur's camera, film, area lights and path-tracing integrator:
//   src/filming/{projective,perspective,film}.rs, src/component/{shape,transformed}.rs (Light impls),
//   src/lighting/mod.rs, src/renderer/{pt,scene}.rs
#pragma once
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <atomic>
#include <vector>
#include "scene.hpp"
#include "shading.hpp"

namespace orc {

// ------------------------------------------------------------------ camera
// PerspecCam::perspective_transform (filming/perspective.rs:93-107)
inline M4 perspective_transform(Float fov, Float znear, Float zfar) {
    M4 persp; std::memset(&persp, 0, sizeof persp);
    persp.m[0] = 1.f; persp.m[5] = 1.f;
    persp.m[10] = zfar / (zfar - znear); persp.m[11] = 1.f;
    persp.m[14] = -zfar * znear / (zfar - znear); persp.m[15] = 0.f;
    Float inv_tan = 1.f / ftan(fov * 0.5f);
    return m4_mul(m4_from_nonuniform_scale(inv_tan, inv_tan, 1.f), persp);
}
// ProjCameraInfo::new (filming/projective.rs:24-45) + PerspecCam::new (perspective.rs:42-90)
inline bool camera_make(const Float* parent_view16, const Float screen[4] /*pminx,pminy,pmaxx,pmaxy*/,
                        Float znear, Float zfar, Float fov, int has_lens, Float lens_r, Float lens_d,
                        Float res_x, Float res_y, arn_camera* out) {
    M4 parent_view = m4_from_cols(parent_view16);
    M4 view_parent; if (!m4_invert(parent_view, &view_parent)) return false;
    M4 view_screen = perspective_transform(fov, znear, zfar);
    M4 raster_screen = m4_mul(m4_from_translation(v3(screen[0], screen[3], 0.f)),
                              m4_from_nonuniform_scale((screen[2] - screen[0]) / res_x, (screen[1] - screen[3]) / res_y, 1.f));
    M4 inv_vs; if (!m4_invert(view_screen, &inv_vs)) return false;
    M4 screen_raster; if (!m4_invert(raster_screen, &screen_raster)) return false;   // .unwrap() in the source
    M4 raster_view = m4_mul(inv_vs, raster_screen);
    std::memcpy(out->raster_view, raster_view.m, 64);
    std::memcpy(out->view_parent, view_parent.m, 64);
    out->has_lens = has_lens ? 1u : 0u; out->lens_radius = lens_r; out->focal_distance = lens_d; out->ortho = 0u;
    return true;
}
// OrthoCam::new (filming/ortho.rs:30-55) with ortho_transform (:57-66)
inline bool ortho_camera_make(const Float* view_parent16, const Float screen[4], Float znear, Float zfar,
                              int has_lens, Float lens_r, Float lens_d, Float res_x, Float res_y, arn_camera* out) {
    M4 view_parent = m4_from_cols(view_parent16);
    M4 parent_view; if (!m4_invert(view_parent, &parent_view)) return false;
    M4 view_screen = m4_mul(m4_from_nonuniform_scale(1.f, 1.f, 1.f / (zfar - znear)), m4_from_translation(v3(0.f, 0.f, -znear)));
    M4 raster_screen = m4_mul(m4_from_translation(v3(screen[0], screen[3], 0.f)),
                              m4_from_nonuniform_scale((screen[2] - screen[0]) / res_x, (screen[1] - screen[3]) / res_y, 1.f));
    M4 inv_vs; if (!m4_invert(view_screen, &inv_vs)) return false;
    M4 screen_raster; if (!m4_invert(raster_screen, &screen_raster)) return false;
    M4 raster_view = m4_mul(inv_vs, raster_screen);
    std::memcpy(out->raster_view, raster_view.m, 64);
    std::memcpy(out->view_parent, view_parent.m, 64);
    out->has_lens = has_lens ? 1u : 0u; out->lens_radius = lens_r; out->focal_distance = lens_d; out->ortho = 1u;
    return true;
}
// PerspecCam::generate_path_differential (perspective.rs:292-320), main ray only
inline RawRay camera_generate(const arn_camera& cam, V2 pfilm, V2 plens) {
    M4 rv = m4_from_cols(cam.raster_view), vp = m4_from_cols(cam.view_parent);
    V3 pview = transform_point(rv, v3(pfilm.x, pfilm.y, 0.f));
    RawRay ray = cam.ortho ? ray_from_od(pview, v3(0.f, 0.f, 1.f))                 // OrthoCam::generate_path (ortho.rs:180-198)
                           : ray_from_od(v3(0.f, 0.f, 0.f), normalize(pview));
    if (cam.has_lens) {                                                            // same lines in both cameras (sic for ortho: the origin forgets pview)
        V2 pl = cam.lens_radius * sample_concentric_disk(plens);
        Float ft = cam.focal_distance / ray.dir.z;
        V3 pfocus = ray_evaluate(ray, ft);
        V3 new_origin = v3(pl.x, pl.y, 0.f);
        ray = ray_from_od(new_origin, normalize(pfocus - new_origin));
    }
    return ray_apply_transform(ray, vp);
}

// PerspecCam / OrthoCam::generate_path_differential with the offset rays (perspective.rs:292-320 with :68-73, ortho.rs:203-228 with :46-47)
inline RayDifferential camera_generate_differential(const arn_camera& cam, V2 pfilm, V2 plens) {
    M4 rv = m4_from_cols(cam.raster_view), vp = m4_from_cols(cam.view_parent);
    V3 pview = transform_point(rv, v3(pfilm.x, pfilm.y, 0.f));
    RawRay ray = cam.ortho ? ray_from_od(pview, v3(0.f, 0.f, 1.f)) : ray_from_od(v3(0.f, 0.f, 0.f), normalize(pview));
    if (cam.has_lens) {
        V2 pl = cam.lens_radius * sample_concentric_disk(plens);
        Float ft = cam.focal_distance / ray.dir.z;
        V3 pfocus = ray_evaluate(ray, ft);
        V3 new_origin = v3(pl.x, pl.y, 0.f);
        ray = ray_from_od(new_origin, normalize(pfocus - new_origin));
    }
    RayDifferential r; r.ray = ray; r.has_diffs = true;
    if (cam.ortho) {
        V3 dx = transform_vector(rv, v3(1.f, 0.f, 0.f)), dy = transform_vector(rv, v3(0.f, 1.f, 0.f));
        r.rx = ray_from_od(ray.origin + dx, ray.dir); r.ry = ray_from_od(ray.origin + dy, ray.dir);
    } else {
        V3 or2v = transform_point(rv, v3(1.f, 0.f, 0.f));                        // sic: dx = T(1,0,0) - T(1,0,0) = 0 (perspective.rs:68-73)
        V3 dx = transform_point(rv, v3(1.f, 0.f, 0.f)) - or2v, dy = transform_point(rv, v3(0.f, 1.f, 0.f)) - or2v;
        r.rx = ray_from_od(ray.origin, normalize(pview + dx)); r.ry = ray_from_od(ray.origin, normalize(pview + dy));   // "TODO: account for lens"
    }
    r.ray = ray_apply_transform(r.ray, vp); r.rx = ray_apply_transform(r.rx, vp); r.ry = ray_apply_transform(r.ry, vp);   // transform_ray_differential
    return r;
}

// ------------------------------------------------------------------ area light = emissive sphere primitive
struct LightSample { RGB radiance; Float pdf; V3 pfrom, pto; };

inline RGB sphere_emission(const arn_sphere& sp) { return rgb(sp.emission[0], sp.emission[1], sp.emission[2]); }
// ShapedPrimitive::evaluate_path (component/shape.rs:91-103), in the sphere's local frame
inline RGB shaped_evaluate_path(const arn_sphere& sp, V3 pos, V3 dir) {
    if (sp.emissive) {
        V3 p = pos + dir;
        RawRay ray = ray_from_od(p, -dir);
        Float t; SurfaceInteraction si;
        if (sphere_intersect(sp, ray, &t, &si)) return sphere_emission(sp);
    }
    return grey(0.f);
}
// Light::evaluate_path via TransformedComposable (transformed.rs:113-117)
inline RGB light_evaluate_path(const arn_sphere& sp, V3 pos, V3 dir) {
    if (sp.has_transform) {
        M4 pl = m4_from_cols(sp.parent_local);
        pos = transform_point(pl, pos); dir = transform_vector(pl, dir);
    }
    return shaped_evaluate_path(sp, pos, dir);
}
// ShapedPrimitive::evaluate_sampled (component/shape.rs:108-130) (+ transformed.rs:120-124)
inline LightSample light_evaluate_sampled(const arn_sphere& sp, V3 pos, V2 sample) {
    M4 pl, lp;
    if (sp.has_transform) { pl = m4_from_cols(sp.parent_local); lp = m4_from_cols(sp.local_parent); pos = transform_point(pl, pos); }
    V3 l_pos, l_norm; Float l_pdf;
    sphere_sample_wrt(sp, pos, sample, &l_pos, &l_norm, &l_pdf);
    LightSample ret; ret.radiance = grey(0.f); ret.pdf = l_pdf; ret.pfrom = l_pos; ret.pto = pos;
    if (sp.emissive) {
        V3 ldir = pos - l_pos;
        if (dot(ldir, l_norm) > 0.f) {
            RawRay ray = ray_from_od(pos, -ldir);
            Float t; SurfaceInteraction si;
            if (sphere_intersect(sp, ray, &t, &si)) ret.radiance = sphere_emission(sp);
        }
    }
    if (sp.has_transform) { ret.pfrom = transform_point(lp, ret.pfrom); ret.pto = transform_point(lp, ret.pto); }   // lighting/mod.rs:136-146
    return ret;
}
// Light::pdf (component/shape.rs:155-157, transformed.rs:142-146)
inline Float light_pdf(const arn_sphere& sp, V3 pos, V3 wi) {
    if (sp.has_transform) { M4 pl = m4_from_cols(sp.parent_local); pos = transform_point(pl, pos); wi = transform_vector(pl, wi); }
    return sphere_pdf_wrt(sp, pos, wi);
}
// ---- PointLight / SpotLight / DistantLight (lighting/pointlights.rs, lighting/distantlight.rs)
// SpotLight::falloff (pointlights.rs:147-158): cos_theta = parent_local.transform_vector(dir).z
inline Float spot_falloff(const arn_analytic_light& l, V3 dir) {
    M4 pl = m4_from_cols(l.parent_local);
    Float cos_theta = transform_vector(pl, dir).z;
    if (cos_theta < l.cost) return 0.f;
    else if (cos_theta > l.cosf) return 1.f;
    else {
        Float delta = (cos_theta - l.cost) / (l.cosf - l.cost);
        Float delta2 = delta * delta;
        return delta2 * delta2;
    }
}
// Light::evaluate_sampled: pointlights.rs:50-61 (Point), :181-194 (Spot), distantlight.rs:68-80 (Distant)
inline LightSample analytic_evaluate_sampled(const arn_analytic_light& l, V3 pos) {
    LightSample ret; ret.pdf = 1.0f; ret.pto = pos;
    RGB intensity = rgb(l.intensity[0], l.intensity[1], l.intensity[2]);
    if (l.type == ARN_LIGHT_POINT) {
        ret.pfrom = v3(l.pos[0], l.pos[1], l.pos[2]);
        ret.radiance = intensity / magnitude2(ret.pto - ret.pfrom);
    } else if (l.type == ARN_LIGHT_SPOT) {
        ret.pfrom = v3(l.pos[0], l.pos[1], l.pos[2]);
        V3 dir = ret.pto - ret.pfrom;
        Float mag2 = magnitude2(dir);
        ret.radiance = intensity * spot_falloff(l, dir / std::sqrt(mag2)) / mag2;
    } else {
        V3 d = v3(l.dir[0], l.dir[1], l.dir[2]);
        ret.radiance = intensity;
        ret.pfrom = pos + d * (-2.0f * l.world_radius);
    }
    return ret;
}
inline bool analytic_is_delta(const arn_analytic_light& l) { return l.type != ARN_LIGHT_DISTANT; }   // LIGHT_DPOS vs LIGHT_INFINITE
// Light::power: pointlights.rs:79-81, :222-226, distantlight.rs:108-110
inline RGB analytic_power(const arn_analytic_light& l) {
    RGB intensity = rgb(l.intensity[0], l.intensity[1], l.intensity[2]);
    if (l.type == ARN_LIGHT_POINT) return intensity * (pi() * 4.0f);
    if (l.type == ARN_LIGHT_SPOT) return intensity * (pi() * 2.0f) * (1.0f - 0.5f * (l.cosf - l.cost));
    return intensity * (l.world_radius * l.world_radius * pi());
}

// Light::power (component/shape.rs:160-167): mean * area * pi
inline RGB light_power(const arn_sphere& sp) {
    if (!sp.emissive) return grey(0.f);
    return sphere_emission(sp) * sphere_surface_area(sp) * pi();
}
// SurfaceInteraction::le (geometry/interaction.rs:254-261)
inline RGB si_le(const Scene& s, const SurfaceInteraction& si, V3 dir) {
    if (si.primitive_hit >= 0 && s.prim_is_sphere((uint32_t)si.primitive_hit)) {
        const arn_sphere& sp = s.spheres[s.prim_index((uint32_t)si.primitive_hit)];
        if (sp.emissive) return light_evaluate_path(sp, si.basic.pos, dir);
    }
    return grey(0.f);   // triangles never carry a lighting profile (component/mod.rs:177)
}

struct RayStats { uint64_t camera = 0, extend = 0, shadow = 0, mis = 0, invalid = 0, extend_bounce = 0; TraversalCounters trav; };

// LightSample::occluded (lighting/mod.rs:125-133) -> Composable::can_intersect default (component/mod.rs:35-38)
inline bool ls_occluded(const Scene& s, const LightSample& ls, RayStats* st) {
    Float eps = epsilon() * 2.0f;
    V3 dir = ls.pto - ls.pfrom;
    V3 pfrom = ls.pfrom + dir * eps;
    V3 pto = ls.pto + (-dir * eps);
    RawRay ray = ray_spawn(pfrom, pto);
    if (st) st->shadow++;
    return bvh_intersect(s, ray, nullptr, nullptr, st ? &st->trav : nullptr, false);
}

// Scene::evaluate_direct (renderer/scene.rs:83-167)
// Debugging aid of the test infrastructure: ARN_ORACLE_TRACE="x,y,s" makes the render print what that camera sample does, bounce by bounce.
inline thread_local bool g_trace = false;
#define ORC_TRACE(...) do { if (g_trace) std::fprintf(stderr, __VA_ARGS__); } while (0)

inline RGB evaluate_direct(const Scene& s, uint32_t light_comp, V2 ulight, V2 uscattering,
                           const SurfaceInteraction& si, const Bsdf& bsdf, RayStats* st) {
    if (light_comp & ARN_LIGHT_ANALYTIC) {
        const arn_analytic_light& al = s.analytic_lights[light_comp & ~ARN_LIGHT_ANALYTIC];
        RGB ret = grey(0.f);
        LightSample ls = analytic_evaluate_sampled(al, si.basic.pos);
        V3 wi = normalize(ls.pfrom - ls.pto);
        bool no_effect = ls.pdf == 0.f || is_black(ls.radiance);
        if (!no_effect) {
            RGB f = bsdf_evaluate(bsdf, si.basic.wo, wi, BXDF_ALL) * std::fabs(dot(wi, si.shading_norm));
            Float spdf = bsdf_pdf(bsdf, si.basic.wo, wi, BXDF_ALL);
            if (spdf == 0.f) f = grey(0.f);
            if (!is_black(f) && ls_occluded(s, ls, st)) f = grey(0.f);
            if (analytic_is_delta(al)) ret = ret + ls.radiance * f / ls.pdf;                       // scene.rs:107-115
            else ret = ret + ls.radiance * f * power_heuristic(ls.pdf, spdf) / ls.pdf;            // :116-124
        }
        // scene.rs:128-165 for a non-delta light without Light::pdf / as_light: the BSDF-sampled ray can only
        // add `li = black` — non-specular lobes stop at `lpdf == 0` (lighting/mod.rs:64-66), specular ones
        // trace a ray whose hit is never `ptr::eq` to this light.  The ray is still traced (and counted).
        if (!analytic_is_delta(al)) {
            Sampled bs = bsdf_evaluate_sampled(bsdf, si.basic.wo, uscattering, BXDF_ALL);
            RGB f = bs.f * std::fabs(dot(bs.wi, si.shading_norm));
            if (!is_black(f) && bs.pdf > 0.f && (bs.type & BXDF_SPECULAR)) {
                RawRay ray = si_spawn_ray(si, bs.wi);
                SurfaceInteraction lsi; int lprim;
                if (st) st->mis++;
                bvh_intersect(s, ray, &lsi, &lprim, st ? &st->trav : nullptr, true);
            }
        }
        return ret;
    }
    const arn_sphere& light = s.spheres[s.prim_index(light_comp)];
    RGB ret = grey(0.f);
    LightSample ls = light_evaluate_sampled(light, si.basic.pos, ulight);
    V3 wi = normalize(ls.pfrom - ls.pto);                                   // LightSample::wi, lighting/mod.rs:118-120
    bool no_effect = ls.pdf == 0.f || is_black(ls.radiance);                // :148-150
    if (!no_effect) {
        RGB f = bsdf_evaluate(bsdf, si.basic.wo, wi, BXDF_ALL) * std::fabs(dot(wi, si.shading_norm));
        Float spdf = bsdf_pdf(bsdf, si.basic.wo, wi, BXDF_ALL);
        if (spdf == 0.f) f = grey(0.f);
        if (!is_black(f) && ls_occluded(s, ls, st)) f = grey(0.f);
        Float weight = power_heuristic(ls.pdf, spdf);                       // area lights are never delta
        RGB addition = ls.radiance * f * weight / ls.pdf;
        ORC_TRACE("    light sample: radiance %.9g %.9g %.9g pdf %.9g wi %.9g %.9g %.9g f %.9g %.9g %.9g spdf %.9g weight %.9g add %.9g %.9g %.9g\n", ls.radiance.x, ls.radiance.y, ls.radiance.z, ls.pdf, wi.x, wi.y, wi.z, f.x, f.y, f.z, spdf, weight, addition.x, addition.y, addition.z);
        ret = ret + addition;
    }
    // sample BSDF with multiple importance sampling
    Sampled bs = bsdf_evaluate_sampled(bsdf, si.basic.wo, uscattering, BXDF_ALL);
    RGB f = bs.f * std::fabs(dot(bs.wi, si.shading_norm));
    if (!is_black(f) && bs.pdf > 0.f) {
        Float weight = 1.f;
        if (!(bs.type & BXDF_SPECULAR)) {
            Float lpdf = light_pdf(light, si.basic.pos, bs.wi);
            ORC_TRACE("    light_pdf %.9g\n", lpdf);
            if (lpdf == 0.f) return ret;
            weight = power_heuristic(bs.pdf, lpdf);
        }
        RawRay ray = si_spawn_ray(si, bs.wi);
        RGB li = grey(0.f);
        SurfaceInteraction lsi; int lprim;
        if (st) st->mis++;
        if (bvh_intersect(s, ray, &lsi, &lprim, st ? &st->trav : nullptr, true)) {
            if ((uint32_t)lprim == light_comp) li = si_le(s, lsi, -bs.wi);  // ptr::eq(light, primitive.as_light())
        }
        ORC_TRACE("    bsdf sample: wi %.9g %.9g %.9g f*cos %.9g %.9g %.9g pdf %.9g type %u weight %.9g hit %d li %.9g %.9g %.9g\n", bs.wi.x, bs.wi.y, bs.wi.z, f.x, f.y, f.z, bs.pdf, (unsigned)bs.type, weight, lprim, li.x, li.y, li.z);
        if (!is_black(li)) {
            RGB addition = f * li * weight / bs.pdf;
            ret = ret + addition;
        }
    }
    return ret;
}

// calculate_lighting (renderer/pt.rs:55-125)
// `rd` (scenes with image textures): the camera's ray differential; compute_dxy / textured materials / spawn_ray_differential
// then run as in the source.  Without textures the differentials cannot reach any result and are not tracked.
inline RGB calculate_lighting(const Scene& s, RawRay ray, ParitySampler& sampler, uint32_t max_depth,
                              uint32_t min_depth, Float rr_threshold, RayStats* st, const RayDifferential* rd_in = nullptr) {
    const bool textured = rd_in != nullptr;
    RayDifferential rd; if (textured) { rd = *rd_in; ray = rd.ray; }
    TexTable tex; tex.textures = s.textures.data(); tex.texels = s.texels.data();
    RGB ret = grey(0.f);
    RGB beta = grey(1.f);
    bool specular_bounce = false;
    uint32_t bounces = 0;
    for (;;) {
        SurfaceInteraction si; int prim;
        if (st) { st->extend++; if (bounces > 0) st->extend_bounce++; }
        if (bvh_intersect(s, ray, &si, &prim, st ? &st->trav : nullptr, true)) {
            if (bounces == 0 || specular_bounce) {
                RGB term = si_le(s, si, -ray.dir);
                ret = ret + beta * term;
            }
            uint32_t mat = s.prim_is_sphere((uint32_t)prim) ? s.spheres[s.prim_index((uint32_t)prim)].material
                                                            : s.meshes[s.tri_mesh[s.prim_index((uint32_t)prim)]].material;
            DxyInfo dxy = dxy_default();
            if (textured) { rd.ray = ray; dxy = compute_dxy(si, rd); }             // pt.rs:80 (the traversal may have rewritten `ray`)
            Bsdf bsdf = textured ? compute_scattering(s.materials[mat], si, &dxy, &tex) : compute_scattering(s.materials[mat], si);
            if (bsdf_have_n(bsdf, BXDF_ALL & ~BXDF_SPECULAR) > 0) {
                // Scene::uniform_sample_one_light (scene.rs:58-66): next(), next_2d(), next_2d()
                uint32_t lidx; Float lightpdf;
                sample_discrete(s.light_func.data(), s.light_cdf.data(), (uint32_t)s.light_func.size(), s.light_func_integral,
                                sampler.next(), &lidx, &lightpdf);
                V2 ulight = sampler.next_2d();
                V2 uscattering = sampler.next_2d();
                ORC_TRACE("  si: wo %.9g %.9g %.9g ng %.9g %.9g %.9g ns %.9g %.9g %.9g dpdu %.9g %.9g %.9g ulight %.9g %.9g uscatter %.9g %.9g\n", si.basic.wo.x, si.basic.wo.y, si.basic.wo.z, si.basic.norm.x, si.basic.norm.y, si.basic.norm.z,
                          si.shading_norm.x, si.shading_norm.y, si.shading_norm.z, si.shading_duv.dpdu.x, si.shading_duv.dpdu.y, si.shading_duv.dpdu.z, ulight.x, ulight.y, uscattering.x, uscattering.y);
                RGB term = evaluate_direct(s, s.light_prims[lidx], ulight, uscattering, si, bsdf, st) / lightpdf;
                ORC_TRACE("  bounce %u prim %d material %u (type %u) pos %.9g %.9g %.9g light %u (comp 0x%x) lightpdf %.9g direct %.9g %.9g %.9g beta %.9g %.9g %.9g\n", bounces, prim, mat, (unsigned)s.materials[mat].type, si.basic.pos.x, si.basic.pos.y, si.basic.pos.z, lidx, s.light_prims[lidx], lightpdf, term.x, term.y, term.z, beta.x, beta.y, beta.z);
                ret = ret + beta * term;
            }
            V3 wo = -ray.dir;
            Sampled bs = bsdf_evaluate_sampled(bsdf, wo, sampler.next_2d(), BXDF_ALL);
            specular_bounce = (bs.type & BXDF_SPECULAR) != 0;
            if (is_black(bs.f) || bs.pdf == 0.f) break;
            beta = beta * (bs.f * (std::fabs(dot(bs.wi, si.shading_norm)) / bs.pdf));
            if (!rgb_valid(beta)) break;
            if (textured) { rd = spawn_ray_differential(si, bs.wi, &dxy); ray = rd.ray; }   // pt.rs:103
            else ray = si_spawn_ray(si, bs.wi);
        } else break;
        bounces += 1;
        if (bounces >= max_depth) break;
        if (rgb_y(beta) < rr_threshold && bounces >= min_depth) {
            Float q = fmax_(rr_threshold, 0.05f);
            if (sampler.next() < q) break;
            beta = beta / (1.f - q);
        }
    }
    return ret;
}

// ------------------------------------------------------------------ film (filming/film.rs)
struct IBox { long x0, y0, x1, y1; };   // BBox2<isize>, max exclusive
inline bool ibox_intersect(const IBox& a, const IBox& b, IBox* o) {           // bbox.rs BBox2::intersect
    o->x0 = a.x0 > b.x0 ? a.x0 : b.x0; o->y0 = a.y0 > b.y0 ? a.y0 : b.y0;
    o->x1 = a.x1 < b.x1 ? a.x1 : b.x1; o->y1 = a.y1 < b.y1 ? a.y1 : b.y1;
    return !(o->x0 > o->x1 || o->y0 > o->y1);
}
struct FilmTile { IBox bounding, sink; std::vector<Float> px; };              // px: 4 floats per sink pixel
// Film::spawn_tiles (film.rs:104-135); tiles in ix-major order
// `ok` turns false where the reference panics: `.intersect(&self.crop_window).unwrap()` (film.rs:129) on a tile that,
// grown by the filter radius, does not meet the crop window — tiles are laid out from (0, 0), NOT from crop.pmin (sic).
inline std::vector<FilmTile> spawn_tiles(const arn_film& film, long nx, long ny, bool* ok = nullptr) {
    if (ok) *ok = true;
    IBox crop = {film.crop_min_x, film.crop_min_y, film.crop_max_x, film.crop_max_y};
    long ex = crop.x1 - crop.x0, ey = crop.y1 - crop.y0;
    long dx = ex / nx, dy = ey / ny;
    long lastx = dx + ex % dx, lasty = dy + ey % dy;
    long rx = (long)film.filter_radius_x, ry = (long)film.filter_radius_y;     // filter_radius.cast()
    std::vector<FilmTile> ret;
    for (long ix = 0; ix < nx; ix++) {
        long cdx = ix == nx - 1 ? lastx : dx;
        for (long iy = 0; iy < ny; iy++) {
            long cdy = iy == ny - 1 ? lasty : dy;
            FilmTile t; t.bounding = IBox{ix * dx, iy * dy, ix * dx + cdx, iy * dy + cdy};
            IBox grown = {t.bounding.x0 - rx, t.bounding.y0 - ry, t.bounding.x1 + rx, t.bounding.y1 + ry};
            if (!ibox_intersect(grown, crop, &t.sink) && ok) *ok = false;
            ret.push_back(t);
        }
    }
    return ret;
}
// FilmTile::add_sample (film.rs:297-319)
inline void tile_add_sample(FilmTile& t, const arn_film& film, V2 pos, RGB spectrum) {
    if (t.px.empty()) t.px.assign((size_t)(t.sink.x1 - t.sink.x0) * (size_t)(t.sink.y1 - t.sink.y0) * 4, 0.f);
    V2 fr = v2(film.filter_radius_x, film.filter_radius_y);
    V2 ceil = pos - fr + v2(0.5f, 0.5f);
    V2 floor = pos + fr - v2(0.5f, 0.5f);
    IBox fb = {(long)ceil.x, (long)ceil.y, (long)floor.x + 1, (long)floor.y + 1};   // cast: truncation toward zero
    // BBox2::new orders the corners
    if (fb.x0 > fb.x1) { long q = fb.x0; fb.x0 = fb.x1; fb.x1 = q; }
    if (fb.y0 > fb.y1) { long q = fb.y0; fb.y0 = fb.y1; fb.y1 = q; }
    IBox rb;
    if (!ibox_intersect(fb, t.sink, &rb)) return;
    long sw = t.sink.x1 - t.sink.x0;
    for (long y = rb.y0; y < rb.y1; y++) for (long x = rb.x0; x < rb.x1; x++) {     // row-major iteration, bbox.rs:603-620
        V2 pixel_pos = v2((Float)x + 0.5f, (Float)y + 0.5f);                         // pidx_to_pcenter :23-28
        V2 offset = pixel_pos - pos;
        Float weight = filter_evaluate(film, offset);
        Float* p = &t.px[((size_t)(x - t.sink.x0) + (size_t)(y - t.sink.y0) * (size_t)sw) * 4];
        RGB c = spectrum * weight;
        p[0] += c.x; p[1] += c.y; p[2] += c.z; p[3] += weight;
    }
}

// PTRenderer::render (renderer/pt.rs:128-176) up to collect_into's merge (film.rs:171-183, 82-101)
// radiance_out (diagnostic) is indexed by the TILE pixel (x, y) in [0, crop width) x [0, crop height).
inline bool render_pt(const Scene& s, const arn_camera& cam, const arn_film& film, const arn_sampler& smp,
                      const arn_pt_params& prm, Float* film_out, RayStats* stats, int nthreads, Float* radiance_out = nullptr) {
    long nx = prm.tiles_x ? prm.tiles_x : 16, ny = prm.tiles_y ? prm.tiles_y : 16;
    bool tiles_ok = true;
    std::vector<FilmTile> tiles = spawn_tiles(film, nx, ny, &tiles_ok);
    if (!tiles_ok) return false;
    uint32_t spp = smp.sampledx * smp.sampledy;
    uint32_t s0 = prm.spp_begin, s1 = prm.spp_end ? prm.spp_end : spp;
    uint32_t world = prm.world_size ? prm.world_size : 1;
    std::atomic<size_t> next_tile(0);
    if (nthreads < 1) nthreads = 1;
    std::vector<RayStats> tstats((size_t)nthreads);
    long trace_x = -1, trace_y = -1, trace_s = -1;
    const bool trace_on = std::getenv("ARN_ORACLE_TRACE") && std::sscanf(std::getenv("ARN_ORACLE_TRACE"), "%ld,%ld,%ld", &trace_x, &trace_y, &trace_s) == 3;
    auto worker = [&](int tid) {
        for (;;) {
            size_t ti = next_tile.fetch_add(1);
            if (ti >= tiles.size()) break;
            const uint32_t sub = prm.partition_subdiv > 1 ? prm.partition_subdiv : 1;     // arn_pt_params.partition_subdiv (include/arn.h)
            if (sub == 1 && ((ti / (size_t)ny) + (ti % (size_t)ny)) % world != prm.rank) continue;   // tile (ix, iy) -> rank (ix + iy) % world: diagonal interleave
            FilmTile& tile = tiles[ti];
            const long tw = tile.bounding.x1 - tile.bounding.x0, thh = tile.bounding.y1 - tile.bounding.y0;
            const long ncx = std::min<long>(sub, tw), ncy = std::min<long>(sub, thh), sdx = ncx > 0 ? tw / ncx : 1, sdy = ncy > 0 ? thh / ncy : 1;
            ParitySampler sampler; sampler.seed = smp.seed; sampler.spp = spp;
            sampler.mode = smp.mode; sampler.sampledx = smp.sampledx; sampler.sampledy = smp.sampledy; sampler.ndim = smp.ndim;
            for (long y = tile.bounding.y0; y < tile.bounding.y1; y++) for (long x = tile.bounding.x0; x < tile.bounding.x1; x++) {
                if (sub > 1) {      // cell (jx, jy) of tile (ix, iy) -> rank (ix*sub + jx + iy*sub + jy) % world
                    long jx = std::min((x - tile.bounding.x0) / sdx, ncx - 1), jy = std::min((y - tile.bounding.y0) / sdy, ncy - 1);
                    if (((ti / (size_t)ny) * sub + (size_t)jx + (ti % (size_t)ny) * sub + (size_t)jy) % world != prm.rank) continue;
                }
                sampler.start_pixel((uint32_t)x, (uint32_t)y);
                for (uint32_t si = s0; si < s1; si++) {
                    sampler.set_sample_index(si);
                    // Sampler::get_camera_sample (sample/mod.rs:38-43)
                    V2 j = sampler.next_2d();
                    V2 pfilm = j + v2((Float)(uint32_t)x, (Float)(uint32_t)y);
                    V2 plens = sampler.next_2d();
                    RawRay ray = camera_generate(cam, pfilm, plens);
                    tstats[tid].camera++;
                    RGB L;
                    g_trace = trace_on && x == trace_x && y == trace_y && (long)si == trace_s;
                    if (!s.textures.empty()) {                                   // pt.rs:141-142
                        RayDifferential rd = camera_generate_differential(cam, pfilm, plens);
                        scale_differentials(rd, 1.f / (Float)spp);
                        L = calculate_lighting(s, rd.ray, sampler, prm.max_depth, prm.min_depth, prm.rr_threshold, &tstats[tid], &rd);
                    } else
                    L = calculate_lighting(s, ray, sampler, prm.max_depth, prm.min_depth, prm.rr_threshold, &tstats[tid]);
                    if (radiance_out) {
                        size_t ri = (((size_t)y * (size_t)(film.crop_max_x - film.crop_min_x) + (size_t)x) * (s1 - s0) + (si - s0)) * 4;
                        radiance_out[ri] = L.x; radiance_out[ri + 1] = L.y; radiance_out[ri + 2] = L.z; radiance_out[ri + 3] = 0.f;
                    }
                    if (rgb_valid(L)) tile_add_sample(tile, film, pfilm, L);
                    else { tile_add_sample(tile, film, pfilm, grey(0.f)); tstats[tid].invalid++; }
                }
            }
        }
    };
    std::vector<std::thread> th;
    for (int i = 1; i < nthreads; i++) th.emplace_back(worker, i);
    worker(0);
    for (auto& t : th) t.join();
    // Film::collect_into: merge tile sinks, in tile order, into the crop-window sink
    long cw = film.crop_max_x - film.crop_min_x, chh = film.crop_max_y - film.crop_min_y;
    for (size_t i = 0; i < (size_t)cw * (size_t)chh * 4; i++) film_out[i] = 0.f;
    for (FilmTile& t : tiles) {
        if (t.px.empty()) continue;
        long sw = t.sink.x1 - t.sink.x0;
        for (long y = t.sink.y0; y < t.sink.y1; y++) for (long x = t.sink.x0; x < t.sink.x1; x++) {
            const Float* p = &t.px[((size_t)(x - t.sink.x0) + (size_t)(y - t.sink.y0) * (size_t)sw) * 4];
            Float* o = &film_out[((size_t)(x - film.crop_min_x) + (size_t)(y - film.crop_min_y) * (size_t)cw) * 4];
            o[0] += p[0]; o[1] += p[1]; o[2] += p[2]; o[3] += p[3];
        }
    }
    if (stats) for (auto& t : tstats) {
        stats->camera += t.camera; stats->extend += t.extend; stats->shadow += t.shadow; stats->mis += t.mis;
        stats->invalid += t.invalid; stats->extend_bounce += t.extend_bounce;
        stats->trav.nodes += t.trav.nodes; stats->trav.tris += t.trav.tris; stats->trav.spheres += t.trav.spheres;
    }
    return true;
}

// TilePixel::finalize + ToNorm<u8>::from_norm (film.rs:338-344, spectrum/macros.rs:164-180)
inline void film_finalize(const Float* film, size_t n, Float* rgb_out, uint8_t* rgb8_out) {
    for (size_t i = 0; i < n; i++) {
        const Float* p = film + 4 * i; Float c[3];
        if (p[3] == 0.f) { c[0] = c[1] = c[2] = 0.f; } else { c[0] = p[0] / p[3]; c[1] = p[1] / p[3]; c[2] = p[2] / p[3]; }
        for (int k = 0; k < 3; k++) {
            if (rgb_out) rgb_out[3 * i + k] = c[k];
            if (rgb8_out) { Float f = clampf(c[k], 0.f, 1.f); rgb8_out[3 * i + k] = (uint8_t)(f * 255.f); }
        }
    }
}

}  // namespace orc
