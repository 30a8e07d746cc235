// ORACLE — test infrastructure only.  C entry points (ctypes-callable) over the CPU
// restatement of arendur's hot path.  Twin of include/arn.h with an `arn_oracle_` prefix.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may load this library; the product (libarn_b200.so) never does.
//
// PARITY PINNING: the reference's own tests cover only BBox2 algebra and three sphere
// known-answers (src/geometry/tests.rs, src/shape/tests.rs) — restated in
// tests/test_oracle_reference_tests.py.  The reference is Rust (2017 nightly) and cannot be
// built here, so for triangles, the BVH, BxDFs, the integrator and the film this oracle is
// a line-by-line restatement with PARITY UNPINNED.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "render.hpp"

using namespace orc;

struct arn_oracle_scene { Scene s; };

extern "C" {

int arn_oracle_bvh_build(uint32_t n, const float* bounds6, const float* costs, int strategy,
                         arn_node* nodes_out, uint32_t* order_out, uint32_t* n_nodes_out) {
    if (n == 0 || !bounds6 || !costs || !nodes_out || !order_out || !n_nodes_out) return ARN_E_INVALID;
    std::vector<arn_node> nodes; std::vector<uint32_t> order;
    bvh_build(n, bounds6, costs, strategy, nodes, order);
    std::memcpy(nodes_out, nodes.data(), nodes.size() * sizeof(arn_node));
    std::memcpy(order_out, order.data(), order.size() * 4);
    *n_nodes_out = (uint32_t)nodes.size();
    return ARN_OK;
}

// ComponentInfo::new (component/bvh.rs:24-35): bbox_parent + intersection_cost of every component.
int arn_oracle_prim_bounds(const arn_scene_desc* d, float* bounds6_out, float* costs_out) {
    if (!d || !bounds6_out || !costs_out) return ARN_E_INVALID;
    Scene s; s.load(*d);
    for (uint32_t i = 0; i < d->n_prims; i++) {
        BBox3 b; Float c; component_info(s, i, &b, &c);
        bounds6_out[6*i] = b.pmin.x; bounds6_out[6*i+1] = b.pmin.y; bounds6_out[6*i+2] = b.pmin.z;
        bounds6_out[6*i+3] = b.pmax.x; bounds6_out[6*i+4] = b.pmax.y; bounds6_out[6*i+5] = b.pmax.z;
        costs_out[i] = c;
    }
    return ARN_OK;
}

int arn_oracle_light_distribution(uint32_t n, const float* func, float* cdf_out, float* integral_out) {
    distribution_new(func, n, cdf_out, integral_out); return ARN_OK;
}
// Light::power().to_xyz().y for a sphere primitive (component/shape.rs:160-167, renderer/scene.rs:38-40)
float arn_oracle_light_power_y(const arn_sphere* sp) { return rgb_y(light_power(*sp)); }
// the same for Point / Spot / Distant lights (pointlights.rs:79-81,222-226, distantlight.rs:108-110)
float arn_oracle_analytic_power_y(const arn_analytic_light* l) { return rgb_y(analytic_power(*l)); }
// Light::evaluate_sampled of an analytic light: out7 = radiance rgb, pdf, pfrom xyz
void arn_oracle_analytic_sample(const arn_analytic_light* l, const float* pos3, float* out7) {
    LightSample ls = analytic_evaluate_sampled(*l, v3(pos3[0], pos3[1], pos3[2]));
    out7[0] = ls.radiance.x; out7[1] = ls.radiance.y; out7[2] = ls.radiance.z; out7[3] = ls.pdf;
    out7[4] = ls.pfrom.x; out7[5] = ls.pfrom.y; out7[6] = ls.pfrom.z;
}

int arn_oracle_scene_create(const arn_scene_desc* d, arn_oracle_scene** out) {
    if (!d || !out) return ARN_E_INVALID;
    arn_oracle_scene* h = new arn_oracle_scene; h->s.load(*d); *out = h; return ARN_OK;
}
void arn_oracle_scene_destroy(arn_oracle_scene* h) { delete h; }

// counters_out (may be NULL): [nodes tested, triangles tested, spheres tested] summed over the batch
int arn_oracle_intersect_closest(arn_oracle_scene* h, const arn_ray* rays, size_t n, arn_hit* hits, uint64_t* counters_out, int nthreads) {
    if (!h || !rays || !hits) return ARN_E_INVALID;
    if (nthreads < 1) nthreads = 1;
    std::vector<TraversalCounters> ctr((size_t)nthreads);
    auto work = [&](int tid) {
        size_t lo = n * (size_t)tid / (size_t)nthreads, hi = n * (size_t)(tid + 1) / (size_t)nthreads;
        for (size_t i = lo; i < hi; i++) {
            RawRay r = ray_new(v3(rays[i].o[0], rays[i].o[1], rays[i].o[2]), v3(rays[i].d[0], rays[i].d[1], rays[i].d[2]), rays[i].tmax);
            int prim;
            if (bvh_intersect(h->s, r, nullptr, &prim, &ctr[tid], false)) { hits[i].prim_id = prim; hits[i].t = r.tmax; }
            else { hits[i].prim_id = -1; hits[i].t = infinity(); }
        }
    };
    std::vector<std::thread> th; for (int i = 1; i < nthreads; i++) th.emplace_back(work, i);
    work(0); for (auto& t : th) t.join();
    if (counters_out) { counters_out[0] = counters_out[1] = counters_out[2] = 0;
        for (auto& c : ctr) { counters_out[0] += c.nodes; counters_out[1] += c.tris; counters_out[2] += c.spheres; } }
    return ARN_OK;
}
int arn_oracle_intersect_any(arn_oracle_scene* h, const arn_ray* rays, size_t n, uint8_t* out) {
    if (!h || !rays || !out) return ARN_E_INVALID;
    for (size_t i = 0; i < n; i++) {
        RawRay r = ray_new(v3(rays[i].o[0], rays[i].o[1], rays[i].o[2]), v3(rays[i].d[0], rays[i].d[1], rays[i].d[2]), rays[i].tmax);
        out[i] = bvh_intersect(h->s, r, nullptr, nullptr, nullptr, false) ? 1 : 0;
    }
    return ARN_OK;
}

int arn_oracle_camera_make(const float* parent_view16, const float* screen4, float znear, float zfar, float fov,
                           int has_lens, float lens_radius, float focal_distance, float res_x, float res_y, arn_camera* out) {
    return camera_make(parent_view16, screen4, znear, zfar, fov, has_lens, lens_radius, focal_distance, res_x, res_y, out) ? ARN_OK : ARN_E_INVALID;
}
int arn_oracle_ortho_camera_make(const float* view_parent16, const float* screen4, float znear, float zfar,
                                 int has_lens, float lens_radius, float focal_distance, float res_x, float res_y, arn_camera* out) {
    return ortho_camera_make(view_parent16, screen4, znear, zfar, has_lens, lens_radius, focal_distance, res_x, res_y, out) ? ARN_OK : ARN_E_INVALID;
}
// Camera rays for a list of film positions (pfilm.xy, plens.xy per ray): PerspecCam::generate_path
int arn_oracle_camera_rays(const arn_camera* cam, const float* pfilm_plens4, size_t n, arn_ray* out) {
    for (size_t i = 0; i < n; i++) {
        RawRay r = camera_generate(*cam, v2(pfilm_plens4[4*i], pfilm_plens4[4*i+1]), v2(pfilm_plens4[4*i+2], pfilm_plens4[4*i+3]));
        out[i].o[0] = r.origin.x; out[i].o[1] = r.origin.y; out[i].o[2] = r.origin.z;
        out[i].d[0] = r.dir.x; out[i].d[1] = r.dir.y; out[i].d[2] = r.dir.z; out[i].tmax = r.tmax;
    }
    return ARN_OK;
}

// stats_out: arn_stats (ray counters only) ; trav_out (may be NULL): [nodes, tris, spheres] over ALL traversals
int arn_oracle_render_pt(arn_oracle_scene* h, const arn_camera* cam, const arn_film* film, const arn_sampler* smp,
                         const arn_pt_params* prm, float* film_out, arn_stats* stats_out, uint64_t* trav_out, int nthreads) {
    if (!h || !cam || !film || !smp || !prm || !film_out) return ARN_E_INVALID;
    if (h->s.light_prims.empty()) return ARN_E_INVALID;   // the reference panics (index out of bounds, scene.rs:53-55)
    RayStats st;
    if (!render_pt(h->s, *cam, *film, *smp, *prm, film_out, &st, nthreads)) return ARN_E_INVALID;   // spawn_tiles' unwrap panics (film.rs:129)
    if (stats_out) {
        std::memset(stats_out, 0, sizeof *stats_out);
        stats_out->camera_rays = st.camera; stats_out->extend_rays = st.extend; stats_out->shadow_rays = st.shadow;
        stats_out->mis_rays = st.mis; stats_out->invalid_samples = st.invalid; stats_out->extend_bounce_rays = st.extend_bounce;
    }
    if (trav_out) { trav_out[0] = st.trav.nodes; trav_out[1] = st.trav.tris; trav_out[2] = st.trav.spheres; }
    return ARN_OK;
}

int arn_oracle_render_pt_samples(arn_oracle_scene* h, const arn_camera* cam, const arn_film* film, const arn_sampler* smp,
                                 const arn_pt_params* prm, float* film_out, float* radiance_out, int nthreads) {
    if (!h || !cam || !film || !smp || !prm || !film_out || !radiance_out || h->s.light_prims.empty()) return ARN_E_INVALID;
    RayStats st;
    if (!render_pt(h->s, *cam, *film, *smp, *prm, film_out, &st, nthreads, radiance_out)) return ARN_E_INVALID;
    return ARN_OK;
}

int arn_oracle_film_finalize(const float* film, size_t n_pixels, float* rgb_out, uint8_t* rgb8_out) {
    film_finalize(film, n_pixels, rgb_out, rgb8_out); return ARN_OK;
}

// TriangleMesh::from_model_transformed (shape/triangle.rs:120-160): positions through
// Matrix4::transform_point (homogeneous divide), normals through transform_norm.
int arn_oracle_mesh_transform(const float* transform16, uint32_t n, const float* pos_in, const float* nrm_in,
                              float* pos_out, float* nrm_out) {
    M4 t = m4_from_cols(transform16);
    for (uint32_t i = 0; i < n; i++) {
        V3 p = transform_point(t, v3(pos_in[3*i], pos_in[3*i+1], pos_in[3*i+2]));
        pos_out[3*i] = p.x; pos_out[3*i+1] = p.y; pos_out[3*i+2] = p.z;
        if (nrm_in && nrm_out) {
            V3 q = transform_norm(t, v3(nrm_in[3*i], nrm_in[3*i+1], nrm_in[3*i+2]));
            nrm_out[3*i] = q.x; nrm_out[3*i+1] = q.y; nrm_out[3*i+2] = q.z;
        }
    }
    return ARN_OK;
}
int arn_oracle_m4_invert(const float* m16, float* out16) {
    M4 o; if (!m4_invert(m4_from_cols(m16), &o)) return ARN_E_INVALID; std::memcpy(out16, o.m, 64); return ARN_OK;
}
// Sphere::new (shape/sphere.rs:133-156): clamps and theta range
int arn_oracle_sphere_new(float radius, float zmin, float zmax, float phimax, arn_sphere* out) {
    if (!(radius > 0.f) || !(zmin < zmax)) return ARN_E_INVALID;    // assert! in the source
    if (zmin < -radius) zmin = -radius;
    if (zmax > radius) zmax = radius;
    if (phimax < 0.f) phimax = 0.f;
    Float twopi = pi() * 2.f;
    if (phimax > twopi) phimax = twopi;
    out->radius = radius; out->zmin = zmin; out->zmax = zmax; out->phimax = phimax;
    out->thetamin = facos(zmin / radius); out->thetamax = facos(zmax / radius);
    return ARN_OK;
}

// ---------------------------------------------------------------- unit-test hooks
// Sphere::intersect_ray in local space: returns 1 on hit and fills t, pos, norm, wo.
int arn_oracle_sphere_intersect(const arn_sphere* sp, const arn_ray* ray, float* t, float* pos3, float* norm3, float* wo3) {
    RawRay r = ray_new(v3(ray->o[0], ray->o[1], ray->o[2]), v3(ray->d[0], ray->d[1], ray->d[2]), ray->tmax);
    Float tt; SurfaceInteraction si;
    if (!sphere_intersect(*sp, r, &tt, &si)) return 0;
    *t = tt; pos3[0] = si.basic.pos.x; pos3[1] = si.basic.pos.y; pos3[2] = si.basic.pos.z;
    norm3[0] = si.basic.norm.x; norm3[1] = si.basic.norm.y; norm3[2] = si.basic.norm.z;
    wo3[0] = si.basic.wo.x; wo3[1] = si.basic.wo.y; wo3[2] = si.basic.wo.z;
    return 1;
}
// BBox2<isize> algebra (geometry/bbox.rs:21-232). op codes in tests/test_oracle_reference_tests.py
int arn_oracle_bbox2i(int op, const long* a4, const long* b4, long* out4) {
    BBox2<long> a = BBox2<long>::make(a4[0], a4[1], a4[2], a4[3]);
    BBox2<long> r = a; int flag = 1;
    switch (op) {
    case 0: break;                                                    // new
    case 1: { long x, y; a.corner((int)b4[0], &x, &y); out4[0] = x; out4[1] = y; return 1; }
    case 2: r = a.extend(b4[0], b4[1]); break;
    case 3: { BBox2<long> b = BBox2<long>::make(b4[0], b4[1], b4[2], b4[3]); r = a.unite(b); break; }
    case 4: { BBox2<long> b = BBox2<long>::make(b4[0], b4[1], b4[2], b4[3]); flag = a.intersect(b, &r) ? 1 : 0; break; }
    case 5: { BBox2<long> b = BBox2<long>::make(b4[0], b4[1], b4[2], b4[3]); return a.overlap(b) ? 1 : 0; }
    case 6: return a.contain(b4[0], b4[1]) ? 1 : 0;
    case 7: return a.contain_lb(b4[0], b4[1]) ? 1 : 0;
    case 8: r = a.expand_by(b4[0]); break;
    case 9: out4[0] = a.surface_area(); return 1;
    case 10: return a.max_extent();
    default: return -1;
    }
    out4[0] = r.x0; out4[1] = r.y0; out4[2] = r.x1; out4[3] = r.y1;
    return flag;
}
int arn_oracle_bbox2f_lerp(const float* a4, float tx, float ty, float* out2) {
    BBox2<float> a = BBox2<float>::make(a4[0], a4[1], a4[2], a4[3]); a.lerp(tx, ty, &out2[0], &out2[1]); return 1;
}
// ParitySampler draws: n1 1-D then n2 2-D draws of sample `s` at pixel (px,py)
int arn_oracle_sampler_draws(uint32_t seed, uint32_t px, uint32_t py, uint32_t s, uint32_t n1, uint32_t n2, float* out) {
    ParitySampler sp; sp.seed = seed; sp.spp = s + 1; sp.start_pixel(px, py); sp.set_sample_index(s);
    for (uint32_t i = 0; i < n1; i++) out[i] = sp.next();
    for (uint32_t i = 0; i < n2; i++) { V2 v = sp.next_2d(); out[n1 + 2*i] = v.x; out[n1 + 2*i + 1] = v.y; }
    return ARN_OK;
}
// the same for a full sampler description (ARN_SAMPLER_STRATIFIED: sample s of sampledx * sampledy)
int arn_oracle_sampler_draws2(const arn_sampler* smp, uint32_t px, uint32_t py, uint32_t s, uint32_t n1, uint32_t n2, float* out) {
    ParitySampler sp; sp.seed = smp->seed; sp.spp = smp->sampledx * smp->sampledy;
    sp.mode = smp->mode; sp.sampledx = smp->sampledx; sp.sampledy = smp->sampledy; sp.ndim = smp->ndim;
    sp.start_pixel(px, py); sp.set_sample_index(s);
    for (uint32_t i = 0; i < n1; i++) out[i] = sp.next();
    for (uint32_t i = 0; i < n2; i++) { V2 v = sp.next_2d(); out[n1 + 2*i] = v.x; out[n1 + 2*i + 1] = v.y; }
    return ARN_OK;
}
float arn_oracle_lanczos(float dx, float dy) { return lanczos_evaluate(v2(dx, dy), 1.f / 3.f); }
// Filter::evaluate_unsafe of the film's filter at a signed offset (sample/filters.rs)
float arn_oracle_filter(const arn_film* film, float dx, float dy) { return filter_evaluate(*film, v2(dx, dy)); }
float arn_oracle_roughness_to_alpha(float r) { return roughness_to_alpha(r); }

// BSDF probe for kernel-level parity: evaluate_sampled / evaluate / pdf of a material at a fixed
// local frame (ts, bs, ns = x, y, z; ng = z).  out: f(3), wi(3), pdf, type, feval(3), pdfeval
int arn_oracle_bsdf_probe2(const arn_material* m, const float* wo3, const float* u2, const float* wi_eval3, const float* frame9, float* out12);
int arn_oracle_bsdf_probe(const arn_material* m, const float* wo3, const float* u2, const float* wi_eval3, float* out12) {
    return arn_oracle_bsdf_probe2(m, wo3, u2, wi_eval3, nullptr, out12);
}
// frame9 (optional): dpdu, shading normal, geometric normal
int arn_oracle_bsdf_probe2(const arn_material* m, const float* wo3, const float* u2, const float* wi_eval3, const float* frame9, float* out12) {
    SurfaceInteraction si; std::memset(&si, 0, sizeof si);
    si.shading_duv.dpdu = v3(1, 0, 0); si.shading_norm = v3(0, 0, 1); si.basic.norm = v3(0, 0, 1);
    if (frame9) { si.shading_duv.dpdu = v3(frame9[0], frame9[1], frame9[2]); si.shading_norm = v3(frame9[3], frame9[4], frame9[5]); si.basic.norm = v3(frame9[6], frame9[7], frame9[8]); }
    Bsdf b = compute_scattering(*m, si);
    Sampled s = bsdf_evaluate_sampled(b, v3(wo3[0], wo3[1], wo3[2]), v2(u2[0], u2[1]), BXDF_ALL);
    out12[0] = s.f.x; out12[1] = s.f.y; out12[2] = s.f.z; out12[3] = s.wi.x; out12[4] = s.wi.y; out12[5] = s.wi.z;
    out12[6] = s.pdf; out12[7] = (float)s.type;
    RGB f = bsdf_evaluate(b, v3(wo3[0], wo3[1], wo3[2]), v3(wi_eval3[0], wi_eval3[1], wi_eval3[2]), BXDF_ALL);
    out12[8] = f.x; out12[9] = f.y; out12[10] = f.z;
    out12[11] = bsdf_pdf(b, v3(wo3[0], wo3[1], wo3[2]), v3(wi_eval3[0], wi_eval3[1], wi_eval3[2]), BXDF_ALL);
    return ARN_OK;
}

// MipMap::look_up through UVMapping (texturing/textures/image.rs:447-476, mappings.rs:21-30): dxy4 = dudx, dvdx, dudy, dvdy
int arn_oracle_texture_lookup(const arn_texture* tex, const float* texels, const float* uv2, const float* dxy4, float* out3) {
    TexView v; v.t = tex; v.texels = texels;
    DxyInfo d = dxy_default(); d.dudx = dxy4[0]; d.dvdx = dxy4[1]; d.dudy = dxy4[2]; d.dvdy = dxy4[3];
    Texel t = texture_evaluate(v, v2(uv2[0], uv2[1]), d);
    out3[0] = t.c[0]; out3[1] = t.c[1]; out3[2] = t.c[2];
    return ARN_OK;
}
// SurfaceInteraction::compute_dxy on a bare plane (pos, n, dpdu, dpdv) with offset rays (rxo, rxd, ryo, ryd): out10 = dpdx, dpdy, dudx, dvdx, dudy, dvdy
int arn_oracle_compute_dxy(const float* pos3, const float* n3, const float* dpdu3, const float* dpdv3, const float* rays12, float* out10) {
    SurfaceInteraction si; si.basic.pos = v3(pos3[0], pos3[1], pos3[2]); si.basic.norm = v3(n3[0], n3[1], n3[2]);
    si.duv.dpdu = v3(dpdu3[0], dpdu3[1], dpdu3[2]); si.duv.dpdv = v3(dpdv3[0], dpdv3[1], dpdv3[2]);
    RayDifferential rd; rd.has_diffs = true;
    rd.rx = ray_from_od(v3(rays12[0], rays12[1], rays12[2]), v3(rays12[3], rays12[4], rays12[5]));
    rd.ry = ray_from_od(v3(rays12[6], rays12[7], rays12[8]), v3(rays12[9], rays12[10], rays12[11]));
    DxyInfo d = compute_dxy(si, rd);
    out10[0] = d.dpdx.x; out10[1] = d.dpdx.y; out10[2] = d.dpdx.z; out10[3] = d.dpdy.x; out10[4] = d.dpdy.y; out10[5] = d.dpdy.z;
    out10[6] = d.dudx; out10[7] = d.dvdx; out10[8] = d.dudy; out10[9] = d.dvdy;
    return ARN_OK;
}

const char* arn_oracle_version(void) { return "arendur oracle (CPU restatement) 0.1"; }

}  // extern "C"
