// ORACLE — test infrastructure only (see geom.hpp header).
// CPU restatement of arendur's shapes, components and BVH:
//   src/shape/{mod,triangle,sphere}.rs, src/component/{mod,bvh,shape,transformed}.rs
#pragma once
#include <vector>
#include <cassert>
#include "../include/arn.h"
#include "geom.hpp"

namespace orc {

// ------------------------------------------------------------------ flattened scene copy
struct Scene {
    std::vector<V3> positions, normals; std::vector<V2> uvs;
    bool have_normals = false, have_uvs = false;
    std::vector<uint32_t> indices, tri_mesh;
    std::vector<arn_mesh> meshes;
    std::vector<arn_sphere> spheres;
    std::vector<arn_material> materials;
    std::vector<uint32_t> prims;          // component list handed to BVH::new
    std::vector<arn_node> nodes;
    std::vector<uint32_t> order;          // ordered slot -> component index
    std::vector<uint32_t> light_prims;    // Scene.lights: component index, or ARN_LIGHT_ANALYTIC | index below
    std::vector<arn_analytic_light> analytic_lights;
    std::vector<Float> light_func, light_cdf;
    Float light_func_integral = 0.f;
    std::vector<arn_texture> textures; std::vector<Float> texels;      // image textures (N4)

    void load(const arn_scene_desc& d) {
        positions.resize(d.n_vertices);
        for (uint32_t i = 0; i < d.n_vertices; i++) positions[i] = v3(d.positions[3*i], d.positions[3*i+1], d.positions[3*i+2]);
        have_normals = d.normals != nullptr; have_uvs = d.uvs != nullptr;
        if (have_normals) { normals.resize(d.n_vertices); for (uint32_t i = 0; i < d.n_vertices; i++) normals[i] = v3(d.normals[3*i], d.normals[3*i+1], d.normals[3*i+2]); }
        if (have_uvs) { uvs.resize(d.n_vertices); for (uint32_t i = 0; i < d.n_vertices; i++) uvs[i] = v2(d.uvs[2*i], d.uvs[2*i+1]); }
        indices.assign(d.indices, d.indices + 3 * (size_t)d.n_triangles);
        tri_mesh.assign(d.tri_mesh, d.tri_mesh + d.n_triangles);
        meshes.assign(d.meshes, d.meshes + d.n_meshes);
        spheres.assign(d.spheres, d.spheres + d.n_spheres);
        materials.assign(d.materials, d.materials + d.n_materials);
        prims.assign(d.prims, d.prims + d.n_prims);
        if (d.nodes) nodes.assign(d.nodes, d.nodes + d.n_nodes);
        if (d.order) order.assign(d.order, d.order + d.n_prims);
        if (d.n_lights) {
            light_prims.assign(d.light_prims, d.light_prims + d.n_lights);
            light_func.assign(d.light_func, d.light_func + d.n_lights);
            light_cdf.assign(d.light_cdf, d.light_cdf + d.n_lights + 1);
        }
        light_func_integral = d.light_func_integral;
        if (d.n_analytic_lights) analytic_lights.assign(d.analytic_lights, d.analytic_lights + d.n_analytic_lights);
        if (d.n_textures) { textures.assign(d.textures, d.textures + d.n_textures); texels.assign(d.texels, d.texels + d.n_texel_floats); }
    }
    bool prim_is_sphere(uint32_t comp) const { return (prims[comp] & ARN_PRIM_SPHERE) != 0; }
    uint32_t prim_index(uint32_t comp) const { return prims[comp] & ~ARN_PRIM_SPHERE; }
};

// ------------------------------------------------------------------ triangle (shape/triangle.rs)
struct TriVerts { V3 p0, p1, p2; };
inline TriVerts tri_verts(const Scene& s, uint32_t tri) {                   // :265-283, Index impl :379-387
    TriVerts t; t.p0 = s.positions[s.indices[3*tri]]; t.p1 = s.positions[s.indices[3*tri+1]]; t.p2 = s.positions[s.indices[3*tri+2]];
    return t;
}
inline BBox3 tri_bbox(const Scene& s, uint32_t tri) {                       // bbox_local :391-394
    TriVerts t = tri_verts(s, tri);
    return BBox3::make(t.p0, t.p1).extend(t.p2);
}
inline void tri_uvs(const Scene& s, uint32_t tri, V2* a, V2* b, V2* c) {    // :286-297
    if (s.meshes[s.tri_mesh[tri]].has_uvs) {
        *a = s.uvs[s.indices[3*tri]]; *b = s.uvs[s.indices[3*tri+1]]; *c = s.uvs[s.indices[3*tri+2]];
    } else { *a = v2(0.f, 0.f); *b = v2(1.f, 0.f); *c = v2(1.f, 1.f); }
}
// Matrix3::look_at(dir, up) of cgmath 0.14: transpose of [side, up', dir]; returns columns x and z
inline void m3_look_at_xz(V3 dir, V3 up, V3* colx, V3* colz) {
    V3 d = normalize(dir);
    V3 side = normalize(cross(up, d));
    V3 u = normalize(cross(d, side));
    *colx = v3(side.x, u.x, d.x);
    *colz = v3(side.z, u.z, d.z);
}
inline void computedpduv(V3 p0, V3 p1, V3 p2, V2 uv0, V2 uv1, V2 uv2, V3* dpdu, V3* dpdv) { // :308-331
    V2 duv02 = uv0 - uv2, duv12 = uv1 - uv2;
    V3 dp02 = p0 - p2, dp12 = p1 - p2;
    Float determinant = duv02.x * duv12.y - duv02.y * duv12.x;
    if (determinant == 0.f) {
        V3 up = cross(dp02, p0 - p1);
        m3_look_at_xz(dp02, up, dpdu, dpdv);
    } else {
        Float inv = 1.f / determinant;
        *dpdu = (duv12.y * dp02 - duv02.y * dp12) * inv;
        *dpdv = (-duv12.x * dp02 + duv02.x * dp12) * inv;
    }
}
inline DuvInfo compute_shading_at(const Scene& s, uint32_t tri, V3 b, V3 dpdu) {            // :333-376
    TriVerts t = tri_verts(s, tri);
    V3 shading_normal, dndu, dndv;
    if (s.meshes[s.tri_mesh[tri]].has_normals) {
        V3 n0 = s.normals[s.indices[3*tri]], n1 = s.normals[s.indices[3*tri+1]], n2 = s.normals[s.indices[3*tri+2]];
        shading_normal = normalize(b.x * n0 + b.y * n1 + b.z * n2);
        V2 a, bb, c; tri_uvs(s, tri, &a, &bb, &c);
        computedpduv(n0, n1, n2, a, bb, c, &dndu, &dndv);
    } else {
        shading_normal = normalize(cross(t.p2 - t.p0, t.p1 - t.p0));
        dndu = v3(0, 0, 0); dndv = v3(0, 0, 0);
    }
    V3 shading_tangent = normalize(dpdu);            // mesh.tangents is always None (:104,156)
    V3 shading_bitangent = cross(shading_tangent, shading_normal);
    if (magnitude2(shading_bitangent) > 0.f) {
        shading_bitangent = normalize(shading_bitangent);
        shading_tangent = cross(shading_bitangent, shading_normal);
    } else {
        nrm::get_basis_from(shading_normal, &shading_tangent, &shading_bitangent);
    }
    DuvInfo d; d.dpdu = shading_tangent; d.dpdv = shading_bitangent; d.dndu = dndu; d.dndv = dndv; return d;
}
// Shape::intersect_ray for TriangleInstance (:396-484).  `full` = false stops after the
// acceptance test (what the GPU does for non-final candidates; identical t).
inline bool tri_intersect(const Scene& s, uint32_t tri, const RawRay& ray, Float* t_out, SurfaceInteraction* si, bool full = true) {
    TriVerts tv = tri_verts(s, tri);
    V3 p0 = tv.p0, p1 = tv.p1, p2 = tv.p2;
    const Stc& stc = ray.stc;
    V3 p0t, p1t, p2t; stc_apply(stc, p0, p1, p2, &p0t, &p1t, &p2t);
    Float e0 = p1t.x * p2t.y - p1t.y * p2t.x;
    Float e1 = p2t.x * p0t.y - p2t.y * p0t.x;
    Float e2 = p0t.x * p1t.y - p0t.y * p1t.x;
    if ((e0 < 0.f || e1 < 0.f || e2 < 0.f) && (e0 > 0.f || e1 > 0.f || e2 > 0.f)) return false;
    Float det = e0 + e1 + e2;
    if (det == 0.f) return false;
    p0t.z *= stc.shear.z; p1t.z *= stc.shear.z; p2t.z *= stc.shear.z;
    Float tscaled = e0 * p0t.z + e1 * p1t.z + e2 * p2t.z;
    if (det < 0.f && (tscaled >= 0.f || tscaled < ray.tmax * det)) return false;
    else if (det > 0.f && (tscaled <= 0.f || tscaled > ray.tmax * det)) return false;
    Float inv_det = 1.f / det;
    Float b0 = e0 * inv_det, b1 = e1 * inv_det, b2 = e2 * inv_det;
    Float t = tscaled * inv_det;
    // conservative intersection — plain max, no abs (quirk A-18)
    Float maxxt = fmax_(fmax_(p0t.x, p1t.x), p2t.x);
    Float maxyt = fmax_(fmax_(p0t.y, p1t.y), p2t.y);
    Float maxzt = fmax_(fmax_(p0t.z, p1t.z), p2t.z);
    Float maxe = fmax_(fmax_(e0, e1), e2);
    Float deltax = maxxt * eb_term(5.f);
    Float deltay = maxyt * eb_term(5.f);
    Float deltaz = maxzt * eb_term(3.f);
    Float delta_err = 2.f * (eb_term(2.f) * maxxt * maxyt + deltay * maxxt + deltax * maxyt);
    Float delta_t = 3.f * (eb_term(3.f) * maxe * maxzt + delta_err * maxzt + deltaz * maxe) * std::fabs(inv_det);
    if (t <= delta_t) return false;
    *t_out = t;
    if (!full || !si) return true;

    V2 uv0, uv1, uv2; tri_uvs(s, tri, &uv0, &uv1, &uv2);
    V3 phit = b0 * p0 + b1 * p1 + b2 * p2;
    V3 perr = eb_term(7.f) * v3(
        std::fabs(b0 * p0.x) + std::fabs(b1 * p1.x) + std::fabs(b2 * p2.x),
        std::fabs(b0 * p0.y) + std::fabs(b1 * p1.y) + std::fabs(b2 * p2.y),
        std::fabs(b0 * p0.z) + std::fabs(b1 * p1.z) + std::fabs(b2 * p2.z));
    V2 uvhit = b0 * uv0 + b1 * uv1 + b2 * uv2;
    V3 dpdu, dpdv; computedpduv(p0, p1, p2, uv0, uv1, uv2, &dpdu, &dpdv);
    DuvInfo d; d.dpdu = dpdu; d.dpdv = dpdv; d.dndu = v3(0, 0, 0); d.dndv = v3(0, 0, 0);
    *si = si_new(phit, perr, -ray.dir, uvhit, d);
    si_set_shading(*si, compute_shading_at(s, tri, v3(b0, b1, b2), dpdu), true);
    return true;
}

// ------------------------------------------------------------------ sphere (shape/sphere.rs)
inline BBox3 sphere_bounding(const arn_sphere& sp) {                          // :165-170
    return BBox3::make(v3(-sp.radius, -sp.radius, sp.zmin), v3(sp.radius, sp.radius, sp.zmax));
}
inline bool sphere_intersect_full(Float radius, const RawRay& ray, Float* t) { // :193-221
    V3 origin = ray.origin, direction = ray.dir;
    Float a = magnitude2(direction);
    V3 m = mul_elem(direction, origin) * 2.f;
    Float b = m.x + m.y + m.z;
    Float c = magnitude2(origin) - radius * radius;
    Float delta = b * b - 4.f * a * c;
    if (delta < 0.f) return false;
    Float invert_2a = 1.f / (2.f * a);
    Float d1 = std::sqrt(delta) * invert_2a;
    Float d0 = -b * invert_2a;
    Float t0, t1;
    if (invert_2a > 0.f) { t0 = d0 - d1; t1 = d0 + d1; } else { t0 = d0 + d1; t1 = d0 - d1; }
    Float tmax = ray.tmax;
    if (t0 > tmax || t1 < 0.f) return false;
    if (t0 > 0.f) { *t = t0; return true; }
    else if (t1 > tmax) return false;
    else { *t = t1; return true; }
}
inline bool sphere_intersect(const arn_sphere& sp, const RawRay& ray, Float* t_out, SurfaceInteraction* si) { // :231-297
    Float t;
    if (!sphere_intersect_full(sp.radius, ray, &t)) return false;
    V3 p = ray_evaluate(ray, t);
    p = p * sp.radius / magnitude(p);
    if (p.x == 0.f && p.y == 0.f) p.x = 1e-5f * sp.radius;
    Float phi = fatan2(p.y, p.x);
    if (phi < 0.f) phi += 2.f * pi();
    if (p.z < sp.zmin || p.z > sp.zmax || phi > sp.phimax) return false;
    *t_out = t;
    if (!si) return true;
    Float phimax = sp.phimax, thetamax = sp.thetamax, thetamin = sp.thetamin;
    Float thetadelta = thetamax - thetamin;
    Float u = phi / phimax;
    Float theta = facos(p.z / sp.radius);
    Float v = (theta - thetamin) / thetadelta;
    Float inv_z_radius = 1.f / std::sqrt(p.x * p.x + p.y * p.y);
    Float cos_phi = p.x * inv_z_radius, sin_phi = p.y * inv_z_radius;
    V3 dpdu = v3(-phimax * p.y, phimax * p.x, 0.f);
    V3 dpdv = thetadelta * v3(p.z * cos_phi, p.z * sin_phi, -sp.radius * fsin(theta));
    V3 dppduu = -phimax * phimax * v3(p.x, p.y, 0.f);
    V3 dppduv = thetadelta * p.z * phimax * v3(-sin_phi, cos_phi, 0.f);
    V3 dppdvv = -thetadelta * thetadelta * v3(p.x, p.y, p.z);
    Float e = dot(dpdu, dpdu), f = dot(dpdu, dpdv), g = dot(dpdv, dpdv);
    V3 n = normalize(cross(dpdu, dpdv));
    Float ee = dot(n, dppduu), ff = dot(n, dppduv), gg = dot(n, dppdvv);
    Float inv = 1.f / (e * g - f * f);
    DuvInfo d; d.dpdu = dpdu; d.dpdv = dpdv;
    d.dndu = (ff * f - ee * g) * inv * dpdu + (ee * f - ff * e) * inv * dpdv;
    d.dndv = (gg * f - ff * g) * inv * dpdu + (ff * f - gg * e) * inv * dpdv;
    *si = si_new(p, v3(0, 0, 0), -ray.dir, v2(u, v), d);    // pos_err = 0 ("FIXME: wrong", :281-282)
    return true;
}
inline Float sphere_surface_area(const arn_sphere& sp) { return sp.phimax * sp.radius * (sp.zmax - sp.zmin); } // :300-302
inline void sphere_sample(const arn_sphere& sp, V2 sample, V3* pos, V3* n, Float* pdf) {       // :304-311
    Float phi = sample.x * sp.phimax;
    Float theta = sample.y * (sp.thetamax - sp.thetamin) + sp.thetamin;
    V3 dir = spherical_to_vec(theta, phi);
    *pos = dir * sp.radius; *n = dir; *pdf = 1.f / sphere_surface_area(sp);
}
// Shape::sample_wrt default (shape/mod.rs:52-64)
inline void sphere_sample_wrt(const arn_sphere& sp, V3 pref, V2 sample, V3* lp, V3* lnorm, Float* lpdf) {
    sphere_sample(sp, sample, lp, lnorm, lpdf);
    V3 wi = *lp - pref;
    Float distance2 = magnitude2(wi);
    if (relative_eq(distance2, 0.f)) *lpdf = 0.f;
    else {
        V3 w = wi / std::sqrt(distance2);
        *lpdf *= distance2 / std::fabs(dot(*lnorm, w));
        if (std::isinf(*lpdf)) *lpdf = 0.f;
    }
}
// Shape::pdf_wrt default (shape/mod.rs:67-75)
inline Float sphere_pdf_wrt(const arn_sphere& sp, V3 pos_ref, V3 wi) {
    RawRay ray = ray_from_od(pos_ref, wi);
    Float t; SurfaceInteraction si;
    if (sphere_intersect(sp, ray, &t, &si))
        return magnitude2(si.basic.pos - pos_ref) / (std::fabs(dot(wi, si.basic.norm)) * sphere_surface_area(sp));
    return 0.f;
}

// ------------------------------------------------------------------ components
// Composable::bbox_parent / intersection_cost per component (component/bvh.rs:24-35):
//   triangle: bbox_local, cost 3.0 (triangle.rs:506-538)
//   ShapedPrimitive: shape.bbox_local, cost 1.0 (component/shape.rs:46-48, mod.rs:44-47)
//   TransformedComposable: inner bbox .apply_transform(local_parent), cost 1 + inner (transformed.rs:42-51)
inline void component_info(const Scene& s, uint32_t comp, BBox3* bound, Float* cost) {
    if (s.prim_is_sphere(comp)) {
        const arn_sphere& sp = s.spheres[s.prim_index(comp)];
        BBox3 b = sphere_bounding(sp);
        if (sp.has_transform) { *bound = b.apply_transform(m4_from_cols(sp.local_parent)); *cost = 1.f + 1.f; }
        else { *bound = b; *cost = 1.f; }
    } else { *bound = tri_bbox(s, s.prim_index(comp)); *cost = 3.f; }
}

// Composable::intersect_ray for one component: updates ray (tmax; for transformed spheres
// also origin/direction through the round trip) and returns the SurfaceInteraction.
inline bool component_intersect(const Scene& s, uint32_t comp, RawRay& ray, SurfaceInteraction* si, bool full_si) {
    if (!s.prim_is_sphere(comp)) {                           // triangle.rs:513-523
        Float t;
        if (tri_intersect(s, s.prim_index(comp), ray, &t, si, full_si)) {
            ray_set_tmax(ray, t);
            if (si) si->primitive_hit = (int)comp;
            return true;
        }
        return false;
    }
    const arn_sphere& sp = s.spheres[s.prim_index(comp)];
    if (!sp.has_transform) {                                 // component/shape.rs:52-61
        Float t;
        if (sphere_intersect(sp, ray, &t, si)) { ray_set_tmax(ray, t); if (si) si->primitive_hit = (int)comp; return true; }
        return false;
    }
    // TransformedComposable<T: Primitive>::intersect_ray (transformed.rs:73-83)
    M4 pl = m4_from_cols(sp.parent_local), lp = m4_from_cols(sp.local_parent);
    ray = ray_apply_transform(ray, pl);
    Float t; SurfaceInteraction lsi; bool hit = sphere_intersect(sp, ray, &t, &lsi);
    if (hit) {
        ray_set_tmax(ray, t);
        if (si) { *si = si_apply_transform(lsi, lp); si->primitive_hit = (int)comp; }
    }
    ray = ray_apply_transform(ray, lp);
    return hit;
}

// ------------------------------------------------------------------ BVH build (component/bvh.rs)
struct ComponentInfo { BBox3 bound; V3 centroid; Float cost; uint32_t idx; };   // :15-21
struct BuildNode {                                                             // :166-176
    BBox3 bound; int child0, child1, axis; bool leaf; uint32_t offset, len;
};
struct Bucket { uint32_t count; Float cost; BBox3 bound; bool initialized; };   // :319-325
inline Bucket bucket_union(const Bucket& a, const Bucket& b) {                 // :360-374
    if (!a.initialized) return b;
    else if (!b.initialized) return a;
    Bucket r; r.count = a.count + b.count; r.cost = a.cost + b.cost; r.bound = a.bound.unite(b.bound); r.initialized = true;
    return r;
}

struct BvhBuilder {
    std::vector<BuildNode> arena;
    size_t node_count = 0;

    uint32_t node_length(int n) const { return arena[n].leaf ? 1u : arena[n].len; }   // :193-199
    void to_leaf(int n, uint32_t offset, uint32_t len, BBox3 bound) {                  // :202-207
        arena[n].bound = bound; arena[n].offset = offset; arena[n].len = len; arena[n].leaf = true;
    }
    void to_interior(int n, int c0, int c1, int axis) {                                // :210-219
        arena[n].bound = arena[c0].bound.unite(arena[c1].bound);
        arena[n].child0 = c0; arena[n].child1 = c1; arena[n].axis = axis; arena[n].leaf = false;
        arena[n].offset = node_length(c0) + 1;
        arena[n].len = node_length(c0) + node_length(c1) + 1;
    }
    int alloc_node() { BuildNode b; std::memset(&b, 0, sizeof b); b.leaf = true; arena.push_back(b); return (int)arena.size() - 1; }

    static V3 sah_midpoint(const ComponentInfo* comps, size_t n, int split_axis, BBox3 cb, Float inv_area) { // :377-415
        const int BUCKETS = 32;
        Bucket buckets[BUCKETS];
        for (int i = 0; i < BUCKETS; i++) { buckets[i].count = 0; buckets[i].cost = 0.f; buckets[i].initialized = false; }
        V3 diagonal = cb.diagonal();
        for (size_t k = 0; k < n; k++) {                                               // partition() :328-345
            const ComponentInfo& c = comps[k];
            V3 dif = c.centroid - cb.pmin;
            size_t idx = (size_t)(dif[split_axis] / diagonal[split_axis] * (Float)BUCKETS);
            if (idx == (size_t)BUCKETS) idx -= 1;
            Bucket& b = buckets[idx];
            if (!b.initialized) { b.count = 1; b.cost = c.cost; b.bound = c.bound; b.initialized = true; }
            else { b.count += 1; b.cost += c.cost; b.bound = b.bound.unite(c.bound); }
        }
        Bucket accum[BUCKETS], accum_rev[BUCKETS];
        for (int i = 0; i < BUCKETS; i++) { accum[i].count = 0; accum[i].cost = 0.f; accum[i].initialized = false; accum_rev[i] = accum[i]; }
        accum[0] = buckets[0];
        accum_rev[BUCKETS - 1] = buckets[BUCKETS - 1];
        for (int i = 1; i < BUCKETS - 1; i++) {                                        // quirk A-1, verbatim
            accum[i] = bucket_union(accum[i - 1], buckets[i]);
            accum_rev[BUCKETS - 1 - i] = bucket_union(accum_rev[BUCKETS - i], buckets[BUCKETS - i]);
        }
        accum_rev[0] = bucket_union(accum_rev[1], buckets[0]);
        int boundary_idx = BUCKETS - 1;
        Float min_cost = accum_rev[0].cost;
        for (int i = 0; i < BUCKETS - 1; i++) {
            Float cost = 0.125f + (accum[i].cost * accum[i].bound.surface_area()
                                   + accum_rev[i + 1].cost * accum_rev[i + 1].bound.surface_area()) * inv_area;
            if (cost < min_cost) { boundary_idx = i; min_cost = cost; }
        }
        return cb.pmin + cb.diagonal() * ((Float)(boundary_idx + 1) / (Float)BUCKETS);
    }

    void handle_tails(ComponentInfo* comps, size_t n, uint32_t offset, ComponentInfo* ordered, int strategy,
                      size_t i, int split_axis, int ret, BBox3 bound) {                // :445-465
        if (i == 0 || i == n) to_leaf(ret, offset, (uint32_t)n, bound);
        else {
            int c0 = recursive_build(comps, i, offset, ordered, strategy);
            int c1 = recursive_build(comps + i, n - i, offset + (uint32_t)i, ordered + i, strategy);
            to_interior(ret, c0, c1, split_axis);
        }
    }
    void sort_mid(ComponentInfo* comps, size_t n, uint32_t offset, ComponentInfo* ordered, int strategy,
                  Float mid, int split_axis, int ret, BBox3 bound) {                   // :417-442
        size_t j = n, i = 0;
        for (size_t k = 0; k < n; k++) {
            if (comps[k].centroid[split_axis] < mid) { ordered[i] = comps[k]; i += 1; }
            else { j -= 1; ordered[j] = comps[k]; }
        }
        assert(j == i);
        std::memcpy(comps, ordered, n * sizeof(ComponentInfo));
        handle_tails(comps, n, offset, ordered, strategy, i, split_axis, ret, bound);
    }
    int recursive_build(ComponentInfo* comps, size_t n, uint32_t offset, ComponentInfo* ordered, int strategy) { // :246-316
        assert(n != 0);
        node_count += 1;
        int ret = alloc_node();
        if (n == 1) { to_leaf(ret, offset, 1, comps[0].bound); ordered[0] = comps[0]; return ret; }
        BBox3 b = comps[0].bound;
        BBox3 cb = BBox3::make(comps[0].centroid, comps[0].centroid);
        for (size_t k = 1; k < n; k++) { b = b.unite(comps[k].bound); cb = cb.extend(comps[k].centroid); }
        int split_axis = cb.max_extent();
        if (cb.pmin[split_axis] == cb.pmax[split_axis]) { to_leaf(ret, offset, (uint32_t)n, b); return ret; }
        switch (strategy) {
        case ARN_BVH_SAH:
            if (n <= 4) {
                // `ret = recursive_build(.., MidPoint)`: the node allocated above is abandoned
                ret = recursive_build(comps, n, offset, ordered, ARN_BVH_MIDPOINT);
            } else {
                Float inv_area = 1.f / b.surface_area();
                V3 midpoint = sah_midpoint(comps, n, split_axis, cb, inv_area);
                sort_mid(comps, n, offset, ordered, strategy, midpoint[split_axis], split_axis, ret, b);
            }
            break;
        case ARN_BVH_MIDDLECOUNT:
            handle_tails(comps, n, offset, ordered, strategy, n >> 1, split_axis, ret, b);
            break;
        default: {
            Float mid = (cb.pmax[split_axis] + cb.pmin[split_axis]) / 2.f;
            sort_mid(comps, n, offset, ordered, strategy, mid, split_axis, ret, b);
        } }
        return ret;
    }
    // BuildNode::flatten (:219-243): pre-order, second child at idx + offset
    void flatten(int root, std::vector<arn_node>& out) const {
        std::vector<int> stack; stack.push_back(root);
        while (!stack.empty()) {
            int n = stack.back(); stack.pop_back();
            const BuildNode& bn = arena[n];
            arn_node ln;
            ln.bmin[0] = bn.bound.pmin.x; ln.bmin[1] = bn.bound.pmin.y; ln.bmin[2] = bn.bound.pmin.z;
            ln.bmax[0] = bn.bound.pmax.x; ln.bmax[1] = bn.bound.pmax.y; ln.bmax[2] = bn.bound.pmax.z;
            ln.offset = bn.offset;
            if (!bn.leaf) { stack.push_back(bn.child1); stack.push_back(bn.child0); ln.len_axis = (uint32_t)bn.axis; }
            else ln.len_axis = (bn.len << 2) | 3u;
            out.push_back(ln);
        }
    }
};

// BVH::new (:58-79)
inline void bvh_build(uint32_t n, const Float* bounds6, const Float* costs, int strategy,
                      std::vector<arn_node>& nodes, std::vector<uint32_t>& order) {
    std::vector<ComponentInfo> cinfo(n);
    for (uint32_t i = 0; i < n; i++) {
        ComponentInfo& c = cinfo[i];
        c.bound.pmin = v3(bounds6[6*i], bounds6[6*i+1], bounds6[6*i+2]);
        c.bound.pmax = v3(bounds6[6*i+3], bounds6[6*i+4], bounds6[6*i+5]);
        c.centroid = (c.bound.pmin + c.bound.pmax) / 2.f;
        c.cost = costs[i]; c.idx = i;
    }
    std::vector<ComponentInfo> ordered = cinfo;
    BvhBuilder b; b.arena.reserve(2 * (size_t)n + 64);
    int root = b.recursive_build(cinfo.data(), n, 0, ordered.data(), strategy);
    nodes.clear(); nodes.reserve(b.node_count);
    b.flatten(root, nodes);
    order.resize(n);
    for (uint32_t i = 0; i < n; i++) order[i] = ordered[i].idx;
}

// ------------------------------------------------------------------ BVH traversal (:97-128)
struct TraversalCounters { uint64_t nodes = 0, tris = 0, spheres = 0; };
inline bool bvh_intersect(const Scene& s, RawRay& ray, SurfaceInteraction* final_si, int* prim_out,
                          TraversalCounters* ctr, bool full_si) {
    std::vector<uint32_t> stack; stack.reserve(64); stack.push_back(0);
    bool found = false; SurfaceInteraction fsi; int fprim = -1;
    RayCache cache = construct_ray_cache(ray);
    while (!stack.empty()) {
        uint32_t idx = stack.back(); stack.pop_back();
        const arn_node& node = s.nodes[idx];
        BBox3 nb; nb.pmin = v3(node.bmin[0], node.bmin[1], node.bmin[2]); nb.pmax = v3(node.bmax[0], node.bmax[1], node.bmax[2]);
        if (ctr) ctr->nodes++;
        if (!intersect_ray_cached(nb, cache)) continue;
        uint32_t len = node.len_axis >> 2;
        if (len > 0) {
            for (uint32_t k = node.offset; k < node.offset + len; k++) {
                uint32_t comp = s.order[k];
                RawRay iray = ray;
                SurfaceInteraction tsi;
                if (ctr) { if (s.prim_is_sphere(comp)) ctr->spheres++; else ctr->tris++; }
                bool hit = component_intersect(s, comp, iray, &tsi, full_si);
                if (ray.tmax > iray.tmax) {          // strict: ties keep the first found (:110)
                    ray = iray;
                    cache.tmax = ray.tmax;           // origin / inv_dir of the cache are NOT refreshed (:112)
                    found = hit; fsi = tsi; fprim = (int)comp;
                }
            }
        } else {
            uint32_t axis = node.len_axis & 3u;
            if (cache.neg[axis]) { stack.push_back(idx + 1); stack.push_back(idx + node.offset); }
            else { stack.push_back(idx + node.offset); stack.push_back(idx + 1); }
        }
    }
    if (found) { if (final_si) *final_si = fsi; if (prim_out) *prim_out = fprim; }
    else if (prim_out) *prim_out = -1;
    return found;
}

}  // namespace orc
