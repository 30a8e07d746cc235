// FlatScene — host-side scene assembly and flattening for the B200 path-tracing core.
// Plays the role of arendur's component list + `BVH::new` + `Scene::new`
// (examples/arencli.rs:186-197, src/component/bvh.rs:58-79, src/renderer/scene.rs:31-51)
// and produces the POD `arn_scene_desc` that crosses the C-ABI (include/arn.h).
#pragma once
#include <cmath>
#include <string>
#include <vector>
#include "../../../include/arn.h"
#include "hostmath.hpp"

namespace arnhost {

// bxdf/microfacet.rs:57-63
inline float roughness_to_alpha(float roughness) {
    float r = roughness > 1e-3f ? roughness : 1e-3f;   // f32::max
    float x = (float)std::log((double)r);              // correctly-rounded f32 (see kernels/dev_math.cuh)
    return 1.62142f + 0.819955f * x + 0.1734f * x * x + 0.0171201f * x * x * x + 0.000640711f * x * x * x * x;
}

class FlatScene {
public:
    std::vector<float> positions, normals, uvs;
    std::vector<uint32_t> indices, tri_mesh;
    std::vector<arn_mesh> meshes;
    std::vector<arn_sphere> spheres;
    std::vector<arn_material> materials;
    std::vector<uint32_t> prims;
    std::vector<arn_node> nodes;
    std::vector<uint32_t> order;
    std::vector<uint32_t> light_prims;            // emissive primitives, in component order
    std::vector<arn_analytic_light> analytic;     // Point / Spot / Distant lights (come first in Scene.lights)
    std::vector<uint32_t> light_list;             // Scene.lights: analytic lights, then light_prims
    std::vector<float> light_func, light_cdf;
    float light_integral = 0.f;
    std::vector<arn_texture> textures;            // image textures (N4): finished pyramids handed in by the caller
    std::vector<float> texels;
    bool any_normals = false, any_uvs = false, built = false;
    arn_scene_desc desc;
    std::string err;

    FlatScene() { std::memset(&desc, 0, sizeof desc); }

    int fail(int code, const std::string& msg) { err = msg; return code; }

    int add_material(const arn_material& m_in) {
        arn_material m = m_in;
        if (m.type > ARN_MAT_TRANSLUCENT) return fail(ARN_E_INVALID, "unknown material type");
        if (m.kd_tex > textures.size() || m.ks_tex > textures.size() || m.aux_tex > textures.size() || m.bump_tex > textures.size()) return fail(ARN_E_INVALID, "material references a texture that was not added");
        if (m.type == ARN_MAT_MATTE) {
            // MatteMaterial::compute_scattering clamps sigma to [0, 90] (material/matte.rs:50-54)
            if (m.sigma < 0.f) m.sigma = 0.f; else if (!(m.sigma < 90.f)) m.sigma = 90.f;
        }
        m.alpha = roughness_to_alpha(m.roughness);
        materials.push_back(m); built = false;
        return (int)materials.size() - 1;
    }

    // TriangleMesh::from_model[_transformed] (shape/triangle.rs:82-160)
    int add_mesh(const float* pos, uint32_t nv, const uint32_t* idx, uint32_t ni, const float* nrm, const float* uv,
                 const float* transform16, uint32_t material) {
        if (!pos || !idx || nv == 0) return fail(ARN_E_INVALID, "mesh without vertices or indices");
        if (material >= materials.size()) return fail(ARN_E_INVALID, "mesh material id out of range");
        uint32_t ntri = 0; for (uint32_t i = 0; i + 2 < ni; i += 3) ntri++;     // iterator: while idx + 2 < len
        for (uint32_t i = 0; i < ntri * 3; i++) if (idx[i] >= nv) return fail(ARN_E_INVALID, "mesh index out of range");
        uint32_t base = (uint32_t)(positions.size() / 3);
        bool xf = transform16 != nullptr;
        Mat4 t = xf ? Mat4::from_array(transform16) : Mat4::identity();
        Mat4 inv_t; bool have_inv = false;
        if (xf && nrm) { have_inv = invert(t, &inv_t); if (!have_inv) return fail(ARN_E_INVALID, "mesh transform is not invertible"); inv_t = transpose(inv_t); }
        positions.resize((size_t)(base + nv) * 3);
        normals.resize((size_t)(base + nv) * 3, 0.f);
        uvs.resize((size_t)(base + nv) * 2, 0.f);
        for (uint32_t i = 0; i < nv; i++) {
            Vec3 p{pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]};
            if (xf) p = transform_point(t, p);                                   // homogeneous divide (quirk A-12)
            float* o = &positions[(size_t)(base + i) * 3]; o[0] = p.x; o[1] = p.y; o[2] = p.z;
            if (nrm) {
                Vec3 n{nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2]};
                if (xf) n = normalize(transform_vector(inv_t, n));                // transform_norm
                float* q = &normals[(size_t)(base + i) * 3]; q[0] = n.x; q[1] = n.y; q[2] = n.z;
            }
            if (uv) { uvs[(size_t)(base + i) * 2] = uv[2 * i]; uvs[(size_t)(base + i) * 2 + 1] = uv[2 * i + 1]; }
        }
        arn_mesh m; m.material = material; m.has_normals = nrm ? 1u : 0u; m.has_uvs = uv ? 1u : 0u; m.reserved = 0;
        uint32_t mesh_id = (uint32_t)meshes.size();
        meshes.push_back(m);
        any_normals |= nrm != nullptr; any_uvs |= uv != nullptr;
        for (uint32_t k = 0; k < ntri; k++) {
            uint32_t tri = (uint32_t)tri_mesh.size();
            indices.push_back(base + idx[3 * k]); indices.push_back(base + idx[3 * k + 1]); indices.push_back(base + idx[3 * k + 2]);
            tri_mesh.push_back(mesh_id);
            prims.push_back(tri);
        }
        built = false;
        return (int)mesh_id;
    }

    // Sphere::new (shape/sphere.rs:133-156) + ShapedPrimitive (+ TransformedComposable)
    // `in_lights` = false: an emissive primitive that arencli does not push to `lights` (an instance made by
    // ComponentDesc::Transformed, arencli.rs:162-181): it shows its emission when hit, but is never sampled.
    int add_sphere(float radius, float zmin, float zmax, float phimax, uint32_t material, const float* emission3,
                   const float* transform16, bool in_lights = true, bool* transform_kept = nullptr) {
        if (!(radius > 0.f)) return fail(ARN_E_INVALID, "Sphere radius should be positive");
        if (!(zmin < zmax)) return fail(ARN_E_INVALID, "zmin should be lower than zmax");
        if (material >= materials.size()) return fail(ARN_E_INVALID, "sphere material id out of range");
        arn_sphere s; std::memset(&s, 0, sizeof s);
        if (zmin < -radius) zmin = -radius;
        if (zmax > radius) zmax = radius;
        if (phimax < 0.f) phimax = 0.f;
        const float twopi = 3.14159265358979323846f * 2.0f;
        if (phimax > twopi) phimax = twopi;
        s.radius = radius; s.zmin = zmin; s.zmax = zmax; s.phimax = phimax;
        s.thetamin = (float)std::acos((double)(zmin / radius)); s.thetamax = (float)std::acos((double)(zmax / radius));
        s.material = material;
        if (emission3) { s.emissive = 1; s.emission[0] = emission3[0]; s.emission[1] = emission3[1]; s.emission[2] = emission3[2]; }
        Mat4 lp = Mat4::identity(), pl = Mat4::identity();
        if (transform16) {
            lp = Mat4::from_array(transform16);
            // arencli.rs:133-146: a non-invertible transform silently degrades to the bare primitive
            if (invert(lp, &pl)) s.has_transform = 1; else { lp = Mat4::identity(); pl = Mat4::identity(); }
        }
        if (transform_kept) *transform_kept = s.has_transform != 0;
        lp.to_array(s.local_parent); pl.to_array(s.parent_local);
        uint32_t sid = (uint32_t)spheres.size();
        spheres.push_back(s);
        uint32_t comp = (uint32_t)prims.size();
        prims.push_back(ARN_PRIM_SPHERE | sid);
        if (s.emissive && in_lights) light_prims.push_back(comp);                // `lights.push(sp.clone())`
        built = false;
        return (int)comp;
    }

    // ImageTexture::new's result, flattened: `t.level_offset` is relative to `level_texels` on entry.  Returns the id materials use (k + 1).
    int add_texture(const arn_texture& t_in, const float* level_texels, uint64_t n_floats) {
        arn_texture t = t_in;
        if ((t.channels != 1 && t.channels != 3) || t.n_levels < 1 || t.n_levels > ARN_TEX_MAX_LEVELS || t.wrapping > ARN_WRAP_CLAMP || !level_texels)
            return fail(ARN_E_INVALID, "texture: channels must be 1 or 3, 1..16 levels, a valid wrap mode");
        for (uint32_t l = 0; l < t.n_levels; l++) {
            if (!t.level_w[l] || !t.level_h[l] || (uint64_t)t.level_offset[l] + (uint64_t)t.level_w[l] * t.level_h[l] * t.channels > n_floats) return fail(ARN_E_INVALID, "texture level outside the texel array");
            t.level_offset[l] += (uint32_t)texels.size();
        }
        texels.insert(texels.end(), level_texels, level_texels + n_floats);
        textures.push_back(t); built = false;
        return (int)textures.size();
    }

    int add_light(const arn_analytic_light& l) {
        if (l.type > ARN_LIGHT_DISTANT) return fail(ARN_E_INVALID, "unknown light type");
        analytic.push_back(l);
        built = false;
        return (int)analytic.size() - 1;
    }

    // Composable::bbox_parent + intersection_cost of component `i` (ComponentInfo::new, bvh.rs:24-35)
    void component_bounds(uint32_t i, float* b6, float* cost) const {
        uint32_t ref = prims[i];
        if (ref & ARN_PRIM_SPHERE) {
            const arn_sphere& s = spheres[ref & ~ARN_PRIM_SPHERE];
            float lo[3] = {-s.radius, -s.radius, s.zmin}, hi[3] = {s.radius, s.radius, s.zmax};   // Sphere::bounding
            if (s.has_transform) {
                // BBox3::apply_transform (geometry/bbox.rs:481-499): pmin as point, diagonal as vector
                Mat4 lp = Mat4::from_array(s.local_parent);
                Vec3 p = transform_point(lp, Vec3{lo[0], lo[1], lo[2]});
                Vec3 d = transform_vector(lp, Vec3{hi[0] - lo[0], hi[1] - lo[1], hi[2] - lo[2]});
                float q[3] = {p.x + d.x, p.y + d.y, p.z + d.z}, pp[3] = {p.x, p.y, p.z};
                for (int k = 0; k < 3; k++) { b6[k] = pp[k] < q[k] ? pp[k] : q[k]; b6[3 + k] = pp[k] > q[k] ? pp[k] : q[k]; }
                *cost = 1.0f + 1.0f;                                              // transformed.rs:49-51
            } else { for (int k = 0; k < 3; k++) { b6[k] = lo[k]; b6[3 + k] = hi[k]; } *cost = 1.0f; }
        } else {
            const float* p0 = &positions[(size_t)indices[3 * (size_t)ref] * 3];
            const float* p1 = &positions[(size_t)indices[3 * (size_t)ref + 1] * 3];
            const float* p2 = &positions[(size_t)indices[3 * (size_t)ref + 2] * 3];
            for (int k = 0; k < 3; k++) {                                         // BBox3::new(x, y).extend(z)
                float lo = p0[k] < p1[k] ? p0[k] : p1[k], hi = p0[k] > p1[k] ? p0[k] : p1[k];
                b6[k] = lo < p2[k] ? lo : p2[k]; b6[3 + k] = hi > p2[k] ? hi : p2[k];
            }
            *cost = 3.0f;                                                         // triangle.rs:536-538
        }
    }

    float last_build_ms = 0.f;       // device time of the last GPU build
    // BVH::new + Scene::new.  `gpu` != NULL: the tree comes from arn_bvh_build_gpu (LBVH, not the reference's
    // topology) instead of the reference's host builder.
    int build(int strategy, arn_ctx* gpu = nullptr) {
        uint32_t n = (uint32_t)prims.size();
        if (n == 0) return fail(ARN_E_INVALID, "scene has no components");
        std::vector<float> b6((size_t)n * 6), cost(n);
        for (uint32_t i = 0; i < n; i++) component_bounds(i, &b6[(size_t)i * 6], &cost[i]);
        nodes.resize(2 * (size_t)n); order.resize(n);
        uint32_t nn = 0;
        int rc = gpu ? arn_bvh_build_gpu(gpu, n, b6.data(), nodes.data(), order.data(), &nn, &last_build_ms)
                     : arn_bvh_build(n, b6.data(), cost.data(), strategy, nodes.data(), order.data(), &nn);
        if (rc != ARN_OK) return fail(rc, gpu ? std::string("arn_bvh_build_gpu failed: ") + arn_last_error(gpu) : std::string("arn_bvh_build failed"));
        nodes.resize(nn);
        // Scene::new: power().to_xyz().y per light (renderer/scene.rs:36-41, component/shape.rs:160-167)
        light_func.clear(); light_list.clear();
        const float pi = 3.14159265358979323846f;
        for (size_t k = 0; k < analytic.size(); k++) {
            const arn_analytic_light& l = analytic[k];
            // Light::power: pointlights.rs:79-81 (I * (pi * 4)), :222-226 (I * (pi * 2) * (1 - 0.5 * (cosf - cost))),
            // distantlight.rs:108-110 (I * (r * r * pi))
            float scale = l.type == ARN_LIGHT_POINT ? pi * 4.0f
                        : l.type == ARN_LIGHT_SPOT ? 0.f : l.world_radius * l.world_radius * pi;
            float r = l.intensity[0] * scale, g = l.intensity[1] * scale, b = l.intensity[2] * scale;
            if (l.type == ARN_LIGHT_SPOT) {
                float c = 1.0f - 0.5f * (l.cosf - l.cost);
                r = l.intensity[0] * (pi * 2.0f) * c; g = l.intensity[1] * (pi * 2.0f) * c; b = l.intensity[2] * (pi * 2.0f) * c;
            }
            light_func.push_back(0.212671f * r + 0.715160f * g + 0.072169f * b);
            light_list.push_back(ARN_LIGHT_ANALYTIC | (uint32_t)k);
        }
        for (uint32_t lp : light_prims) {
            const arn_sphere& s = spheres[prims[lp] & ~ARN_PRIM_SPHERE];
            float area = s.phimax * s.radius * (s.zmax - s.zmin);                 // Sphere::surface_area
            float r = s.emission[0] * area * pi, g = s.emission[1] * area * pi, b = s.emission[2] * area * pi;
            light_func.push_back(0.212671f * r + 0.715160f * g + 0.072169f * b);
            light_list.push_back(lp);
        }
        light_cdf.assign(light_func.size() + 1, 0.f);
        rc = arn_light_distribution((uint32_t)light_func.size(), light_func.data(), light_cdf.data(), &light_integral);
        if (rc != ARN_OK) return fail(rc, "light power distribution is invalid");
        fill_desc();
        built = true;
        return ARN_OK;
    }

    void fill_desc() {
        std::memset(&desc, 0, sizeof desc);
        desc.n_vertices = (uint32_t)(positions.size() / 3);
        desc.positions = positions.data();
        desc.normals = any_normals ? normals.data() : nullptr;
        desc.uvs = any_uvs ? uvs.data() : nullptr;
        desc.n_triangles = (uint32_t)tri_mesh.size();
        desc.indices = indices.data(); desc.tri_mesh = tri_mesh.data();
        desc.n_meshes = (uint32_t)meshes.size(); desc.meshes = meshes.data();
        desc.n_spheres = (uint32_t)spheres.size(); desc.spheres = spheres.data();
        desc.n_materials = (uint32_t)materials.size(); desc.materials = materials.data();
        desc.n_prims = (uint32_t)prims.size(); desc.prims = prims.data();
        desc.n_nodes = (uint32_t)nodes.size(); desc.nodes = nodes.data(); desc.order = order.data();
        desc.n_lights = (uint32_t)light_list.size(); desc.light_prims = light_list.data();
        desc.n_analytic_lights = (uint32_t)analytic.size(); desc.analytic_lights = analytic.data();
        desc.light_func = light_func.data(); desc.light_cdf = light_cdf.data(); desc.light_func_integral = light_integral;
        desc.n_textures = (uint32_t)textures.size(); desc.textures = textures.empty() ? nullptr : textures.data();
        desc.n_texel_floats = texels.size(); desc.texels = texels.empty() ? nullptr : texels.data();
    }
};

}  // namespace arnhost
