// BVH traversal + ray/primitive intersection for sm_100a (kernels K2 / K4 of SURVEY.md §2.1).
//
// Replaces, per ray: BVH::intersect_ray (src/component/bvh.rs:97-128),
// BBox3f::intersect_ray_cached (src/geometry/bbox.rs:549-580), ShearingTransformCache
// (src/geometry/ray.rs:190-235), TriangleInstance::intersect_ray up to the acceptance test
// (src/shape/triangle.rs:396-451), Sphere::intersect_ray (src/shape/sphere.rs:193-297) and the
// TransformedComposable ray round trip (src/component/transformed.rs:73-83).
//
// Layout in HBM (built at upload):
//   nodes : 32 B each, 32-B aligned, read as two 128-bit loads
//           q0 = (bmin.x, bmin.y, bmin.z, bmax.x)  q1 = (bmax.y, bmax.z, offset, len_axis)
//   tris  : one 48-B record per ORDERED primitive slot (BVH leaf order), three 128-bit loads
//           (p0.xyz, component id) (p1.xyz, -) (p2.xyz, -); sphere slots carry the component id
//           with ARN_PRIM_SPHERE set and no vertices.
// Traversal order, the strict `<` acceptance and the `t0 < tmax` re-check at pop time
// reproduce the reference exactly; the slab test of a child is evaluated when its parent is
// expanded (its t0 travels on the stack) instead of after the pop — the outcome is identical
// because only the final `t0 < tmax` comparison depends on the shrinking tmax.
#pragma once
#include "dev_math.cuh"
#include "../../../include/arn.h"

namespace arn {

typedef arn_sphere DevSphere;   // same POD on host and device (176 B)

struct DevScene {
    const float4* __restrict__ nodes;     // 2 per node
    const float4* __restrict__ tris;      // 3 per ordered slot
    const DevSphere* __restrict__ spheres;
    // shading data, indexed by triangle id (input order)
    const uint32_t* __restrict__ indices;    // 3 per triangle
    const float* __restrict__ positions;     // 3 per vertex
    const float* __restrict__ normals;       // 3 per vertex (or null)
    const float* __restrict__ uvs;           // 2 per vertex (or null)
    const uint32_t* __restrict__ tri_mesh;   // triangle -> mesh
    const arn_mesh* __restrict__ meshes;
    const arn_material* __restrict__ materials;
    const uint32_t* __restrict__ prims;      // component -> triangle id | sphere bit
    const uint32_t* __restrict__ light_prims;
    const arn_analytic_light* __restrict__ analytic;   // Point / Spot / Distant lights
    const float* __restrict__ light_func;
    const float* __restrict__ light_cdf;
    float light_integral;
    uint32_t n_lights, n_nodes, n_prims, n_spheres;
    // 4-wide collapse of `nodes` (built at upload, see traverse4): 4 records of 32 B per wide node
    const float4* __restrict__ wide;
    float4 root0, root1;                     // record of the root (bounds of nodes[0] + its reference words)
};

#define ARN_STACK 64           /* upload rejects trees deeper than this */
#define ARN_STACK4 96          /* wide traversal: <= 3 pushes per wide level, ARN_STACK/2 wide levels */
// reference words of a wide record: w0 = q1.z, w1 = q1.w
//   w1 & 3 : 0 empty slot, 1 interior (w0 = wide node index, (w1 >> 2) & 63 = split axes of that node:
//            its own | first child's << 2 | second child's << 4), 2 leaf (w0 = first slot, w1 >> 8 = count)
#define ARN_W_EMPTY 0u
#define ARN_W_INNER 1u
#define ARN_W_LEAF 2u

// Ray state used by traversal: the slab cache keeps the ORIGINAL origin / 1/dir
// (bvh.rs:101,112 refreshes only tmax), the shear cache follows the CURRENT ray
// (ray.rs:136-139), which differs only after a transformed-sphere hit.
struct TravRay {
    float3 co, inv;            // cache: origin, 1/dir
    float3 o, d;               // current ray
    float tmax;
    int kz;                    // 0 = XZ perm, 1 = YZ, 2 = ZZ
    float3 shear;
};

ARN_DEV void shear_setup(TravRay& r) {                     // ShearingTransformCache::from_ray
    float ax = fabsf(r.d.x), ay = fabsf(r.d.y), az = fabsf(r.d.z);
    float3 dd;
    if (ax > ay && ax > az) { r.kz = 0; dd = f3(r.d.y, r.d.z, r.d.x); }
    else if (ay > az)       { r.kz = 1; dd = f3(r.d.z, r.d.x, r.d.y); }
    else                    { r.kz = 2; dd = r.d; }
    r.shear = f3(-dd.x / dd.z, -dd.y / dd.z, 1.f / dd.z);
}
ARN_DEV void trav_init(TravRay& r, float3 o, float3 d, float tmax) {
    r.o = o; r.d = d; r.tmax = tmax; r.co = o;
    r.inv = f3(1.f / d.x, 1.f / d.y, 1.f / d.z);            // construct_ray_cache (bbox.rs:583-592)
    shear_setup(r);
}

// Slab test without the tmax comparison: returns false on a definite miss, else t0.
// Branch-free form of BBox3f::intersect_ray_cached (bbox.rs:549-580): the reference's two early
// `return None` become predicates (the updates they skip cannot change a `false` result), the
// comparisons and selects are the reference's own (NaNs propagate identically; no fmin/fmax).
ARN_DEV bool slab(const float4 q0, const float4 q1, const TravRay& r, float& t0_out) {
    const float k = 1.f + 2.f * gamma_n(3.f);
    const bool nx = r.inv.x < 0.f, ny = r.inv.y < 0.f, nz = r.inv.z < 0.f;
    const float bminx = q0.x, bminy = q0.y, bminz = q0.z, bmaxx = q0.w, bmaxy = q1.x, bmaxz = q1.y;
    float t0 = ((nx ? bmaxx : bminx) - r.co.x) * r.inv.x;
    float t1 = ((nx ? bminx : bmaxx) - r.co.x) * r.inv.x;
    float ty0 = ((ny ? bmaxy : bminy) - r.co.y) * r.inv.y;
    float ty1 = ((ny ? bminy : bmaxy) - r.co.y) * r.inv.y;
    float tz0 = ((nz ? bmaxz : bminz) - r.co.z) * r.inv.z;
    float tz1 = ((nz ? bminz : bmaxz) - r.co.z) * r.inv.z;
    t1 *= k; ty1 *= k; tz1 *= k;
    const bool miss_xy = (t0 > ty1) | (ty0 > t1);
    t0 = ty0 > t0 ? ty0 : t0;
    t1 = ty1 < t1 ? ty1 : t1;
    const bool miss_z = (t0 > tz1) | (tz0 > t1);
    t0 = tz0 > t0 ? tz0 : t0;
    t1 = tz1 < t1 ? tz1 : t1;
    t0_out = t0;
    return !miss_xy & !miss_z & (t1 > 0.f);     // caller adds `t0 < tmax` (NaN t0 fails it, as in the reference)
}

ARN_DEV float3 perm_point(float3 p, int kz) {
    return kz == 0 ? f3(p.y, p.z, p.x) : (kz == 1 ? f3(p.z, p.x, p.y) : p);
}

// Watertight ray/triangle acceptance test (triangle.rs:398-451).  Returns true and t, b0..b2.
ARN_DEV bool tri_test(float3 p0, float3 p1, float3 p2, const TravRay& r, float& t, float& b0, float& b1, float& b2) {
    float3 no = -r.o;
    float3 q0 = perm_point(p0 + no, r.kz), q1 = perm_point(p1 + no, r.kz), q2 = perm_point(p2 + no, r.kz);
    q0.x += r.shear.x * q0.z; q0.y += r.shear.y * q0.z;
    q1.x += r.shear.x * q1.z; q1.y += r.shear.y * q1.z;
    q2.x += r.shear.x * q2.z; q2.y += r.shear.y * q2.z;
    float e0 = q1.x * q2.y - q1.y * q2.x;
    float e1 = q2.x * q0.y - q2.y * q0.x;
    float e2 = q0.x * q1.y - q0.y * q1.x;
    if ((e0 < 0.f || e1 < 0.f || e2 < 0.f) && (e0 > 0.f || e1 > 0.f || e2 > 0.f)) return false;
    float det = e0 + e1 + e2;
    if (det == 0.f) return false;
    q0.z *= r.shear.z; q1.z *= r.shear.z; q2.z *= r.shear.z;
    float ts = e0 * q0.z + e1 * q1.z + e2 * q2.z;
    if (det < 0.f && (ts >= 0.f || ts < r.tmax * det)) return false;
    else if (det > 0.f && (ts <= 0.f || ts > r.tmax * det)) return false;
    float inv_det = 1.f / det;
    b0 = e0 * inv_det; b1 = e1 * inv_det; b2 = e2 * inv_det;
    t = ts * inv_det;
    float maxxt = fmaxf(fmaxf(q0.x, q1.x), q2.x);
    float maxyt = fmaxf(fmaxf(q0.y, q1.y), q2.y);
    float maxzt = fmaxf(fmaxf(q0.z, q1.z), q2.z);
    float maxe = fmaxf(fmaxf(e0, e1), e2);
    float deltax = maxxt * gamma_n(5.f);
    float deltay = maxyt * gamma_n(5.f);
    float deltaz = maxzt * gamma_n(3.f);
    float delta_err = 2.f * (gamma_n(2.f) * maxxt * maxyt + deltay * maxxt + deltax * maxyt);
    float delta_t = 3.f * (gamma_n(3.f) * maxe * maxzt + delta_err * maxzt + deltaz * maxe) * fabsf(inv_det);
    if (t <= delta_t) return false;
    return true;
}

// Sphere::intersect_ray_full + refinement + clipping (sphere.rs:193-250), local space.
// On a hit returns t and the refined local hit point p.
ARN_NOINL bool sphere_test(const DevSphere& sp, float3 o, float3 d, float tmax, float& t_out, float3& p_out) {
    float a = dot(d, d);
    float3 m = (d * o) * 2.f;
    float b = m.x + m.y + m.z;
    float c = dot(o, o) - sp.radius * sp.radius;
    float delta = b * b - 4.f * a * c;
    if (delta < 0.f) return false;
    float invert_2a = 1.f / (2.f * a);
    float d1 = sqrtf(delta) * invert_2a;
    float d0 = -b * invert_2a;
    float t0, t1;
    if (invert_2a > 0.f) { t0 = d0 - d1; t1 = d0 + d1; } else { t0 = d0 + d1; t1 = d0 - d1; }
    if (t0 > tmax || t1 < 0.f) return false;
    float t;
    if (t0 > 0.f) t = t0; else if (t1 > tmax) return false; else t = t1;
    float3 p = o + d * t;
    p = p * sp.radius / length(p);
    if (p.x == 0.f && p.y == 0.f) p.x = 1e-5f * sp.radius;
    // phi > phimax clip: decided with the cheap f32 atan2f unless phi lies within 1e-4 of phimax or of the
    // 0 / 2pi seam, where the reference's correctly-rounded value is computed (identical decisions:
    // libdevice atan2f is accurate to ~1e-6, far inside the 1e-4 band).
    float phi = atan2f(p.y, p.x);
    if (phi < 0.f) phi += 2.f * ARN_PI;
    if (fabsf(phi - sp.phimax) < 1e-4f || phi < 1e-4f || phi > 2.f * ARN_PI - 1e-4f) {
        phi = cr_atan2f(p.y, p.x);
        if (phi < 0.f) phi += 2.f * ARN_PI;
    }
    if (p.z < sp.zmin || p.z > sp.zmax || phi > sp.phimax) return false;
    t_out = t; p_out = p;
    return true;
}

struct HitRec {               // what shading needs from the final hit
    int prim;                 // component index, -1 = miss
    float t;
    float a, b, c;            // triangle: b0,b1,b2; sphere: refined local hit point
};

// Component test for a sphere slot inside traversal.  Mirrors `iray = ray.clone();
// element.intersect_ray(&mut iray); if ray.tmax > iray.tmax { *ray = iray }` (bvh.rs:108-113)
// through TransformedComposable (transformed.rs:73-83): on an accepted hit the traversal ray
// becomes the round-tripped one.
ARN_DEV void sphere_slot(const DevScene& sc, uint32_t comp, TravRay& r, HitRec& h) {
    const DevSphere& sp = sc.spheres[sc.prims[comp] & ~ARN_PRIM_SPHERE];
    float3 lo = r.o, ld = r.d;
    if (sp.has_transform) { lo = xform_point(sp.parent_local, r.o); ld = xform_vector(sp.parent_local, r.d); }
    float t; float3 p;
    if (!sphere_test(sp, lo, ld, r.tmax, t, p)) return;
    if (!(r.tmax > t)) return;
    if (sp.has_transform) {
        r.o = xform_point(sp.local_parent, lo); r.d = xform_vector(sp.local_parent, ld);
        shear_setup(r);
    }
    r.tmax = t;
    h.prim = (int)comp; h.t = t; h.a = p.x; h.b = p.y; h.c = p.z;
}

// Closest hit (ANY = false) or any hit (ANY = true: stops at the first accepted primitive —
// the reference's shadow rays run the closest-hit query and only look at is_some(),
// component/mod.rs:35-38, so the boolean is identical).
// COUNT: accumulate nodes/primitives tested into ctr[0..2] (for the algorithmic-bytes figure).
// pop the next stack entry whose entry distance is still below tmax; false when the stack is empty
ARN_DEV bool trav_pop(const DevScene& sc, const TravRay& r, const uint2* stack, int& sp,
                      uint32_t& idx, uint32_t& offset, uint32_t& len_axis) {
    for (;;) {
        if (sp == 0) return false;
        uint2 e = stack[--sp];
        if (__uint_as_float(e.y) < r.tmax) {
            idx = e.x;
            float4 n1 = __ldg(&sc.nodes[2 * idx + 1]);
            offset = __float_as_uint(n1.z); len_axis = __float_as_uint(n1.w);
            return true;
        }
    }
}

// "while-while" form (Aila & Laine): all lanes of a warp first descend through interior nodes,
// then test leaf primitives together, so the long triangle test runs with many lanes active.
template <bool ANY, bool COUNT>
ARN_DEV void traverse(const DevScene& sc, TravRay& r, HitRec& h, uint32_t* ctr) {
    h.prim = -1; h.t = ARN_INF; h.a = h.b = h.c = 0.f;
    uint2 stack[ARN_STACK];
    int sp = 0;
    // root: tested like any popped node
    float4 q0 = __ldg(&sc.nodes[0]), q1 = __ldg(&sc.nodes[1]);
    float t0;
    if (COUNT) ctr[0]++;
    if (!slab(q0, q1, r, t0) || !(t0 < r.tmax)) return;
    uint32_t idx = 0;
    uint32_t offset = __float_as_uint(q1.z), len_axis = __float_as_uint(q1.w);
    for (;;) {
        // ---- interior nodes: expand both children (first child = idx+1, second = idx+offset)
        bool alive = true;
        while ((len_axis >> 2) == 0) {
            uint32_t ia = idx + 1, ib = idx + offset;
            float4 a0 = __ldg(&sc.nodes[2 * ia]), a1 = __ldg(&sc.nodes[2 * ia + 1]);
            float4 b0 = __ldg(&sc.nodes[2 * ib]), b1 = __ldg(&sc.nodes[2 * ib + 1]);
            if (COUNT) ctr[0] += 2;
            float ta, tb;
            bool ha = slab(a0, a1, r, ta) && ta < r.tmax;
            bool hb = slab(b0, b1, r, tb) && tb < r.tmax;
            bool neg = axis_of(r.inv, (int)(len_axis & 3u)) < 0.f;     // dir_is_neg[split_axis]: second child first
            if (ha && hb) {
                bool first_b = neg;
                stack[sp++] = first_b ? make_uint2(ia, __float_as_uint(ta)) : make_uint2(ib, __float_as_uint(tb));
                idx = first_b ? ib : ia;
                offset = __float_as_uint(first_b ? b1.z : a1.z); len_axis = __float_as_uint(first_b ? b1.w : a1.w);
            } else if (ha || hb) {
                idx = ha ? ia : ib;
                offset = __float_as_uint(ha ? a1.z : b1.z); len_axis = __float_as_uint(ha ? a1.w : b1.w);
            } else if (!trav_pop(sc, r, stack, sp, idx, offset, len_axis)) { alive = false; break; }
        }
        if (!alive) return;
        // ---- leaf: primitives in slot order, strict `<` acceptance
        const uint32_t end = offset + (len_axis >> 2);
        for (uint32_t k = offset; k < end; k++) {
            float4 v0 = __ldg(&sc.tris[3 * k]);
            uint32_t comp = __float_as_uint(v0.w);          // component id, sphere bit set for sphere slots
            if (comp & ARN_PRIM_SPHERE) { if (COUNT) ctr[2]++; sphere_slot(sc, comp & ~ARN_PRIM_SPHERE, r, h); }
            else {
                float4 v1 = __ldg(&sc.tris[3 * k + 1]), v2 = __ldg(&sc.tris[3 * k + 2]);
                if (COUNT) ctr[1]++;
                float t, b0, b1, b2;
                if (tri_test(f3(v0.x, v0.y, v0.z), f3(v1.x, v1.y, v1.z), f3(v2.x, v2.y, v2.z), r, t, b0, b1, b2) && r.tmax > t) {
                    r.tmax = t; h.prim = (int)comp; h.t = t; h.a = b0; h.b = b1; h.c = b2;
                }
            }
            if (ANY && h.prim >= 0) return;
        }
        if (!trav_pop(sc, r, stack, sp, idx, offset, len_axis)) return;
    }
}

// ---- 4-wide traversal ------------------------------------------------------------------------
// The wide tree is the binary tree with every second level removed: wide node = (children of the
// first child, children of the second child), a leaf child occupying one slot of its pair.  The four
// records are visited in the order the binary depth-first walk would reach them (pair order from the
// node's split axis, order inside a pair from the child's split axis — signs of the ray direction, not
// distances), entry distances travel on the stack and are re-checked against the shrinking tmax at pop
// time.  Result: the same leaves are visited in the same order as in BVH::intersect_ray
// (bvh.rs:97-128), so hits, ids and ties are bit-identical to the binary walk:
//   * a grandchild's slab interval is contained in its parent's (same monotone f32 operations on
//     nested bounds), so skipping the parent's test never admits a leaf the binary walk would reject
//     — the leaf's own test, made with the same tmax as in the binary walk, decides;
//   * only the `t0 < tmax` part of a slab test depends on when it is evaluated, and that part is
//     re-evaluated when the record is popped.
// tests/test_gpu_parity.py::test_wide_equals_binary checks it ray by ray against the binary kernel.
ARN_DEV bool trav_pop4(const TravRay& r, const uint4* stack, int& sp, uint32_t& w0, uint32_t& w1) {
    for (;;) {
        if (sp == 0) return false;
        uint4 e = stack[--sp];
        if (__uint_as_float(e.z) < r.tmax) { w0 = e.x; w1 = e.y; return true; }
    }
}
template <bool ANY>
ARN_DEV void traverse4(const DevScene& sc, TravRay& r, HitRec& h) {
    h.prim = -1; h.t = ARN_INF; h.a = h.b = h.c = 0.f;
    uint4 stack[ARN_STACK4];
    int sp = 0;
    float t0;
    if (!slab(sc.root0, sc.root1, r, t0) || !(t0 < r.tmax)) return;
    uint32_t w0 = __float_as_uint(sc.root1.z), w1 = __float_as_uint(sc.root1.w);
    const uint32_t negbits = (r.inv.x < 0.f ? 1u : 0u) | (r.inv.y < 0.f ? 2u : 0u) | (r.inv.z < 0.f ? 4u : 0u);
    for (;;) {
        bool alive = true;
        while ((w1 & 3u) == ARN_W_INNER) {
            const uint32_t ax = w1 >> 2;
            const uint32_t sN = (negbits >> (ax & 3u)) & 1u, sA = (negbits >> ((ax >> 2) & 3u)) & 1u, sB = (negbits >> ((ax >> 4) & 3u)) & 1u;
            const uint32_t g1 = sN << 1, g2 = 2u - g1;                 // pair visited first / second
            const uint32_t s1 = sN ? sB : sA, s2 = sN ? sA : sB;
            const float4* __restrict__ rec = sc.wide + (size_t)w0 * 8u;
            const uint32_t o0 = (g1 + s1) * 2u, o1 = (g1 + (s1 ^ 1u)) * 2u, o2 = (g2 + s2) * 2u, o3 = (g2 + (s2 ^ 1u)) * 2u;
            const float4 a0 = __ldg(rec + o0), a1 = __ldg(rec + o0 + 1);
            const float4 b0 = __ldg(rec + o1), b1 = __ldg(rec + o1 + 1);
            const float4 c0 = __ldg(rec + o2), c1 = __ldg(rec + o2 + 1);
            const float4 d0 = __ldg(rec + o3), d1 = __ldg(rec + o3 + 1);
            float ta, tb, tc, td;
            const bool ha = slab(a0, a1, r, ta) && ta < r.tmax && (__float_as_uint(a1.w) & 3u);
            const bool hb = slab(b0, b1, r, tb) && tb < r.tmax && (__float_as_uint(b1.w) & 3u);
            const bool hc = slab(c0, c1, r, tc) && tc < r.tmax && (__float_as_uint(c1.w) & 3u);
            const bool hd = slab(d0, d1, r, td) && td < r.tmax && (__float_as_uint(d1.w) & 3u);
            // walk the visiting order backwards: every hit record but the first goes on the stack
            bool have = hd;
            uint32_t n0 = __float_as_uint(d1.z), n1 = __float_as_uint(d1.w); float nt = td;
            if (hc) { if (have) stack[sp++] = make_uint4(n0, n1, __float_as_uint(nt), 0u); n0 = __float_as_uint(c1.z); n1 = __float_as_uint(c1.w); nt = tc; have = true; }
            if (hb) { if (have) stack[sp++] = make_uint4(n0, n1, __float_as_uint(nt), 0u); n0 = __float_as_uint(b1.z); n1 = __float_as_uint(b1.w); nt = tb; have = true; }
            if (ha) { if (have) stack[sp++] = make_uint4(n0, n1, __float_as_uint(nt), 0u); n0 = __float_as_uint(a1.z); n1 = __float_as_uint(a1.w); nt = ta; have = true; }
            if (have) { w0 = n0; w1 = n1; }
            else if (!trav_pop4(r, stack, sp, w0, w1)) { alive = false; break; }
        }
        if (!alive) return;
        const uint32_t end = w0 + (w1 >> 8);
        for (uint32_t k = w0; k < end; k++) {
            float4 v0 = __ldg(&sc.tris[3 * k]);
            uint32_t comp = __float_as_uint(v0.w);
            if (comp & ARN_PRIM_SPHERE) sphere_slot(sc, comp & ~ARN_PRIM_SPHERE, r, h);
            else {
                float4 v1 = __ldg(&sc.tris[3 * k + 1]), v2 = __ldg(&sc.tris[3 * k + 2]);
                float t, b0, b1, b2;
                if (tri_test(f3(v0.x, v0.y, v0.z), f3(v1.x, v1.y, v1.z), f3(v2.x, v2.y, v2.z), r, t, b0, b1, b2) && r.tmax > t) {
                    r.tmax = t; h.prim = (int)comp; h.t = t; h.a = b0; h.b = b1; h.c = b2;
                }
            }
            if (ANY && h.prim >= 0) return;
        }
        if (!trav_pop4(r, stack, sp, w0, w1)) return;
    }
}

// What the kernels call.  The counted mode always walks the binary nodes: its counters report the
// reference algorithm's node / primitive tests.
#define ARN_TRAV_BINARY 0
#define ARN_TRAV_COUNTED 1
#define ARN_TRAV_WIDE 2
template <bool ANY, int MODE>
ARN_DEV void trace_ray(const DevScene& sc, TravRay& r, HitRec& h, uint32_t* ctr) {
    if (MODE == ARN_TRAV_WIDE) traverse4<ANY>(sc, r, h);
    else traverse<ANY, MODE == ARN_TRAV_COUNTED>(sc, r, h, ctr);
}

}  // namespace arn
