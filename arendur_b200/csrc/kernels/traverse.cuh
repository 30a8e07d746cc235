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

#ifndef ARN_BLOCK
#define ARN_BLOCK 256          /* threads per block of every kernel that traces rays */
#endif

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
    float3 absmax;                           // max(|bmin|, |bmax|) of the root per axis: scale of the conservative slab test's slack
    const uint4* __restrict__ cw8;           // compressed 8-wide nodes, 8 uint4 each (kernels/cw8_build.cuh); null unless built
    const float4* __restrict__ blob;         // leaf blob of the 8-wide walk: [exact bounds, count] + primitive records per leaf
    const float4* __restrict__ pairs;        // pair records of a small tree (traverse2p), 8 float4 per interior node; null when the tree is too large
    uint32_t n_pairs, root_axis;
    const arn_texture* __restrict__ textures;   // image textures (N4); null without
    const float* __restrict__ texels;
    uint32_t n_textures;
};

#define ARN_STACK 64           /* upload rejects trees deeper than this */
// -DARN_DEBUG_STACK: a device-side check on every push (the stacks are sized from the depth the upload measures; a build with
// this flag traps instead of corrupting local memory if that reasoning were ever wrong).  Off in the product build.
#ifdef ARN_DEBUG_STACK
#define ARN_STACK_CHECK(sp, cap) do { if ((sp) >= (cap)) __trap(); } while (0)
#else
#define ARN_STACK_CHECK(sp, cap) ((void)0)
#endif
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
    float3 co, inv;            // cache: origin (read by a leaf's exact slab test only), 1/dir
    float3 o, d;               // current ray (d is read by sphere slots only)
    float tmax;
    int kz;                    // 0 = XZ perm, 1 = YZ, 2 = ZZ
    float3 shear;
};
// Per-ray constants of the CONSERVATIVE slab test used on interior nodes (cull_setup / slab_cull below).
struct CullRay {
    float3 oi0, oi1;           // -o/d -+ slack: entry / exit distances are fma(plane, 1/d, oi0 / oi1)
    uint32_t negbits;          // bit a = 1/d[a] < 0
};

ARN_DEV void shear_setup(TravRay& r, float3 d) {           // ShearingTransformCache::from_ray
    float ax = fabsf(d.x), ay = fabsf(d.y), az = fabsf(d.z);
    float3 dd;
    if (ax > ay && ax > az) { r.kz = 0; dd = f3(d.y, d.z, d.x); }
    else if (ay > az)       { r.kz = 1; dd = f3(d.z, d.x, d.y); }
    else                    { r.kz = 2; dd = d; }
    r.shear = f3(-dd.x / dd.z, -dd.y / dd.z, 1.f / dd.z);
}
ARN_DEV void trav_init(TravRay& r, float3 o, float3 d, float tmax) {
    r.o = o; r.co = o; r.d = d; r.tmax = tmax;
    r.inv = f3(1.f / d.x, 1.f / d.y, 1.f / d.z);            // construct_ray_cache (bbox.rs:583-592)
    shear_setup(r, d);
}

// Slab test without the tmax comparison: returns false on a definite miss, else t0.
// Branch-free form of BBox3f::intersect_ray_cached (bbox.rs:549-580): the reference's two early
// `return None` become predicates (the updates they skip cannot change a `false` result), the
// comparisons and selects are the reference's own (NaNs propagate identically; no fmin/fmax).
ARN_DEV bool slab(const float4 q0, const float4 q1, const TravRay& r, float& t0_out) {
    const float k = 1.f + 2.f * gamma_n(3.f);
    const bool nx = r.inv.x < 0.f, ny = r.inv.y < 0.f, nz = r.inv.z < 0.f;
    const float bminx = q0.x, bminy = q0.y, bminz = q0.z, bmaxx = q0.w, bmaxy = q1.x, bmaxz = q1.y;
    const float3 co = r.co;
    float t0 = ((nx ? bmaxx : bminx) - co.x) * r.inv.x;
    float t1 = ((nx ? bminx : bmaxx) - co.x) * r.inv.x;
    float ty0 = ((ny ? bmaxy : bminy) - co.y) * r.inv.y;
    float ty1 = ((ny ? bminy : bmaxy) - co.y) * r.inv.y;
    float tz0 = ((nz ? bmaxz : bminz) - co.z) * r.inv.z;
    float tz1 = ((nz ? bminz : bmaxz) - co.z) * r.inv.z;
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

// ---- conservative slab test for INTERIOR nodes ------------------------------------------------
// Why interior nodes need not run the reference's slab arithmetic.  For a ray whose three 1/d are finite and non-zero
// ("regular": no inf * 0 = NaN can arise) every f32 operation of BBox3f::intersect_ray_cached (bbox.rs:549-580) is
// monotone in the bounds, and a child's bounds are nested in its parent's: the child's [t0, t1] interval lies inside the
// parent's, and tmax only shrinks between the parent's pop and the child's.  So "leaf L passes its own slab test at
// its turn" implies that every ancestor passed at its turn — the set of leaves the reference processes is
//     { L : slab(L) passes with the tmax current when L's turn comes },   in depth-first order (near child by the
// sign of d[split axis], a per-ray constant).  Which interior nodes were culled on the way does not enter.  Hence any
// interior test C with  slab(N) passes => C(N) passes  visits exactly the reference's leaves in the reference's order
// (a few extra interior nodes are expanded; their leaves then fail their own exact test, which is still run, with the
// reference's arithmetic, before a leaf's primitives are touched).
//
// C(N): entry = max_a fma(near_a, 1/d_a, oi0_a), exit = min_a fma(far_a, 1/d_a, oi1_a), pass iff max(entry, 0) <= min(exit, tmax),
// with oi0/1 = -(o/d) -+ E and E_a = 2^-19 * (|o_a/d_a| + absmax_a * |1/d_a|).  Against the reference's
// t0 = fl(fl(near - o) * inv) (relative error <= 2.1 u, u = 2^-24) and t1 = fl(fl(fl(far - o) * inv) * (1 + 2 gamma_3))
// (<= real value + 10 u |.|), the fma form errs by at most u (3 |o/d| + |plane/d| + 2 E): a slack of 32 u (|o/d| + |plane/d|)
// covers both with a margin > 2.  In space the slack is ~2e-6 of the scene size: the extra interior visits are negligible.
// Rays that are not regular (a direction component 0, denormal or infinite) take the exact walk (`traverse`).
struct Node8 { float4 q0, q1; };
ARN_DEV Node8 ld_node(const float4* __restrict__ p) {      // one 256-bit load per 32-byte record (LDG.E.ENL2.256.CONSTANT)
    Node8 n;
    asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=f"(n.q0.x), "=f"(n.q0.y), "=f"(n.q0.z), "=f"(n.q0.w), "=f"(n.q1.x), "=f"(n.q1.y), "=f"(n.q1.z), "=f"(n.q1.w) : "l"(p));
    return n;
}
ARN_DEV float fmax3(float a, float b, float c) { float d; asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }   // FMNMX3
ARN_DEV float fmin3(float a, float b, float c) { float d; asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
ARN_DEV bool ray_is_regular(const DevScene& sc, const TravRay& r) {
    const float big = 1.0e15f, small = 1.0e-15f;        // products of two such magnitudes stay finite: no inf - inf, no inf * 0
    float ax = fabsf(r.inv.x), ay = fabsf(r.inv.y), az = fabsf(r.inv.z);
    return ax < big && ay < big && az < big && ax > small && ay > small && az > small
        && fabsf(r.o.x) < big && fabsf(r.o.y) < big && fabsf(r.o.z) < big          // NaNs fail every comparison (o == co before any hit)
        && sc.absmax.x < big && sc.absmax.y < big && sc.absmax.z < big;
}
ARN_DEV void cull_setup(const DevScene& sc, const TravRay& r, CullRay& c) {
    const float k = 1.9073486328125e-6f;                                           // 2^-19
    float px = r.o.x * r.inv.x, py = r.o.y * r.inv.y, pz = r.o.z * r.inv.z;       // called before any hit: o == co
    float ex = fmaxf(k * (fabsf(px) + sc.absmax.x * fabsf(r.inv.x)), 1.0e-30f);
    float ey = fmaxf(k * (fabsf(py) + sc.absmax.y * fabsf(r.inv.y)), 1.0e-30f);
    float ez = fmaxf(k * (fabsf(pz) + sc.absmax.z * fabsf(r.inv.z)), 1.0e-30f);
    c.oi0 = f3(-px - ex, -py - ey, -pz - ez);
    c.oi1 = f3(-px + ex, -py + ey, -pz + ez);
    c.negbits = (r.inv.x < 0.f ? 1u : 0u) | (r.inv.y < 0.f ? 2u : 0u) | (r.inv.z < 0.f ? 4u : 0u);
}
// true unless the box is certainly missed; lo = conservative entry distance (>= 0), re-checked against tmax at pop time
ARN_DEV bool slab_cull(const float4 q0, const float4 q1, const TravRay& r, const CullRay& c, float& lo) {
    const bool nx = c.negbits & 1u, ny = c.negbits & 2u, nz = c.negbits & 4u;
    const float l = fmax3(__fmaf_rn(nx ? q0.w : q0.x, r.inv.x, c.oi0.x), __fmaf_rn(ny ? q1.x : q0.y, r.inv.y, c.oi0.y),
                          fmaxf(__fmaf_rn(nz ? q1.y : q0.z, r.inv.z, c.oi0.z), 0.f));
    const float h = fmin3(__fmaf_rn(nx ? q0.x : q0.w, r.inv.x, c.oi1.x), __fmaf_rn(ny ? q0.y : q1.x, r.inv.y, c.oi1.y),
                          fminf(__fmaf_rn(nz ? q0.z : q1.y, r.inv.z, c.oi1.z), r.tmax));
    lo = l;
    return l <= h;
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

struct HitRec {               // what shading needs from the final hit; its distance is the ray's final tmax
    int prim;                 // component index, -1 = miss
    float a, b, c;            // triangle: b0,b1,b2; sphere: refined local hit point
};

// Component test for a sphere slot inside traversal.  Mirrors `iray = ray.clone();
// element.intersect_ray(&mut iray); if ray.tmax > iray.tmax { *ray = iray }` (bvh.rs:108-113)
// through TransformedComposable (transformed.rs:73-83): on an accepted hit the traversal ray
// becomes the round-tripped one.
ARN_DEV void sphere_slot(const DevScene& sc, uint32_t comp, TravRay& r, HitRec& h) {
    const DevSphere& sp = sc.spheres[sc.prims[comp] & ~ARN_PRIM_SPHERE];
    const float3 d = r.d;
    float3 lo = r.o, ld = d;
    if (sp.has_transform) { lo = xform_point(sp.parent_local, r.o); ld = xform_vector(sp.parent_local, d); }
    float t; float3 p;
    if (!sphere_test(sp, lo, ld, r.tmax, t, p)) return;
    if (!(r.tmax > t)) return;
    if (sp.has_transform) {
        r.o = xform_point(sp.local_parent, lo);
        const float3 nd = xform_vector(sp.local_parent, ld);
        r.d = nd;
        shear_setup(r, nd);
    }
    r.tmax = t;
    h.prim = (int)comp; h.a = p.x; h.b = p.y; h.c = p.z;
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
    h.prim = -1; h.a = h.b = h.c = 0.f;
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
                ARN_STACK_CHECK(sp, ARN_STACK);
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
                    r.tmax = t; h.prim = (int)comp; h.a = b0; h.b = b1; h.c = b2;
                }
            }
            if (ANY && h.prim >= 0) return;
        }
        if (!trav_pop(sc, r, stack, sp, idx, offset, len_axis)) return;
    }
}

// ---- fast walks: conservative culling at interior nodes, the reference's slab test at leaves -------------
// Both walks visit the leaves BVH::intersect_ray (bvh.rs:97-128) visits, in its order (see slab_cull above), so hits,
// ids and ties are bit-identical to `traverse`; tests/test_gpu_round2.py checks them ray by ray against the exact walk
// and against the oracle.  A leaf's own slab test is the reference's (`slab`), evaluated when the leaf's turn comes,
// i.e. with the reference's tmax.

// leaf primitives in slot order, strict `<` acceptance (bvh.rs:104-114); returns true when an any-hit query is done
ARN_DEV bool leaf_prims(const DevScene& sc, uint32_t first, uint32_t count, TravRay& r, HitRec& h, const bool ANY, const float4* __restrict__ recs = nullptr) {
    const float4* __restrict__ base = recs ? recs : sc.tris + 3 * (size_t)first;        // `recs`: the leaf's records in the 8-wide walk's leaf blob
    for (uint32_t k = 0; k < count; k++) {
        float4 v0 = __ldg(&base[3 * k]);
        uint32_t comp = __float_as_uint(v0.w);          // component id, sphere bit set for sphere slots
        if (comp & ARN_PRIM_SPHERE) sphere_slot(sc, comp & ~ARN_PRIM_SPHERE, r, h);
        else {
            float4 v1 = __ldg(&base[3 * k + 1]), v2 = __ldg(&base[3 * k + 2]);
            float t, b0, b1, b2;
            if (tri_test(f3(v0.x, v0.y, v0.z), f3(v1.x, v1.y, v1.z), f3(v2.x, v2.y, v2.z), r, t, b0, b1, b2) && r.tmax > t) {
                r.tmax = t; h.prim = (int)comp; h.a = b0; h.b = b1; h.c = b2;
            }
        }
        if (ANY && h.prim >= 0) return true;
    }
    return false;
}

// Binary walk over the 32-byte pre-order nodes (cache-resident trees).  Stack entry = (node, conservative entry distance).
// `any` (warp-uniform at every call site): stop at the first accepted primitive
#ifndef ARN_SPEC_LEAF
#define ARN_SPEC_LEAF 0                /* 1 (measured, slower: DESIGN.md §4): a lane that reaches a leaf postpones it and keeps walking until it holds a second one */
#endif
#if ARN_SPEC_LEAF
// Speculative while-while (Aila & Laine): the first leaf a lane reaches is POSTPONED and the lane goes on through interior nodes
// until it stands on a second leaf (or its stack is empty); the leaf phase then runs the postponed leaf, then the second one —
// the reference's order.  What the walk does between the two with a tmax the postponed leaf might have shortened is interior
// culling only, which merely has to be conservative (a larger tmax culls less); every leaf still takes the exact test with the
// tmax current at ITS turn.  The warp switches phases half as often and both phases run with more lanes.
ARN_DEV void traverse2(const DevScene& sc, TravRay& r, const CullRay& c, HitRec& h, const bool any) {
    h.prim = -1; h.a = h.b = h.c = 0.f;
    uint2 stack[ARN_STACK];
    int sp = 0;
    uint32_t idx = 0, offset, len_axis;
    {
        const Node8 n = ld_node(sc.nodes);
        float lo;
        if (!slab_cull(n.q0, n.q1, r, c, lo)) return;
        offset = __float_as_uint(n.q1.z); len_axis = __float_as_uint(n.q1.w);
    }
    const uint32_t NONE = 0xffffffffu;
    uint32_t pleaf = NONE;
    bool more = true;                                       // (idx, offset, len_axis) is a node still to be visited
    for (;;) {
        while (more) {
            if ((len_axis >> 2) != 0) {
                if (pleaf != NONE) break;                   // second leaf: leaf phase
                pleaf = idx;
                more = trav_pop(sc, r, stack, sp, idx, offset, len_axis);
                continue;
            }
            const uint32_t ia = idx + 1, ib = idx + offset;
            const Node8 a = ld_node(sc.nodes + 2 * ia), b = ld_node(sc.nodes + 2 * ib);
            float la, lb;
            const bool ha = slab_cull(a.q0, a.q1, r, c, la), hb = slab_cull(b.q0, b.q1, r, c, lb);
            const bool first_b = (c.negbits >> (len_axis & 3u)) & 1u;           // dir_is_neg[split_axis]: second child first
            if (ha && hb) {
                ARN_STACK_CHECK(sp, ARN_STACK);
                stack[sp++] = first_b ? make_uint2(ia, __float_as_uint(la)) : make_uint2(ib, __float_as_uint(lb));
                idx = first_b ? ib : ia;
                offset = __float_as_uint(first_b ? b.q1.z : a.q1.z); len_axis = __float_as_uint(first_b ? b.q1.w : a.q1.w);
            } else if (ha || hb) {
                idx = ha ? ia : ib;
                offset = __float_as_uint(ha ? a.q1.z : b.q1.z); len_axis = __float_as_uint(ha ? a.q1.w : b.q1.w);
            } else more = trav_pop(sc, r, stack, sp, idx, offset, len_axis);
        }
        // ---- leaf phase: the postponed leaf, then the one the lane stands on; each takes the reference's own slab test on its bounds
        // (they come back from L1) with the tmax current at its turn, then its primitives
#pragma unroll 1
        for (int k = 0; k < 2; k++) {
            uint32_t li;
            if (k == 0) { li = pleaf; pleaf = NONE; if (li == NONE) continue; }
            else { if (!more) return; li = idx; }
            const Node8 n = ld_node(sc.nodes + 2 * li);
            float t0;
            if (slab(n.q0, n.q1, r, t0) && t0 < r.tmax) {
                if (leaf_prims(sc, __float_as_uint(n.q1.z), __float_as_uint(n.q1.w) >> 2, r, h, any)) return;
            }
        }
        more = trav_pop(sc, r, stack, sp, idx, offset, len_axis);
        if (!more) return;
    }
}
#else
ARN_DEV void traverse2(const DevScene& sc, TravRay& r, const CullRay& c, HitRec& h, const bool any) {
    h.prim = -1; h.a = h.b = h.c = 0.f;
    uint2 stack[ARN_STACK];
    int sp = 0;
    uint32_t idx = 0, offset, len_axis;
    uint32_t nb = c.negbits;
    asm volatile("" : "+r"(nb));        // pinned: left alone, the compiler rebuilds the three sign bits from 1/d (ten instructions) on every push
    {
        const Node8 n = ld_node(sc.nodes);
        float lo;
        if (!slab_cull(n.q0, n.q1, r, c, lo)) return;
        offset = __float_as_uint(n.q1.z); len_axis = __float_as_uint(n.q1.w);
    }
    for (;;) {
        bool alive = true;
        // ---- interior nodes: cull both children (first child = idx+1, second = idx+offset) conservatively
        while ((len_axis >> 2) == 0) {
            const uint32_t ia = idx + 1, ib = idx + offset;
            const Node8 a = ld_node(sc.nodes + 2 * ia), b = ld_node(sc.nodes + 2 * ib);
            float la, lb;
            const bool ha = slab_cull(a.q0, a.q1, r, c, la), hb = slab_cull(b.q0, b.q1, r, c, lb);
            const bool first_b = ((nb >> len_axis) & 1u) != 0u;                 // dir_is_neg[split_axis]: second child first (interior: len_axis == axis)
            if (ha && hb) {
                ARN_STACK_CHECK(sp, ARN_STACK);
                stack[sp++] = first_b ? make_uint2(ia, __float_as_uint(la)) : make_uint2(ib, __float_as_uint(lb));
                idx = first_b ? ib : ia;
                offset = __float_as_uint(first_b ? b.q1.z : a.q1.z); len_axis = __float_as_uint(first_b ? b.q1.w : a.q1.w);
            } else if (ha || hb) {
                idx = ha ? ia : ib;
                offset = __float_as_uint(ha ? a.q1.z : b.q1.z); len_axis = __float_as_uint(ha ? a.q1.w : b.q1.w);
            } else if (!trav_pop(sc, r, stack, sp, idx, offset, len_axis)) { alive = false; break; }
        }
        if (!alive) return;
        // ---- leaf: the reference's own slab test (its bounds come back from L1), then the primitives
        {
            const Node8 n = ld_node(sc.nodes + 2 * idx);
            float t0;
            if (slab(n.q0, n.q1, r, t0) && t0 < r.tmax) {
                if (leaf_prims(sc, offset, len_axis >> 2, r, h, any)) return;
            }
        }
        if (!trav_pop(sc, r, stack, sp, idx, offset, len_axis)) return;
    }
}
#endif


// ---- binary walk of a SMALL tree from shared memory (ARN_TRAV_BINARY_SMEM; k_trace stages sc.pairs) -----------------------
// Cache-resident trees are bound by instruction issue in the interior loop (profiles/r02_trace_hotloop_source.txt), so the
// records the loop reads are laid out for the fewest instructions per step — and for the fewest SHARED-MEMORY instructions: a first
// layout with one 8-byte load per (child, axis) ran fewer instructions and was slower (8 LDS per step against 4).
// One PAIR RECORD per interior node, 128 bytes; A = first child (node + 1), B = second child (node + offset):
//     +0   x, direction >= 0: (A.min, A.max, B.min, B.max)     +16  x, direction < 0: (A.max, A.min, B.max, B.min)
//     +32  y, ...                                               +48  y, ...
//     +64  z, ...                                               +80  z, ...
//     +96  A.w0, A.w1, B.w0, B.w1                               +112 the same four words again
//          w1 = the child's len_axis; w0 = byte offset of the child's own pair record (interior child) or its first primitive slot (leaf child)
// A lane reads the (near, far) planes of BOTH children on one axis with ONE 16-byte load at +0 or +16 according to the sign of
// 1/d — no selects — and two packed FFMA2 (sm_100 fma.rn.f32x2) turn them into the conservative (entry, exit) distances of
// slab_cull: the same fma(plane, 1/d, oi0 / oi1) values, so the walk visits exactly what traverse2 visits.  The three per-ray
// bases (shared window + axis block + 16 for a negative direction) are all the addressing state; the reference words are stored
// twice so that they too can be read relative to the x base.  Stack entry = (record offset | child slot, entry distance); a
// leaf's exact test reads its (near, far) pairs back from its parent's record and runs the reference's arithmetic on them (slab_nf).
extern __shared__ __align__(16) float4 arn_spairs[];
#define ARN_PAIR_BYTES 128u
ARN_DEV float2 lds64(uint32_t addr) { float2 v; asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr)); return v; }
ARN_DEV float4 lds128(uint32_t addr) { float4 v; asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr)); return v; }
ARN_DEV uint4 lds128u(uint32_t addr) { uint4 v; asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr)); return v; }
ARN_DEV uint2 lds64u(uint32_t addr) { uint2 v; asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr)); return v; }
ARN_DEV unsigned long long pack2(float lo, float hi) { unsigned long long v; asm("mov.b64 %0, {%1,%2};" : "=l"(v) : "f"(lo), "f"(hi)); return v; }
ARN_DEV float2 ffma2(float2 a, unsigned long long b, unsigned long long c) {           // (a.x * b.lo + c.lo, a.y * b.hi + c.hi), each an IEEE fma.rn
    unsigned long long d; float2 v;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(pack2(a.x, a.y)), "l"(b), "l"(c));
    asm("mov.b64 {%0,%1}, %2;" : "=f"(v.x), "=f"(v.y) : "l"(d));
    return v;
}
// BBox3f::intersect_ray_cached (bbox.rs:549-580) on planes already ordered by the sign of 1/d: slab() after its selects
ARN_DEV bool slab_nf(const float2 X, const float2 Y, const float2 Z, const TravRay& r, float& t0_out) {
    const float k = 1.f + 2.f * gamma_n(3.f);
    const float3 co = r.co;
    float t0 = (X.x - co.x) * r.inv.x, t1 = (X.y - co.x) * r.inv.x;
    float ty0 = (Y.x - co.y) * r.inv.y, ty1 = (Y.y - co.y) * r.inv.y;
    float tz0 = (Z.x - co.z) * r.inv.z, tz1 = (Z.y - co.z) * r.inv.z;
    t1 *= k; ty1 *= k; tz1 *= k;
    const bool miss_xy = (t0 > ty1) | (ty0 > t1);
    t0 = ty0 > t0 ? ty0 : t0;
    t1 = ty1 < t1 ? ty1 : t1;
    const bool miss_z = (t0 > tz1) | (tz0 > t1);
    t0 = tz0 > t0 ? tz0 : t0;
    t1 = tz1 < t1 ? tz1 : t1;
    t0_out = t0;
    return !miss_xy & !miss_z & (t1 > 0.f);
}
// 4-byte stack entries: the entry distance truncated to its upper 16 bits (it is >= 0, so truncation rounds DOWN: the pop-time
// re-check stays conservative) over (record index << 1 | child slot) — a warp's stack level is one 128-byte line of local memory
static_assert(ARN_SMEM_NODE_BYTES / ARN_PAIR_BYTES <= 32768u, "a stack entry keeps the record index in 15 bits");
ARN_DEV uint32_t pair_entry(uint32_t ref, float lo) { return (__float_as_uint(lo) & 0xffff0000u) | (ref >> 6) | (ref & 1u); }      // ref = record offset (multiple of 128) | slot
ARN_DEV bool trav_pop_p(const TravRay& r, const uint32_t* stack, int& sp, uint32_t bx, uint32_t& ref, uint32_t& w0, uint32_t& w1) {
    for (;;) {
        if (sp == 0) return false;
        const uint32_t e = stack[--sp];
        if (__uint_as_float(e & 0xffff0000u) < r.tmax) {
            ref = ((e & 0xfffeu) << 6) | (e & 1u);
            const uint2 m = lds64u(bx + (ref & ~15u) + 96u + (ref & 1u) * 8u);
            w0 = m.x; w1 = m.y;
            return true;
        }
    }
}
ARN_DEV void traverse2p(const DevScene& sc, TravRay& r, const CullRay& c, HitRec& h, const bool any) {
    h.prim = -1; h.a = h.b = h.c = 0.f;
    uint32_t stack[ARN_STACK];
    int sp = 0;
    {
        float lo;
        if (!slab_cull(sc.root0, sc.root1, r, c, lo)) return;
    }
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(arn_spairs);
    uint32_t nb = c.negbits;
    uint32_t bx = sbase + ((nb & 1u) ? 16u : 0u), by = sbase + 32u + ((nb & 2u) ? 16u : 0u), bz = sbase + 64u + ((nb & 4u) ? 16u : 0u);
    asm volatile("" : "+r"(nb), "+r"(bx), "+r"(by), "+r"(bz));     // kept in registers: left alone, the compiler rebuilds them in front of every use
    const unsigned long long ix = pack2(r.inv.x, r.inv.x), iy = pack2(r.inv.y, r.inv.y), iz = pack2(r.inv.z, r.inv.z);
    const unsigned long long ox = pack2(c.oi0.x, c.oi1.x), oy = pack2(c.oi0.y, c.oi1.y), oz = pack2(c.oi0.z, c.oi1.z);
    uint32_t w0 = 0u, w1 = sc.root_axis, ref = 0u;              // the root is interior (the kernel is not chosen otherwise); its record is the first
    for (;;) {
        bool alive = true;
        while ((w1 >> 2) == 0u) {
            const uint32_t pa = w0, px = bx + pa;
            uint32_t fb;                                         // dir_is_neg[split_axis] (interior: w1 == axis), taken BEFORE the loads so that
            asm volatile("shr.u32 %0, %1, %2;" : "=r"(fb) : "r"(nb), "r"(w1));    // a spilled `nb` comes back under their latency, not after the test
            const float4 X = lds128(px), Y = lds128(by + pa), Z = lds128(bz + pa);
            const uint4 m = lds128u(px + 96u);
            const uint2 ma = make_uint2(m.x, m.y), mb = make_uint2(m.z, m.w);
            const float2 tax = ffma2(make_float2(X.x, X.y), ix, ox), tay = ffma2(make_float2(Y.x, Y.y), iy, oy), taz = ffma2(make_float2(Z.x, Z.y), iz, oz);
            const float2 tbx = ffma2(make_float2(X.z, X.w), ix, ox), tby = ffma2(make_float2(Y.z, Y.w), iy, oy), tbz = ffma2(make_float2(Z.z, Z.w), iz, oz);
            const float la = fmax3(tax.x, tay.x, fmaxf(taz.x, 0.f)), ua = fmin3(tax.y, tay.y, fminf(taz.y, r.tmax));
            const float lb = fmax3(tbx.x, tby.x, fmaxf(tbz.x, 0.f)), ub = fmin3(tbx.y, tby.y, fminf(tbz.y, r.tmax));
            const bool ha = la <= ua, hb = lb <= ub;
            const bool first_b = (fb & 1u) != 0u;                // second child first
            if (ha && hb) {
                ARN_STACK_CHECK(sp, ARN_STACK);
                stack[sp++] = pair_entry(pa | (first_b ? 0u : 1u), first_b ? la : lb);
                ref = pa | (first_b ? 1u : 0u);
                w0 = first_b ? mb.x : ma.x; w1 = first_b ? mb.y : ma.y;
            } else if (ha || hb) {
                ref = pa | (ha ? 0u : 1u);
                w0 = ha ? ma.x : mb.x; w1 = ha ? ma.y : mb.y;
            } else if (!trav_pop_p(r, stack, sp, bx, ref, w0, w1)) { alive = false; break; }
        }
        if (!alive) return;
        // ---- leaf: the reference's own slab test on its bounds, then the primitives
        {
            const uint32_t lp = (ref & ~15u) + (ref & 1u) * 8u;
            const float2 X = lds64(bx + lp), Y = lds64(by + lp), Z = lds64(bz + lp);
            float t0;
            if (slab_nf(X, Y, Z, r, t0) && t0 < r.tmax) {
                if (leaf_prims(sc, w0, w1 >> 2, r, h, any)) return;
            }
        }
        if (!trav_pop_p(r, stack, sp, bx, ref, w0, w1)) return;
    }
}

// ---- 4-wide walk (trees that do not fit the caches) ------------------------------------------------------------
// The wide tree is the binary tree with every second level removed: wide node = (children of the first child, children
// of the second child), a leaf child occupying one slot of its pair; 4 records of 32 B (bounds, w0, w1) = one 128-B
// line.  The four records are visited in the order the binary depth-first walk would reach them (pair order from the
// node's split axis, order inside a pair from the child's split axis — signs of the ray direction, not distances).
// Stack entry = (record index, conservative entry distance): 8 bytes, so a warp's stack level is two 128-B lines of
// local memory instead of four; the record is fetched again when it is popped (bounds for a leaf's exact test, the
// reference words for an interior record).  Empty slots carry inverted infinite bounds and fail every test.
#define ARN_REC_LEAF 0x80000000u       /* stack entry: the record is a leaf (its reference words need no fetch before the leaf phase) */
#ifndef ARN_WIDE_STACK16
#define ARN_WIDE_STACK16 1             /* 1: 16-byte entries that also carry the record's reference words (no dependent fetch when an interior record is popped):
                                          C4 k_trace 136.7 -> 134.4 ms; 0: 8-byte entries (record index, entry distance) */
#endif
#if ARN_WIDE_STACK16
typedef uint4 WideEntry;
ARN_DEV WideEntry wide_entry(uint32_t rec, float lo, uint32_t w0, uint32_t w1) { return make_uint4(rec, __float_as_uint(lo), w0, w1); }
#else
typedef uint2 WideEntry;
ARN_DEV WideEntry wide_entry(uint32_t rec, float lo, uint32_t, uint32_t) { return make_uint2(rec, __float_as_uint(lo)); }
#endif
ARN_DEV bool trav_pop4(const DevScene& sc, const TravRay& r, const WideEntry* stack, int& sp, uint32_t& rec, uint32_t& w0, uint32_t& w1) {
    for (;;) {
        if (sp == 0) return false;
        const WideEntry e = stack[--sp];
        if (__uint_as_float(e.y) < r.tmax) {
            rec = e.x;
#if ARN_WIDE_STACK16
            w0 = e.z; w1 = e.w;
#else
            if (!(rec & ARN_REC_LEAF)) { const float4 q1 = __ldg(sc.wide + 2 * (size_t)rec + 1); w0 = __float_as_uint(q1.z); w1 = __float_as_uint(q1.w); }
#endif
            return true;
        }
    }
}
#ifndef ARN_WIDE_SPEC_PREFETCH
#define ARN_WIDE_SPEC_PREFETCH 0
#endif
ARN_DEV void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" :: "l"(p)); }
ARN_DEV void traverse4(const DevScene& sc, TravRay& r, const CullRay& c, HitRec& h, const bool any) {
    h.prim = -1; h.a = h.b = h.c = 0.f;
    WideEntry stack[ARN_STACK4];
    int sp = 0;
    {
        float lo;
        if (!slab_cull(sc.root0, sc.root1, r, c, lo)) return;
    }
    // current record: its index (leaf bit set for leaves) and, for an interior record, its reference words
    uint32_t w0 = __float_as_uint(sc.root1.z), w1 = __float_as_uint(sc.root1.w), rec = 0;
    if ((w1 & 3u) != ARN_W_INNER) {             // a one-leaf tree: the root record is the leaf
        float t0;
        if (slab(sc.root0, sc.root1, r, t0) && t0 < r.tmax) leaf_prims(sc, w0, w1 >> 8, r, h, any);
        return;
    }
    for (;;) {
        bool alive = true;
        while (!(rec & ARN_REC_LEAF)) {
            const uint32_t ax = w1 >> 2;
            const uint32_t sN = (c.negbits >> (ax & 3u)) & 1u, sA = (c.negbits >> ((ax >> 2) & 3u)) & 1u, sB = (c.negbits >> ((ax >> 4) & 3u)) & 1u;
            const uint32_t g1 = sN << 1, g2 = 2u - g1;                 // pair visited first / second
            const uint32_t s1 = sN ? sB : sA, s2 = sN ? sA : sB;
            const uint32_t base = w0 * 4u;
            const uint32_t r0 = base + g1 + s1, r1 = base + g1 + (s1 ^ 1u), r2 = base + g2 + s2, r3 = base + g2 + (s2 ^ 1u);
            const Node8 na = ld_node(sc.wide + 2 * (size_t)r0), nb = ld_node(sc.wide + 2 * (size_t)r1);
            const Node8 nc = ld_node(sc.wide + 2 * (size_t)r2), nd = ld_node(sc.wide + 2 * (size_t)r3);
#if ARN_WIDE_SPEC_PREFETCH
            // speculative: ask L2 for the wide nodes of all interior children before their boxes are tested, so that the fetch of the
            // child visited next overlaps the tests instead of following them
            if ((__float_as_uint(na.q1.w) & 3u) == ARN_W_INNER) prefetch_l2(sc.wide + 8 * (size_t)__float_as_uint(na.q1.z));
            if ((__float_as_uint(nb.q1.w) & 3u) == ARN_W_INNER) prefetch_l2(sc.wide + 8 * (size_t)__float_as_uint(nb.q1.z));
            if ((__float_as_uint(nc.q1.w) & 3u) == ARN_W_INNER) prefetch_l2(sc.wide + 8 * (size_t)__float_as_uint(nc.q1.z));
            if ((__float_as_uint(nd.q1.w) & 3u) == ARN_W_INNER) prefetch_l2(sc.wide + 8 * (size_t)__float_as_uint(nd.q1.z));
#endif
            float ta, tb, tc, td;
            const bool ha = slab_cull(na.q0, na.q1, r, c, ta), hb = slab_cull(nb.q0, nb.q1, r, c, tb);
            const bool hc = slab_cull(nc.q0, nc.q1, r, c, tc), hd = slab_cull(nd.q0, nd.q1, r, c, td);
            const uint32_t ka = __float_as_uint(na.q1.w), kb = __float_as_uint(nb.q1.w), kc = __float_as_uint(nc.q1.w), kd = __float_as_uint(nd.q1.w);
            const uint32_t ea = r0 | ((ka & 3u) == ARN_W_LEAF ? ARN_REC_LEAF : 0u), eb = r1 | ((kb & 3u) == ARN_W_LEAF ? ARN_REC_LEAF : 0u);
            const uint32_t ec = r2 | ((kc & 3u) == ARN_W_LEAF ? ARN_REC_LEAF : 0u), ed = r3 | ((kd & 3u) == ARN_W_LEAF ? ARN_REC_LEAF : 0u);
            // every surviving record but the first in visiting order goes on the stack, last one first (branch-free)
            const bool have = ha | hb | hc | hd;
            ARN_STACK_CHECK(sp + 2, ARN_STACK4);
            if (hd & (ha | hb | hc)) stack[sp++] = wide_entry(ed, td, __float_as_uint(nd.q1.z), kd);
            if (hc & (ha | hb)) stack[sp++] = wide_entry(ec, tc, __float_as_uint(nc.q1.z), kc);
            if (hb & ha) stack[sp++] = wide_entry(eb, tb, __float_as_uint(nb.q1.z), kb);
            const uint32_t nr = ha ? ea : (hb ? eb : (hc ? ec : ed));
            const uint32_t n0 = __float_as_uint(ha ? na.q1.z : (hb ? nb.q1.z : (hc ? nc.q1.z : nd.q1.z)));
            const uint32_t n1 = ha ? ka : (hb ? kb : (hc ? kc : kd));
            if (have) { rec = nr; w0 = n0; w1 = n1; }
            else if (!trav_pop4(sc, r, stack, sp, rec, w0, w1)) { alive = false; break; }
        }
        if (!alive) return;
        // ---- leaf record: the reference's own slab test on its bounds (fetched again: one line, usually still in L1), then the primitives
        {
            const Node8 n = ld_node(sc.wide + 2 * (size_t)(rec & ~ARN_REC_LEAF));
            float t0;
            if (slab(n.q0, n.q1, r, t0) && t0 < r.tmax) {
                if (leaf_prims(sc, __float_as_uint(n.q1.z), __float_as_uint(n.q1.w) >> 8, r, h, any)) return;
            }
        }
        if (!trav_pop4(sc, r, stack, sp, rec, w0, w1)) return;
    }
}

// ---- compressed 8-wide walk (trees far larger than the caches) --------------------------------------------------
// Node layout and the argument for quantised, outward-rounded child boxes: kernels/cw8_build.cuh.  A node's eight child boxes
// are tested with t = q * (2^e / d) + ((origin - o) / d -+ slack): one byte-permute, one FADD and one FFMA per plane.  The
// children that survive are kept as ONE stack entry per node — (node | leaf mask, pending list in the reference's visiting
// order for this ray's sign octant) — and the next child's reference is read from the node when its turn comes.  Nothing
// but leaves decides a hit: a leaf's exact bounds sit in front of its primitives in the leaf blob and take the reference's
// slab test with the tmax current at its turn.
#define ARN_STACK8 32
ARN_DEV float cw8_qf(uint32_t word, uint32_t sel) { return __uint_as_float(__byte_perm(word, 0x4B000000u, sel)) - 8388608.f; }   // byte -> float, exact
struct Cw8Axis { float a, b0, b1; uint32_t n_lo, n_hi, f_lo, f_hi; };      // scale / d, near / far offsets, near / far plane bytes of slots 0-3 / 4-7
ARN_DEV void cw8_expand(const DevScene& sc, const TravRay& r, const CullRay& c, uint32_t w, uint32_t perm_off, uint32_t& gnode, uint32_t& gpend) {
    const uint4* __restrict__ nd = sc.cw8 + 8 * (size_t)w;
    const uint4 q0 = __ldg(nd), p0 = __ldg(nd + 2), p1 = __ldg(nd + 3), p2 = __ldg(nd + 4);
    const uint32_t order = __ldg(reinterpret_cast<const uint32_t*>(nd) + perm_off);
    const uint32_t meta = q0.w;
    Cw8Axis X, Y, Z;
    {
        const float ox = __uint_as_float(q0.x), oy = __uint_as_float(q0.y), oz = __uint_as_float(q0.z);
        X.a = __uint_as_float((meta & 0xffu) << 23) * r.inv.x; Y.a = __uint_as_float(((meta >> 8) & 0xffu) << 23) * r.inv.y; Z.a = __uint_as_float(((meta >> 16) & 0xffu) << 23) * r.inv.z;
        X.b0 = __fmaf_rn(ox, r.inv.x, c.oi0.x); X.b1 = __fmaf_rn(ox, r.inv.x, c.oi1.x);
        Y.b0 = __fmaf_rn(oy, r.inv.y, c.oi0.y); Y.b1 = __fmaf_rn(oy, r.inv.y, c.oi1.y);
        Z.b0 = __fmaf_rn(oz, r.inv.z, c.oi0.z); Z.b1 = __fmaf_rn(oz, r.inv.z, c.oi1.z);
        const bool nx = c.negbits & 1u, ny = c.negbits & 2u, nz = c.negbits & 4u;
        // planes: p0 = (lo.x[0-3], lo.x[4-7], lo.y[0-3], lo.y[4-7]), p1 = (lo.z, lo.z, hi.x, hi.x), p2 = (hi.y, hi.y, hi.z, hi.z)
        X.n_lo = nx ? p1.z : p0.x; X.n_hi = nx ? p1.w : p0.y; X.f_lo = nx ? p0.x : p1.z; X.f_hi = nx ? p0.y : p1.w;
        Y.n_lo = ny ? p2.x : p0.z; Y.n_hi = ny ? p2.y : p0.w; Y.f_lo = ny ? p0.z : p2.x; Y.f_hi = ny ? p0.w : p2.y;
        Z.n_lo = nz ? p2.z : p1.x; Z.n_hi = nz ? p2.w : p1.y; Z.f_lo = nz ? p1.x : p2.z; Z.f_hi = nz ? p1.y : p2.w;
    }
    uint32_t hitmask = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const uint32_t sel = 0x7440u | (uint32_t)(k & 3);
        const float tnx = __fmaf_rn(cw8_qf(k < 4 ? X.n_lo : X.n_hi, sel), X.a, X.b0), tfx = __fmaf_rn(cw8_qf(k < 4 ? X.f_lo : X.f_hi, sel), X.a, X.b1);
        const float tny = __fmaf_rn(cw8_qf(k < 4 ? Y.n_lo : Y.n_hi, sel), Y.a, Y.b0), tfy = __fmaf_rn(cw8_qf(k < 4 ? Y.f_lo : Y.f_hi, sel), Y.a, Y.b1);
        const float tnz = __fmaf_rn(cw8_qf(k < 4 ? Z.n_lo : Z.n_hi, sel), Z.a, Z.b0), tfz = __fmaf_rn(cw8_qf(k < 4 ? Z.f_lo : Z.f_hi, sel), Z.a, Z.b1);
        const float lo = fmax3(tnx, tny, fmaxf(tnz, 0.f)), hi = fmin3(tfx, tfy, fminf(tfz, r.tmax));
        hitmask |= (lo <= hi ? 1u : 0u) << k;
    }
    // pending list: the node's visiting order for this sign octant, a valid bit on the slots that survived
    uint32_t pend = order;
#pragma unroll
    for (int q = 0; q < 8; q++) pend |= ((hitmask >> ((order >> (4 * q)) & 7u)) & 1u) << (4 * q + 3);
    gnode = w | (meta & 0xff000000u); gpend = pend;
}
ARN_DEV void traverse8(const DevScene& sc, TravRay& r, const CullRay& c, HitRec& h, const bool any) {
    h.prim = -1; h.a = h.b = h.c = 0.f;
    uint2 stack[ARN_STACK8];
    int sp = 0;
    const uint32_t perm_off = 4u + c.negbits + ((c.negbits & 4u) ? 12u : 0u);      // word offset of this octant's visiting order inside a node
    uint32_t gnode, gpend;
    cw8_expand(sc, r, c, 0u, perm_off, gnode, gpend);
    for (;;) {
        // ---- interior: take children in order, expand interior ones, until a leaf's turn comes
        uint32_t leaf_ref = 0; bool have_leaf = false;
        while (!have_leaf) {
            const uint32_t valid = gpend & 0x88888888u;
            if (!valid) {
                if (sp == 0) return;
                const uint2 e = stack[--sp]; gnode = e.x; gpend = e.y;
                continue;
            }
            const uint32_t k4 = (uint32_t)__ffs((int)valid) - 4u;                    // bit position of the first pending nibble
            const uint32_t slot = (gpend >> k4) & 7u;
            gpend &= ~(8u << k4);
            const uint32_t ref = __ldg(reinterpret_cast<const uint32_t*>(sc.cw8 + 8 * (size_t)(gnode & 0xffffffu)) + 24u + slot);
            if ((gnode >> (24u + slot)) & 1u) { leaf_ref = ref; have_leaf = true; }
            else {
                ARN_STACK_CHECK(sp, ARN_STACK8);
                if (gpend & 0x88888888u) stack[sp++] = make_uint2(gnode, gpend);
                cw8_expand(sc, r, c, ref, perm_off, gnode, gpend);
            }
        }
        // ---- leaf: the reference's slab test on its exact bounds, then its primitives
        {
            const float4 b0 = __ldg(sc.blob + leaf_ref), b1 = __ldg(sc.blob + leaf_ref + 1);      // 16-byte aligned only: leaves are 2 + 3 * count float4 long
            float t0;
            if (slab(b0, b1, r, t0) && t0 < r.tmax) {
                if (leaf_prims(sc, 0u, __float_as_uint(b1.z), r, h, any, sc.blob + leaf_ref + 2)) return;
            }
        }
    }
}

// What the kernels call.  The counted mode always walks the binary nodes: its counters report the
// reference algorithm's node / primitive tests.
#define ARN_TRAV_BINARY 0
#define ARN_TRAV_COUNTED 1
#define ARN_TRAV_WIDE 2
#define ARN_TRAV_CW8 3
#define ARN_TRAV_BINARY_SMEM 4        /* binary walk over pair records staged in shared memory by the calling kernel (k_trace only) */
// out-of-line exact walks for the rare rays the conservative test does not cover (one copy per kernel)
ARN_NOINL void traverse_exact_closest(const DevScene& sc, TravRay& r, HitRec& h) { traverse<false, false>(sc, r, h, nullptr); }
ARN_NOINL void traverse_exact_any(const DevScene& sc, TravRay& r, HitRec& h) { traverse<true, false>(sc, r, h, nullptr); }
template <int MODE>
ARN_DEV void trace_ray(const DevScene& sc, TravRay& r, HitRec& h, uint32_t* ctr, const bool any) {
    if (MODE == ARN_TRAV_COUNTED) {                 // the reference walk, every node with the reference's arithmetic; counts its tests
        if (any) traverse<true, true>(sc, r, h, ctr); else traverse<false, true>(sc, r, h, ctr);
        return;
    }
    if (!ray_is_regular(sc, r)) {                   // axis-parallel / degenerate directions: the reference's arithmetic at every node
        if (any) traverse_exact_any(sc, r, h); else traverse_exact_closest(sc, r, h);
        return;
    }
    if (MODE == ARN_TRAV_WIDE || MODE == ARN_TRAV_CW8) {                    // large scenes are often seen from outside: rays that miss the root (the reference's own
        float t0;                                   // first test, bvh.rs:100-102) leave before the conservative-test constants are computed
        if (!slab(sc.root0, sc.root1, r, t0) || !(t0 < r.tmax)) { h.prim = -1; h.a = h.b = h.c = 0.f; return; }
    }
    CullRay c; cull_setup(sc, r, c);
    if (MODE == ARN_TRAV_CW8) traverse8(sc, r, c, h, any);
    else if (MODE == ARN_TRAV_WIDE) traverse4(sc, r, c, h, any);
    else if (MODE == ARN_TRAV_BINARY_SMEM) traverse2p(sc, r, c, h, any);
    else traverse2(sc, r, c, h, any);
}

}  // namespace arn
