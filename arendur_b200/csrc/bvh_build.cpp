// Host BVH builder of the B200 path-tracing core: arn_bvh_build (include/arn.h).
//
// Drop-in for arendur's `BVH::new` (src/component/bvh.rs:58-79): same top-down splits
// (recursive_build :246-316, sah_midpoint :377-415, sort_mid :417-442, handle_tails :445-465)
// and the same pre-order layout as `BuildNode::flatten` (:219-243), so the tree, the
// primitive order and hence every traversal tie-break equal the reference's.  The SAH
// bucket-scan quirks are reproduced on purpose.
//
// Design (differs from the reference's arena + flatten): nodes are emitted straight into
// the final 32-byte pre-order layout.  Interior nodes store a RELATIVE second-child offset
// and leaves an ABSOLUTE slot in the ordered primitive list, so a subtree's node array is
// position independent: large right subtrees are built concurrently into their own vectors
// and spliced in afterwards (20 M triangles: seconds instead of a minute).
#include <cstdint>
#include <cstring>
#include <future>
#include <thread>
#include <vector>
#include <atomic>
#include "../../include/arn.h"

namespace {

struct Box { float lo[3], hi[3]; };
struct Item { Box b; float c[3]; float cost; uint32_t idx; };   // ComponentInfo (bvh.rs:15-21)

inline float fmin2(float a, float b) { return a < b ? a : b; }  // cgmath partial_min
inline float fmax2(float a, float b) { return a > b ? a : b; }  // cgmath partial_max
inline void box_union(Box& a, const Box& b) {
    for (int k = 0; k < 3; k++) { a.lo[k] = fmin2(a.lo[k], b.lo[k]); a.hi[k] = fmax2(a.hi[k], b.hi[k]); }
}
inline float box_area(const Box& b) {                           // BBox3::surface_area (bbox.rs:399-411)
    float dx = b.hi[0] - b.lo[0], dy = b.hi[1] - b.lo[1], dz = b.hi[2] - b.lo[2];
    if (dx < 0.f) dx = 0.f;
    if (dy < 0.f) dy = 0.f;
    if (dz < 0.f) dz = 0.f;
    return 2.f * (dx * dy + dx * dz + dy * dz);
}

struct Bin { uint32_t count; float cost; Box b; bool init; };
inline Bin bin_merge(const Bin& a, const Bin& b) {              // Bucket::union (bvh.rs:360-374)
    if (!a.init) return b;
    if (!b.init) return a;
    Bin r; r.count = a.count + b.count; r.cost = a.cost + b.cost; r.b = a.b; box_union(r.b, b.b); r.init = true;
    return r;
}

const size_t kParallelCutoff = 1u << 16;

struct Builder {
    int strategy;
    std::atomic<int> tasks_left{0};      // helper threads this build may still start (per build: concurrent builds do not share it)

    void emit_leaf(std::vector<arn_node>& out, const Box& b, uint32_t offset, uint32_t len) {
        arn_node n;
        std::memcpy(n.bmin, b.lo, 12); std::memcpy(n.bmax, b.hi, 12);
        n.offset = offset; n.len_axis = (len << 2) | 3u;
        out.push_back(n);
    }

    // sah_midpoint (bvh.rs:377-415), returns the split coordinate on `axis`
    static float sah_split(const Item* it, size_t n, int axis, const Box& cb, float inv_area) {
        const int NB = 32;
        Bin bins[NB];
        for (int i = 0; i < NB; i++) { bins[i].count = 0; bins[i].cost = 0.f; bins[i].init = false; }
        const float diag = cb.hi[axis] - cb.lo[axis];
        for (size_t k = 0; k < n; k++) {
            float dif = it[k].c[axis] - cb.lo[axis];
            size_t bi = (size_t)(dif / diag * (float)NB);
            if (bi == (size_t)NB) bi -= 1;
            Bin& b = bins[bi];
            if (!b.init) { b.count = 1; b.cost = it[k].cost; b.b = it[k].b; b.init = true; }
            else { b.count += 1; b.cost += it[k].cost; box_union(b.b, it[k].b); }
        }
        Bin fwd[NB], rev[NB];
        for (int i = 0; i < NB; i++) { fwd[i].count = 0; fwd[i].cost = 0.f; fwd[i].init = false; rev[i] = fwd[i]; }
        fwd[0] = bins[0];
        rev[NB - 1] = bins[NB - 1];
        for (int i = 1; i < NB - 1; i++) {
            fwd[i] = bin_merge(fwd[i - 1], bins[i]);
            rev[NB - 1 - i] = bin_merge(rev[NB - i], bins[NB - i]);     // sic: bucket 31 counted twice, bucket i+1 skipped
        }
        rev[0] = bin_merge(rev[1], bins[0]);
        int best = NB - 1;
        float min_cost = rev[0].cost;                                   // sic: raw cost sum as the initial bound
        for (int i = 0; i < NB - 1; i++) {
            float cost = 0.125f + (fwd[i].cost * box_area(fwd[i].b) + rev[i + 1].cost * box_area(rev[i + 1].b)) * inv_area;
            if (cost < min_cost) { best = i; min_cost = cost; }
        }
        // cb.pmin + cb.diagonal() * ((best+1)/32), component `axis`
        return cb.lo[axis] + diag * ((float)(best + 1) / (float)NB);
    }

    // Builds the subtree over it[0..n) (ordered slots offset..offset+n) and appends it to `out`.
    // `scratch` is the matching window of the `ordered` array of the reference.
    void build(Item* it, size_t n, uint32_t offset, Item* scratch, int strat, std::vector<arn_node>& out) {
        if (n == 1) { emit_leaf(out, it[0].b, offset, 1); return; }
        Box b = it[0].b, cb;
        for (int k = 0; k < 3; k++) { cb.lo[k] = it[0].c[k]; cb.hi[k] = it[0].c[k]; }
        for (size_t k = 1; k < n; k++) {
            box_union(b, it[k].b);
            for (int a = 0; a < 3; a++) { cb.lo[a] = fmin2(cb.lo[a], it[k].c[a]); cb.hi[a] = fmax2(cb.hi[a], it[k].c[a]); }
        }
        float dx = cb.hi[0] - cb.lo[0], dy = cb.hi[1] - cb.lo[1], dz = cb.hi[2] - cb.lo[2];
        int axis = (dx > dy && dx > dz) ? 0 : (dy > dz ? 1 : 2);        // BBox3::max_extent (bbox.rs:432-441)
        if (cb.lo[axis] == cb.hi[axis]) { emit_leaf(out, b, offset, (uint32_t)n); return; }
        size_t mid_count;
        if (strat == ARN_BVH_SAH && n <= 4) strat = ARN_BVH_MIDPOINT;   // bvh.rs:279-283 (whole subtree)
        if (strat == ARN_BVH_MIDDLECOUNT) mid_count = n >> 1;           // no sorting (bvh.rs:295-300)
        else {
            float mid = (strat == ARN_BVH_SAH) ? sah_split(it, n, axis, cb, 1.f / box_area(b))
                                               : (cb.hi[axis] + cb.lo[axis]) / 2.f;
            // sort_mid: "< mid" fill from the front in order, the rest from the back, backwards
            size_t i = 0, j = n;
            for (size_t k = 0; k < n; k++) {
                if (it[k].c[axis] < mid) scratch[i++] = it[k]; else scratch[--j] = it[k];
            }
            std::memcpy(it, scratch, n * sizeof(Item));
            mid_count = i;
        }
        if (mid_count == 0 || mid_count == n) { emit_leaf(out, b, offset, (uint32_t)n); return; }

        size_t self = out.size();
        arn_node placeholder; std::memset(&placeholder, 0, sizeof placeholder);
        out.push_back(placeholder);
        bool spawn = n >= kParallelCutoff && tasks_left.fetch_sub(1) > 0;
        if (spawn) {
            std::vector<arn_node> right;
            auto fut = std::async(std::launch::async, [&]() {
                build(it + mid_count, n - mid_count, offset + (uint32_t)mid_count, scratch + mid_count, strat, right);
            });
            build(it, mid_count, offset, scratch, strat, out);
            fut.get();
            tasks_left.fetch_add(1);
            size_t second = out.size();
            out.insert(out.end(), right.begin(), right.end());
            finish_interior(out, self, second, axis);
        } else {
            if (n >= kParallelCutoff) tasks_left.fetch_add(1);
            build(it, mid_count, offset, scratch, strat, out);
            size_t second = out.size();
            build(it + mid_count, n - mid_count, offset + (uint32_t)mid_count, scratch + mid_count, strat, out);
            finish_interior(out, self, second, axis);
        }
    }
    // BuildNode::to_interior (bvh.rs:210-219): bound = union of the children's bounds
    static void finish_interior(std::vector<arn_node>& out, size_t self, size_t second, int axis) {
        arn_node& n = out[self];
        const arn_node& a = out[self + 1]; const arn_node& c = out[second];
        for (int k = 0; k < 3; k++) { n.bmin[k] = fmin2(a.bmin[k], c.bmin[k]); n.bmax[k] = fmax2(a.bmax[k], c.bmax[k]); }
        n.offset = (uint32_t)(second - self);
        n.len_axis = (uint32_t)axis;
    }
};

}  // namespace

extern "C" int arn_bvh_build(uint32_t n, const float* bounds6, const float* costs, int strategy,
                             arn_node* nodes_out, uint32_t* order_out, uint32_t* n_nodes_out) {
    if (n == 0 || !bounds6 || !costs || !nodes_out || !order_out || !n_nodes_out) return ARN_E_INVALID;
    if (strategy != ARN_BVH_SAH && strategy != ARN_BVH_MIDDLECOUNT && strategy != ARN_BVH_MIDPOINT) return ARN_E_INVALID;
    std::vector<Item> items, scratch;
    try { items.resize(n); scratch.resize(n); } catch (...) { return ARN_E_OOM; }
    for (uint32_t i = 0; i < n; i++) {                              // ComponentInfo::new (bvh.rs:24-35)
        Item& it = items[i];
        for (int k = 0; k < 3; k++) {
            it.b.lo[k] = bounds6[6 * (size_t)i + k]; it.b.hi[k] = bounds6[6 * (size_t)i + 3 + k];
            it.c[k] = (it.b.lo[k] + it.b.hi[k]) / 2.0f;
        }
        it.cost = costs[i]; it.idx = i;
    }
    unsigned hw = std::thread::hardware_concurrency();
    std::vector<arn_node> nodes;
    try {
        nodes.reserve(2 * (size_t)n);
        Builder b; b.strategy = strategy; b.tasks_left.store(hw > 1 ? (int)hw - 1 : 0);
        b.build(items.data(), n, 0, scratch.data(), strategy, nodes);
    } catch (...) { return ARN_E_OOM; }
    std::memcpy(nodes_out, nodes.data(), nodes.size() * sizeof(arn_node));
    for (uint32_t i = 0; i < n; i++) order_out[i] = items[i].idx;
    *n_nodes_out = (uint32_t)nodes.size();
    return ARN_OK;
}

// Distribution1D::new (src/sample/distribution.rs:25-63)
extern "C" int arn_light_distribution(uint32_t n, const float* func, float* cdf_out, float* integral_out) {
    if ((n && !func) || !cdf_out || !integral_out) return ARN_E_INVALID;
    for (uint32_t i = 0; i < n; i++) if (!(func[i] >= 0.f)) return ARN_E_INVALID;   // assert!(curfunc >= 0)
    cdf_out[0] = 0.f;
    for (uint32_t i = 0; i < n; i++) cdf_out[i + 1] = cdf_out[i] + func[i];
    float total = cdf_out[n];
    if (total == 0.f) { for (uint32_t i = 1; i <= n; i++) cdf_out[i] = (float)i / (float)(n + 1); }
    else { for (uint32_t i = 1; i <= n; i++) cdf_out[i] /= total; }
    *integral_out = total;
    return ARN_OK;
}
