"""Device BVH build (arn_bvh_build_gpu): structural validity of the emitted reference layout, and
hit parity.  The LBVH topology is not the reference's, so the rule is:
  * on the SAME tree the GPU equals the oracle bit for bit (the tree is an input of both);
  * against the reference's SAH tree, `t` is bit-identical for every ray and the primitive id is identical except
    where two primitives are hit at exactly the same `t` (the first one met wins, bvh.rs:111)."""
import math

import numpy as np
import pytest

import oracle_lib as O
from arendur_b200 import api, scenes, _lib as L

pytestmark = pytest.mark.gpu


def _check_tree(nodes, order, bounds6):
    """Pre-order LinearNode invariants (component/bvh.rs:136-146,219-243)."""
    n = bounds6.shape[0]
    f = nodes.view(np.float32)
    assert nodes.shape[0] == 2 * n - 1
    assert sorted(order.tolist()) == list(range(n)), "every component in exactly one slot"
    seen_slots = np.zeros(n, bool)
    stack = [(0, 1)]
    max_depth, visited = 0, 0
    while stack:
        i, depth = stack.pop()
        visited += 1
        max_depth = max(max_depth, depth)
        ln = nodes[i, 7] >> 2
        if ln:
            off = nodes[i, 6]
            assert ln == 1 and not seen_slots[off]
            seen_slots[off] = True
            assert np.array_equal(f[i, :6], bounds6[order[off]]), "leaf bounds = component bounds"
        else:
            a, b = i + 1, i + nodes[i, 6]
            assert nodes[i, 6] >= 2 and b < nodes.shape[0] and (nodes[i, 7] & 3) < 3
            lo = np.minimum(f[a, :3], f[b, :3]); hi = np.maximum(f[a, 3:6], f[b, 3:6])
            assert np.array_equal(f[i, :3], lo) and np.array_equal(f[i, 3:6], hi), "interior bounds = union of the children's"
            stack.append((b, depth + 1)); stack.append((a, depth + 1))
    assert visited == nodes.shape[0] and seen_slots.all()
    return max_depth


def _bounds_of(hs):
    d = hs.desc()
    b6, _ = O.prim_bounds(d)
    return b6


def test_gpu_build_structure_and_same_tree_parity(ctx):
    """Cornell (1112 triangles + 2 transformed partial spheres): valid tree; GPU == oracle on that tree, for rays
    and for the whole bounce loop."""
    hs, cam, film, smp, prm = scenes.cornell_scene(64, 48, 2, 2)
    b6 = _bounds_of(hs)                                    # from the reference-tree build; bounds do not depend on the tree
    d, ms = hs.build_gpu(ctx)
    nodes, order = hs.nodes(), hs.order()
    depth = _check_tree(nodes, order, b6)
    assert depth <= 64 and ms > 0
    sc = ctx.upload(d); osc = O.OracleScene(d)
    rng = np.random.default_rng(3)
    n = 50_000
    rays = np.zeros(n, api.RAY_DTYPE)
    rays["o"] = rng.uniform([-1.8, -1.3, 2.2], [1.8, 2.2, 5.8], (n, 3)).astype(np.float32)
    v = rng.normal(size=(n, 3)); v /= np.linalg.norm(v, axis=1, keepdims=True)
    rays["d"] = v.astype(np.float32); rays["tmax"] = np.inf
    gh, oh = sc.intersect_closest(rays), osc.intersect_closest(rays)
    assert np.array_equal(gh["prim_id"], oh["prim_id"]) and gh["t"].tobytes() == oh["t"].tobytes()
    _, grad, st = sc.render_pt_samples(cam, film, smp, prm)
    _, orad = osc.render_pt_samples(cam, film, smp, prm)
    same = np.all(grad.view(np.uint32) == orad.view(np.uint32), axis=-1)
    assert same.all()
    sc.close(); osc.close()


def test_gpu_build_vs_reference_tree(ctx):
    """131 072-triangle height field: t bit-identical to the reference-tree result on every ray; ids equal except exact ties."""
    hs = api.HostScene()
    mat = hs.add_material(api.material(L.ARN_MAT_MATTE, kd=(0.7, 0.7, 0.7)))
    pos, idx = scenes.heightfield(256, -2.0, 2.0, 4.0, 0.15, 0x5EED)
    hs.add_mesh(pos, idx, mat)
    d_ref = hs.build()
    w, h = 640, 360
    cam = api.make_camera(api.IDENTITY, (-16.0 / 9.0, -1.0, 16.0 / 9.0, 1.0), 0.1, 1000.0, math.pi / 4, w, h)
    rays = scenes.pixel_center_rays(cam, w, h)
    rng = np.random.default_rng(9)
    inc = np.zeros(100_000, api.RAY_DTYPE)
    inc["o"] = rng.uniform([-2, -2, 1.0], [2, 2, 3.5], (inc.size, 3)).astype(np.float32)
    v = rng.normal(size=(inc.size, 3)); v[:, 2] = np.abs(v[:, 2]) * 0.5 + 0.05; v /= np.linalg.norm(v, axis=1, keepdims=True)
    inc["d"] = v.astype(np.float32); inc["tmax"] = np.inf
    rays = np.concatenate([rays, inc])
    sc = ctx.upload(d_ref)
    ref = sc.intersect_closest(rays)
    sc.close()
    b6 = _bounds_of(hs)
    d_gpu, ms = hs.build_gpu(ctx)
    assert _check_tree(hs.nodes(), hs.order(), b6) <= 64
    sc = ctx.upload(d_gpu)
    got = sc.intersect_closest(rays)
    assert (ref["prim_id"] >= 0).sum() > 100_000
    assert got["t"].tobytes() == ref["t"].tobytes()
    differ = got["prim_id"] != ref["prim_id"]
    assert differ.mean() < 1e-3                     # only exact-t ties on shared edges may pick the other triangle
    # and any-hit agrees
    assert np.array_equal(sc.intersect_any(rays) != 0, ref["prim_id"] >= 0)
    sc.close()


def test_gpu_build_degenerate_inputs(ctx):
    """One component; all components at the same place (identical Morton keys)."""
    hs = api.HostScene()
    mat = hs.add_material(api.material(L.ARN_MAT_MATTE, kd=(0.7, 0.7, 0.7)))
    tri = np.float32([[0, 0, 3], [1, 0, 3], [0, 1, 3]])
    hs.add_mesh(tri, np.uint32([0, 1, 2]), mat)
    d, _ = hs.build_gpu(ctx)
    assert d.n_nodes == 1
    sc = ctx.upload(d)
    r = np.zeros(1, api.RAY_DTYPE); r["o"] = (0.2, 0.2, 0); r["d"] = (0, 0, 1); r["tmax"] = np.inf
    one = sc.intersect_closest(r)
    assert one["prim_id"][0] == 0 and abs(one["t"][0] - 3.0) < 1e-5
    sc.close()
    hs = api.HostScene()
    mat = hs.add_material(api.material(L.ARN_MAT_MATTE, kd=(0.7, 0.7, 0.7)))
    for k in range(37):
        hs.add_mesh(tri + np.float32([0, 0, 0]), np.uint32([0, 1, 2]), mat)       # 37 coincident triangles
    d, _ = hs.build_gpu(ctx)
    b6 = np.tile(np.float32([0, 0, 3, 1, 1, 3]), (37, 1))
    _check_tree(hs.nodes(), hs.order(), b6)
    sc = ctx.upload(d)
    hit = sc.intersect_closest(r)
    assert hit["prim_id"][0] >= 0 and hit["t"][0] == one["t"][0]
    sc.close()
