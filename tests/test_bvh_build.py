"""arn_bvh_build (product host builder) vs the oracle's restatement of BVH::new
(src/component/bvh.rs:58-79,246-465): node arrays and primitive order must be bit-identical,
including the SAH bucket-scan quirks."""
import numpy as np
import pytest

import oracle_lib as O
from arendur_b200 import api, scenes, _lib as L


def _tri_bounds(pos, idx):
    p = pos[idx]                                  # (T, 3, 3)
    return np.concatenate([p.min(axis=1), p.max(axis=1)], axis=1).astype(np.float32)


def _same(b, c, strategy=L.ARN_BVH_SAH):
    pn, po = api.bvh_build(b, c, strategy)
    on, oo = O.bvh_build(b, c, strategy)
    assert pn.shape == on.shape
    assert np.array_equal(pn, on), f"{(pn != on).any(axis=1).sum()} nodes differ"
    assert np.array_equal(po, oo)
    return pn, po


def _check_tree(nodes, order, n):
    """Structural invariants: pre-order layout, every primitive in exactly one leaf."""
    assert sorted(order.tolist()) == list(range(n))
    seen = np.zeros(n, bool)
    stack = [0]
    visited = 0
    while stack:
        i = stack.pop(); visited += 1
        off, la = int(nodes[i, 6]), int(nodes[i, 7])
        if la >> 2 == 0:
            assert (la & 3) < 3 and off >= 2
            stack += [i + off, i + 1]
        else:
            assert (la & 3) == 3
            assert not seen[off:off + (la >> 2)].any()
            seen[off:off + (la >> 2)] = True
    assert visited == nodes.shape[0] and seen.all()


@pytest.mark.parametrize("strategy", [L.ARN_BVH_SAH, L.ARN_BVH_MIDPOINT, L.ARN_BVH_MIDDLECOUNT])
def test_random_boxes(strategy):
    rng = np.random.default_rng(3)
    for n in (1, 2, 3, 4, 5, 7, 33, 257, 5000):
        lo = rng.uniform(-10, 10, (n, 3)).astype(np.float32)
        hi = lo + rng.uniform(0, 2, (n, 3)).astype(np.float32)
        b = np.concatenate([lo, hi], axis=1)
        c = rng.choice([3.0, 2.0, 1.0], n).astype(np.float32)
        nodes, order = _same(b, c, strategy)
        _check_tree(nodes, order, n)


def test_heightfield_and_cornell():
    pos, idx = scenes.heightfield(96, -2.0, 2.0, 4.0, 0.15, 0x5EED)
    b = _tri_bounds(pos, idx)
    nodes, order = _same(b, np.full(b.shape[0], 3.0, np.float32))
    _check_tree(nodes, order, b.shape[0])
    hs, *_ = scenes.cornell_scene(64, 48, 1, 1)
    d = hs.desc()
    ob, oc = O.prim_bounds(d)                     # oracle's ComponentInfo::new on the same components
    on, oo = O.bvh_build(ob, oc)
    assert np.array_equal(on, hs.nodes()) and np.array_equal(oo, hs.order())
    assert set(oc.tolist()) == {3.0, 2.0}         # triangles 3.0, transformed spheres 1 + 1


def test_degenerate_inputs():
    # identical centroids -> one multi-primitive leaf (bvh.rs:273-275)
    b = np.tile(np.float32([0, 0, 0, 1, 1, 1]), (9, 1))
    nodes, order = _same(b, np.full(9, 3.0, np.float32))
    assert nodes.shape[0] == 1 and int(nodes[0, 7]) >> 2 == 9
    # one-sided partition -> leaf with reversed order (sort_mid fills the right side backwards)
    b = np.zeros((6, 6), np.float32)
    b[:, 0] = [0, 0, 0, 0, 0, 1e-30]; b[:, 3] = b[:, 0]
    nodes, order = _same(b, np.full(6, 3.0, np.float32))
    _check_tree(nodes, order, 6)
    # large coordinates, zero-area boxes, many duplicates
    rng = np.random.default_rng(5)
    p = rng.integers(-3, 3, (4000, 3)).astype(np.float32) * 1e6
    nodes, order = _same(np.concatenate([p, p], axis=1), np.full(4000, 3.0, np.float32))
    _check_tree(nodes, order, 4000)


def test_big_parallel_build_matches_oracle():
    """> 65 536 primitives: the product builder builds subtrees concurrently; result unchanged."""
    pos, idx = scenes.heightfield(260, -2.0, 2.0, 4.0, 0.15, 7)       # 135 200 triangles
    b = _tri_bounds(pos, idx)
    nodes, order = _same(b, np.full(b.shape[0], 3.0, np.float32))
    _check_tree(nodes, order, b.shape[0])


def test_invalid_arguments():
    lib = L.load()
    assert lib.arn_bvh_build(0, None, None, 0, None, None, None) == L.ARN_E_INVALID
    with pytest.raises(api.ArnError):
        api.bvh_build(np.zeros((2, 6), np.float32), np.ones(2, np.float32), strategy=9)
