"""Soak run of the traversal fuzz: python tools/trav_soak.py FIRST COUNT [smem] — random degenerate triangle soups with clipped / affine spheres
at five coordinate scales, through the binary, 4-wide and compressed 8-wide walks; ids and t against the oracle.  `smem`: soups of < 1200 triangles
and batches of 35 000 rays, i.e. through the shared-memory pair walk (traverse2p) of the batched queries."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from arendur_b200 import api, _lib as L
import oracle_lib as O
import test_gpu_round2 as T
first, count = int(sys.argv[1]), int(sys.argv[2])
smem = len(sys.argv) > 3 and sys.argv[3] == "smem"
widths = (0,) if smem else (2, 4, 8)
ctx = api.Context(0)
bad = []; nrays = 0; t0 = time.time()
for seed in range(first, first + count):
    rng = np.random.default_rng(seed)
    scale = [1e-3, 1.0, 3e4, 1e-6, 1e7][seed % 5]
    h, d, rays = T._soup_case(rng, scale, n_tri=int(rng.integers(1, 1200 if smem else 1500)), n=7000 if smem else 2000)
    if smem: assert d.n_nodes < 3 or ((d.n_nodes - 1) // 2) * 128 <= L.ARN_SMEM_NODE_BYTES
    osc = O.OracleScene(d); oh = osc.intersect_closest(rays)
    for width in widths:
        ctx.set_option(L.ARN_OPT_BVH_WIDTH, width)
        try:
            sc = ctx.upload(d)
            gh, ga = sc.intersect_closest(rays), sc.intersect_any(rays)
            sc.close()
        finally:
            ctx.set_option(L.ARN_OPT_BVH_WIDTH, 0)
        mism = np.nonzero((gh["prim_id"] != oh["prim_id"]) | (gh["t"] != oh["t"]))[0]
        anyok = np.array_equal(ga != 0, oh["prim_id"] >= 0)
        nrays += rays.shape[0]
        if mism.size or not anyok:
            bad.append((seed, width))
            print(f"seed {seed} scale {scale} width {width}: {mism.size} mismatches (any-hit equal {anyok}); first o {rays['o'][mism[:2]]} d {rays['d'][mism[:2]]} gpu {gh[mism[:2]]} oracle {oh[mism[:2]]}", flush=True)
    osc.close()
print(f"traversal soak: {count} scenes x {len(widths)} walks ({'shared-memory pair walk' if smem else 'widths 2, 4, 8'}), {nrays} rays, {len(bad)} (scene, width) pairs with differences {bad[:10]}, {time.time() - t0:.1f} s")
