"""Locate a GPU-vs-oracle per-sample difference of the shading fuzz (tests/test_gpu_round2.py::_random_scene):
python tools/fuzz_debug.py SEED  — for every max_depth 1..D prints the samples that differ."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from arendur_b200 import api, _lib as L
import oracle_lib as O
import test_gpu_round2 as T
seed = int(sys.argv[1])
hs, cam, film, smp, prm = T._random_scene(seed)
d = hs.desc()
ctx = api.Context(0); sc = ctx.upload(d); osc = O.OracleScene(d)
D = prm.max_depth
for depth in range(1, D + 1):
    p = api.make_pt_params(max_depth=depth, min_depth=prm.min_depth, rr_threshold=prm.rr_threshold)
    _, g, _ = sc.render_pt_samples(cam, film, smp, p)
    _, o = osc.render_pt_samples(cam, film, smp, p)
    bad = np.argwhere(~np.all(g.view(np.uint32) == o.view(np.uint32), axis=-1))
    print("depth", depth, "min_depth", p.min_depth, "differ", len(bad), [(tuple(b), g[tuple(b)][:3].tolist(), o[tuple(b)][:3].tolist()) for b in bad[:4]])
