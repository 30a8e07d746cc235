"""Bisect a per-sample GPU-vs-oracle difference by patching the material of the hit (debugging aid)."""
import os, sys, ctypes as C
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from arendur_b200 import api, _lib as L
import oracle_lib as O
import test_gpu_round2 as T
seed, y, x, smp_i, mat = 7, 17, 7, 3, 1
ctx = api.Context(0)
np.set_printoptions(precision=9, linewidth=200)
def run(tag, patch):
    hs, cam, film, smp, prm = T._random_scene(seed)
    d = hs.desc()
    m = C.cast(d.materials, C.POINTER(L.Material))
    patch(m[mat])
    sc = ctx.upload(d); osc = O.OracleScene(d)
    p = api.make_pt_params(max_depth=1, min_depth=prm.min_depth, rr_threshold=prm.rr_threshold)
    _, g, _ = sc.render_pt_samples(cam, film, smp, p)
    _, o = osc.render_pt_samples(cam, film, smp, p)
    nbad = (~np.all(g.view(np.uint32) == o.view(np.uint32), axis=-1)).sum()
    print(f"{tag:28s} differ {nbad:4d}  gpu {g[y, x, smp_i][:3]} oracle {o[y, x, smp_i][:3]}", flush=True)
    sc.close(); osc.close()
def setv(m, **kw):
    for k, v in kw.items():
        if k in ("kd", "ks"): getattr(m, k)[:] = v
        else: setattr(m, k, v)
run("as is", lambda m: None)
run("ks = 0", lambda m: setv(m, ks=(0, 0, 0)))
run("kd = 0", lambda m: setv(m, kd=(0, 0, 0)))
run("kd = ks = grey", lambda m: setv(m, kd=(0.5, 0.5, 0.5), ks=(0.5, 0.5, 0.5)))
run("alpha 0.3", lambda m: setv(m, alpha=0.3))
run("dissolve 0.5", lambda m: setv(m, dissolve=0.5))
run("plastic", lambda m: setv(m, type=L.ARN_MAT_PLASTIC))
