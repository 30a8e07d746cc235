import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from arendur_b200 import api, scenes, _lib as L
hs, cam, film, smp, prm = scenes.cornell_scene(96, 72, 2, 2)
d = hs.desc()
ctx = api.Context(0)
ctx.set_option(L.ARN_OPT_BVH_WIDTH, 8)
sc = ctx.upload(d)
rng = np.random.default_rng(21)
n = 100_000
rays = np.zeros(n, api.RAY_DTYPE)
rays["o"] = rng.uniform([-1.8, -1.3, 2.2], [1.8, 2.2, 5.8], (n, 3)).astype(np.float32)
v = rng.normal(size=(n, 3)); v /= np.linalg.norm(v, axis=1, keepdims=True)
rays["d"] = v.astype(np.float32)
rays["tmax"] = np.where(rng.random(n) < 0.3, rng.uniform(0.1, 4.0, n), np.inf).astype(np.float32)
h8 = sc.intersect_closest(rays)
ctx.set_option(L.ARN_OPT_BVH_WIDTH, 2)
h2 = sc.intersect_closest(rays)
bad = np.nonzero((h8["prim_id"] != h2["prim_id"]) | (h8["t"] != h2["t"]))[0]
print("mismatches", bad.size, "of", n)
for i in bad[:12]:
    print(i, "o", rays["o"][i], "d", rays["d"][i], "tmax", rays["tmax"][i], "cw8", h8[i], "bin", h2[i])
