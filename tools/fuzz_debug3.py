import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from arendur_b200 import api
import test_gpu_round2 as T
hs, cam, film, smp, prm = T._random_scene(7)
d = hs.desc()
ctx = api.Context(0); sc = ctx.upload(d)
p = api.make_pt_params(max_depth=1, min_depth=prm.min_depth, rr_threshold=prm.rr_threshold)
_, g, _ = sc.render_pt_samples(cam, film, smp, p)
ctx.synchronize()
print(g[17, 7, 3])
