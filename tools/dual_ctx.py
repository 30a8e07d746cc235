"""Experiment: do two concurrent wave pipelines (two contexts = two streams on one GPU) fill the tails of the
late-bounce launches?  Renders the same total work once on one context and once split over two threads."""
import os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from arendur_b200 import api, scenes
res, sx = int(sys.argv[1]), int(sys.argv[2])
hs, cam, film, smp, prm = scenes.cornell_scene(res, res, sx, sx)
spp = sx * sx
ctxs = [api.Context(0), api.Context(0)]
scs = [c.upload(hs.desc()) for c in ctxs]
films = [torch.zeros(res * res * 4, dtype=torch.float32, device="cuda") for _ in range(2)]

def run(i, b, e):
    p = api.make_pt_params(max_depth=8, spp_begin=b, spp_end=e)
    scs[i].render_pt_dev(cam, film, smp, p, films[i].data_ptr(), want_stats=False)
    ctxs[i].synchronize()

for rep in range(3):
    torch.cuda.synchronize(); t = time.time(); run(0, 0, spp); one = time.time() - t
    torch.cuda.synchronize(); t = time.time()
    th = [threading.Thread(target=run, args=(i, i * spp // 2, (i + 1) * spp // 2)) for i in range(2)]
    [x.start() for x in th]; [x.join() for x in th]
    two = time.time() - t
    print(f"one pipeline {one*1e3:.1f} ms, two concurrent pipelines {two*1e3:.1f} ms ({one/two:.3f}x)")
