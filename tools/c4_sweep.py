"""Pipelines x wave capacity sweep on C4 (one scene build)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from arendur_b200 import api, scenes, _lib as L
hs, cam, film, smp, prm = scenes.c4_box_scene(res=1024, sampledx=4, sampledy=4)
ctx = api.Context(0); sc = ctx.upload(hs.desc())
for pipes, wave in ((8, 1 << 18), (8, 1 << 17), (4, 1 << 18), (8, 1 << 16), (6, 1 << 18)):
    ctx.set_option(L.ARN_OPT_PIPELINES, pipes); ctx.set_option(L.ARN_OPT_WAVE_CAPACITY, wave)
    best = 1e9
    for _ in range(3):
        f, st = sc.render_pt(cam, film, smp, prm); best = min(best, st.gpu_ms)
    rays = st.extend_rays + st.shadow_rays + st.mis_rays
    print(f"pipes {pipes} wave {wave}: {best:.1f} ms, {rays/best/1e3:.1f} Mrays/s", flush=True)
