"""One C4 render for profiling: python tools/prof_c4.py WIDTH [SPPX]   (WIDTH 4 or 8; ARN_PIPES=1 for serial launches)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from arendur_b200 import api, scenes, _lib as L
width = int(sys.argv[1]); sx = int(sys.argv[2]) if len(sys.argv) > 2 else 1
hs, cam, film, smp, prm = scenes.c4_box_scene(sampledx=sx, sampledy=sx)
ctx = api.Context(0)
ctx.set_option(L.ARN_OPT_BVH_WIDTH, width)
sc = ctx.upload(hs.desc())
for _ in range(2):
    f, st = sc.render_pt(cam, film, smp, prm)
rays = st.extend_rays + st.shadow_rays + st.mis_rays
print(f"c4 width {width}: {st.gpu_ms:.2f} ms, {rays/st.gpu_ms/1e3:.1f} Mrays/s, trace {st.extend_ms:.2f} ms")
