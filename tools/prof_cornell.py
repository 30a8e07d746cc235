"""One Cornell render for profiling: python tools/prof_cornell.py RES SPPX [DEPTH]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from arendur_b200 import api, scenes
res, sx = int(sys.argv[1]), int(sys.argv[2])
hs, cam, film, smp, prm = scenes.cornell_scene(res, res, sx, sx)
if len(sys.argv) > 3:
    prm = api.make_pt_params(max_depth=int(sys.argv[3]))
ctx = api.Context(0)
if os.environ.get("GPU_BUILD") == "1":        # tree from arn_bvh_build_gpu (LBVH) instead of the reference's SAH tree
    hs.build_gpu(ctx)
sc = ctx.upload(hs.desc())
reps = int(os.environ.get("REPS", "1"))
for _ in range(reps):
    f, st = sc.render_pt(cam, film, smp, prm)
rays = st.extend_rays + st.shadow_rays + st.mis_rays
print(f"cornell {res}^2 x {sx*sx}spp: {st.gpu_ms:.2f} ms, {rays/st.gpu_ms/1e3:.1f} Mrays/s, extend {st.extend_ms:.2f} ms (needs ARN_PIPES=1) / {st.extend_rays} rays, shadow {st.shadow_rays} mis {st.mis_rays}")
