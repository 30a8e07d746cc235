"""A/B of library builds on the same workloads (run on a GPU box): for every ARN_LIB_PATH given, a fresh process renders
C3-like steps and (optionally) C4 and prints frame times.  usage: python tools/ab_trace.py lib1.so lib2.so ... [--c4]"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, json, time
sys.path.insert(0, os.environ["ARN_ROOT"])
import numpy as np, torch
from arendur_b200 import api, scenes, _lib as L
out = {"lib": os.environ.get("ARN_LIB_PATH", "default") + ("+refill" if os.environ.get("ARN_REFILL") else "") + ("+w" + os.environ["AB_WIDTH"] if os.environ.get("AB_WIDTH") else "") + ("+nosmem" if os.environ.get("ARN_SMEM_NODES") == "0" else "") + ("+pdl" if os.environ.get("ARN_PDL") == "1" else "")}
ctx = api.Context(0)
if os.environ.get('AB_WIDTH'): ctx.set_option(L.ARN_OPT_BVH_WIDTH, int(os.environ['AB_WIDTH']))
def run(name, hs, cam, film, smp, prm, reps):
    sc = ctx.upload(hs.desc())
    res = {}
    for pipes in (0, 1):
        ctx.set_option(L.ARN_OPT_PIPELINES, pipes)
        best = None
        for k in range(reps + 1):
            f, st = sc.render_pt(cam, film, smp, prm)
            if k and (best is None or st.gpu_ms < best.gpu_ms): best = st
        rays = best.extend_rays + best.shadow_rays + best.mis_rays
        res["auto" if pipes == 0 else "serial"] = {"ms": best.gpu_ms, "mrays_s": rays / best.gpu_ms / 1e3, "trace_ms": best.extend_ms, "trace_inc_ms": best.extend_bounce_ms,
                    "trace_mrays_s": rays / best.extend_ms / 1e3 if best.extend_ms else None, "inc_mrays_s": best.extend_bounce_rays / best.extend_bounce_ms / 1e3 if best.extend_bounce_ms else None,
                    "launches": best.kernel_launches}
    ctx.set_option(L.ARN_OPT_PIPELINES, 0)
    res["film_mean"] = [float(v) for v in f.reshape(-1, 4).mean(0)]
    sc.close()
    out[name] = res
hs, cam, film, smp, prm = scenes.cornell_scene(1024, 1024, 32, 32)
run("c3_16spp", hs, cam, film, smp, api.make_pt_params(max_depth=8, spp_begin=0, spp_end=16), 3)
if os.environ.get("AB_C4") == "1":
    hs2, cam2, film2 = scenes.c2_heightfield_scene()
    sc2 = ctx.upload(hs2.desc())
    rays = scenes.pixel_center_rays(cam2, 1920, 1080); n = rays.shape[0]
    rd = torch.from_numpy(rays.view("u1").reshape(n, 28)).cuda(); hd = torch.empty((n, 8), dtype=torch.uint8, device="cuda")
    ms = []
    for k in range(8):
        st = L.Stats(); sc2.intersect_closest_dev(rd.data_ptr(), n, hd.data_ptr(), st); ms.append(st.gpu_ms)
    out["c2_primary_ms"] = min(ms[3:]); out["c2_mrays_s"] = n / min(ms[3:]) / 1e3
    sc2.close()
    hs, cam, film, smp, prm = scenes.c4_box_scene()
    run("c4", hs, cam, film, smp, prm, 3)
print("AB " + json.dumps(out))
'''
libs = [a for a in sys.argv[1:] if not a.startswith("--")]
for lib in libs:
    env = dict(os.environ, ARN_ROOT=ROOT, AB_C4="1" if "--c4" in sys.argv else "0")
    if lib.endswith("+w4"):                     # force the 4-wide walk (large trees default to the compressed 8-wide walk)
        env["AB_WIDTH"] = "4"; lib = lib[:-3]
    if lib.endswith("+nosmem"):                 # small trees walked from global memory (the round-2 default until the shared-memory walk)
        env["ARN_SMEM_NODES"] = "0"; lib = lib[:-7]
    if lib.endswith("+pdl"):                    # programmatic dependent launch along the kernel chain (ARN_OPT_PDL)
        env["ARN_PDL"] = "1"; lib = lib[:-4]
    if lib.endswith("+refill"):                 # the lane-refilling trace of the same build
        env["ARN_REFILL"] = "1"; lib = lib[:-7]
    if lib != "default":
        env["ARN_LIB_PATH"] = os.path.abspath(lib)
    r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
    lines = [l for l in r.stdout.splitlines() if l.startswith("AB ")]
    print(lines[-1] if lines else f"AB-FAILED {lib}: {r.stderr[-1500:]}", flush=True)
