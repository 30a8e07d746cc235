"""Wave-size / pipeline sweep on C3 (16 spp of 1024^2): python tools/wave_sweep.py   (run on a GPU box)
env SWEEP_SPP (16), SWEEP_WAVES ("16,17,18,19,20": log2 of the wave sizes), SWEEP_PIPES ("2,4,8")"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys
sys.path.insert(0, os.environ["ARN_ROOT"])
from arendur_b200 import api, scenes, _lib as L
hs, cam, film, smp, prm = scenes.cornell_scene(1024, 1024, 32, 32)
ctx = api.Context(0); sc = ctx.upload(hs.desc())
p = api.make_pt_params(max_depth=8, spp_begin=0, spp_end=int(os.environ.get("SWEEP_SPP", "16")))
best = 1e9
for k in range(4):
    f, st = sc.render_pt(cam, film, smp, p)
    if k: best = min(best, st.gpu_ms)
print(f"SWEEP wave={os.environ.get('ARN_WAVE')} pipes={os.environ.get('ARN_PIPES')} ms={best:.2f} launches={st.kernel_launches}")
'''
for wave in [1 << int(v) for v in os.environ.get("SWEEP_WAVES", "16,17,18,19,20").split(",")]:
    for pipes in [int(v) for v in os.environ.get("SWEEP_PIPES", "2,4,8").split(",")]:
        env = dict(os.environ, ARN_ROOT=ROOT, ARN_WAVE=str(wave), ARN_PIPES=str(pipes))
        r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
        print([l for l in r.stdout.splitlines() if l.startswith("SWEEP")] or r.stderr[-500:], flush=True)
