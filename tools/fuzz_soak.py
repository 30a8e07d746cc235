"""Soak run of the shading fuzz (tests/test_gpu_round2.py::_random_scene) over many seeds: python tools/fuzz_soak.py FIRST COUNT
Every scene is rendered per sample on the device and in the oracle; prints the seeds whose samples or ray counts differ."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from arendur_b200 import api, _lib as L
import oracle_lib as O
import test_gpu_round2 as T
first, count = int(sys.argv[1]), int(sys.argv[2])
ctx = api.Context(0)
bad = []; t0 = time.time(); samples = 0
for seed in range(first, first + count):
    textured, wild, strat = seed % 3 == 0, seed % 2 == 0, seed % 5 == 0
    hs, cam, film, smp, prm = T._random_scene(seed, textured=textured, wild=wild)
    if strat:
        smp = api.make_sampler(smp.sampledx, smp.sampledy, smp.ndim, smp.seed, mode=L.ARN_SAMPLER_STRATIFIED)
    d = hs.desc()
    sc = ctx.upload(d); osc = O.OracleScene(d)
    _, g, st = sc.render_pt_samples(cam, film, smp, prm)
    _, o = osc.render_pt_samples(cam, film, smp, prm)
    _, ost, _ = osc.render_pt(cam, film, smp, prm)
    same = np.all(g.view(np.uint32) == o.view(np.uint32), axis=-1) | (np.isnan(g).any(-1) & np.isnan(o).any(-1))
    counts = (st.extend_rays, st.shadow_rays, st.mis_rays, st.invalid_samples) == (ost.extend_rays, ost.shadow_rays, ost.mis_rays, ost.invalid_samples)
    samples += same.size
    if not same.all() or not counts:
        bad.append(seed)
        print(f"seed {seed} textured {textured} wild {wild} strat {strat}: {(~same).sum()} of {same.size} samples differ, counts equal {counts}, first {np.argwhere(~same)[:3].tolist()}", flush=True)
    sc.close(); osc.close()
print(f"soak: {count} scenes, {samples} samples, {len(bad)} scenes with differences {bad}, {time.time() - t0:.1f} s")
