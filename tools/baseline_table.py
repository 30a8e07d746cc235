"""Measures BASELINE.md §5's C1 and C2 rows on one GPU, with the CPU restatement (the oracle) timed on the box's
host cores beside them.  C3: bench.py.  C4: tools/c4.py.  C5: tools/c5.py."""
import json, math, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import oracle_lib as O
from arendur_b200 import api, scenes, _lib as L
from bench import ClockSampler

cores = os.cpu_count() or 1
ctx = api.Context(0)
clk = ClockSampler(0); clk.start()
out = {}
# ---- C1: Cornell 256 x 256, 16 spp
hs, cam, film, smp, prm = scenes.cornell_scene(256, 256, 4, 4)
sc = ctx.upload(hs.desc()); osc = O.OracleScene(hs.desc())
best = None
for _ in range(6):
    f, st = sc.render_pt(cam, film, smp, prm)
    if best is None or st.gpu_ms < best.gpu_ms: best = st
t = time.perf_counter(); rf, ost, _ = osc.render_pt(cam, film, smp, prm, nthreads=cores); cpu_s = time.perf_counter() - t
_, grad, _ = sc.render_pt_samples(cam, film, smp, prm)
_, orad = osc.render_pt_samples(cam, film, smp, prm)
same = np.all(grad.view(np.uint32) == orad.view(np.uint32), axis=-1)
g, _ = api.film_finalize(f); r, _ = O.film_finalize(rf)
ok = rf[..., 3] > 0.05 * np.median(rf[..., 3])
rays = best.extend_rays + best.shadow_rays + best.mis_rays
out["C1"] = {"gpu_ms": best.gpu_ms, "mrays_s": rays / best.gpu_ms / 1e3, "spp_s": best.camera_rays / best.gpu_ms * 1e3, "cpu_s": cpu_s, "cpu_mrays_s": rays / cpu_s / 1e6, "cores": cores,
             "samples_bit_exact": float(same.mean()), "image_rel_rmse": float(np.sqrt(np.mean((g[ok] - r[ok]) ** 2)) / np.mean(r[ok]))}
sc.close(); osc.close()
# ---- C2: 1 002 528-triangle height field, 1920 x 1080 primary rays
hs, cam, film = scenes.c2_heightfield_scene()
d = hs.desc()
w, h = film.res_x, film.res_y
xs, ys = np.meshgrid(np.arange(w) + 0.5, np.arange(h) + 0.5, indexing="xy")
pf = np.zeros((w * h, 4), np.float32); pf[:, 0], pf[:, 1] = xs.reshape(-1), ys.reshape(-1)
rays_h = O.camera_rays(cam, pf)
sc = ctx.upload(d); osc = O.OracleScene(d)
dr = torch.from_numpy(rays_h.view(np.uint8).reshape(-1)).cuda()
hits = torch.empty(rays_h.shape[0] * api.HIT_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
best_ms = 1e9
for _ in range(6):
    st = L.Stats(); sc.intersect_closest_dev(dr.data_ptr(), rays_h.shape[0], hits.data_ptr(), st); best_ms = min(best_ms, st.gpu_ms)
ctr = sc.intersect_closest_counted_dev(dr.data_ptr(), rays_h.shape[0], hits.data_ptr())
t = time.perf_counter(); oh = osc.intersect_closest(rays_h); cpu_s = time.perf_counter() - t
gh = hits.cpu().numpy().view(api.HIT_DTYPE)
bytes_per_ray = (32.0 * ctr[0] + 36.0 * ctr[1]) / rays_h.shape[0] + 36
out["C2"] = {"gpu_ms": best_ms, "mrays_s": rays_h.shape[0] / best_ms / 1e3, "cpu_s": cpu_s, "cpu_mrays_s": rays_h.shape[0] / cpu_s / 1e6, "cores": cores,
             "id_mismatches": int((gh["prim_id"] != oh["prim_id"]).sum()), "t_bit_exact": bool(gh["t"].tobytes() == oh["t"].tobytes()), "hit_fraction": float((oh["prim_id"] >= 0).mean()),
             "bytes_per_ray": bytes_per_ray, "algorithmic_gb_s": bytes_per_ray * rays_h.shape[0] / best_ms / 1e6, "frac_of_measured_6457": bytes_per_ray * rays_h.shape[0] / best_ms / 1e6 / 6457.1}
out["clocks"] = clk.stop()
print(json.dumps(out))
