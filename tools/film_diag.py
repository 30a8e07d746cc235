import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from arendur_b200 import api, scenes
import oracle_lib as O
hs, cam, film, smp, prm = scenes.cornell_scene(96, 72, 2, 2)
d = hs.desc(); ctx = api.Context(0); sc = ctx.upload(d); osc = O.OracleScene(d)
gf, st = sc.render_pt(cam, film, smp, prm); rf, ost, _ = osc.render_pt(cam, film, smp, prm)
g, _ = api.film_finalize(gf); r, _ = O.film_finalize(rf)
print("sum maxdiff rel", np.abs(gf - rf).max() / np.abs(rf).max())
print("rmse all", np.sqrt(np.mean((g - r) ** 2)) / np.mean(r))
w = np.abs(rf[..., 3]); med = np.median(w)
for frac in (0.001, 0.01, 0.05, 0.2):
    ok = w >= frac * med
    print(frac, "ok frac", ok.mean(), "rmse", np.sqrt(np.mean((g[ok] - r[ok]) ** 2)) / np.mean(r[ok]))
print("median w", med, "min |w|", w.min(), "n w<0", (rf[..., 3] < 0).sum(), "rows with zero weight", (w.sum(1) == 0).sum())
rel = np.abs(g - r).max(-1) / np.maximum(np.abs(r).max(-1), 1e-6)
print("per-pixel rel err percentiles", np.percentile(rel, [50, 90, 99, 99.9, 100]))
