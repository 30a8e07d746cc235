"""Generates tests/golden/bench_oracle_stats.json: statistics of the ORACLE's film for sample 0 of every pixel of the
bench workloads (C3: Cornell 1024x1024; C5: Cornell 3840x2160).  bench.py renders the same sample on the GPU
and asserts that its film has the same means (the committed parity statistic of the film_check).
Run from the repo root: python tests/golden/make_bench_golden.py"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O

out = {}
flat = O.OracleFlatScene()
osc = O.OracleScene(flat.desc)
for name, (w, h, sx) in {"c3": (1024, 1024, 32), "c5": (3840, 2160, 64)}.items():
    t = time.time()
    film, st, _ = osc.render_pt(flat.camera(w, h), O.make_film(w, h), O.make_sampler(sx, sx), O.make_pt_params(flat.max_depth(), 0, 1))
    f64 = film.astype(np.float64)
    out[name] = {"width": w, "height": h, "sampled": [sx, sx], "spp_range": [0, 1], "seed": 0, "max_depth": flat.max_depth(),
                 "mean_rgbw": [float(v) for v in f64.reshape(-1, 4).mean(0)],
                 "rays": int(st.extend_rays + st.shadow_rays + st.mis_rays), "camera_rays": int(st.camera_rays)}
    print(name, out[name], f"{time.time() - t:.1f} s")
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "bench_oracle_stats.json"), "w"), indent=1)
