import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from arendur_b200 import api, scenes
import oracle_lib as O
res = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (64, 48)
hs, cam, film, smp, prm = scenes.cornell_scene(res[0], res[1], 2, 2)
depth = int(sys.argv[3]) if len(sys.argv) > 3 else 8
prm = api.make_pt_params(max_depth=depth)
d = hs.desc()
ctx = api.Context(0); sc = ctx.upload(d); osc = O.OracleScene(d)
gf, grad, st = sc.render_pt_samples(cam, film, smp, prm)
of, orad = osc.render_pt_samples(cam, film, smp, prm)
g, o = grad[..., :3], orad[..., :3]
diff = np.abs(g - o).max(-1)
scale = np.maximum(np.abs(o).max(-1), 1e-3)
rel = diff / scale
print("samples", rel.size, "exact", (diff == 0).sum(), "rel>1e-5", (rel > 1e-5).sum(), "rel>1e-3", (rel > 1e-3).sum(), "rel>1e-1", (rel > 1e-1).sum())
bad = np.argwhere(rel > 1e-3)
for (y, x, s) in bad[:20]:
    print("pixel", x, y, "sample", s, "gpu", g[y, x, s], "oracle", o[y, x, s])
print("film w max diff", np.abs(gf[..., 3] - of[..., 3]).max(), "film rgb max rel", (np.abs(gf[..., :3] - of[..., :3]).max() / of[..., :3].max()))
