import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from arendur_b200 import api, scenes
os.environ["ARN_PROBE"] = "1"
hs, cam, film, smp, prm = scenes.cornell_scene(256, 256, 1, 1)
ctx = api.Context(0); sc = ctx.upload(hs.desc())
rng = np.random.default_rng(7); n = 1 << 20
rays = np.zeros(n, api.RAY_DTYPE)
rays["o"] = rng.uniform([-1.8, -1.3, 2.2], [1.8, 2.2, 5.8], (n, 3)).astype(np.float32)
v = rng.normal(size=(n, 3)); v /= np.linalg.norm(v, axis=1, keepdims=True)
rays["d"] = v.astype(np.float32); rays["tmax"] = np.inf
r = torch.from_numpy(rays.view(np.uint8).reshape(-1, 28)).cuda(); h = torch.empty((n, 8), dtype=torch.uint8, device="cuda")
print("random rays:", sc.intersect_closest_counted_dev(r.data_ptr(), n, h.data_ptr()))
# sorted by origin cell (coherent origins, random directions)
key = (np.floor((rays["o"] - [-1.8, -1.3, 2.2]) / 0.25).astype(np.int64) * [1, 64, 4096]).sum(1)
rs = rays[np.argsort(key, kind="stable")]
r2 = torch.from_numpy(rs.view(np.uint8).reshape(-1, 28)).cuda()
print("origin-sorted:", sc.intersect_closest_counted_dev(r2.data_ptr(), n, h.data_ptr()))
from arendur_b200 import _lib as L
for name, buf in (("random", r), ("origin-sorted", r2)):
    for _ in range(3):
        st = L.Stats(); sc.intersect_closest_dev(buf.data_ptr(), n, h.data_ptr(), st)
    print(name, f"{st.gpu_ms:.3f} ms -> {n/st.gpu_ms/1e3:.0f} Mrays/s")
