"""First GPU contact: smoke + quick timings (not a bench; bench.py is the contract)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as G
G.smoke()
from arendur_b200 import api, scenes
import torch
ctx = api.Context(0)
# C1: 256x256, 16 spp
for (res, sx) in ((256, 4), (512, 8)):
    hs, cam, film, smp, prm = scenes.cornell_scene(res, res, sx, sx)
    sc = ctx.upload(hs.desc())
    for it in range(3):
        t = time.time(); f, st = sc.render_pt(cam, film, smp, prm); dt = time.time() - t
    rays = st.extend_rays + st.shadow_rays + st.mis_rays
    print(f"cornell {res}x{res}x{sx*sx}spp: gpu {st.gpu_ms:.2f} ms wall {dt*1e3:.1f} ms rays {rays} -> {rays/st.gpu_ms/1e3:.1f} Mrays/s, "
          f"spp/s {res*res*sx*sx/st.gpu_ms*1e3:.3e}, extend_ms {st.extend_ms:.2f} ({st.extend_rays} rays), launches {st.kernel_launches}")
    sc.close()
# C2 full
t = time.time(); hs, cam, film = scenes.c2_heightfield_scene(); print("c2 build %.2fs" % (time.time() - t))
d = hs.desc(); print("c2 nodes", d.n_nodes, "prims", d.n_prims)
sc = ctx.upload(d)
rays = scenes.pixel_center_rays(cam, 1920, 1080)
r_dev = torch.from_numpy(rays.view(np.uint8).reshape(-1, 28)).cuda()
h_dev = torch.empty((rays.shape[0], 8), dtype=torch.uint8, device="cuda")
from arendur_b200 import _lib as L
for it in range(5):
    st = L.Stats(); sc.intersect_closest_dev(r_dev.data_ptr(), rays.shape[0], h_dev.data_ptr(), st)
    print(f"c2 closest: {st.gpu_ms:.3f} ms -> {rays.shape[0]/st.gpu_ms/1e3:.1f} Mrays/s")
nn, nt, ns = sc.intersect_closest_counted_dev(r_dev.data_ptr(), rays.shape[0], h_dev.data_ptr())
hits = h_dev.cpu().numpy().view(api.HIT_DTYPE).reshape(-1)
print("c2 counters nodes", nn, "tris", nt, "hit fraction", (hits["prim_id"] >= 0).mean(), "bytes/ray", (32*nn + 36*nt)/rays.shape[0] + 36)
t = time.time(); hh = sc.intersect_closest(rays); print("c2 e2e host buffers: %.2f ms" % ((time.time() - t) * 1e3))
