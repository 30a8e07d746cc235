"""A/B of the binary and the 4-wide BVH walk on height fields of growing size (primary + incoherent rays)."""
import math, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from arendur_b200 import api, scenes, _lib as L

ctx = api.Context(0)
for cells in [int(a) for a in (sys.argv[1:] or ["32", "64", "128", "256", "708"])]:
    hs = api.HostScene()
    mat = hs.add_material(api.material(L.ARN_MAT_MATTE, kd=(0.7, 0.7, 0.7)))
    pos, idx = scenes.heightfield(cells, -2.0, 2.0, 4.0, 0.15, 0x5EED)
    hs.add_mesh(pos, idx, mat)
    d = hs.build()
    sc = ctx.upload(d)
    w, h = 1920, 1080
    cam = api.make_camera(api.IDENTITY, (-16.0 / 9.0, -1.0, 16.0 / 9.0, 1.0), 0.1, 1000.0, math.pi / 2, w, h)
    prim = scenes.pixel_center_rays(cam, w, h)
    rng = np.random.default_rng(3)
    n = prim.shape[0]
    inc = np.zeros(n, api.RAY_DTYPE)
    inc["o"] = rng.uniform([-2, -2, 1.0], [2, 2, 3.5], (n, 3)).astype(np.float32)
    v = rng.normal(size=(n, 3)); v[:, 2] = np.abs(v[:, 2]) * 0.5 + 0.05; v /= np.linalg.norm(v, axis=1, keepdims=True)
    inc["d"] = v.astype(np.float32); inc["tmax"] = np.inf
    out = []
    for name, rays in (("primary", prim), ("incoherent", inc)):
        dr = torch.from_numpy(rays.view(np.uint8).reshape(-1)).cuda()
        res = {}
        for width in (2, 4):
            ctx.set_option(L.ARN_OPT_BVH_WIDTH, width)
            hits = torch.empty(n * api.HIT_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
            best = 1e9
            for rep in range(5):
                st = L.Stats(); sc.intersect_closest_dev(dr.data_ptr(), n, hits.data_ptr(), st); best = min(best, st.gpu_ms)
            res[width] = (best, hits)
        assert torch.equal(res[2][1], res[4][1])
        out.append(f"{name}: binary {n/res[2][0]/1e6:.2f} Grays/s, wide {n/res[4][0]/1e6:.2f} Grays/s")
    print(f"cells {cells}: {d.n_triangles} tris, {d.n_nodes} nodes | " + " | ".join(out), flush=True)
    ctx.set_option(L.ARN_OPT_BVH_WIDTH, 0)
    sc.close()
