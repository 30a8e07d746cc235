"""C4: 20 000 172-triangle closed box, incoherent diffuse bounces, depth 8.
usage: python tools/c4.py [cells] [res] [sppx]   (no GPU: build only)"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from arendur_b200 import api, scenes, _lib as L
cells = int(sys.argv[1]) if len(sys.argv) > 1 else 1291
res = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
sx = int(sys.argv[3]) if len(sys.argv) > 3 else 4
gpu_build = os.environ.get("C4_GPU_BUILD") == "1"     # tree from arn_bvh_build_gpu instead of the reference's host SAH build
t = time.time()
hs, cam, film, smp, prm = scenes.c4_box_scene(cells=cells, res=res, sampledx=sx, sampledy=sx, build=not gpu_build)
if gpu_build:
    t1 = time.time()
    ctx0 = api.Context(0)
    t2 = time.time()
    d, ms = hs.build_gpu(ctx0)
    print(f"c4 cells={cells}: {d.n_triangles} triangles, {d.n_nodes} nodes; mesh generation {t1-t:.1f} s, GPU LBVH build {ms:.1f} ms on the device, "
          f"{time.time()-t2:.2f} s incl. bounds, copies and light table", flush=True)
else:
    d = hs.desc()
    print(f"c4 cells={cells}: {d.n_triangles} triangles, {d.n_nodes} nodes, build+flatten {time.time()-t:.1f} s", flush=True)
import torch
if not torch.cuda.is_available():
    sys.exit(0)
ctx = ctx0 if gpu_build else api.Context(0)
t = time.time(); sc = ctx.upload(d); print(f"upload {time.time()-t:.1f} s")
for rep in range(2):
    f, st_fast = sc.render_pt(cam, film, smp, prm)
print(f"c4 render {res}^2 x {sx*sx} spp, default pipelines: {st_fast.gpu_ms:.1f} ms, all rays {(st_fast.extend_rays + st_fast.shadow_rays + st_fast.mis_rays)/st_fast.gpu_ms/1e3:.1f} Mrays/s")
ctx.set_option(L.ARN_OPT_PIPELINES, 1)          # per-kernel timings need serial launches
for rep in range(2):
    f, st = sc.render_pt(cam, film, smp, prm)
rays = st.extend_rays + st.shadow_rays + st.mis_rays
print(f"c4 render {res}^2 x {sx*sx} spp: {st.gpu_ms:.1f} ms, all rays {rays/st.gpu_ms/1e3:.1f} Mrays/s; extend {st.extend_rays} rays in {st.extend_ms:.1f} ms = {st.extend_rays/st.extend_ms/1e3:.1f} Mrays/s; "
      f"incoherent (bounce>=1) {st.extend_bounce_rays} rays in {st.extend_bounce_ms:.1f} ms = {st.extend_bounce_rays/max(st.extend_bounce_ms,1e-9)/1e3:.1f} Mrays/s; invalid {st.invalid_samples}")
ctx.set_option(L.ARN_OPT_COUNT_TRAVERSAL, 1)
p1 = api.make_pt_params(max_depth=8, spp_begin=0, spp_end=1)
f, stc = sc.render_pt(cam, film, smp, p1)
rays_c = stc.extend_rays + stc.shadow_rays + stc.mis_rays        # the counters cover ALL traversals (path, shadow and light rays): divide by all of them, as bench.py does
bpr = (32.0 * stc.extend_nodes + 36.0 * stc.extend_tris + 152.0 * stc.extend_spheres) / rays_c + 36
print(f"Nn/ray {stc.extend_nodes/rays_c:.1f} Nt/ray {stc.extend_tris/rays_c:.2f} bytes/ray {bpr:.0f} (per traversal, all kinds) -> k_trace algorithmic {bpr*rays/st.extend_ms/1e6:.0f} GB/s of 6457 measured")
g, _ = api.film_finalize(f)
print("image mean", g.reshape(-1, 3).mean(0), "finite", np.isfinite(g).all())
if os.environ.get("C4_CPU") == "1":        # CPU restatement on a bounded sample: the same view at 256 x 256, 1 spp
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    osc = O.OracleScene(d)
    cam2 = api.make_camera(np.float32([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0], [0, 0, 3.5, 1]]), (-1.0, -1.0, 1.0, 1.0), 0.1, 1000.0, 1.2707964, 256, 256)
    t = time.time(); _, ost, _ = osc.render_pt(cam2, api.make_film(256, 256), api.make_sampler(1, 1, 8, 0), prm, nthreads=os.cpu_count()); dt = time.time() - t
    orays = ost.extend_rays + ost.shadow_rays + ost.mis_rays
    print(f"c4 CPU restatement, {os.cpu_count()} threads, 256^2 x 1 spp sample: {dt:.2f} s, {orays/dt/1e6:.2f} Mrays/s")
