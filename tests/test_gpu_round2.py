"""GPU tests added in round 2: roofline inputs pinned to the oracle's counters, the in-library film merge
(arn_film_merge / arn_film_reduce), the finer rank partition, and the conservative interior-node walk."""
import math
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle_lib as O
from arendur_b200 import api, scenes, _lib as L

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _camera_grid_rays(cam, w, h, step=1):
    xs, ys = np.meshgrid(np.arange(0, w, step) + 0.5, np.arange(0, h, step) + 0.5, indexing="xy")
    pf = np.zeros((xs.size, 4), np.float32)
    pf[:, 0], pf[:, 1] = xs.reshape(-1), ys.reshape(-1)
    return O.camera_rays(cam, pf)


def _dev(rays):
    import torch
    n = rays.shape[0]
    return torch.from_numpy(np.ascontiguousarray(rays).view(np.uint8).reshape(n, 28)).cuda(), torch.empty((n, 8), dtype=torch.uint8, device="cuda")


def test_traversal_counters_equal_the_oracles(ctx, cornell_small):
    """The Nn / Nt of the algorithmic-bytes figure are the REFERENCE traversal's node and primitive tests:
    the instrumented kernel must count exactly what the instrumented oracle counts (Cornell incl. spheres, and a height field)."""
    hs, cam, film, smp, prm = cornell_small
    cases = [(hs.desc(), _camera_grid_rays(cam, film.res_x, film.res_y))]
    h2 = api.HostScene()
    mat = h2.add_material(api.material(L.ARN_MAT_MATTE, kd=(0.7, 0.7, 0.7)))
    pos, idx = scenes.heightfield(96, -2.0, 2.0, 4.0, 0.15, 0x5EED)
    h2.add_mesh(pos, idx, mat)
    cam2 = api.make_camera(api.IDENTITY, (-16.0 / 9.0, -1.0, 16.0 / 9.0, 1.0), 0.1, 1000.0, math.pi / 2, 320, 180)
    cases.append((h2.build(), _camera_grid_rays(cam2, 320, 180)))
    rng = np.random.default_rng(3)
    inc = np.zeros(8000, api.RAY_DTYPE)
    inc["o"] = rng.uniform([-1.8, -1.3, 2.2], [1.8, 2.2, 5.8], (8000, 3)).astype(np.float32)
    v = rng.normal(size=(8000, 3)); inc["d"] = (v / np.linalg.norm(v, axis=1, keepdims=True)).astype(np.float32); inc["tmax"] = np.inf
    cases.append((hs.desc(), inc))
    for d, rays in cases:
        sc = ctx.upload(d); osc = O.OracleScene(d)
        rd, hd = _dev(rays)
        g = sc.intersect_closest_counted_dev(rd.data_ptr(), rays.shape[0], hd.data_ptr())
        oh, octr = osc.intersect_closest(rays, counters=True)
        assert tuple(int(x) for x in octr) == g, f"gpu counters {g} != oracle counters {tuple(int(x) for x in octr)}"
        gh = hd.cpu().numpy().view(api.HIT_DTYPE).reshape(-1)
        assert np.array_equal(gh["prim_id"], oh["prim_id"]) and np.array_equal(gh["t"], oh["t"])
        sc.close(); osc.close()


def test_render_counters_equal_the_oracles(ctx, cornell_small):
    """Same for the counted render pass bench.py takes bytes_per_ray from: nodes / triangles / spheres tested by ALL traversals."""
    hs, cam, film, smp, prm = cornell_small
    sc = ctx.upload(hs.desc()); osc = O.OracleScene(hs.desc())
    ctx.set_option(L.ARN_OPT_COUNT_TRAVERSAL, 1)
    try:
        _, st = sc.render_pt(cam, film, smp, prm)
    finally:
        ctx.set_option(L.ARN_OPT_COUNT_TRAVERSAL, 0)
    _, ost, trav = osc.render_pt(cam, film, smp, prm)
    assert (st.extend_nodes, st.extend_tris, st.extend_spheres) == tuple(int(x) for x in trav)
    assert (st.extend_rays, st.shadow_rays, st.mis_rays) == (ost.extend_rays, ost.shadow_rays, ost.mis_rays)
    sc.close(); osc.close()


def test_film_merge_on_device(ctx):
    import torch
    a = torch.rand((37, 53, 4), device="cuda"); b = torch.rand((37, 53, 4), device="cuda")
    want = (a + b).cpu()
    api.film_merge(ctx, a.data_ptr(), b.data_ptr(), 37 * 53)
    ctx.synchronize()
    assert torch.equal(a.cpu(), want)


@pytest.mark.parametrize("world,subdiv", [(2, 0), (3, 4), (8, 4), (5, 7)])
def test_subdivided_partition_sums_to_full_frame(ctx, cornell_small, world, subdiv):
    """arn_pt_params.partition_subdiv: the ranks' films add up to the one-rank film; every rank's film equals the oracle's
    film for the same (rank, world, subdiv) — same cells on both sides."""
    hs, cam, film, smp, prm = cornell_small
    sc = ctx.upload(hs.desc()); osc = O.OracleScene(hs.desc())
    full, st_full = sc.render_pt(cam, film, smp, api.make_pt_params(max_depth=4))
    acc = np.zeros_like(full, dtype=np.float64); cams = 0
    for r in range(world):
        p = api.make_pt_params(max_depth=4, rank=r, world_size=world, subdiv=subdiv)
        f, st = sc.render_pt(cam, film, smp, p)
        acc += f; cams += st.camera_rays
        if r in (0, world - 1):
            of, ost, _ = osc.render_pt(cam, film, smp, p)
            assert st.camera_rays == ost.camera_rays
            assert np.allclose(f, of, rtol=2e-5, atol=2e-5)
    assert cams == st_full.camera_rays
    assert np.allclose(acc, full, rtol=2e-5, atol=2e-5)
    sc.close(); osc.close()


def test_film_reduce_over_nccl_two_ranks():
    """arn_film_reduce (ncclReduce inside the C-ABI) + arn_film_merge across two GPUs == the one-GPU film."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    port = 23000 + os.getpid() % 3000
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", str(port),
                        os.path.join(ROOT, "tests", "mp_film_reduce.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "film reduce OK" in r.stdout
