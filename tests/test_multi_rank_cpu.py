"""N > 1 path on CPU: world_size-2 gloo processes render their tile partitions (rank, world) and sum
the films with torch.distributed.reduce — Film::merge_into semantics (film.rs:82-101).  The renderer
behind each rank is the oracle here (no GPU in this container); the partition arithmetic and the
reduce are what bench.py runs over NCCL."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    from arendur_b200 import api, scenes
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    hs, cam, film, smp, prm = scenes.cornell_scene(64, 64, 2, 1)
    osc = O.OracleScene(hs.desc())
    p = api.make_pt_params(max_depth=3, rank=rank, world_size=world)
    f, st, _ = osc.render_pt(cam, film, smp, p, nthreads=2)
    t = torch.from_numpy(f.copy())
    cnt = torch.tensor([float(st.camera_rays)])
    dist.reduce(t, dst=0); dist.reduce(cnt, dst=0)
    if rank == 0:
        full, fst, _ = osc.render_pt(cam, film, smp, api.make_pt_params(max_depth=3), nthreads=2)
        np.savez(out_path, merged=t.numpy(), full=full, cnt=cnt.numpy(), full_cnt=float(fst.camera_rays))
    dist.destroy_process_group()


def test_two_rank_tile_partition_and_film_reduce(tmp_path):
    out = str(tmp_path / "r.npz")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    z = np.load(out)
    assert z["cnt"][0] == z["full_cnt"] == 64 * 64 * 2
    # every pixel's (sum, weight) equals the single-rank render up to the order of float additions
    assert np.allclose(z["merged"], z["full"], rtol=1e-5, atol=1e-6)
    assert (z["full"][..., 3] > 0).all()


def _worker_steps(rank, world, port, out_path):
    """bench.py's N > 1 step structure: every step renders its sample slice of this rank's cells (partition_subdiv = 4) into a
    zeroed per-step film, the step films are summed on rank 0 and merged into the running film there."""
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    flat = O.OracleFlatScene()
    osc = O.OracleScene(flat.desc)
    cam, film, smp = flat.camera(72, 56), O.make_film(72, 56), O.make_sampler(2, 2)
    acc = torch.zeros((56, 72, 4))
    for step in range(2):
        p = O.make_pt_params(3, spp_begin=2 * step, spp_end=2 * step + 2, rank=rank, world_size=world, subdiv=4)
        f, st, _ = osc.render_pt(cam, film, smp, p, nthreads=2)
        t = torch.from_numpy(f.copy())
        dist.reduce(t, dst=0)                      # every step reduces ITS OWN film, never the running one
        if rank == 0:
            acc += t
    if rank == 0:
        full, fst, _ = osc.render_pt(cam, film, smp, O.make_pt_params(3), nthreads=2)
        np.savez(out_path, merged=acc.numpy(), full=full)
    dist.destroy_process_group()


def test_per_step_reduce_of_subdivided_tiles_equals_the_full_render(tmp_path):
    out = str(tmp_path / "s.npz")
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_worker_steps, args=(3, port, out), nprocs=3, join=True)
    z = np.load(out)
    assert np.allclose(z["merged"], z["full"], rtol=1e-5, atol=1e-6)
    w = z["full"][..., 3].mean()
    assert abs(z["merged"][..., 3].mean() / w - 1.0) < 1e-6        # no sample counted twice (round 1's bench bug)
