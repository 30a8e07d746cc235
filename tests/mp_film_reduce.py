"""Launched by tests/test_gpu_round2.py::test_film_reduce_over_nccl_two_ranks under torchrun (one rank per GPU): every rank
renders its cells of two sample slices, arn_film_reduce sums each slice on rank 0, arn_film_merge keeps the running film; rank 0
compares with its own one-GPU render of all samples."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist
from arendur_b200 import api, scenes

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
hs, cam, film, smp, prm = scenes.cornell_scene(160, 120, 2, 2)
ctx = api.Context(local)
sc = ctx.upload(hs.desc())


def exchange(b):
    box = [b]
    dist.broadcast_object_list(box, src=0)
    return box[0]


comm = api.FilmComm(ctx, rank, world, exchange)
ext = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", local))
npix = 160 * 120
acc = torch.zeros((120, 160, 4), device="cuda"); step = torch.zeros_like(acc)
with torch.cuda.stream(ext):
    for k in range(2):
        step.zero_()
        sc.render_pt_dev(cam, film, smp, api.make_pt_params(max_depth=8, rank=rank, world_size=world, spp_begin=2 * k, spp_end=2 * k + 2, subdiv=4), step.data_ptr())
        comm.reduce(step.data_ptr(), npix, 0)
        if rank == 0:
            api.film_merge(ctx, acc.data_ptr(), step.data_ptr(), npix)
ctx.synchronize()
ok = True
if rank == 0:
    full, st = sc.render_pt(cam, film, smp, api.make_pt_params(max_depth=8))
    got = acc.cpu().numpy()
    ok = bool(np.allclose(got, full, rtol=2e-5, atol=2e-5)) and abs(got[..., 3].mean() / full[..., 3].mean() - 1.0) < 1e-6
    print("film reduce OK" if ok else f"film reduce MISMATCH: max abs diff {np.abs(got - full).max()}")
comm.close(); sc.close(); ctx.close()
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
