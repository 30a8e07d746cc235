"""BASELINE C5: the Cornell scene at 3840 x 2160, 64 x 64 = 4096 spp, tile-partitioned over the ranks of one node,
NCCL film reduce.  Launch: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1
--master-port P tools/c5.py [--spp-slices K]   (K < 16 renders only the first K of the 16 slices of 256 spp and says so)."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from arendur_b200 import api, scenes

ap = argparse.ArgumentParser(); ap.add_argument("--spp-slices", type=int, default=16); args = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
W, H, SX = 3840, 2160, 64
hs, cam, film, smp, prm0 = scenes.cornell_scene(W, H, SX, SX)
ctx = api.Context(local); scene = ctx.upload(hs.desc())
ext = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", local))
film_dev = torch.zeros((H, W, 4), dtype=torch.float32, device="cuda")
slice_spp = SX * SX // 16
comm = None
if world > 1:
    def exchange(b):
        box = [b]; dist.broadcast_object_list(box, src=0); return box[0]
    comm = api.FilmComm(ctx, rank, world, exchange)
rays = samples = 0
def sync():
    if dist is not None: dist.barrier()
    torch.cuda.synchronize()
with torch.cuda.stream(ext):
    scene.render_pt_dev(cam, film, smp, api.make_pt_params(max_depth=prm0.max_depth, rank=rank, world_size=world, spp_begin=0, spp_end=2), torch.zeros_like(film_dev).data_ptr())   # warm-up
    sync(); t0 = time.perf_counter()
    for k in range(args.spp_slices):
        st = scene.render_pt_dev(cam, film, smp, api.make_pt_params(max_depth=prm0.max_depth, rank=rank, world_size=world, spp_begin=k * slice_spp, spp_end=(k + 1) * slice_spp, subdiv=4), film_dev.data_ptr())
        rays += st.extend_rays + st.shadow_rays + st.mis_rays; samples += st.camera_rays
    if comm is not None: comm.reduce(film_dev.data_ptr(), W * H, 0)      # arn_film_reduce: ncclReduce inside the C-ABI, once, after the last slice
    ctx.synchronize()
    sync(); dt = time.perf_counter() - t0
tot = torch.tensor([rays, samples], dtype=torch.float64, device="cuda")
if dist is not None: dist.all_reduce(tot)
if rank == 0:
    f = film_dev.cpu().numpy()
    g, _ = api.film_finalize(f)
    print(json.dumps({"config": f"C5: Cornell {W}x{H}, {args.spp_slices * slice_spp} of {SX*SX} spp, depth 8, {world} GPU(s), 16x16 tiles cut into 4x4 cells, cell -> rank interleave, arn_film_reduce (NCCL)",
                      "seconds": dt, "spp_per_s": tot[1].item() / dt, "mrays_per_s": tot[0].item() / dt / 1e6, "samples": tot[1].item(),
                      "image_mean_rgb": [float(v) for v in g.reshape(-1, 3).mean(0)], "finite": bool(np.isfinite(g).all())}))
if dist is not None: dist.destroy_process_group()
