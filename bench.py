#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 path-tracing core (contract: task prompt (4)).

Workload (BASELINE.json configs[2], "C3"): Cornell box (examples/cornellbox/cb.json), 1024x1024,
full PT bounce loop (depth 8, area-light NEE + MIS, Russian roulette), 1024 spp.  One STEP = one
pass of the whole hot path (generate -> trace -> shade -> resolve -> accumulate) over a batch of
64 spp x 1024 x 1024 camera samples (sample indices [64k, 64k+64)); the default 16 timed steps
are exactly the 1024 spp of C3 accumulated into one film.  At N GPUs the same frame is
tile-partitioned and every step renders 64*N spp, i.e. per-GPU work is fixed ("weak"); each step's
per-rank films are summed on rank 0 by arn_film_reduce (ncclReduce inside the C-ABI) and merged into
the running film (Film::merge_into semantics, filming/film.rs:82-101), all inside the timed region.
`--workload c5` runs BASELINE.json configs[4] the same way (3840x2160, 4096 spp, 16*N spp per step).

  value  = Mrays/s, rays = every BVH traversal (path + shadow + light rays), film resident in HBM
  e2e    = same metric through the host-buffer C-ABI: scene upload (H2D) + render + reduce + film D2H
  roofline: k_trace (all BVH traversals), algorithmic bytes = 32*Nn + 36*Nt + 152*Ns + 36 per ray
  extra  = (N = 1) C2 (1 M triangles, 1920x1080 primary rays) and C4 (20 M triangles, depth 8) once each
  cpu_baseline / --impl reference: the oracle (CPU restatement of arendur; the Rust original cannot
  be built here) on the host cores, bounded sample of the same workload; that arm never loads the
  product library.

The run FAILS (non-zero exit, no JSON line) when the film it produced is wrong: see film_checks().
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (width, height, sampled per axis, spp per step per GPU, description)
    "c3": (1024, 1024, 32, 64, "C3: Cornell box 1024x1024, PT depth 8 + area-light NEE/MIS, 1024 spp (64 spp per step per GPU)"),
    "c5": (3840, 2160, 64, 16, "C5: Cornell box 3840x2160, PT depth 8 + area-light NEE/MIS, 4096 spp (16 spp per step per GPU), tile-partitioned, NCCL film reduce"),
}
TILES = tuple(int(v) for v in os.environ.get("BENCH_TILES", "16,16").split(","))   # Film::spawn_tiles grid (pt.rs:131 uses 16 x 16)
SUBDIV = int(os.environ.get("BENCH_SUBDIV", "4"))    # N > 1: every tile is cut into SUBDIV x SUBDIV cells for the rank interleave (arn_pt_params.partition_subdiv)
PIPELINES = 4            # concurrent wave pipelines of the timed region (the library default; ARN_OPT_PIPELINES)
METRIC = "Mrays/s (primary + incoherent bounce: every BVH traversal) on the Cornell box"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------- CPU arm (oracle only; never touches libarn_b200.so)
class OracleWorkload:
    def __init__(self, wl):
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_lib as O
        self.O = O
        self.w, self.h, self.sx = WORKLOADS[wl][0], WORKLOADS[wl][1], WORKLOADS[wl][2]
        self.flat = O.OracleFlatScene()                       # fixture -> oracle's own mesh transform, BVH::new, Scene::new
        self.osc = O.OracleScene(self.flat.desc)
        self.cam, self.film, self.smp = self.flat.camera(self.w, self.h), O.make_film(self.w, self.h), O.make_sampler(self.sx, self.sx)

    def run(self, s0, s1, threads):
        p = self.O.make_pt_params(self.flat.max_depth(), s0, s1)
        t = time.perf_counter()
        _, st, _ = self.osc.render_pt(self.cam, self.film, self.smp, p, nthreads=threads)
        dt = time.perf_counter() - t
        return dt, int(st.extend_rays + st.shadow_rays + st.mis_rays), int(st.camera_rays)


def oracle_sample(wl, budget_s, threads):
    """Times the oracle on a bounded sample of the workload: the whole frame, as many samples per pixel as fit
    the time budget (calibrated with a 1-spp pass)."""
    ow = OracleWorkload(wl)
    dt1, _, _ = ow.run(0, 1, threads)
    n = max(1, min(64, int(budget_s / max(dt1, 1e-3))))
    dt, rays, samples = ow.run(1, 1 + n, threads)
    return {"seconds": dt, "rays": rays, "samples": samples, "spp": n, "mrays_s": rays / dt / 1e6, "spp_s": samples / dt, "w": ow.w, "h": ow.h}


def reference_arm(args):
    """--impl reference: the reference's CPU implementation of the path.  arendur is Rust (2017 nightly) and cannot
    be built in this image, so this is the oracle port, all host threads.  Scene assembly, BVH build and rendering
    all run in oracle/liboracle.so; the product library is not loaded."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    ow = OracleWorkload(args.workload)
    per_step, rays_tot, samples_tot = [], 0, 0
    spp_total = ow.sx * ow.sx
    # bounded sample: each step = 1 spp of the full frame (the GPU arm's step is 64 spp (C3) / 16 spp (C5) per GPU)
    for i in range(args.warmup + args.steps):
        dt, rays, samples = ow.run(i % spp_total, i % spp_total + 1, threads)
        if i >= args.warmup:
            per_step.append(dt); rays_tot += rays; samples_tot += samples
    total = sum(per_step)
    value = rays_tot / total / 1e6
    assert "libarn_b200" not in open("/proc/self/maps").read(), "the reference arm must not load the product library"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Mrays/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / max(1, len(per_step)),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOADS[args.workload][4], "note": f"reference arm step = 1 spp of the {ow.w}x{ow.h} frame (bounded sample of the GPU arm's step; a rate metric)"},
        "spp_per_s": samples_tot / total,
        "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": threads, "kind": "port",
                         "sample": f"{args.steps} steps x ({ow.w}x{ow.h} px x 1 spp), oracle/ C++ restatement of arendur with std::thread over the 16x16 tile grid"},
        "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ---------------------------------------------------------------- extra legs (rank 0, one GPU): C2 and C4 under the driver's clock
def extra_c2(ctx, torch, peak):
    """C2: 1 002 528-triangle height field, one closest-hit query per pixel of 1920x1080 (Composable::intersect_ray batched)."""
    from arendur_b200 import api, scenes
    t0 = time.perf_counter()
    hs, cam, film = scenes.c2_heightfield_scene()
    d = hs.desc()
    build_s = time.perf_counter() - t0
    sc = ctx.upload(d)
    rays = scenes.pixel_center_rays(cam, 1920, 1080)
    n = rays.shape[0]
    rays_dev = torch.from_numpy(rays.view("u1").reshape(n, 28)).cuda()
    hits_dev = torch.empty((n, 8), dtype=torch.uint8, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    from arendur_b200 import _lib as L
    ms = []
    for k in range(8):
        flush.fill_(k); torch.cuda.synchronize()
        st = L.Stats()
        sc.intersect_closest_dev(rays_dev.data_ptr(), n, hits_dev.data_ptr(), st)
        if k >= 3:
            ms.append(st.gpu_ms)
    nn, nt, ns = sc.intersect_closest_counted_dev(rays_dev.data_ptr(), n, hits_dev.data_ptr())
    hits = hits_dev.cpu().numpy().view(api.HIT_DTYPE).reshape(-1)
    bpr = (32.0 * nn + 36.0 * nt + 152.0 * ns) / n + 36
    best = min(ms)
    out = {"workload": "C2: 1 002 528-triangle height field, 1920x1080 primary rays, closest hit only (k_closest_batch)", "triangles": int(d.n_triangles), "nodes": int(d.n_nodes),
           "rays": n, "ms": best, "mrays_s": n / best / 1e3, "hit_fraction": float((hits["prim_id"] >= 0).mean()), "nodes_per_ray": nn / n, "tris_per_ray": nt / n,
           "bytes_per_ray": bpr, "achieved_gbs": bpr * n / (best * 1e-3) / 1e9, "frac": bpr * n / (best * 1e-3) / 1e9 / peak, "host_build_s": build_s,
           "timing": "best of 5 single launches after 3 warm-ups, 256 MB L2 flush before each, CUDA events on the context's stream"}
    sc.close(); hs.close()
    return out


def extra_c4(ctx, torch, peak, clocks_index):
    """C4: 20 000 172-triangle closed box, 1024x1024 x 16 spp, depth 8, incoherent diffuse bounces; the reference's SAH tree."""
    from arendur_b200 import api, scenes, _lib as L
    t0 = time.perf_counter()
    hs, cam, film, smp, prm = scenes.c4_box_scene()
    d = hs.desc()
    build_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    sc = ctx.upload(d)
    upload_s = time.perf_counter() - t0
    film_dev = torch.zeros((1024, 1024, 4), dtype=torch.float32, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    clk = ClockSampler(clocks_index); clk.start()
    frames = []
    for k in range(5):
        flush.fill_(k); film_dev.zero_(); torch.cuda.synchronize()
        st = sc.render_pt_dev(cam, film, smp, prm, film_dev.data_ptr())
        ctx.synchronize()
        if k >= 2:
            frames.append((st.gpu_ms, st.extend_rays + st.shadow_rays + st.mis_rays, st.kernel_launches))
    clocks = clk.stop()
    finite = bool(torch.isfinite(film_dev).all().item())
    ctx.set_option(L.ARN_OPT_PIPELINES, 1)                  # per-kernel timings need serial launches
    ser = None
    for k in range(2):
        film_dev.zero_(); torch.cuda.synchronize()
        ser = sc.render_pt_dev(cam, film, smp, prm, film_dev.data_ptr())
    ctx.set_option(L.ARN_OPT_PIPELINES, 0)
    ctx.set_option(L.ARN_OPT_COUNT_TRAVERSAL, 1)
    stc = sc.render_pt_dev(cam, film, smp, api.make_pt_params(max_depth=8, spp_begin=0, spp_end=1), film_dev.data_ptr())
    ctx.set_option(L.ARN_OPT_COUNT_TRAVERSAL, 0)
    rays_c = stc.extend_rays + stc.shadow_rays + stc.mis_rays
    bpr = (32.0 * stc.extend_nodes + 36.0 * stc.extend_tris + 152.0 * stc.extend_spheres) / rays_c + 36       # / ALL traversals, as for C3
    ms, rays, launches = sorted(frames)[len(frames) // 2]
    ser_rays = ser.extend_rays + ser.shadow_rays + ser.mis_rays
    ach = bpr * ser_rays / (ser.extend_ms * 1e-3) / 1e9
    out = {"workload": "C4: 20 000 172-triangle closed box, 1024x1024 x 16 spp, depth 8, two sphere lights, reference SAH tree, 4-wide walk",
           "triangles": int(d.n_triangles), "nodes": int(d.n_nodes), "ms_per_frame": ms, "rays_per_frame": int(rays), "mrays_s": rays / ms / 1e3,
           "incoherent_mrays_s": ser.extend_bounce_rays / max(ser.extend_bounce_ms, 1e-9) / 1e3, "trace_mrays_s": ser_rays / max(ser.extend_ms, 1e-9) / 1e3,
           "nodes_per_ray": stc.extend_nodes / rays_c, "tris_per_ray": stc.extend_tris / rays_c, "bytes_per_ray": bpr,
           "achieved_gbs": ach, "frac": ach / peak, "launches_per_frame": int(launches), "finite": finite, "clocks": clocks,
           "host_build_s": build_s, "upload_s": upload_s,
           "timing": "mrays_s: median of 3 frames (after 2 warm-ups) with the default 8 wave pipelines, all traversals / CUDA-event frame time; "
                     "incoherent_mrays_s / trace_mrays_s / frac: k_trace launches of a serial (1-pipeline) frame, bounces >= 1 only for 'incoherent'"}
    sc.close(); hs.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=16)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="arendur_b200")
    ap.add_argument("--workload", default=os.environ.get("BENCH_WORKLOAD", "c3"), choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="time budget of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the C2 / C4 legs (they only run at N = 1)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return reference_arm(args)

    import numpy as np
    import torch
    from arendur_b200 import api, scenes, _lib as L

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the path-tracing core has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))

    W, H, SX, SPP_GPU, WORKLOAD = WORKLOADS[args.workload]
    NPIX = W * H
    hs, cam, film, smp, prm0 = scenes.cornell_scene(W, H, SX, SX)
    desc = hs.desc()
    ctx = api.Context(local)
    scene = ctx.upload(desc)
    ext = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", local))
    film_acc = torch.zeros((H, W, 4), dtype=torch.float32, device="cuda")        # the running film (meaningful on rank 0)
    film_step = torch.zeros((H, W, 4), dtype=torch.float32, device="cuda") if world > 1 else None   # this step's per-rank film
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")      # > 126 MB L2
    spp_step = SPP_GPU * world
    n_slices = (SX * SX) // spp_step
    subdiv = SUBDIV if world > 1 else 0

    comm = None
    if world > 1:       # the library's own communicator: the unique id travels over the launcher's rendezvous (plumbing)
        def exchange(b):
            box = [b]
            dist.broadcast_object_list(box, src=0)
            return box[0]
        comm = api.FilmComm(ctx, rank, world, exchange)

    def params(step, r=rank, w=world, sd=None):
        k = step % n_slices
        return api.make_pt_params(max_depth=prm0.max_depth, rank=r, world_size=w, spp_begin=k * spp_step, spp_end=(k + 1) * spp_step, tiles=TILES,
                                  subdiv=subdiv if sd is None else sd)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def device_step(step, want_stats=True):
        """One step on the context's stream.  1 GPU: accumulate straight into the running film.  N GPUs: render this step's
        samples of this rank's cells into a zeroed film, sum the films on rank 0 (arn_film_reduce), merge into the running film."""
        if world == 1:
            return scene.render_pt_dev(cam, film, smp, params(step), film_acc.data_ptr(), want_stats=want_stats), 1
        film_step.zero_()
        st = scene.render_pt_dev(cam, film, smp, params(step), film_step.data_ptr(), want_stats=want_stats)
        comm.reduce(film_step.data_ptr(), NPIX, 0)
        if rank == 0:
            api.film_merge(ctx, film_acc.data_ptr(), film_step.data_ptr(), NPIX)
        return st, 2 if rank == 0 else 1          # + the memset, + the merge kernel

    # ---- warm-up
    warm_rays = []
    with torch.cuda.stream(ext):
        for w in range(args.warmup):
            stw, _ = device_step(w)
            warm_rays.append(int(stw.extend_rays + stw.shadow_rays + stw.mis_rays))
    barrier()
    film_acc.zero_()
    # ---- timed: device-resident film
    clocks = ClockSampler(local)
    clocks.start()
    step_ms, render_ms, ext_ms, ext_rays, rays, samples, launches, inc_ms, inc_rays = [], [], 0.0, 0, 0, 0, 0, 0.0, 0
    with torch.cuda.stream(ext):
        for k in range(args.steps):
            flush.fill_(k & 0xFF)                      # L2 flush between timed iterations (untimed)
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(ext)
            st, extra_launches = device_step(k)
            e1.record(ext)
            barrier()
            step_ms.append(e0.elapsed_time(e1)); render_ms.append(st.gpu_ms)
            rays += st.extend_rays + st.shadow_rays + st.mis_rays
            samples += st.camera_rays
            launches += st.kernel_launches + extra_launches
    clk = clocks.stop()
    total_ms = sum(step_ms)
    ctx.synchronize()

    # ---- film checks (the run fails if the film is wrong)
    checks = film_checks(args, np, torch, api, scene, ctx, cam, film, smp, params, film_acc, film_step, rank, world, spp_step, NPIX, barrier, prm0, ext)

    # ---- kernel-timing pass for the roofline: the same steps once more with ONE wave pipeline.  The timed region above
    # runs several wave pipelines concurrently (ARN_OPT_PIPELINES), where a kernel's launch-to-finish time is shared with
    # other kernels; CUDA events around every k_trace launch only measure that kernel when launches are serial.
    timing_steps = max(1, min(args.steps, 4))
    serial_ms = 0.0
    ctx.set_option(L.ARN_OPT_PIPELINES, 1)
    scratch_t = torch.zeros_like(film_acc)
    with torch.cuda.stream(ext):
        for k in range(timing_steps):
            flush.fill_(k & 0xFF)
            torch.cuda.synchronize()
            st = scene.render_pt_dev(cam, film, smp, params(k), scratch_t.data_ptr(), want_stats=True)
            ext_ms += st.extend_ms; ext_rays += st.extend_rays + st.shadow_rays + st.mis_rays
            inc_ms += st.extend_bounce_ms; inc_rays += st.extend_bounce_rays
            serial_ms += st.gpu_ms
    ctx.set_option(L.ARN_OPT_PIPELINES, 0)              # back to auto (= PIPELINES for this scene)
    torch.cuda.synchronize()

    # ---- instrumented pass (untimed): Nn / Nt of all traversals -> algorithmic bytes per ray
    ctx.set_option(L.ARN_OPT_COUNT_TRAVERSAL, 1)
    scratch_t.zero_()
    with torch.cuda.stream(ext):
        stc = scene.render_pt_dev(cam, film, smp, api.make_pt_params(max_depth=prm0.max_depth, rank=rank, world_size=world, spp_begin=0, spp_end=min(8, spp_step), subdiv=subdiv), scratch_t.data_ptr())
    ctx.set_option(L.ARN_OPT_COUNT_TRAVERSAL, 0)
    torch.cuda.synchronize()
    del scratch_t
    rays_counted = stc.extend_rays + stc.shadow_rays + stc.mis_rays
    bytes_per_ray = (32.0 * stc.extend_nodes + 36.0 * stc.extend_tris + 152.0 * stc.extend_spheres) / max(1, rays_counted) + 28 + 8

    # ---- e2e: host buffers through the C-ABI, copies inside the timed region
    scene_bytes = int(desc.n_nodes * 32 + desc.n_prims * 48 + desc.n_spheres * 176 + desc.n_triangles * 16 + desc.n_vertices * 32
                      + desc.n_meshes * 16 + desc.n_materials * 64 + desc.n_prims * 4 + desc.n_lights * 12 + 4)
    e2e_rays, e2e_ms = 0, 0.0
    host_film_t = torch.empty((H, W, 4), dtype=torch.float32, pin_memory=True)      # pinned: the D2H read of the result
    host_film = host_film_t.numpy()
    e2e_steps = max(1, min(args.steps, 4))
    for k in range(-1, e2e_steps):                                       # k = -1: one untimed warm-up of this path (first-use allocations)
        barrier()
        t0 = time.perf_counter()
        sc2 = ctx.upload(desc)                                           # H2D of the flattened scene
        if world == 1:
            _, st = sc2.render_pt(cam, film, smp, params(max(k, 0)), out=host_film)   # render + film D2H into pinned memory (synchronous)
        else:
            with torch.cuda.stream(ext):
                film_step.zero_()
                st = sc2.render_pt_dev(cam, film, smp, params(max(k, 0)), film_step.data_ptr())
                comm.reduce(film_step.data_ptr(), NPIX, 0)               # films gathered on rank 0 over NVLink
                if rank == 0:
                    host_film_t.copy_(film_step, non_blocking=True)      # D2H of the merged film
            ctx.synchronize()
        sc2.close()
        barrier()
        if k < 0:
            continue
        e2e_ms += (time.perf_counter() - t0) * 1e3
        e2e_rays += st.extend_rays + st.shadow_rays + st.mis_rays

    # ---- reduce over ranks: MAX time, SUM work
    def allmax(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); return float(t.item())

    def allmin(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MIN); return float(t.item())

    def allsum(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.SUM); return float(t.item())
    total_ms_max, e2e_ms_max = allmax(total_ms), allmax(e2e_ms)
    rays_all, samples_all, launches_all, e2e_rays_all = allsum(rays), allsum(samples), allsum(launches), allsum(e2e_rays)
    rank_render_ms = sum(render_ms) / max(1, len(render_ms))
    rank_ms_min, rank_ms_max = allmin(rank_render_ms), allmax(rank_render_ms)

    extra = None
    if rank == 0 and world == 1 and not args.no_extra:
        peak0, _ = load_peaks()
        scene.close(); scene = None                        # free the wave buffers' neighbours before the 20 M-triangle scene
        extra = {}
        for name, fn in (("c2", lambda: extra_c2(ctx, torch, peak0)), ("c4", lambda: extra_c4(ctx, torch, peak0, local))):
            try:
                extra[name] = fn()
            except Exception as e:                          # an extra leg never takes the headline down with it
                extra[name] = {"error": f"{type(e).__name__}: {e}"}

    if rank == 0:
        peak, peak_src = load_peaks()
        cap = 1 << 19
        n_ext_launch = timing_steps * (prm0.max_depth + 1) * max(1, (NPIX * spp_step // world + cap - 1) // cap)
        achieved = bytes_per_ray * ext_rays / (ext_ms * 1e-3) / 1e9 if ext_ms > 0 else 0.0
        traffic, issue = None, None
        tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tp):
            try:
                tj = json.load(open(tp))
                traffic = tj.get("k_trace_dram_bytes_per_launch")
                # the limit that actually binds on the cache-resident Cornell scene: FP32 issue.  Instructions per
                # ray come from the committed ncu capture of this command, the rate from the live timing above.
                if ext_ms > 0 and tj.get("k_trace_thread_inst_per_ray"):
                    rays_s = ext_rays / (ext_ms * 1e-3)
                    issue = {"thread_inst_per_ray": tj["k_trace_thread_inst_per_ray"], "achieved_thread_inst_per_s": tj["k_trace_thread_inst_per_ray"] * rays_s,
                             "peak_thread_inst_per_s": tj["fp32_issue_peak_thread_inst_per_s"], "frac": tj["k_trace_thread_inst_per_ray"] * rays_s / tj["fp32_issue_peak_thread_inst_per_s"],
                             "warp_issue_frac": tj["k_trace_warp_inst_per_ray"] * rays_s / (148 * 4 * 1.965e9),
                             "lanes_per_inst": tj.get("k_trace_lanes_per_inst"), "source": tj.get("source")}
            except Exception:
                traffic = None
        line = {
            "metric": METRIC, "value": rays_all / (total_ms_max * 1e-3) / 1e6, "unit": "Mrays/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "spp_per_step": spp_step, "tiles": f"{TILES[0]}x{TILES[1]}" + (f", each cut into {subdiv}x{subdiv} cells, cell -> rank (ix*k + jx + iy*k + jy) % N" if world > 1 else ""),
                       "l2": "256 MB flush write between timed steps; wave buffers (4 pipelines x 124 MB, 248 MB when a rank has >= 8 waves of 2^20 samples per step) also exceed L2", "wave_pipelines": PIPELINES,
                       "film_reduce": "per step: arn_film_reduce (ncclReduce, sum to rank 0) of the step's films + arn_film_merge into the running film, inside the timed region" if world > 1 else "none (1 GPU)"},
            "spp_per_s": samples_all / (total_ms_max * 1e-3),
            "rays_per_sample": rays_all / max(1.0, samples_all),
            "rays_warmup_step0_rank0": warm_rays[0],     # the step profiles/roofline_traffic.json's ncu capture covers (tools/roofline_from_ncu.py)
            "e2e": {"value": e2e_rays_all / (e2e_ms_max * 1e-3) / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": scene_bytes,
                    "d2h_bytes_per_step": NPIX * 16, "steps": e2e_steps, "ms_per_step": e2e_ms_max / e2e_steps},
            "gpu_launches": int(launches_all),
            "rank_render_ms": {"min": rank_ms_min, "max": rank_ms_max, "note": "mean device time of a step's arn_render_pt_dev per rank, min / max over ranks: the load imbalance of the partition"},
            "roofline": {"kernel": "k_trace (closest hit of path rays + any hit of shadow rays + closest hit of light rays, one launch)", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "peak_source": peak_src, "bytes_per_ray": bytes_per_ray, "rays_per_launch": ext_rays / max(1, n_ext_launch),
                         "trace_mrays_s": ext_rays / (ext_ms * 1e-3) / 1e6 if ext_ms > 0 else 0.0, "trace_share_of_step": ext_ms / serial_ms if serial_ms > 0 else 0.0,
                         "timing": f"CUDA events around every k_trace launch of {timing_steps} steps re-run with ONE wave pipeline (serial launches; {serial_ms / timing_steps:.1f} ms per step); value and e2e run {PIPELINES} concurrent wave pipelines",
                         "incoherent_mrays_s": inc_rays / (inc_ms * 1e-3) / 1e6 if inc_ms > 0 else 0.0, "fp32_issue": issue,
                         "note": "Cornell scene (0.2 MB) is cache resident: the HBM roofline is the contract's denominator, not the binding limit (DESIGN.md)"},
            "clocks": clk,
            "film_check": checks,
        }
        if extra is not None:
            line["extra"] = extra
        if not args.no_cpu_baseline:
            ob = oracle_sample(args.workload, args.cpu_seconds, os.cpu_count() or 1)
            line["cpu_baseline"] = {"value": ob["mrays_s"], "unit": "Mrays/s", "cores": os.cpu_count() or 1, "kind": "port",
                                    "sample": f"{ob['w']}x{ob['h']} px x {ob['spp']} spp of the same workload ({ob['seconds']:.1f} s), oracle/ C++ restatement of arendur, std::thread over 16x16 tiles",
                                    "spp_per_s": ob["spp_s"]}
        print(json.dumps(line))
    if scene is not None:
        scene.close()
    if comm is not None:
        comm.close()
    if dist is not None:
        dist.destroy_process_group()


def film_checks(args, np, torch, api, scene, ctx, cam, film, smp, params, film_acc, film_step, rank, world, spp_step, NPIX, barrier, prm0, ext):
    """Asserts that the film the timed region produced is the right one; raises SystemExit otherwise.
      1. every accumulator is finite;
      2. mean filter-weight sum per pixel = (samples per pixel rendered) x E[w]: the one-sided-Lanczos weight sum of a sample is
         20.23 on average (committed oracle statistic), so a film that counted a rank's samples twice (round 1's bug) is caught;
      3. N > 1: the last step's REDUCED film equals the same sample range rendered by rank 0 alone (world_size = 1) to float
         summation order — partition + ncclReduce == one GPU;
      4. the GPU film of sample 0 has the oracle's committed means (tests/golden/bench_oracle_stats.json) to 1e-5."""
    out, fails = {}, []
    golden = json.load(open(os.path.join(ROOT, "tests", "golden", "bench_oracle_stats.json")))[args.workload]
    total_spp = args.steps * spp_step
    scratch = torch.zeros_like(film_acc)
    if rank == 0:
        f64 = film_acc.double()
        out["finite_frac"] = float(torch.isfinite(film_acc).double().mean().item())
        out["mean_weight"] = float(f64[..., 3].mean().item())
        out["spp_accumulated"] = total_spp
        out["mean_weight_per_sample"] = out["mean_weight"] / total_spp
        out["expected_mean_weight_per_sample"] = golden["mean_rgbw"][3]
        if out["finite_frac"] != 1.0:
            fails.append("non-finite film accumulators")
        if abs(out["mean_weight_per_sample"] / golden["mean_rgbw"][3] - 1.0) > 3e-3:
            fails.append(f"mean weight per sample {out['mean_weight_per_sample']:.4f} != {golden['mean_rgbw'][3]:.4f}: samples were lost or counted twice")
    if world > 1:
        # 3. rank 0 alone renders the last step's sample range; the other ranks wait
        if rank == 0:
            with torch.cuda.stream(ext):
                scene.render_pt_dev(cam, film, smp, params(args.steps - 1, r=0, w=1, sd=0), scratch.data_ptr())
            ctx.synchronize()
            a, b = film_step.double(), scratch.double()
            den = b.abs().mean(dim=(0, 1))
            out["reduced_vs_single_gpu"] = {"rel_err_of_means": [float(v) for v in ((a - b).mean(dim=(0, 1)).abs() / den).tolist()],
                                            "max_abs_pixel_diff_over_mean": float(((a - b).abs().amax() / den.max()).item())}
            if max(out["reduced_vs_single_gpu"]["rel_err_of_means"]) > 1e-5 or out["reduced_vs_single_gpu"]["max_abs_pixel_diff_over_mean"] > 1e-3:
                fails.append("the reduced multi-GPU film of a step differs from the single-GPU render of the same samples")
        barrier()
    if rank == 0:
        scratch.zero_()
        with torch.cuda.stream(ext):
            st = scene.render_pt_dev(cam, film, smp, api.make_pt_params(max_depth=prm0.max_depth, spp_begin=0, spp_end=1), scratch.data_ptr())
        ctx.synchronize()
        m = scratch.double().reshape(-1, 4).mean(0).tolist()
        rel = [abs(g / o - 1.0) for g, o in zip(m, golden["mean_rgbw"])]
        gpu_rays = int(st.extend_rays + st.shadow_rays + st.mis_rays)
        out["sample0_vs_oracle"] = {"gpu_mean_rgbw": m, "oracle_mean_rgbw": golden["mean_rgbw"], "rel_err": rel, "gpu_rays": gpu_rays, "oracle_rays": golden["rays"]}
        if max(rel) > 1e-5 or gpu_rays != golden["rays"]:
            fails.append("the film of sample 0 does not match the committed oracle statistic")
        out["passed"] = not fails
    flag = torch.tensor([1.0 if fails else 0.0], device="cuda")
    if world > 1:
        import torch.distributed as dist
        dist.broadcast(flag, src=0)
    if flag.item() != 0.0:
        if rank == 0:
            sys.stderr.write("bench.py: FILM CHECK FAILED: " + "; ".join(fails) + "\n" + json.dumps(out) + "\n")
        raise SystemExit(3)
    return out


if __name__ == "__main__":
    main()
