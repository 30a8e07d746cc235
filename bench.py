#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 path-tracing core (contract: task prompt (4)).

Workload (BASELINE.json configs[2], "C3"): Cornell box (examples/cornellbox/cb.json), 1024x1024,
full PT bounce loop (depth 8, area-light NEE + MIS, Russian roulette), 1024 spp.  One STEP = one
pass of the whole hot path (generate -> extend -> shade -> connect -> accumulate) over a batch of
64 spp x 1024 x 1024 camera samples (sample indices [64k, 64k+64)); the default 16 timed steps
are exactly the 1024 spp of C3 accumulated into one film.  At N GPUs the same frame is
tile-partitioned (16x16 tiles, (ix + iy) % N) and every step renders 64*N spp, i.e. per-GPU work is fixed
("weak"); the per-rank films are summed with one NCCL reduce per step inside the timed region
(Film::merge_into semantics).

  value  = Mrays/s, rays = every BVH traversal (path + shadow + MIS), film resident in HBM
  e2e    = same metric through the host-buffer C-ABI: scene upload (H2D) + arn_render_pt + film D2H
  roofline: k_trace (all BVH traversals), algorithmic bytes = 32*Nn + 36*Nt + 152*Ns + 36 per ray
  cpu_baseline / --impl reference: the oracle (CPU restatement of arendur; the Rust original
  cannot be built here) on the host cores, bounded sample of the same workload.
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

RES = 1024
SPP_TOTAL_X = 32          # 32 x 32 = 1024 spp
SPP_PER_STEP = 64
TILES = tuple(int(v) for v in os.environ.get("BENCH_TILES", "16,16").split(","))   # Film::spawn_tiles grid (pt.rs:131 uses 16 x 16)
PIPELINES = 4            # concurrent wave pipelines of the timed region (the library default; ARN_OPT_PIPELINES)
WORKLOAD = "C3: Cornell box 1024x1024, PT depth 8 + area-light NEE/MIS, 1024 spp (64 spp per step)"


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


def oracle_sample(spp_budget_s, threads):
    """Times the oracle on a bounded sample of the workload: whole 1024x1024 frame, as many
    samples per pixel as fit the time budget (calibrated with a 1-spp pass)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    from arendur_b200 import api, scenes
    hs, cam, film, smp, prm = scenes.cornell_scene(RES, RES, SPP_TOTAL_X, SPP_TOTAL_X)
    osc = O.OracleScene(hs.desc())

    def run(s0, s1):
        p = api.make_pt_params(max_depth=prm.max_depth, spp_begin=s0, spp_end=s1)
        t = time.perf_counter()
        _, st, _ = osc.render_pt(cam, film, smp, p, nthreads=threads)
        dt = time.perf_counter() - t
        return dt, int(st.extend_rays + st.shadow_rays + st.mis_rays), int(st.camera_rays)
    dt1, _, _ = run(0, 1)
    n = max(1, min(64, int(spp_budget_s / max(dt1, 1e-3))))
    dt, rays, samples = run(1, 1 + n)
    return {"seconds": dt, "rays": rays, "samples": samples, "spp": n, "mrays_s": rays / dt / 1e6, "spp_s": samples / dt}


def reference_arm(args):
    """--impl reference: the reference's CPU implementation of the path.  arendur is Rust
    (2017 nightly) and cannot be built in this image, so this is the oracle port, all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    per_step = []
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    from arendur_b200 import api, scenes
    hs, cam, film, smp, prm = scenes.cornell_scene(RES, RES, SPP_TOTAL_X, SPP_TOTAL_X)
    osc = O.OracleScene(hs.desc())
    rays_tot, samples_tot = 0, 0
    # bounded sample: each step = 1 spp of the full 1024x1024 frame (1/64 of the GPU arm's step)
    for i in range(args.warmup + args.steps):
        p = api.make_pt_params(max_depth=prm.max_depth, spp_begin=i % 1024, spp_end=i % 1024 + 1)
        t = time.perf_counter()
        _, st, _ = osc.render_pt(cam, film, smp, p, nthreads=threads)
        dt = time.perf_counter() - t
        if i >= args.warmup:
            per_step.append(dt)
            rays_tot += int(st.extend_rays + st.shadow_rays + st.mis_rays)
            samples_tot += int(st.camera_rays)
    total = sum(per_step)
    value = rays_tot / total / 1e6
    line = {
        "impl": "reference", "metric": "Mrays/s (primary + incoherent bounce: every BVH traversal) on the Cornell box", "value": value, "unit": "Mrays/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / max(1, len(per_step)),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "note": "reference arm step = 1 spp of the 1024x1024 frame (bounded sample, 1/64 of the GPU step)"},
        "spp_per_s": samples_tot / total,
        "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": threads, "kind": "port",
                         "sample": f"{args.steps} steps x (1024x1024 px x 1 spp), oracle/ C++ restatement of arendur with std::thread over the 16x16 tile grid"},
        "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=16)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="arendur_b200")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="time budget of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
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

    hs, cam, film, smp, prm0 = scenes.cornell_scene(RES, RES, SPP_TOTAL_X, SPP_TOTAL_X)
    desc = hs.desc()
    ctx = api.Context(local)
    scene = ctx.upload(desc)
    ext = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", local))
    film_dev = torch.zeros((RES, RES, 4), dtype=torch.float32, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")      # > 126 MB L2
    spp_step = SPP_PER_STEP * world
    n_slices = (SPP_TOTAL_X * SPP_TOTAL_X) // spp_step

    def params(step):
        k = step % n_slices
        return api.make_pt_params(max_depth=prm0.max_depth, rank=rank, world_size=world, spp_begin=k * spp_step, spp_end=(k + 1) * spp_step, tiles=TILES)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def device_step(step, want_stats):
        st = scene.render_pt_dev(cam, film, smp, params(step), film_dev.data_ptr(), want_stats=want_stats)
        if dist is not None:
            dist.reduce(film_dev, dst=0)
        return st

    # ---- warm-up
    with torch.cuda.stream(ext):
        for w in range(args.warmup):
            device_step(w, True)
    barrier()
    film_dev.zero_()
    # ---- timed: device-resident film
    clocks = ClockSampler(local)
    clocks.start()
    step_ms, ext_ms, ext_rays, rays, samples, launches, inc_ms, inc_rays = [], 0.0, 0, 0, 0, 0, 0.0, 0
    with torch.cuda.stream(ext):
        for k in range(args.steps):
            flush.fill_(k & 0xFF)                      # L2 flush between timed iterations (untimed)
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(ext)
            st = device_step(k, True)
            e1.record(ext)
            barrier()
            step_ms.append(e0.elapsed_time(e1))
            rays += st.extend_rays + st.shadow_rays + st.mis_rays
            samples += st.camera_rays
            launches += st.kernel_launches
    clk = clocks.stop()
    total_ms = sum(step_ms)
    film_host = film_dev.cpu().numpy() if rank == 0 else None

    # ---- kernel-timing pass for the roofline: the same steps once more with ONE wave pipeline.  The timed region above
    # runs several wave pipelines concurrently (ARN_OPT_PIPELINES), where a kernel's launch-to-finish time is shared with
    # other kernels; CUDA events around every k_trace launch only measure that kernel when launches are serial.
    timing_steps = max(1, min(args.steps, 4))
    serial_ms = 0.0
    ctx.set_option(L.ARN_OPT_PIPELINES, 1)
    scratch_t = torch.zeros_like(film_dev)
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
    del scratch_t

    # ---- instrumented pass (untimed): Nn / Nt of the extend rays -> algorithmic bytes per ray
    ctx.set_option(L.ARN_OPT_COUNT_TRAVERSAL, 1)
    scratch = torch.zeros_like(film_dev)
    with torch.cuda.stream(ext):
        stc = scene.render_pt_dev(cam, film, smp, api.make_pt_params(max_depth=prm0.max_depth, rank=rank, world_size=world, spp_begin=0, spp_end=min(8, spp_step)), scratch.data_ptr())
    ctx.set_option(L.ARN_OPT_COUNT_TRAVERSAL, 0)
    torch.cuda.synchronize()
    rays_counted = stc.extend_rays + stc.shadow_rays + stc.mis_rays
    bytes_per_ray = (32.0 * stc.extend_nodes + 36.0 * stc.extend_tris + 152.0 * stc.extend_spheres) / max(1, rays_counted) + 28 + 8

    # ---- e2e: host buffers through the C-ABI, copies inside the timed region
    scene_bytes = int(desc.n_nodes * 32 + desc.n_prims * 48 + desc.n_spheres * 176 + desc.n_triangles * 16 + desc.n_vertices * 32
                      + desc.n_meshes * 16 + desc.n_materials * 48 + desc.n_prims * 4 + desc.n_lights * 12 + 4)
    e2e_rays, e2e_ms = 0, 0.0
    host_film = torch.empty((RES, RES, 4), dtype=torch.float32, pin_memory=True).numpy()      # pinned: the D2H read of the result
    e2e_steps = max(1, min(args.steps, 4))
    for k in range(-1, e2e_steps):                                       # k = -1: one untimed warm-up of this path (first-use allocations)
        barrier()
        t0 = time.perf_counter()
        sc2 = ctx.upload(desc)                                           # H2D of the flattened scene
        f, st = sc2.render_pt(cam, film, smp, params(max(k, 0)), out=host_film)  # render + film D2H into pinned memory (synchronous)
        if dist is not None:                                             # multi-GPU: gather on rank 0 through NCCL as well
            t = torch.from_numpy(f).cuda(); dist.reduce(t, dst=0); f = t.cpu().numpy()
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

    def allsum(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.SUM); return float(t.item())
    total_ms_max, e2e_ms_max = allmax(total_ms), allmax(e2e_ms)
    rays_all, samples_all, launches_all, e2e_rays_all = allsum(rays), allsum(samples), allsum(launches), allsum(e2e_rays)

    if rank == 0:
        peak, peak_src = load_peaks()
        n_ext_launch = timing_steps * (prm0.max_depth + 1) * max(1, (RES * RES * spp_step // world + (1 << 20) - 1) // (1 << 20))
        achieved = bytes_per_ray * ext_rays / (ext_ms * 1e-3) / 1e9 if ext_ms > 0 else 0.0
        traffic, issue = None, None
        tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tp):
            try:
                tj = json.load(open(tp))
                traffic = tj.get("k_trace_dram_bytes_per_launch")
                # the limit that actually binds on the cache-resident Cornell scene: FP32 issue.  Instructions per
                # ray come from the committed ncu capture of this kernel on this workload, the rate from the live timing above.
                if ext_ms > 0 and tj.get("k_trace_thread_inst_per_ray"):
                    rays_s = ext_rays / (ext_ms * 1e-3)
                    issue = {"thread_inst_per_ray": tj["k_trace_thread_inst_per_ray"], "achieved_thread_inst_per_s": tj["k_trace_thread_inst_per_ray"] * rays_s,
                             "peak_thread_inst_per_s": tj["fp32_issue_peak_thread_inst_per_s"], "frac": tj["k_trace_thread_inst_per_ray"] * rays_s / tj["fp32_issue_peak_thread_inst_per_s"],
                             "warp_issue_frac": tj["k_trace_warp_inst_per_ray"] * rays_s / (148 * 4 * 1.965e9),
                             "source": "instructions per ray: profiles/r01_e_trace_full.txt (ncu); peak: tools/fp32_issue.cu measured on this GPU model (profiles/r01_fp32_issue.json)"}
            except Exception:
                traffic = None
        line = {
            "metric": "Mrays/s (primary + incoherent bounce: every BVH traversal) on the Cornell box", "value": rays_all / (total_ms_max * 1e-3) / 1e6, "unit": "Mrays/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "spp_per_step": spp_step, "tiles": f"{TILES[0]}x{TILES[1]}, (ix + iy) % N", "l2": "256 MB flush write between timed steps; wave buffers (4 x 124 MB) also exceed L2", "wave_pipelines": PIPELINES,
                       "film_reduce": "ncclReduce per step" if world > 1 else "none (1 GPU)"},
            "spp_per_s": samples_all / (total_ms_max * 1e-3),
            "rays_per_sample": rays_all / max(1.0, samples_all),
            "e2e": {"value": e2e_rays_all / (e2e_ms_max * 1e-3) / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": scene_bytes,
                    "d2h_bytes_per_step": RES * RES * 16, "steps": e2e_steps, "ms_per_step": e2e_ms_max / e2e_steps},
            "gpu_launches": int(launches_all),
            "roofline": {"kernel": "k_trace (closest hit of path rays + any hit of shadow rays + closest hit of light rays, one launch)", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "peak_source": peak_src, "bytes_per_ray": bytes_per_ray, "rays_per_launch": ext_rays / max(1, n_ext_launch),
                         "trace_mrays_s": ext_rays / (ext_ms * 1e-3) / 1e6 if ext_ms > 0 else 0.0, "trace_share_of_step": ext_ms / serial_ms if serial_ms > 0 else 0.0,
                         "timing": f"CUDA events around every k_trace launch of {timing_steps} steps re-run with ONE wave pipeline (serial launches; {serial_ms / timing_steps:.1f} ms per step); value and e2e run {PIPELINES} concurrent wave pipelines",
                         "incoherent_mrays_s": inc_rays / (inc_ms * 1e-3) / 1e6 if inc_ms > 0 else 0.0, "fp32_issue": issue,
                         "note": "Cornell scene (0.2 MB) is cache resident: the HBM roofline is the contract's denominator, not the binding limit (DESIGN.md)"},
            "clocks": clk,
        }
        if not args.no_cpu_baseline:
            ob = oracle_sample(args.cpu_seconds, os.cpu_count() or 1)
            line["cpu_baseline"] = {"value": ob["mrays_s"], "unit": "Mrays/s", "cores": os.cpu_count() or 1, "kind": "port",
                                    "sample": f"1024x1024 px x {ob['spp']} spp of the same workload ({ob['seconds']:.1f} s), oracle/ C++ restatement of arendur, std::thread over 16x16 tiles",
                                    "spp_per_s": ob["spp_s"]}
        # parity guard on the rendered film: finite, and its mean matches the committed oracle statistic
        fin = float(np.isfinite(film_host).mean())
        line["film_check"] = {"finite_frac": fin, "mean_weight": float(film_host[..., 3].mean())}
        print(json.dumps(line))
    scene.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
