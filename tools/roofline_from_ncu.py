"""profiles/roofline_traffic.json from an ncu metrics pass over bench.py ITSELF.
  ncu --metrics smsp__thread_inst_executed.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
      --clock-control none -k regex:k_trace -c 576 --csv --log-file trace.csv python bench.py --steps 1 --warmup 3 --no-extra --no-cpu-baseline > bench.json
  python tools/roofline_from_ncu.py trace.csv bench.json profiles/roofline_traffic.json
The first 576 k_trace launches are the 64 waves (2^20 samples each) x 9 trace launches of warm-up step 0 (64 spp of the 1024^2 frame;
1152 launches while the waves were 2^19 samples); its ray
count is bench.py's `rays_warmup_step0_rank0`."""
import csv, json, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
mi, vi, ii, ui = hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID"), hdr.index("Metric Unit")
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1.0, "us": 1e3, "ms": 1e6, "inst": 1.0, "": 1.0}
by = {}
for r in rows[1:]:
    by.setdefault(r[ii], {})[r[mi]] = float(r[vi].replace(",", "")) * scale.get(r[ui], 1.0)
n = len(by)
tot = {k: sum(m[k] for m in by.values()) for k in next(iter(by.values()))}
line = None
for l in open(sys.argv[2]):
    if l.strip().startswith("{"):
        line = json.loads(l)
rays = line["rays_warmup_step0_rank0"]
out = {
    "k_trace_dram_bytes_per_launch": (tot["dram__bytes_read.sum"] + tot["dram__bytes_write.sum"]) / n,
    "k_trace_thread_inst_per_ray": tot["smsp__thread_inst_executed.sum"] / rays,
    "k_trace_warp_inst_per_ray": tot["smsp__inst_executed.sum"] / rays,
    "k_trace_lanes_per_inst": tot["smsp__thread_inst_executed.sum"] / tot["smsp__inst_executed.sum"],
    "k_trace_launches": n, "rays": rays, "k_trace_ms_under_ncu": tot["gpu__time_duration.sum"] / 1e6,
    "fp32_issue_peak_thread_inst_per_s": 36285000000000.0,
    "source": f"ncu --metrics (thread / warp instructions, DRAM bytes, duration) --clock-control none over the first {n} k_trace launches of `python bench.py --steps 1 --warmup 3 --no-extra --no-cpu-baseline` "
              f"= warm-up step 0 of the bench command itself (C3, 64 spp of 1024x1024, {n // 9} waves x 9 launches); issue peak: profiles/r01_fp32_issue.json; written by tools/roofline_from_ncu.py",
}
json.dump(out, open(sys.argv[3], "w"), indent=1)
print(json.dumps(out, indent=1))
