"""Per-launch key metrics of an ncu report: python tools/ncu_summary.py <file.ncu-rep> [json_out]
(reads `ncu -i <rep> --page raw --csv`; writes one line per launch + the average DRAM bytes per launch)."""
import csv, io, json, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread", "sm__icc_request_hit_rate.pct"]
idx = {k: hdr.index(k) for k in want if k in hdr}
name_i = hdr.index("Kernel Name")
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
tot = 0.0; n = 0
for r in rows[2:]:
    if len(r) != len(hdr):
        continue
    parts = [r[name_i].split("(")[0].replace("void ", "").replace("arn::", "")]
    for k, i in idx.items():
        parts.append(f"{k.split('.')[0].replace('smsp__', '').replace('sm__', '')}={r[i]}{units[i]}")
    print(" ".join(parts))
    rd = float(r[idx["dram__bytes_read.sum"]].replace(",", "")) * scale.get(units[idx["dram__bytes_read.sum"]], 1.0)
    wr = float(r[idx["dram__bytes_write.sum"]].replace(",", "")) * scale.get(units[idx["dram__bytes_write.sum"]], 1.0)
    tot += rd + wr; n += 1
print(f"avg dram bytes per launch {tot / max(n, 1):.1f} over {n} launches")
if len(sys.argv) > 2:
    json.dump({"k_trace_dram_bytes_per_launch": tot / max(n, 1), "source": f"ncu --set full, {n} launches, {rep}"}, open(sys.argv[2], "w"))
