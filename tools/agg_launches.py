"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, agg = None, collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    if 'Kernel Name' in r: hdr = r; continue
    if hdr is None or len(r) != len(hdr): continue
    d = dict(zip(hdr, r))
    val = float(d['Metric Value'].replace(',', '')); unit = d['Metric Unit']
    val = {'us': val / 1e3, 'ns': val / 1e6, 's': val * 1e3}.get(unit, val)
    name = re.sub(r'\(.*', '', d['Kernel Name'])
    agg[name][0] += 1; agg[name][1] += val
tot = sum(v[1] for v in agg.values())
print(f"{'kernel':40s} {'launches':>8s} {'ms':>10s} {'share':>7s}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:40s} {v[0]:8d} {v[1]:10.3f} {100*v[1]/tot:6.1f}%")
print(f"{'total':40s} {sum(v[0] for v in agg.values()):8d} {tot:10.3f}")
