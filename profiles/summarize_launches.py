"""Per-kernel totals from an ncu `--metrics gpu__time_duration.sum --csv` launch list.
usage: summarize_launches.py launches.csv [start-after-last-kernel-substring]"""
import collections, csv, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
rows = list(csv.DictReader(lines))
if len(sys.argv) > 2:
    idx = [i for i, x in enumerate(rows) if sys.argv[2] in x["Kernel Name"]]
    rows = rows[idx[-1]:]
agg = collections.OrderedDict()
for x in rows:
    v = float(x["Metric Value"].replace(",", ""))
    v = v / 1000 if x["Metric Unit"] == "ns" else v * 1000 if x["Metric Unit"] == "ms" else v
    a = agg.setdefault(x["Kernel Name"][:70], [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(v[1] for v in agg.values())
print("kernel,launches,total_us,share")
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f'"{k}",{c},{t:.1f},{t / tot:.3f}')
