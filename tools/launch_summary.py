"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches,
total time and share.  usage: launch_summary.py file.csv [last_n_launches]"""
import collections
import csv
import sys

path = sys.argv[1]
last = int(sys.argv[2]) if len(sys.argv) > 2 else 0
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
rows = []
for row in csv.DictReader(lines):
    try:
        v = float(row["Metric Value"].replace(",", ""))
    except (ValueError, KeyError):
        continue
    u = row["Metric Unit"]
    v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
    name = row["Kernel Name"].split("(")[0].replace("zkodst::<unnamed>::", "").replace("void ", "")
    if "imad_kernel" in name or "fieldmul_kernel" in name or "madd_kernel" in name:
        continue  # the integer-pipe microbenchmark bench.py runs after the timed region
    rows.append((name, v))
if last:
    rows = rows[-last:]
agg = collections.defaultdict(lambda: [0, 0.0])
for name, v in rows:
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
print("launches %d, total %.1f us" % (len(rows), tot))
for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
    print("%-56s %6d %12.1f us %6.2f%%" % (k[:56], v[0], v[1], 100 * v[1] / tot))
