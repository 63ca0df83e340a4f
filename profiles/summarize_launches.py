"""Aggregates an ncu launch list (`--metrics gpu__time_duration.sum --csv`) into one step of bench.py --no-graph:
takes the launches between the last two voxelizer starts (vox_clear_kernel), prints a markdown table.
    python profiles/summarize_launches.py profiles/r1_launches_fp32.csv "fp32 step ..."
"""
import csv
import re
import sys
from collections import OrderedDict

path, title = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else sys.argv[1])
rows = []
with open(path) as f:
    lines = [ln for ln in f if not ln.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    val = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "us")
    us = val / 1000.0 if unit in ("ns", "nsecond") else (val * 1000.0 if unit in ("ms", "msecond") else val)
    rows.append((r["Kernel Name"], us))
starts = [i for i, (k, _) in enumerate(rows) if "vox_clear_kernel" in k]
if len(starts) < 2:
    sys.exit("need at least two steps in the launch list")
step = rows[starts[-2]:starts[-1]]
agg = OrderedDict()
for k, us in step:
    name = re.sub(r"^void ", "", k)
    name = re.sub(r"fv2p::\(anonymous namespace\)::", "", name)
    name = re.sub(r"\(.*$", "", name)
    if "fill_" in name.lower() and us > 50:
        name = "(bench L2 flush, untimed) " + name
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += us
total = sum(v[1] for v in agg.values())
print("### %s\n" % title)
print("| kernel | launches | us | share |\n|---|---:|---:|---:|")
for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("| `%s` | %d | %.1f | %.1f%% |" % (name[:90], n, us, 100.0 * us / total))
print("| total | %d | %.1f | |" % (sum(v[0] for v in agg.values()), total))
