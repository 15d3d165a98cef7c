"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: one PPO iteration (between two GAE launches)."""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
data = [(r[ki], float(r[vi].replace(",", ""))) for r in rows[1:] if r[vi].replace(",", "").replace(".", "").isdigit()]
gae = [i for i, (n, _) in enumerate(data) if "gae" in n]
seg = data[gae[-2] + 1:gae[-1] + 1] if len(gae) > 1 else data
agg = collections.OrderedDict()
for n, t in seg:
    m = re.search(r"(vine_\w+|at::\w+)", n)
    a = agg.setdefault(m.group(1) if m else n[:50], [0, 0.0])
    a[0] += 1
    a[1] += t
tot = sum(v[1] for v in agg.values())
print(f"launches {len(seg)}  total {tot / 1e3:.1f} us")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:24]:
    print(f"  {k:45s} n={v[0]:4d} total={v[1] / 1e3:8.1f} us avg={v[1] / v[0] / 1e3:7.2f} {100 * v[1] / tot:5.1f}%")
