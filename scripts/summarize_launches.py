"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: time share per kernel name.
    python scripts/summarize_launches.py gpurun_out/launches.csv > profiles/r01_ncu_launch_shares.txt"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hi]
iname, ival, iunit = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= ival:
        continue
    v = float(r[ival].replace(",", ""))
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[iunit], 1.0)
    name = re.sub(r"\(.*", "", r[iname]).replace("void ", "").replace("wj::", "")
    c, t = agg.get(name, (0, 0.0))
    agg[name] = (c + 1, t + v)
tot = sum(t for _, t in agg.values())
print(f"{sum(c for c, _ in agg.values())} launches, {tot / 1e3:.2f} ms serialized (cold-cache, per-launch ncu replay: compare SHARES)")
for name, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{100 * t / tot:6.2f} %  {t / 1e3:9.3f} ms  {c:5d} x  {name}")
