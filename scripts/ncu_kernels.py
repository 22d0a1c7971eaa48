"""Per-kernel medians from an `ncu --csv --log-file` launch list with several --metrics (launches that returned at the
state-machine check are dropped: duration < 30 % of the kernel's longest).  python scripts/ncu_kernels.py <file.csv>"""
import csv, statistics as st, sys
from collections import defaultdict
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h, data = rows[hi], rows[hi + 1:]
ix = {n: i for i, n in enumerate(h)}
agg = defaultdict(lambda: defaultdict(dict))
for r in data:
    if len(r) < len(h):
        continue
    v = float(r[ix["Metric Value"]].replace(",", ""))
    u = r[ix["Metric Unit"]]
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0}.get(u, 1.0)
    agg[r[ix["Kernel Name"]][:40]][r[ix["Metric Name"]]][r[ix["ID"]]] = v * scale
for k, d in agg.items():
    t = d["gpu__time_duration.sum"]
    big = [i for i, x in t.items() if x > 0.3 * max(t.values())]
    line = [f"{k:40s} n={len(big):3d}"]
    for m, vals in d.items():
        line.append(f"{m.split('.')[0].replace('__', ':')}={st.median([vals[i] for i in big if i in vals]):.3f}")
    print("  ".join(line))
