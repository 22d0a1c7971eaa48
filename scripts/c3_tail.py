"""Which instance bounds a 65,536-instance batched launch?  Times the kernel on the whole demo grid, then on each of
the candidate tail instances alone (the most inner iterations, the unsolved ones): python scripts/c3_tail.py [name] [k]
With a third argument `only`, runs just the k-th candidate alone (for ncu)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bazinga_b200 as bz
name = sys.argv[1] if len(sys.argv) > 1 else "mpvca"
lo, hi = (-4.0, 8.0) if name == "eitheror" else (-5.0, 20.0)
v = np.linspace(lo, hi, 256)
X0 = np.array([(i, j) for j in v for i in v])
out = bz.alps_batch(name, X0)
out = bz.alps_batch(name, X0)
print(name, "all: kernel_ms", round(out["kernel_ms"], 2))
cand = list(np.argsort(-out["sub_iters"])[:6]) + list(np.flatnonzero(out["status"] != 0)[:4])
cand += list(np.flatnonzero(~np.isfinite(out["x"]).all(axis=1))[:4])
seen = []
for i in cand:
    if i in seen: continue
    seen.append(i)
    o1 = bz.alps_batch(name, X0[[i]])
    print(f"  instance {i} x0={X0[i]} alone: kernel_ms {o1['kernel_ms']:.2f} sub_iters {int(o1['sub_iters'][0])} "
          f"outer {int(o1['iters'][0])} status {int(o1['status'][0])} x {o1['x'][0]}", flush=True)
# time by blocks of 2048 instances to see how the tail is distributed
t = []
for b in range(0, len(X0), 8192):
    t.append(round(bz.alps_batch(name, X0[b:b + 8192])["kernel_ms"], 1))
print("  per 8192-block kernel_ms", t)
