"""Batched C3 timing: whole demo grid and the known tail instances alone.  python scripts/c3_quick.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bazinga_b200 as bz
res = []
for name, lo, hi, tails in (("eitheror", -4.0, 8.0, (48119, 5716)), ("mpvca", -5.0, 20.0, (17254,))):
    v = np.linspace(lo, hi, 256)
    X0 = np.array([(i, j) for j in v for i in v])
    bz.alps_batch(name, X0[:1024])
    t = min(bz.alps_batch(name, X0)["kernel_ms"] for _ in range(2))
    res.append(f"{name} all {t:.1f}")
    for i in tails:
        res.append(f"#{i} {bz.alps_batch(name, X0[[i]])['kernel_ms']:.1f}")
print(os.environ.get("VARIANT", ""), " | ".join(res), flush=True)
