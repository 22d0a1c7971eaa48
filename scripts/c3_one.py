"""One 65,536-instance batched launch of a demo problem (for ncu): python scripts/c3_one.py mpvca|eitheror [reps]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bazinga_b200 as bz
name = sys.argv[1] if len(sys.argv) > 1 else "mpvca"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
lo, hi = (-4.0, 8.0) if name == "eitheror" else (-5.0, 20.0)
v = np.linspace(lo, hi, 256)
X0 = np.array([(i, j) for j in v for i in v])
for _ in range(reps):
    out = bz.alps_batch(name, X0)
    print(name, "kernel_ms", out["kernel_ms"], "max sub_iters", int(out["sub_iters"].max()), "argmax", int(out["sub_iters"].argmax()))
