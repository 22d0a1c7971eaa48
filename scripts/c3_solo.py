"""One instance of a demo problem alone in a batched launch (for ncu): python scripts/c3_solo.py mpvca 17254"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bazinga_b200 as bz
name, idx = sys.argv[1], int(sys.argv[2])
lo, hi = (-4.0, 8.0) if name == "eitheror" else (-5.0, 20.0)
v = np.linspace(lo, hi, 256)
X0 = np.array([(i, j) for j in v for i in v])
o = bz.alps_batch(name, X0[[idx]])
print(name, idx, "kernel_ms", o["kernel_ms"], "sub_iters", int(o["sub_iters"][0]))
