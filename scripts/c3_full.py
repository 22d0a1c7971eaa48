"""The whole 65,536-start demo grid in one batched launch (for ncu): python scripts/c3_full.py eitheror"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bazinga_b200 as bz
name = sys.argv[1] if len(sys.argv) > 1 else "eitheror"
lo, hi = (-4.0, 8.0) if name == "eitheror" else (-5.0, 20.0)
v = np.linspace(lo, hi, 256)
X0 = np.array([(i, j) for j in v for i in v])
o = bz.alps_batch(name, X0)
print(name, "kernel_ms", o["kernel_ms"], "sub_iters", int(o["sub_iters"].sum()))
