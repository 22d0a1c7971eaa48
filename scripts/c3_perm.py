"""Does the instance -> thread mapping matter for the batched launch?  Natural (column-major grid) order vs strided and
random permutations of the same 65,536 starts.  python scripts/c3_perm.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bazinga_b200 as bz
for name, lo, hi in (("eitheror", -4.0, 8.0), ("mpvca", -5.0, 20.0)):
    v = np.linspace(lo, hi, 256)
    X0 = np.array([(i, j) for j in v for i in v])
    B = len(X0)
    bz.alps_batch(name, X0[:1024])
    base = bz.alps_batch(name, X0)
    perms = {"natural": np.arange(B), "stride 2053": (np.arange(B) * 2053) % B, "stride 32*k+..": np.arange(B).reshape(32, -1).T.reshape(-1),
             "random": np.random.default_rng(0).permutation(B),
             "sorted by work": np.argsort(-base["sub_iters"], kind="stable")}
    for k, p in perms.items():
        t = min(bz.alps_batch(name, X0[p])["kernel_ms"] for _ in range(2))
        o = bz.alps_batch(name, X0[p])
        same = np.array_equal(o["x"], base["x"][p], equal_nan=True)
        print(f"{name:9s} {k:16s} {t:7.2f} ms  identical results: {same}", flush=True)
