"""Scheduling experiments on the batched launch: python scripts/c3_order.py [eitheror|mpvca]
   the whole grid in launch order, sorted by work (longest first / last), random, and without its 200 longest instances"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bazinga_b200 as bz
name = sys.argv[1] if len(sys.argv) > 1 else "eitheror"
lo, hi = (-4.0, 8.0) if name == "eitheror" else (-5.0, 20.0)
v = np.linspace(lo, hi, 256)
X0 = np.array([(i, j) for j in v for i in v])
out = bz.alps_batch(name, X0)
t = lambda X: min(bz.alps_batch(name, X)["kernel_ms"] for _ in range(2))
its = out["sub_iters"]
order = np.argsort(-its)
rng = np.random.default_rng(0)
print(name, os.environ.get("VARIANT", ""), "launch order %.1f | longest first %.1f | longest last %.1f | random %.1f | without top 200: %.1f | without top 2000: %.1f | top 200 alone %.1f | top 2000 alone %.1f | first 32768 %.1f"
      % (t(X0), t(X0[order]), t(X0[order[::-1]]), t(X0[rng.permutation(len(X0))]), t(X0[np.sort(order[200:])]), t(X0[np.sort(order[2000:])]),
         t(X0[order[:200]]), t(X0[order[:2000]]), t(X0[:32768])), flush=True)
