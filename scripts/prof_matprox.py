"""A batch of the distance-matrix-completion demo's prox calls for ncu / timing: 4096 nuclear-norm prox evaluations of 20 x 20
matrices (demo/distance_matrix_completion.jl:163: N = 20).  python scripts/prof_matprox.py [N] [batch]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import numpy as np
import bazinga_b200 as bz
import bazinga_oracle as bo
N = int(sys.argv[1]) if len(sys.argv) > 1 else 20
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
rng = np.random.default_rng(0)
X = rng.standard_normal((B, N, N))
f = bz.NuclearNorm(1.0)
f.prox(X[:8], 0.3)
t = time.perf_counter(); Y, gz = f.prox(X, 0.3); dt = time.perf_counter() - t
o = bo.GNuclearNorm(1.0); Yr = np.empty((N, N))
t = time.perf_counter()
for b in range(min(B, 512)): o.prox(Yr, X[b], 0.3)
dc = (time.perf_counter() - t) / min(B, 512)
print({"matrices": B, "N": N, "device_call_ms_with_copies": 1e3 * dt, "prox_per_s_e2e": B / dt, "numpy_lapack_prox_per_s_1core": 1 / dc,
       "max_err_last": float(np.max(np.abs(Y[min(B, 512) - 1] - Yr)))})
