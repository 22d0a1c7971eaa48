"""Batched shared-A AL-gradient at the C2 shape for ncu / timing (python scripts/prof_bdense.py)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import numpy as np
import bazinga_b200 as bz
m, n, B = 4096, 16384, 8
h = bz.SharedMatrixBatch(shape=(m, n), gen_seed=0)
rng = np.random.default_rng(0)
X = rng.standard_normal((B, n)); z = np.zeros((B, m))
out = h.gradient(X, z, np.ones((B, m)), z, reps=20)
byt = 8 * m * n
print({"us_per_batched_gradient": out["us"], "instances": B, "us_per_instance": out["us"] / B,
       "A_bytes_streamed_GBs": 2 * byt / (out["us"] * 1e-6) / 1e9, "algorithmic_GBs_single_pass": byt / (out["us"] * 1e-6) / 1e9,
       "dmma_tflops": 2 * 2.0 * m * n * B / (out["us"] * 1e-6) / 1e12})
h.close()
