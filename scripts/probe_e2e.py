"""Where does the public bz.alps(...) call spend its time on the C2 instance?  (create / upload / solve / download / destroy)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import numpy as np
import torch
import bazinga_b200 as bz
import bazinga_oracle as bo
from bazinga_b200.solver import _options

m, n = 4096, 16384
ctx = bz.Context(0)
A_t = torch.empty((m, n), dtype=torch.float64, pin_memory=True)
A = A_t.numpy()
for r in range(0, m, 1024):
    A[r:r + 1024] = ctx.randn_fill(0, r * n, 1024 * n).reshape(1024, n)
b_t = torch.empty(m, dtype=torch.float64, pin_memory=True)
b = b_t.numpy()
b[:] = A @ bo.bz_sparse_randn(1, n, 0.1)
sub = bz.PANOCplus(directions=bz.LBFGS(5), maxit=100_000, minimum_gamma=2.220446049250313e-16)
o = _options(tol=1e-6, tol_prim=None, tol_dual=None, inner_tol=None, maxit=100, theta_penalty=0.8, kappa_penalty=0.5, kappa_tol=0.1,
             subsolver_maxit=100_000, sub=sub, max_inner_total=1000)
c = bz.ConstraintBasisPursuit(A, b)
x0, y0 = np.zeros(n), np.zeros(m)
for rep in range(3):
    t0 = time.perf_counter()
    p = bz.Problem(bz.Zero(), bz.NormL1(1.0), c, bz.ZeroSet(), ctx=ctx)
    t1 = time.perf_counter()
    out = p.alps(x0, y0, o)
    t2 = time.perf_counter()
    p.close()
    t3 = time.perf_counter()
    p2 = bz.Problem(bz.Zero(), bz.NormL1(1.0), c, bz.ZeroSet(), ctx=ctx)
    t4 = time.perf_counter()
    p2.upload_dense(A, b)
    t5 = time.perf_counter()
    p2.close()
    print(f"rep {rep}: create {1e3*(t1-t0):.1f} ms, solve {1e3*(t2-t1):.1f} ms ({out[4].tot_inner_it} its), destroy {1e3*(t3-t2):.1f} ms, "
          f"create again {1e3*(t4-t3):.1f} ms, upload_dense {1e3*(t5-t4):.1f} ms", flush=True)
A2 = np.array(A)   # pageable copy
t0 = time.perf_counter(); p = bz.Problem(bz.Zero(), bz.NormL1(1.0), bz.ConstraintBasisPursuit(A2, b), bz.ZeroSet(), ctx=ctx); t1 = time.perf_counter()
print(f"create from pageable memory {1e3*(t1-t0):.1f} ms")
p.close()
