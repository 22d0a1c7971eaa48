"""A short C4 run for ncu: obstacle_l1 on a 4096 x 4096 grid, a few PANOC+ iterations (python scripts/prof_c4.py [side] [iters])"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import numpy as np
import bazinga_b200 as bz
from bazinga_b200.solver import _options
side = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 12
N = side * side
ctx = bz.default_context()
p = bz.Problem(bz.SmoothCostObstacleL1(N), bz.NonsmoothCostObstacleL1(N), bz.ConstraintObstacle(N, (side, side)), bz.ZeroSet(), ctx=ctx)
sub = bz.PANOCplus(directions=bz.LBFGS(5), maxit=1_000_000, minimum_gamma=2.220446049250313e-16)
o = _options(tol=1e-4, tol_prim=None, tol_dual=None, inner_tol=None, maxit=100, theta_penalty=0.8, kappa_penalty=0.5, kappa_tol=0.1,
             subsolver_maxit=1_000_000, sub=sub, max_inner_total=iters)
x, y, s, mu, res = p.alps(2.0 * np.ones(3 * N), np.zeros(N), o)
print("iters", res.tot_inner_it, "n_grad", res.n_grad, "device_ms", res.device_ms,
      {k: (1e3 * res.kind_ms[i] / res.kind_launches[i] if res.kind_launches[i] else None) for i, k in enumerate(["G", "L", "X/F", "D"])})
p.close()
