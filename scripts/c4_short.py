"""Short C4 run for profiling: obstacle_l1 on a 4096^2 grid, a few inner iterations (python scripts/c4_short.py [side] [inner])."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bazinga_b200 as bz
from bazinga_b200.solver import _options
side = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
inner = int(sys.argv[2]) if len(sys.argv) > 2 else 12
N = side * side
ctx = bz.default_context()
p = bz.Problem(bz.SmoothCostObstacleL1(N), bz.NonsmoothCostObstacleL1(N), bz.ConstraintObstacle(N, (side, side)), bz.ZeroSet(), ctx=ctx)
sub = bz.PANOCplus(directions=bz.LBFGS(5), maxit=1_000_000, minimum_gamma=2.220446049250313e-16)
o = _options(tol=1e-4, tol_prim=None, tol_dual=None, inner_tol=None, maxit=100, theta_penalty=0.8, kappa_penalty=0.5, kappa_tol=0.1,
             subsolver_maxit=1_000_000, sub=sub, max_inner_total=inner)
x, y, s, mu, res = p.alps(2.0 * np.ones(3 * N), np.zeros(N), o)
print(res.tot_inner_it, res.n_grad, [1e3 * res.kind_ms[k] / max(1, res.kind_launches[k]) for k in range(4)])
