"""Portfolio demo (demo/portfolio.jl) on the device vs the numpy restatement: truncated runs (iterate parity) and the
five full solves.  python scripts/portfolio_probe.py [n]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import bazinga_b200 as bz
import bazinga_oracle as bo
from bazinga_b200 import portfolio as pf
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
Q, rho, mu, ub = bo.make_portfolio(n, seed=3)
f, c, D = bz.SmoothCostP(Q), bz.ConstraintP(mu), bz.SetP(rho)
for inner in (3, 12, 40):
    sub = lambda **kw: bz.PANOCplus(directions=bz.LBFGS(5), maxit=inner, minimum_gamma=1e-32, **kw)
    for gname, g in (("l0", bz.NormL0Box(100.0, u=ub)), ("lp", bz.NormLpPowerBox(0.5, 100.0, u=ub)), ("l1", bz.NormL1Box(100.0, u=ub))):
        out = bz.als(f, g, c, D, np.ones(n), np.zeros(2), maxit=2, subsolver=sub)
        P = bo.problem_portfolio(Q, rho, mu, ub, gname)
        with np.errstate(all="ignore"):
            r = bo.als(P["f"], P["g"], P["c"], P["D"], np.ones(n), np.zeros(2), maxit=2, subsolver=bo.PanocOptions(minimum_gamma=1e-32, maxit=inner))
        print(f"inner {inner:3d} {gname}: it {out[2]}/{r.tot_it} inner {out[3]}/{r.tot_inner_it} status {out[5]}/{r.status} "
              f"dx {np.max(np.abs(out[0]-r.x)):.2e} dy_rel {np.max(np.abs(out[1]-r.y))/(1+np.max(np.abs(r.y))):.2e} "
              f"ds {np.max(np.abs(out[8]-r.s)):.2e} dmu_rel {np.max(np.abs(out[9]-r.mu)/r.mu):.2e} nnz {(out[0]>0).sum()}/{(r.x>0).sum()}", flush=True)
t = time.time()
rows, outs = pf.solve_portfolio(Q, rho, mu, ub)
print("five solves on the device:", round(time.time() - t, 3), "s")
for k in pf.KEYS:
    print(" ", k, rows[k][1:], outs[k][5])
t = time.time()
with np.errstate(all="ignore"):
    ref = {}
    for key, g, start in (("l0", "l0", None), ("lp", "lp", None), ("lp0", "l0", "lp"), ("l1", "l1", None), ("l10", "l0", "l1")):
        P = bo.problem_portfolio(Q, rho, mu, ub, g)
        x0, y0 = (np.ones(n), np.zeros(2)) if start is None else (ref[start].x, ref[start].y)
        ref[key] = bo.als(P["f"], P["g"], P["c"], P["D"], x0, y0, **P["kw"])
print("five solves, numpy restatement:", round(time.time() - t, 3), "s")
for k in pf.KEYS:
    r = ref[k]
    print(" ", k, (r.tot_it, r.tot_inner_it, r.norm_res_prim, 0.5 * r.x @ Q @ r.x, int((r.x > 0).sum())), r.status)
