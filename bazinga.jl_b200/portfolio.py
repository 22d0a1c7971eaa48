"""Host-side mirror of the reference's portfolio driver (demo/portfolio.jl:71-260): read one problem folder
(Q.csv, rho.csv, mu.csv, ub.csv, no headers), run the demo's five `als` solves -- L0, Lp, Lp then L0, L1, L1 then L0 --
on the device, and write the per-key tables `<name>_<key>.csv` with the columns of push_out_to_data.  The solves go
through bz.als (bz_als_solve); this file only arranges them."""
import csv
import os

import numpy as np

from . import oracles as O
from .solver import LBFGS, PANOCplus, als

KEYS = ("l0", "l1", "l10", "lp", "lp0")                                             # demo/portfolio.jl:136
COLUMNS = ("id", "iters", "subiters", "runtime", "cviolation", "objectiveq", "nnz")  # :92-108


def read_FG_problem_data(foldername):
    """demo/portfolio.jl:71-90: returns (Q, rho, mu, ub) and checks the sizes like the reference's @asserts."""
    load = lambda name: np.atleast_2d(np.loadtxt(os.path.join(foldername, name), delimiter=",", dtype=np.float64))  # noqa: E731
    Q = load("Q.csv")
    rho = load("rho.csv").reshape(-1)
    mu = load("mu.csv").reshape(-1)
    ub = load("ub.csv").reshape(-1)
    nx = mu.shape[0]
    if rho.shape[0] != 1 or ub.shape[0] != nx or Q.shape != (nx, nx):
        raise ValueError("portfolio data: expected Q (n x n), rho (scalar), mu (n), ub (n)")
    return Q, float(rho[0]), mu, ub


def _row(pid, out, Q):
    x = out[0]
    return (pid, int(out[2]), int(out[3]), float(out[4]), float(out[7]), 0.5 * float(x @ (Q @ x)), int(np.sum(x > 0)))


def solve_portfolio(Q, rho, mu, ub, *, pid=1, data_a=100.0, pnorm=0.5, ctx=None, **kwargs):
    """The body of the demo's loop over problems (demo/portfolio.jl:140-203) for one data set: dict key -> table row.
    Defaults as in the demo: a = 100, p = 0.5, x0 = ones(n), y0 = zeros(2), PANOCplus(LBFGS(5), maxit = 1000,
    minimum_gamma = 1e-32)."""
    n = mu.shape[0]
    f, c, D = O.SmoothCostP(Q), O.ConstraintP(mu), O.SetP(rho)
    g = {"l1": O.NormL1Box(data_a, u=ub), "lp": O.NormLpPowerBox(pnorm, data_a, u=ub), "l0": O.NormL0Box(data_a, u=ub)}
    subsolver = lambda **kw: PANOCplus(directions=LBFGS(5), maxit=1000, freq=1000, minimum_gamma=1e-32, **kw)  # noqa: E731
    solver = lambda gk, x0, y0: als(f, g[gk], c, D, x0, y0, subsolver=subsolver, ctx=ctx, **kwargs)  # noqa: E731
    X0, y0 = np.ones(n), np.zeros(2)
    rows, outs = {}, {}
    outs["l0"] = solver("l0", X0, y0)
    outs["lp"] = solver("lp", X0, y0)
    outs["lp0"] = solver("l0", outs["lp"][0], outs["lp"][1])      # warm start from the Lp solution (:170-173)
    outs["l1"] = solver("l1", X0, y0)
    outs["l10"] = solver("l0", outs["l1"][0], outs["l1"][1])
    for k in KEYS:
        rows[k] = _row(pid, outs[k], Q)
    return rows, outs


def write_portfolio_results(filepath, tables):
    """`tables`: key -> list of rows; writes `<filepath>_<key>.csv` with a header (demo/portfolio.jl:245-249)."""
    for k in KEYS:
        with open(f"{filepath}_{k}.csv", "w", newline="") as fh:
            w = csv.writer(fh)
            w.writerow(COLUMNS)
            w.writerows(tables.get(k, []))


def summary(tables):
    """The demo's closing five-number lines (demo/portfolio.jl:251-260): key -> {column: (min, max, median)}."""
    out = {}
    for k in KEYS:
        r = tables.get(k, [])
        col = lambda i: np.array([row[i] for row in r], dtype=float)  # noqa: E731
        out[k] = {name: (float(col(i).min()), float(col(i).max()), float(np.median(col(i)))) if r else None
                  for name, i in (("nnz", 6), ("cviolation", 4), ("subiters", 2))}
    return out
