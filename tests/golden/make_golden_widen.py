"""Writes tests/golden/oracle_golden_widen.json from the CPU oracle: the widened rows of SURVEY.md §8(f) (portfolio
demo, obstacle_l2, Lp prox, als).  Like make_golden.py these freeze the ORACLE's outputs (the Julia reference cannot run
in this image); a change in the oracle's arithmetic or in the synthetic data generators shows up in
tests/test_oracle_pins.py before it can silently move the GPU parity target."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import bazinga_oracle as bo  # noqa: E402


def build():
    n = 40
    Q, rho, mu, ub = bo.make_portfolio(n, seed=9)
    P = bo.problem_portfolio(Q, rho, mu, ub, "lp")
    with np.errstate(all="ignore"):
        r = bo.als(P["f"], P["g"], P["c"], P["D"], np.ones(n), np.zeros(2), **P["kw"])
    N = 6
    x = np.linspace(-1.0, 2.0, 3 * N)
    m, y = np.linspace(0.2, 1.5, N), np.linspace(-1.0, 1.0, N)
    Po = bo.problem_obstacle_l2(N)
    al = bo.AugLagFun(Po["f"], Po["c"], Po["D"], m.copy(), y.copy(), x)
    dl = np.empty(3 * N)
    lx = al.gradient(dl, x)
    z = np.empty(3 * N)
    Po["g"].prox(z, x, 0.25)
    xs = np.linspace(0.0, 1.2, 13)
    zl = np.empty(13)
    gl = bo.GNormLpBox(0.5, 0.3, np.full(13, 0.9)).prox(zl, xs, 0.4)
    return {
        "portfolio40": {"Q_row0": Q[0, :4].tolist(), "Q_trace": float(np.trace(Q)), "rho": rho, "mu": mu[:4].tolist(),
                        "ub": ub[:4].tolist(), "lp": {"tot_it": r.tot_it, "tot_inner_it": r.tot_inner_it, "status": r.status,
                                                      "objective_q": 0.5 * float(r.x @ Q @ r.x), "nnz": int((r.x > 0).sum()),
                                                      "x": r.x.tolist(), "y": r.y.tolist()}},
        "obstacle_l2_N6": {"lx": lx, "dlx": dl.tolist(), "prox": z.tolist()},
        "lp_prox": {"x": xs.tolist(), "z": zl.tolist(), "g": gl},
    }


if __name__ == "__main__":
    with open(os.path.join(HERE, "oracle_golden_widen.json"), "w") as fh:
        json.dump(build(), fh)
    print("wrote oracle_golden_widen.json")
