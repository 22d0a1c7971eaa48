"""Per-iteration golden traces of PANOCplus (upstream's verbose table prints k, gamma, ||res||/gamma,
tau -- test/problems/test_verbose.jl:37) from the numpy oracle: gamma, tau, f_x, f_z, g_z, stop norm and the iterate z for
the first 12 iterations of four small problems.  Run from the repo root:  python tests/golden/make_golden_traces.py
The GPU tests replay every prefix (PANOCplus(maxit = k)) and compare to 1e-10."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import bazinga_oracle as bo  # noqa: E402

K = 12


def cases():
    rng = np.random.default_rng(2024)
    # (1) basis pursuit 20 x 100 (demo/basispursuit.jl sizes), L1
    A, _, b = bo.make_basis_pursuit(20, 100)
    yield "bp_20x100", dict(kind="bp", A=A, b=b), bo.FZero(), bo.GNormL1(1.0), bo.CDenseAffine(A, b), bo.ZeroSet(), \
        rng.uniform(0.05, 0.5, 20), 0.1 * rng.standard_normal(20), np.zeros(100)
    # (2) obstacle_l1, 1-D, N = 16 (demo/obstacle.jl)
    N = 16
    P = bo.problem_obstacle(N)
    yield "obstacle_1d_N16", dict(kind="obstacle", N=N, grid=None), P["f"], P["g"], P["c"], P["D"], \
        rng.uniform(0.05, 0.5, N), 0.1 * rng.standard_normal(N), 2.0 * np.ones(3 * N)
    # (3) obstacle_l1 on an 8 x 12 grid (the 2-D extension of config C4)
    N = 96
    P = bo.problem_obstacle(N, (8, 12))
    yield "obstacle_2d_8x12", dict(kind="obstacle", N=N, grid=[8, 12]), P["f"], P["g"], P["c"], P["D"], \
        rng.uniform(0.05, 0.5, N), 0.1 * rng.standard_normal(N), 2.0 * np.ones(3 * N)
    # (4) the tiny nonconvex QP of test/problems/test_nonconvex_qp.jl:9-17
    Q, q = np.diag([-0.5, 1.0]), np.array([0.3, 0.5])
    yield "tiny_qp", dict(kind="qp", Q=Q, q=q, lo=-1.0, hi=1.0), bo.FQuadratic(Q, q), bo.GIndBox(-1.0, 1.0), bo.CIdentity(), \
        bo.BoxSet(-1.0, 1.0), np.array([0.3, 0.7]), np.array([0.2, -0.1]), np.array([0.9, -0.4])


def main():
    out = {}
    for name, meta, f, g, c, D, mu, y, x0 in cases():
        al = bo.AugLagFun(f, c, D, mu.copy(), y.copy(), x0)
        tr = []
        bo.panocplus(al, bo.NonsmoothCostFun(g), x0, bo.PanocOptions(minimum_gamma=bo.EPS, maxit=K, tol=0.0, trace=tr))
        ent = {k: (v.tolist() if isinstance(v, np.ndarray) else v) for k, v in meta.items()}
        ent.update(mu=mu.tolist(), y=y.tolist(), x0=x0.tolist(),
                   trace=[dict(k=t["k"], gamma=t["gamma"], tau=t["tau"], f_x=t["f_x"], f_z=t["f_z"], g_z=t["g_z"],
                               stopnorm=t["stopnorm"], z=t["z"].tolist()) for t in tr])
        out[name] = ent
    with open(os.path.join(ROOT, "tests", "golden", "panoc_traces.json"), "w") as fh:
        json.dump(out, fh)
    print({k: len(v["trace"]) for k, v in out.items()})


if __name__ == "__main__":
    main()
