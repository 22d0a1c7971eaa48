"""Golden vectors for the matrix prox operators (src/proxoperators/rank.jl, nuclearNorm.jl, schattenNormLp.jl): inputs and the
numpy restatement's outputs (LAPACK svd, as the reference's `svd`), written to tests/golden/matrix_prox.json.
    python tests/golden/make_golden_matprox.py"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle"))
import bazinga_oracle as bo  # noqa: E402

rng = np.random.default_rng(20261019)
cases = []
for name, shape in (("sq5", (5, 5)), ("tall7x4", (7, 4)), ("wide3x6", (3, 6)), ("rank2_6x6", (6, 6)), ("dmc20", (20, 20))):
    if name == "rank2_6x6":
        X = rng.standard_normal((6, 2)) @ rng.standard_normal((2, 6))
    elif name == "dmc20":      # demo/distance_matrix_completion.jl:163-167: a squared-distance matrix of 20 points in R^5
        P = rng.standard_normal((20, 5))
        X = ((P[:, None, :] - P[None, :, :]) ** 2).sum(-1)
    else:
        X = rng.standard_normal(shape)
    S = np.linalg.svd(X, compute_uv=False)
    gamma = float(0.11 * np.median(S[S > 1e-9]) ** 2 + 0.01)
    ent = {"name": name, "X": X.tolist(), "gamma": gamma, "sigma": S.tolist(), "out": {}}
    for key, g in (("rank", bo.GRank(0.7)), ("nuclear", bo.GNuclearNorm(0.7)), ("schatten", bo.GSchattenNormLpPower(0.5, 0.8))):
        thr = np.sqrt(2 * gamma * 0.7)
        assert key != "rank" or np.all(np.abs(S - thr) > 0.02 * thr), (name, S, thr)     # keep the hard threshold clear of sigma
        Y = np.empty_like(X)
        val = g.prox(Y, X, gamma)
        ent["out"][key] = {"Y": Y.tolist(), "value": float(val), "f_X": float(g(X))}
    cases.append(ent)
json.dump({"lambda": 0.7, "p": 0.5, "alpha": 0.8, "cases": cases}, open(os.path.join(HERE, "matrix_prox.json"), "w"))
print("wrote", len(cases), "cases")
