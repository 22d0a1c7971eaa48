"""Per-iteration golden traces of the inner solver (tests/golden/panoc_traces.json, made by tests/golden/make_golden_traces.py
from the numpy oracle): gamma, tau, f_x, f_z, g_z, the stop norm and the iterate z for the first 12 PANOCplus iterations of
basis pursuit 20x100, the 1-D and 2-D obstacle problems and the reference's tiny QP.

  * CPU: the oracle still produces exactly these numbers (a pin against silent changes of the restatement);
  * GPU: replaying every prefix -- PANOCplus(maxit = k), k = 1..12 -- through the C ABI reproduces entry k to 1e-10
    (the verbose table of upstream prints k, gamma, ||res||/gamma, tau: test/problems/test_verbose.jl:37)."""
import json
import os

import numpy as np
import pytest

import bazinga_oracle as bo

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "panoc_traces.json")))
KEYS = ("gamma", "tau", "f_x", "f_z", "g_z", "stopnorm")


def _oracle_problem(ent):
    if ent["kind"] == "bp":
        A, b = np.array(ent["A"]), np.array(ent["b"])
        return bo.FZero(), bo.GNormL1(1.0), bo.CDenseAffine(A, b), bo.ZeroSet()
    if ent["kind"] == "obstacle":
        P = bo.problem_obstacle(ent["N"], tuple(ent["grid"]) if ent["grid"] else None)
        return P["f"], P["g"], P["c"], P["D"]
    Q, q = np.array(ent["Q"]), np.array(ent["q"])
    return bo.FQuadratic(Q, q), bo.GIndBox(ent["lo"], ent["hi"]), bo.CIdentity(), bo.BoxSet(ent["lo"], ent["hi"])


@pytest.mark.parametrize("name", sorted(GOLD))
def test_oracle_reproduces_the_golden_traces(name):
    ent = GOLD[name]
    f, g, c, D = _oracle_problem(ent)
    mu, y, x0 = np.array(ent["mu"]), np.array(ent["y"]), np.array(ent["x0"])
    tr = []
    bo.panocplus(bo.AugLagFun(f, c, D, mu.copy(), y.copy(), x0), bo.NonsmoothCostFun(g), x0,
                 bo.PanocOptions(minimum_gamma=bo.EPS, maxit=len(ent["trace"]), tol=0.0, trace=tr))
    assert len(tr) == len(ent["trace"])
    for t, gt in zip(tr, ent["trace"]):
        assert t["k"] == gt["k"]
        for key in KEYS:
            assert t[key] == gt[key], (name, t["k"], key)
        assert np.array_equal(t["z"], np.array(gt["z"]))


def _gpu_problem(bz, ctx, ent):
    if ent["kind"] == "bp":
        return bz.Problem(bz.Zero(), bz.NormL1(1.0), bz.ConstraintBasisPursuit(np.array(ent["A"]), np.array(ent["b"])), bz.ZeroSet(), ctx=ctx)
    if ent["kind"] == "obstacle":
        N = ent["N"]
        return bz.Problem(bz.SmoothCostObstacleL1(N), bz.NonsmoothCostObstacleL1(N),
                          bz.ConstraintObstacle(N, tuple(ent["grid"]) if ent["grid"] else None), bz.ZeroSet(), ctx=ctx)
    return bz.Problem(bz.Quadratic(np.array(ent["Q"]), np.array(ent["q"])), bz.IndBox(ent["lo"], ent["hi"]), bz.IdentityFunction(),
                      bz.ClosedSet(bz.IndBox(ent["lo"], ent["hi"])), nx=2, ny=2, ctx=ctx)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(GOLD))
def test_gpu_replays_the_golden_traces(gpu_ctx, name):
    import bazinga_b200 as bz
    ent = GOLD[name]
    p = _gpu_problem(bz, gpu_ctx, ent)
    mu, y, x0 = np.array(ent["mu"]), np.array(ent["y"]), np.array(ent["x0"])
    for gt in ent["trace"]:
        k = gt["k"]
        sub = bz.PANOCplus(directions=bz.LBFGS(5), maxit=k, tol=0.0, minimum_gamma=bo.EPS)
        z, it, fx, gz = p.panocplus(mu, y, x0, sub)
        st = p.panoc_state()
        assert it == k == st["k"]
        for key in KEYS:
            ref = gt[key]
            assert abs(st[key] - ref) <= 1e-10 * max(1.0, abs(ref)), (name, k, key, st[key], ref)
        zr = np.array(gt["z"])
        assert np.max(np.abs(z - zr)) <= 1e-10 * max(1.0, np.max(np.abs(zr))), (name, k)
        assert abs(gz - gt["g_z"]) <= 1e-10 * max(1.0, abs(gt["g_z"]))
    p.close()
