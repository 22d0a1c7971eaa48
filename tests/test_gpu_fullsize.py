"""Parity at the FULL sizes of BASELINE.json's configs, through properties that need no O(size) oracle run
(prompt ③): identities of the operators, sampled rows against the oracle's generator, determinism, and -- where the
oracle is vectorised enough to finish in seconds -- the direct comparison.  All through the C ABI, -m gpu."""
import numpy as np
import pytest

import bazinga_b200 as bz
import bazinga_oracle as bo
from tiny_oracle_binding import tiny_batch

pytestmark = pytest.mark.gpu
KTOL = 1e-12


def _dense_identities(ctx, m, n, rows_sampled=6):
    """c(x) = A x - b with A generated on the device (never on the host): D = {0}, b = 0."""
    c = bz.ConstraintBasisPursuit(None, np.zeros(m), gen_seed=0, shape=(m, n))
    p = bz.Problem(bz.Zero(), bz.NormL1(1.0), c, bz.ZeroSet(), ctx=ctx)
    rng = np.random.default_rng(m + n)
    mu, y = rng.uniform(0.5, 2.0, m), rng.standard_normal(m)
    x1, x2 = rng.standard_normal(n), rng.standard_normal(n)
    # (1) energy / adjoint identity with mu = 1, y = 0:  x'(A'(A x)) = ||A x||^2,  lx = 0.5 ||A x||^2
    p.auglag_update(np.ones(m), np.zeros(m))
    g1 = p.gradient(x1)
    e = float(np.dot(g1["cx"], g1["cx"]))
    assert abs(float(np.dot(x1, g1["dlx"])) - e) <= 1e-11 * e
    assert abs(g1["lx"] - 0.5 * e) <= 1e-12 * e and np.array_equal(g1["yupd"], g1["cx"])
    # (2) sampled rows against the oracle's generator (the same Philox stream, bit for bit)
    for i in np.linspace(0, m - 1, rows_sampled).astype(int):
        row = bo.bz_randn_like(0, int(i) * n, n)
        assert abs(g1["cx"][i] - row @ x1) <= KTOL * np.linalg.norm(row) * np.linalg.norm(x1)
    # (3) linearity of x -> A'(A x) (b = 0)
    g2, g12 = p.gradient(x2), p.gradient(x1 + x2)
    nrm = np.linalg.norm(g1["dlx"]) + np.linalg.norm(g2["dlx"])
    assert np.linalg.norm(g12["dlx"] - (g1["dlx"] + g2["dlx"])) <= 1e-12 * nrm
    # (4) general mu, y:  dlx = A' yupd with yupd = cx/mu + y  ->  x' dlx = cx' yupd;  lx = 0.5 sum(mu yupd^2) - 0.5 sum(mu y^2)
    p.auglag_update(mu, y)
    g = p.gradient(x1)
    assert np.max(np.abs(g["yupd"] - (g["cx"] / mu + y))) <= 1e-13 * np.max(np.abs(g["yupd"]))
    lhs, rhs = float(np.dot(x1, g["dlx"])), float(np.dot(g["cx"], g["yupd"]))
    assert abs(lhs - rhs) <= 1e-11 * (np.linalg.norm(g["cx"]) * np.linalg.norm(g["yupd"]))
    lx = 0.5 * float(np.sum(mu * g["yupd"] ** 2)) - 0.5 * float(np.sum(mu * y * y))
    assert abs(g["lx"] - lx) <= 1e-12 * float(np.sum(mu * g["yupd"] ** 2))
    # (5) determinism: the same evaluation twice is bit-identical (order-fixed reductions, no float atomics)
    gg = p.gradient(x1)
    assert np.array_equal(gg["dlx"], g["dlx"]) and gg["lx"] == g["lx"]
    p.close()


def test_c2_full_size_dense_identities(gpu_ctx):
    """BASELINE configs[1]: 4096 x 16384 (the fused 2-CTA cluster kernel)."""
    _dense_identities(gpu_ctx, 4096, 16384)


def test_c5_shard_full_width_dense_identities(gpu_ctx):
    """One of the 8 row shards of BASELINE configs[4]: 4096 x 262144 (8.6 GB; the grid-wide single-pass kernel)."""
    _dense_identities(gpu_ctx, 4096, 262144, rows_sampled=3)


def test_c2_full_size_solve_is_deterministic(gpu_ctx):
    c = bz.ConstraintBasisPursuit(None, np.zeros(4096), gen_seed=0, shape=(4096, 16384))
    p = bz.Problem(bz.Zero(), bz.NormL1(1.0), c, bz.ZeroSet(), ctx=gpu_ctx)
    p.auglag_update(np.ones(4096), np.zeros(4096))
    b = p.gradient(bo.bz_sparse_randn(1, 16384, 0.1))["cx"]          # b = A xtrue from the device pass
    p.upload_dense(None, b)
    from bazinga_b200.solver import _options
    sub = bz.PANOCplus(directions=bz.LBFGS(5), maxit=100_000, minimum_gamma=bo.EPS)
    o = _options(tol=1e-6, tol_prim=None, tol_dual=None, inner_tol=None, maxit=100, theta_penalty=0.8, kappa_penalty=0.5,
                 kappa_tol=0.1, subsolver_maxit=100_000, sub=sub, max_inner_total=150)
    runs = [p.alps(np.zeros(16384), np.zeros(4096), o) for _ in range(2)]
    assert runs[0][4].tot_inner_it == runs[1][4].tot_inner_it == 150
    assert np.array_equal(runs[0][0], runs[1][0]) and np.array_equal(runs[0][1], runs[1][1])
    assert runs[0][4].objective == runs[1][4].objective and np.isfinite(runs[0][4].objective)
    p.close()


@pytest.mark.parametrize("kind,name,lo,hi", [(1, "eitheror", -4.0, 8.0), (3, "mpvca", -5.0, 20.0)])
def test_c3_full_batch_bit_exact(gpu_ctx, kind, name, lo, hi):
    """BASELINE configs[2]: 65,536 independent instances (256 x 256 starts on the demo box), bit-identical to the C oracle."""
    v = np.linspace(lo, hi, 256)
    X0 = np.array([(i, j) for j in v for i in v])
    out = bz.alps_batch(name, X0, ctx=gpu_ctx)
    ref = tiny_batch(kind, X0, min_gamma=bo.EPS, sub_maxit=1000, subsolver_maxit=1000)
    assert np.array_equal(out["iters"], ref["iters"]) and np.array_equal(out["sub_iters"], ref["sub_iters"])
    assert np.array_equal(out["status"], ref["status"])
    assert np.array_equal(out["x"], ref["x"], equal_nan=True) and np.array_equal(out["y"], ref["y"], equal_nan=True)


def test_c4_full_grid_gradient_equals_oracle(gpu_ctx):
    """BASELINE configs[3]: the 4096 x 4096 obstacle grid (nx = 50,331,648).  The oracle's stencil is vectorised numpy,
    so the AL-gradient is compared directly: elementwise outputs bit for bit, the scalar to 1e-13."""
    S = 4096
    N = S * S
    p = bz.Problem(bz.SmoothCostObstacleL1(N), bz.NonsmoothCostObstacleL1(N), bz.ConstraintObstacle(N, (S, S)), bz.ZeroSet(), ctx=gpu_ctx)
    rng = np.random.default_rng(4)
    x, mu, y = rng.standard_normal(3 * N), rng.uniform(0.1, 2.0, N), rng.standard_normal(N)
    p.auglag_update(mu, y)
    g = p.gradient(x)
    al = bo.AugLagFun(bo.FObstacleL1(N), bo.CObstacle(N, (S, S)), bo.ZeroSet(), mu.copy(), y.copy(), x)
    dl = np.empty(3 * N)
    lx = al.gradient(dl, x)
    assert np.array_equal(g["cx"], al.cx) and np.array_equal(g["yupd"], al.yupd) and np.array_equal(g["dlx"], dl)
    assert np.array_equal(g["dlx"][:N], -g["dlx"][2 * N:])             # jtv blocks 1 and 3: w and -w
    assert abs(g["lx"] - lx) <= 1e-13 * abs(lx)
    p.close()
