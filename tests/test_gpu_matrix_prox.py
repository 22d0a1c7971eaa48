"""Matrix prox operators (src/proxoperators/rank.jl, nuclearNorm.jl, schattenNormLp.jl) on the device against the numpy
restatement (LAPACK svd, as the reference): singular values to 1e-12 relative, prox output to 1e-10 of ||X||, the returned
value to 1e-12 (the rank count exactly).  Thresholds are placed away from the singular values: the hard threshold of Rank and
the jump of the Lp subproblem are discontinuous, and two factorisations may not fall on opposite sides of them."""
import numpy as np
import pytest

import bazinga_b200 as bz
import bazinga_oracle as bo

pytestmark = pytest.mark.gpu


def _pairs(lam=0.7, p=0.5, alpha=0.8):
    return [(bz.Rank(lam), bo.GRank(lam)), (bz.NuclearNorm(lam), bo.GNuclearNorm(lam)),
            (bz.SchattenNormLpPower(p, alpha), bo.GSchattenNormLpPower(p, alpha))]


def _safe_gamma(S, lam, base):
    """a step for which the Rank threshold sqrt(2 gamma lam) keeps 2 % clear of every singular value"""
    g = base
    for _ in range(50):
        thr = np.sqrt(2 * g * lam)
        if np.all(np.abs(S - thr) > 0.02 * thr):
            return g
        g *= 1.07
    return g


@pytest.mark.parametrize("shape", [(1, 1), (2, 2), (3, 3), (5, 5), (20, 20), (33, 33), (64, 64), (7, 4), (4, 7), (64, 3), (2, 63)])
def test_matrix_prox_matches_the_svd_restatement(gpu_ctx, shape):
    rng = np.random.default_rng(shape[0] * 100 + shape[1])
    X = rng.standard_normal(shape)
    S = np.linalg.svd(X, compute_uv=False)
    sig = bz.NuclearNorm().svdvals(X, ctx=gpu_ctx)
    assert sig.shape == S.shape and np.max(np.abs(sig - S)) <= 1e-12 * S[0]
    for f, o in _pairs():
        gamma = _safe_gamma(S, 0.7, 0.3 * np.median(S) ** 2 + 1e-3)
        Y, gz = f.prox(X, gamma, ctx=gpu_ctx)
        Yr = np.empty_like(X)
        gr = o.prox(Yr, X, gamma)
        assert np.linalg.norm(Y - Yr) <= 1e-10 * max(1.0, np.linalg.norm(X)), type(f).__name__
        assert abs(gz - gr) <= 1e-12 * max(1.0, abs(gr)), (type(f).__name__, gz, gr)
        # f(X) itself (rank.jl:24-26, nuclearNorm.jl:24-27, schattenNormLp.jl:32-34)
        assert abs(f(X, ctx=gpu_ctx) - o(X)) <= 1e-12 * max(1.0, abs(o(X)))


def test_matrix_prox_batch_rank_deficient_zero_and_vector_input(gpu_ctx):
    rng = np.random.default_rng(5)
    N, l = 20, 5                                    # demo/distance_matrix_completion.jl:163-167: N = 20, rank-5 Gram data
    P = rng.standard_normal((N, l))
    G = P @ P.T                                     # rank 5
    batch = np.stack([G, np.zeros((N, N)), rng.standard_normal((N, N)), np.diag(np.arange(N, 0, -1.0))])
    for f, o in _pairs(lam=1.0, p=0.5, alpha=1.0):
        gamma = 0.37
        Y, gz = f.prox(batch, gamma, ctx=gpu_ctx)
        assert Y.shape == batch.shape and gz.shape == (4,)
        for b in range(4):
            Yr = np.empty((N, N))
            gr = o.prox(Yr, batch[b], gamma)
            assert np.linalg.norm(Y[b] - Yr) <= 1e-10 * max(1.0, np.linalg.norm(batch[b])), (type(f).__name__, b)
            assert abs(gz[b] - gr) <= 1e-11 * max(1.0, abs(gr)), (type(f).__name__, b, gz[b], gr)
    assert bz.Rank(2.0)(G, ctx=gpu_ctx) == 2.0 * 5 and bz.Rank()(np.zeros((4, 4)), ctx=gpu_ctx) == 0
    # diagonal input: closed forms (soft / hard threshold of the diagonal)
    d = np.array([3.0, 1.0, 0.2])
    Y, gz = bz.NuclearNorm(0.5).prox(np.diag(d), 1.0, ctx=gpu_ctx)
    assert np.allclose(Y, np.diag([2.5, 0.5, 0.0]), atol=1e-14) and abs(gz - 0.5 * 3.0) <= 1e-14
    Y, gz = bz.Rank(0.5).prox(np.diag(d), 1.0, ctx=gpu_ctx)        # threshold sqrt(2 * 1 * 0.5) = 1: strictly above only
    assert np.allclose(Y, np.diag([3.0, 0.0, 0.0]), atol=1e-14) and gz == 0.5
    # vectorised input (rank.jl:36-58): column-major reshape, vector out
    x = rng.standard_normal(36)
    y, gz = bz.NuclearNorm().prox(x, 0.2, ctx=gpu_ctx)
    Yr = np.empty((6, 6))
    gr = bo.GNuclearNorm().prox(Yr, bo.check_and_reshape_as_matrix(x), 0.2)
    assert y.shape == (36,) and np.max(np.abs(y - Yr.reshape(-1, order="F"))) <= 1e-12 and abs(gz - gr) <= 1e-12
    with pytest.raises(ValueError, match="do not know how to reshape"):
        bz.Rank().prox(np.zeros(35), 1.0, ctx=gpu_ctx)


def test_matrix_prox_constructor_errors_and_limits(gpu_ctx):
    with pytest.raises(ValueError, match="nonnegative"):
        bz.Rank(-1.0)
    with pytest.raises(ValueError, match="nonnegative"):
        bz.NuclearNorm(-0.1)
    with pytest.raises(ValueError, match="p must be positive"):
        bz.SchattenNormLpPower(0.0)
    with pytest.raises(ValueError, match="smaller than one"):
        bz.SchattenNormLpPower(1.0)
    with pytest.raises(ValueError, match="alpha must be nonnegative"):
        bz.SchattenNormLpPower(0.5, -1.0)
    with pytest.raises(bz.BazingaError):
        bz.NuclearNorm().prox(np.zeros((65, 65)), 1.0, ctx=gpu_ctx)     # one CTA per matrix: 64 x 64 at most


def test_matrix_prox_against_the_golden_vectors(gpu_ctx):
    """tests/golden/matrix_prox.json: inputs and LAPACK-based outputs committed with the script that made them (among them a
    20 x 20 squared-distance matrix of rank <= 7 like demo/distance_matrix_completion.jl's, and a rank-2 matrix)."""
    import json
    import os
    G = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "matrix_prox.json")))
    for ent in G["cases"]:
        X = np.array(ent["X"])
        sig = bz.NuclearNorm().svdvals(X, ctx=gpu_ctx)
        assert np.max(np.abs(sig - np.array(ent["sigma"]))) <= 1e-12 * max(ent["sigma"])
        for key, f in (("rank", bz.Rank(G["lambda"])), ("nuclear", bz.NuclearNorm(G["lambda"])),
                       ("schatten", bz.SchattenNormLpPower(G["p"], G["alpha"]))):
            Y, val = f.prox(X, ent["gamma"], ctx=gpu_ctx)
            ref = ent["out"][key]
            assert np.linalg.norm(Y - np.array(ref["Y"])) <= 1e-10 * max(1.0, np.linalg.norm(X)), (ent["name"], key)
            assert abs(val - ref["value"]) <= 1e-11 * max(1.0, abs(ref["value"])), (ent["name"], key)
            tol = 1e-12 if key != "schatten" else 1e-6      # sigma^p of a singular value that is zero up to rounding
            assert abs(f(X, ctx=gpu_ctx) - ref["f_X"]) <= tol * max(1.0, abs(ref["f_X"])), (ent["name"], key)
