"""Batched shared-A mode (csrc/bz_bdense.cu; BASELINE north_star job 2): the AL-gradients of up to eight dense-affine
instances that share A, as two DMMA (fp64 tensor core) passes over A, against `batch` independent evaluations of the
numpy oracle's gradient!(dlx, alFun, x) (auglagfun.jl:73-86 around demo/basispursuit.jl:43,47) and against the
single-instance kernels of the library.  Tolerance 1e-12 normwise (results of dense products)."""
import numpy as np
import pytest

import bazinga_b200 as bz
import bazinga_oracle as bo

pytestmark = pytest.mark.gpu
KTOL = 1e-12


@pytest.mark.parametrize("m,n,B", [(32, 128, 8), (33, 130, 8), (7, 2, 3), (100, 1026, 5), (300, 4096, 8), (64, 16384, 8), (257, 386, 1)])
@pytest.mark.parametrize("dname", ["zero", "box"])
def test_shared_matrix_batch_gradient(gpu_ctx, m, n, B, dname):
    A = bo.bz_randn_like(m + n, 0, m * n).reshape(m, n)
    rng = np.random.default_rng(m * 3 + n + B)
    X = rng.standard_normal((B, n))
    X[:, rng.random(n) < 0.5] = 0.0
    b = rng.standard_normal((B, m))
    mu = rng.uniform(0.05, 3.0, (B, m))
    y = rng.standard_normal((B, m))
    Dg = {"zero": bz.ZeroSet(), "box": bz.ClosedSet(bz.IndBox(-0.5, 0.7))}[dname]
    Do = {"zero": bo.ZeroSet(), "box": bo.BoxSet(-0.5, 0.7)}[dname]
    h = bz.SharedMatrixBatch(A, ctx=gpu_ctx)
    out = h.gradient(X, b, mu, y, D=Dg)
    out2 = h.gradient(X, b, mu, y, D=Dg)
    nA = np.linalg.norm(A, 2)
    for k in range(B):
        al = bo.AugLagFun(bo.FZero(), bo.CDenseAffine(A, b[k]), Do, mu[k].copy(), y[k].copy(), X[k])
        dl = np.empty(n)
        lx = al.gradient(dl, X[k])
        cscale = nA * np.linalg.norm(X[k]) + np.linalg.norm(b[k])
        assert np.max(np.abs(out["cx"][k] - al.cx)) <= KTOL * cscale
        assert np.max(np.abs(out["yupd"][k] - al.yupd)) <= KTOL * cscale / mu[k].min()
        gscale = nA * max(np.linalg.norm(al.yupd), 1e-300)
        assert np.max(np.abs(out["dlx"][k] - dl)) <= KTOL * gscale
        assert abs(out["lx"][k] - lx) <= KTOL * max(abs(lx), np.sum(al.yupd**2 * mu[k]))
    for key in ("dlx", "cx", "yupd", "lx"):
        assert np.array_equal(out[key], out2[key])             # deterministic
    h.close()


def test_shared_matrix_batch_matches_single_instance_kernels(gpu_ctx):
    """eight instances through the batched DMMA passes = eight runs of the single-instance fused GEMV kernel"""
    m, n, B = 128, 2048, 8
    A = bo.bz_randn_like(5, 0, m * n).reshape(m, n)
    rng = np.random.default_rng(1)
    X, b = rng.standard_normal((B, n)), rng.standard_normal((B, m))
    mu, y = rng.uniform(0.1, 2.0, (B, m)), rng.standard_normal((B, m))
    h = bz.SharedMatrixBatch(A, ctx=gpu_ctx)
    out = h.gradient(X, b, mu, y)
    nA = np.linalg.norm(A, 2)
    for k in range(B):
        p = bz.Problem(bz.Zero(), bz.NormL1(1.0), bz.ConstraintBasisPursuit(A, b[k]), bz.ZeroSet(), ctx=gpu_ctx)
        p.auglag_update(mu[k], y[k])
        g = p.gradient(X[k])
        assert np.max(np.abs(out["dlx"][k] - g["dlx"])) <= KTOL * nA * np.linalg.norm(g["yupd"])
        assert np.max(np.abs(out["cx"][k] - g["cx"])) <= KTOL * (nA * np.linalg.norm(X[k]) + np.linalg.norm(b[k]))
        assert abs(out["lx"][k] - g["lx"]) <= KTOL * max(abs(g["lx"]), np.sum(g["yupd"]**2 * mu[k]))
        p.close()
    h.close()


def test_shared_matrix_batch_full_size_identities(gpu_ctx):
    """BASELINE configs[1] shape (4096 x 16384), A generated on the device: x'(A'w) = (Ax)'w per instance, linearity across the
    batch, and the timing line the DMMA profile in profiles/ refers to"""
    m, n, B = 4096, 16384, 8
    h = bz.SharedMatrixBatch(shape=(m, n), gen_seed=0, ctx=gpu_ctx)
    rng = np.random.default_rng(2)
    X = rng.standard_normal((B, n))
    X[1] = 2.0 * X[0]                                          # linearity: instance 1 = 2 x instance 0 (b = 0, y = 0, mu = 1)
    z = np.zeros((B, m))
    out = h.gradient(X, z, np.ones((B, m)), z, reps=5)
    for k in range(B):
        lhs, rhs = X[k] @ out["dlx"][k], out["cx"][k] @ out["yupd"][k]
        assert abs(lhs - rhs) <= 1e-11 * abs(rhs)
        assert abs(out["lx"][k] - 0.5 * np.sum(out["cx"][k]**2)) <= 1e-12 * out["lx"][k]
    assert np.max(np.abs(out["dlx"][1] - 2.0 * out["dlx"][0])) <= 1e-12 * np.max(np.abs(out["dlx"][1]))
    # a sampled row of A X against the oracle's generator
    row = bo.bz_randn_like(0, 1234 * n, n)
    assert abs(out["cx"][3][1234] - row @ X[3]) <= 1e-11 * np.linalg.norm(row) * np.linalg.norm(X[3])
    assert out["us"] > 0
    h.close()
