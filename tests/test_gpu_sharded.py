"""Row-sharded dense path (csrc/bz_comm.cu, the `sharded` branches of csrc/bz_solver.cu) against
the numpy oracle (auglagfun.jl:73-86 on the stacked rows, alps.jl:72-84) and against the unsharded kernels.

  * one GPU: a communicator of ONE rank makes the library take the sharded code path (k_local_reduce, the NCCL
    allreduce entry, the separate outer-logic kernel) on a strict sub-block of the rows; two such blocks emulate two
    ranks and their partial gradients must add up to the oracle's gradient;
  * two GPUs (skipped on a one-GPU box): one process per GPU, the library's own NCCL communicator; gradient, lx,
    multiplier update and the first iterates against the oracle and the unsharded kernels."""
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest

import bazinga_b200 as bz
import bazinga_oracle as bo

pytestmark = pytest.mark.gpu
KTOL = 1e-12
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _instance(m, n, seed):
    A = bo.bz_randn_like(seed, 0, m * n).reshape(m, n)
    b = A @ bo.bz_sparse_randn(seed + 1, n, 0.1)
    rng = np.random.default_rng(seed)
    return A, b, rng.uniform(0.1, 2.0, m), rng.standard_normal(m), rng.standard_normal(n)


@pytest.fixture(scope="module")
def comm1_ctx(gpu_ctx):
    """a second context on device 0 that owns a one-rank NCCL communicator"""
    ctx = bz.Context(0)
    ctx.comm_init(bz.comm_unique_id(), 1, 0)
    yield ctx
    ctx.close()


@pytest.mark.parametrize("m,n,blocks", [(64, 256, [(0, 30), (30, 34)]), (300, 2050, [(0, 150), (150, 150)]),
                                        (128, 16384, [(0, 64), (64, 64)]), (96, 40000, [(0, 40), (40, 56)])])
def test_row_block_gradient_is_the_oracles_gradient_restricted_to_the_rows(comm1_ctx, gpu_ctx, m, n, blocks):
    A, b, mu, y, x = _instance(m, n, m + n)
    al = bo.AugLagFun(bo.FZero(), bo.CDenseAffine(A, b), bo.ZeroSet(), mu.copy(), y.copy(), x)
    dl = np.empty(n)
    lx = al.gradient(dl, x)
    nA = np.linalg.norm(A, 2)
    gscale = nA * np.linalg.norm(al.yupd)
    total, lx_sum, musqy = np.zeros(n), 0.0, 0.0
    for r0, ml in blocks:
        c = bz.ConstraintBasisPursuit(A[r0:r0 + ml], b[r0:r0 + ml], row_begin=r0, ny_global=m)
        p = bz.Problem(bz.Zero(), bz.NormL1(1.0), c, bz.ZeroSet(), ctx=comm1_ctx)
        assert p.ny_local == ml and p.row_begin == r0
        p.auglag_update(mu[r0:r0 + ml], y[r0:r0 + ml])
        g = p.gradient(x)
        # the oracle on this row block alone (the same function on the stacked rows of the block)
        alb = bo.AugLagFun(bo.FZero(), bo.CDenseAffine(A[r0:r0 + ml], b[r0:r0 + ml]), bo.ZeroSet(), mu[r0:r0 + ml].copy(),
                           y[r0:r0 + ml].copy(), x)
        dlb = np.empty(n)
        lxb = alb.gradient(dlb, x)
        assert np.max(np.abs(g["cx"] - al.cx[r0:r0 + ml])) <= KTOL * (nA * np.linalg.norm(x) + np.linalg.norm(b))
        assert np.max(np.abs(g["yupd"] - al.yupd[r0:r0 + ml])) <= KTOL * (nA * np.linalg.norm(x) + np.linalg.norm(b)) / mu.min()
        assert np.max(np.abs(g["dlx"] - dlb)) <= KTOL * gscale
        assert abs(g["lx"] - lxb) <= KTOL * max(abs(lxb), np.sum(alb.yupd**2 * mu[r0:r0 + ml]))
        # k_local_reduce (fixed order p = 0, 1, ...) against the unsharded kernels' blocked assembly of the same partial rows
        p1 = bz.Problem(bz.Zero(), bz.NormL1(1.0), bz.ConstraintBasisPursuit(A[r0:r0 + ml], b[r0:r0 + ml]), bz.ZeroSet(), ctx=gpu_ctx)
        p1.auglag_update(mu[r0:r0 + ml], y[r0:r0 + ml])
        g1 = p1.gradient(x)
        assert np.array_equal(g["cx"], g1["cx"]) and np.array_equal(g["yupd"], g1["yupd"])
        assert np.max(np.abs(g["dlx"] - g1["dlx"])) <= 1e-14 * gscale
        # repeat: deterministic
        assert np.array_equal(p.gradient(x)["dlx"], g["dlx"])
        total += g["dlx"]
        lx_sum += g["lx"] + p.auglag_state()["musqy"]
        musqy += p.auglag_state()["musqy"]
        p.close()
        p1.close()
    assert np.max(np.abs(total - dl)) <= KTOL * gscale                      # the partial gradients add up (the allreduce)
    assert abs((lx_sum - musqy) - lx) <= KTOL * max(abs(lx), np.sum(al.yupd**2 * mu))


def test_one_rank_sharded_solve_follows_the_oracle(comm1_ctx):
    """the whole sharded driver (k_local_reduce + allreduce entry + k_outer_logic) on a one-rank communicator: the
    first iterates of alps on the row block are the oracle's on the same rows"""
    m, n, r0, ml = 80, 512, 16, 48
    A, b, *_ = _instance(m, n, 5)
    As, bs = A[r0:r0 + ml], b[r0:r0 + ml]
    c = bz.ConstraintBasisPursuit(As, bs, row_begin=r0, ny_global=m)
    p = bz.Problem(bz.Zero(), bz.NormL1(1.0), c, bz.ZeroSet(), ctx=comm1_ctx)
    from bazinga_b200.solver import _options
    sub = bz.PANOCplus(directions=bz.LBFGS(5), maxit=100_000, minimum_gamma=bo.EPS)
    for cap in (3, 12):
        o = _options(tol=1e-6, tol_prim=None, tol_dual=None, inner_tol=None, maxit=100, theta_penalty=0.8, kappa_penalty=0.5,
                     kappa_tol=0.1, subsolver_maxit=100_000, sub=sub, max_inner_total=cap)
        x, y, s, mu, res = p.alps(np.zeros(n), np.zeros(m), o)
        ro = bo.solve(bo.problem_basis_pursuit(As, bs), np.zeros(n), max_inner_total=cap)
        assert res.tot_inner_it == ro.tot_inner_it and res.tot_it == ro.tot_it
        sc = max(1e-300, np.max(np.abs(ro.x)))
        assert np.max(np.abs(x - ro.x)) <= 1e-10 * sc
        assert np.max(np.abs(y - ro.y)) <= 1e-9 * max(1.0, np.max(np.abs(ro.y)))
        assert np.array_equal(mu, ro.mu)
    p.close()


def _n_devices():
    import ctypes
    n = ctypes.c_int(0)
    bz.load_library().bz_device_count(ctypes.byref(n))
    return n.value


def test_row_sharded_two_ranks(gpu_ctx):
    """one process per GPU, the library's own NCCL communicator: the row-sharded AL-gradient, multiplier update and the
    first PANOC+ iterates against the numpy oracle and the unsharded kernels"""
    if _n_devices() < 2:
        pytest.skip("needs two GPUs (run under gpurun --gpus 2)")
    m, n, world = 256, 4096, 2
    A, b, mu, y, xp = _instance(m, n, 11)
    caps = [3, 12]
    with tempfile.TemporaryDirectory() as tmp:
        np.savez(os.path.join(tmp, "problem.npz"), A=A, b=b, mu=mu, y=y, xp=xp, caps=np.array(caps))
        procs = [subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "_sharded_worker.py"), str(r), str(world), tmp],
                                  stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(world)]
        outs = [q.communicate(timeout=600)[0] for q in procs]
        assert all(q.returncode == 0 for q in procs), "\n".join(outs)
        R = [dict(np.load(os.path.join(tmp, f"rank{r}.npz"))) for r in range(world)]
    al = bo.AugLagFun(bo.FZero(), bo.CDenseAffine(A, b), bo.ZeroSet(), mu.copy(), y.copy(), xp)
    dl = np.empty(n)
    lx = al.gradient(dl, xp)
    nA = np.linalg.norm(A, 2)
    gscale = nA * np.linalg.norm(al.yupd)
    assert np.array_equal(R[0]["dlx"], R[1]["dlx"]) and R[0]["lx"] == R[1]["lx"]          # bitwise identical on every rank
    assert np.max(np.abs(R[0]["dlx"] - dl)) <= KTOL * gscale
    assert abs(R[0]["lx"] - lx) <= KTOL * max(abs(lx), np.sum(al.yupd**2 * mu))
    cx = np.concatenate([r["cx"] for r in R])
    assert np.max(np.abs(cx - al.cx)) <= KTOL * (nA * np.linalg.norm(xp) + np.linalg.norm(b))
    assert np.max(np.abs(np.concatenate([r["yupd"] for r in R]) - al.yupd)) <= KTOL * (nA * np.linalg.norm(xp) + np.linalg.norm(b)) / mu.min()
    # multiplier update y = (c(x) + mu*y - proj_D(.)) / mu, alps.jl:72-81, row-local
    y_ref = (cx + mu * y - 0.0) / mu
    assert np.max(np.abs(np.concatenate([r["y_mult"] for r in R]) - y_ref)) <= 1e-14 * np.max(np.abs(y_ref))
    # unsharded kernels on one GPU
    pu = bz.Problem(bz.Zero(), bz.NormL1(1.0), bz.ConstraintBasisPursuit(A, b), bz.ZeroSet(), ctx=gpu_ctx)
    pu.auglag_update(mu, y)
    gu = pu.gradient(xp)
    assert np.max(np.abs(R[0]["dlx"] - gu["dlx"])) <= 1e-13 * gscale
    from bazinga_b200.solver import _options
    sub = bz.PANOCplus(directions=bz.LBFGS(5), maxit=100_000, minimum_gamma=bo.EPS)
    for cap in caps:
        o = _options(tol=1e-6, tol_prim=None, tol_dual=None, inner_tol=None, maxit=100, theta_penalty=0.8, kappa_penalty=0.5,
                     kappa_tol=0.1, subsolver_maxit=100_000, sub=sub, max_inner_total=cap)
        xu, yu, su, muu, ru = pu.alps(np.zeros(n), np.zeros(m), o)
        ro = bo.solve(bo.problem_basis_pursuit(A, b), np.zeros(n), max_inner_total=cap)
        for r in R:
            assert list(r[f"it_{cap}"][:2]) == [ro.tot_it, ro.tot_inner_it] == [ru.tot_it, ru.tot_inner_it]
        assert np.array_equal(R[0][f"x_{cap}"], R[1][f"x_{cap}"])                        # lock-step
        sc = np.max(np.abs(ro.x))
        assert np.max(np.abs(R[0][f"x_{cap}"] - ro.x)) <= 1e-10 * sc
        assert np.max(np.abs(R[0][f"x_{cap}"] - xu)) <= 1e-10 * sc
        ysh = np.concatenate([r[f"y_{cap}"] for r in R])
        assert np.max(np.abs(ysh - ro.y)) <= 1e-9 * max(1.0, np.max(np.abs(ro.y)))
        assert np.array_equal(np.concatenate([r[f"mu_{cap}"] for r in R]), ro.mu)
        assert abs(R[0][f"nrp_{cap}"] - ro.norm_res_prim) <= 1e-10 * max(1.0, ro.norm_res_prim)
    pu.close()
