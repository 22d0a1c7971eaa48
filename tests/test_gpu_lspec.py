"""Speculative L step (csrc/bz_lspec.cuh): the L-BFGS update / two-loop direction of the NEXT PANOC+ iteration is
computed beside the dense AL-gradient at z and committed when the decision step accepts the iterate.  The arithmetic is
the plain L kernel's statement by statement, so a solve must give the SAME BITS with the speculation on and off
(BZ_LSPEC is read when a context is created), and both must follow the numpy oracle."""
import os

import numpy as np
import pytest

import bazinga_b200 as bz
import bazinga_oracle as bo

pytestmark = pytest.mark.gpu


def _ctx(lspec):
    old = os.environ.get("BZ_LSPEC")
    os.environ["BZ_LSPEC"] = "1" if lspec else "0"
    try:
        return bz.Context(0)
    finally:
        if old is None:
            os.environ.pop("BZ_LSPEC", None)
        else:
            os.environ["BZ_LSPEC"] = old


@pytest.mark.parametrize("m,n,memory,g", [(20, 100, 5, "l1"), (64, 1000, 5, "l1"), (150, 4096, 3, "l1"), (300, 16384, 5, "l1"),
                                         (40, 2500, 8, "l0"), (33, 257, 0, "l1")])
def test_speculative_l_step_gives_the_same_bits(m, n, memory, g):
    A, _, b = bo.make_basis_pursuit(m, n, seed=m + n)
    gg = bz.NormL1(1.0) if g == "l1" else bz.NormL0(0.01)
    outs = []
    for lspec in (True, False):
        ctx = _ctx(lspec)
        sub = bz.PANOCplus(directions=bz.LBFGS(memory) if memory else bz.NoAcceleration(), maxit=100_000, minimum_gamma=bo.EPS)
        out, res = bz.alps(bz.Zero(), gg, bz.ConstraintBasisPursuit(A, b), bz.ZeroSet(), np.zeros(n), np.zeros(m), tol=1e-6,
                           subsolver=sub, subsolver_maxit=100_000, max_inner_total=400, ctx=ctx, return_result=True)
        outs.append((out, res.n_grad, res.tot_inner_it, res.tot_it))
        ctx.close()
    (o1, g1, i1, t1), (o0, g0, i0, t0) = outs
    assert (g1, i1, t1) == (g0, i0, t0)
    assert np.array_equal(o1[0], o0[0]) and np.array_equal(o1[1], o0[1])       # x and y, bit for bit
    assert np.array_equal(o1[9], o0[9]) and o1[7] == o0[7]                     # mu, ||c(x) - s||_inf


def test_speculative_l_step_follows_the_oracle():
    m, n = 20, 100
    A, _, b = bo.make_basis_pursuit(m, n)
    ctx = _ctx(True)
    p = bz.Problem(bz.Zero(), bz.NormL1(1.0), bz.ConstraintBasisPursuit(A, b), bz.ZeroSet(), ctx=ctx)
    from bazinga_b200.solver import _options
    sub = bz.PANOCplus(directions=bz.LBFGS(5), maxit=100_000, minimum_gamma=bo.EPS)
    for cap in (2, 5, 12, 30):
        o = _options(tol=1e-6, tol_prim=None, tol_dual=None, inner_tol=None, maxit=100, theta_penalty=0.8, kappa_penalty=0.5,
                     kappa_tol=0.1, subsolver_maxit=100_000, sub=sub, max_inner_total=cap)
        x, y, s, mu, res = p.alps(np.zeros(n), np.zeros(m), o)
        r = bo.solve(bo.problem_basis_pursuit(A, b), np.zeros(n), max_inner_total=cap)
        assert (res.tot_it, res.tot_inner_it, res.n_grad) == (r.tot_it, r.tot_inner_it, r.n_grad)
        assert np.max(np.abs(x - r.x)) <= 1e-9 * max(1e-300, np.max(np.abs(r.x)))
    p.close()
    ctx.close()
