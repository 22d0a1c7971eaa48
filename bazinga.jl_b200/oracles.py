"""The oracle catalogue: Python mirrors of the Julia types that ``Bazinga.alps(f, g, c, D, x0, y0)``
dispatches on.  These objects only *describe* a problem; every method that
computes (``prox!``, ``proj!``, ``gradient!`` ...) goes through the C-ABI to the CUDA kernels, see
solver.py.  Constructor guards raise the reference's messages.

Arbitrary user-defined Julia methods cannot cross a C ABI: the GPU path accepts this
closed catalogue and raises ``BazingaError(BZ_ERR_UNSUPPORTED)`` for anything else -- never a CPU
fallback.
"""
from __future__ import annotations

import numpy as np

from . import _capi as K


def _f64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


# ------------------------------------------------------------------ smooth costs f
class Zero:
    """ProximalOperators.Zero() / Bazinga.Zero (src/proxoperators/zero.jl): as f or as g."""
    f_kind = K.F_ZERO
    g_kind = K.G_ZERO


class LeastSquares:
    """ProximalOperators.LeastSquares(A, b): 0.5||Ax-b||^2 (test/problems/test_verbose.jl:22)."""
    f_kind = K.F_LEAST_SQUARES

    def __init__(self, A, b):
        self.A, self.b = _f64(A), _f64(b)
        if self.A.shape[0] != self.b.shape[0]:
            raise ValueError("LeastSquares: size mismatch")


class Quadratic:
    """ProximalOperators.Quadratic(Q, q): 0.5x'Qx + q'x (test/problems/test_nonconvex_qp.jl:14)."""
    f_kind = K.F_QUADRATIC

    def __init__(self, Q, q):
        self.Q, self.q = _f64(Q), _f64(q)


class SmoothCostP(Quadratic):
    """demo/portfolio.jl:33-44 SmoothCostP(Q): f(x) = 0.5 x'Qx, gradient! = mul!(dfx, Q, x) then 0.5 dot(x, dfx)."""

    def __init__(self, Q):
        Q = _f64(Q)
        if Q.ndim != 2 or Q.shape[0] != Q.shape[1]:
            raise ValueError("SmoothCostP: Q must be square")
        Quadratic.__init__(self, Q, np.zeros(Q.shape[0]))


class SmoothCostObstacleL1:
    """demo/obstacle.jl:37-47."""
    f_kind = K.F_OBSTACLE_L1

    def __init__(self, N):
        self.N = int(N)


class SmoothCostObstacleL2:
    """demo/obstacle.jl:21-35: 0.5||x[1:2N]||^2 - sum(x[N+1:2N]), gradient [x1; x2 - 1; 0]."""
    f_kind = K.F_OBSTACLE_L2

    def __init__(self, N):
        self.N = int(N)


# ------------------------------------------------------------------ nonsmooth costs g
class IndFree(Zero):
    """ProximalOperators.IndFree() (test/problems/test_nonconvex_qp.jl:39)."""


class NormL1:
    """ProximalOperators.NormL1(lambda) (demo/basispursuit.jl:63)."""
    g_kind = K.G_NORM_L1

    def __init__(self, lam=1.0):
        if lam < 0:
            raise ValueError("parameter λ must be nonnegative")
        self.lam = float(lam)


class NormL0:
    """ProximalOperators.NormL0(lambda) (demo/basispursuit.jl:64)."""
    g_kind = K.G_NORM_L0

    def __init__(self, lam=1.0):
        if lam < 0:
            raise ValueError("parameter λ must be nonnegative")
        self.lam = float(lam)


class IndBox:
    """ProximalOperators.IndBox(lo, hi) (test/problems/test_nonconvex_qp.jl:15); also usable as
    ``ClosedSet(IndBox(lo, hi))``."""
    g_kind = K.G_IND_BOX

    def __init__(self, lo, hi):
        self.lo, self.hi = float(lo), float(hi)


class NormL1Box:
    """src/proxoperators/normL1Box.jl:11-39."""
    g_kind = K.G_NORML1_BOX

    def __init__(self, lam=1.0, *, u):
        u = _f64(u)
        if lam < 0:
            raise ValueError("parameter λ must be nonnegative")
        if np.any(u < 0):
            raise ValueError("vector u must have nonnegative entries")
        self.lam, self.u = float(lam), u


class NormL1Nonneg:
    """src/proxoperators/normL1Nonneg.jl:11-42."""
    g_kind = K.G_NORML1_NONNEG

    def __init__(self, lam=1.0):
        if lam < 0:
            raise ValueError("λ must be nonnegative")
        self.lam = float(lam)


class NormL0Box:
    """src/proxoperators/normL0Box.jl:12-58."""
    g_kind = K.G_NORML0_BOX

    def __init__(self, lam=1.0, *, u):
        u = _f64(u)
        if lam < 0:
            raise ValueError("parameter λ must be nonnegative")
        if np.any(u < 0):
            raise ValueError("vector u must have nonnegative entries")
        self.lam, self.u = float(lam), u


def _check_lp(p, alpha):
    if p <= 0:
        raise ValueError("p must be positive")                  # normLpBox.jl:17-18
    if p >= 1:
        raise ValueError("p must be smaller than one")          # :19-20
    if alpha < 0:
        raise ValueError("alpha must be nonnegative")           # :22-23


class NormLpPowerBox:
    """src/proxoperators/normLpBox.jl:11-97: alpha * sum_i x_i^p, 0 < p < 1, on the box [0, u]."""
    g_kind = K.G_NORMLP_BOX

    def __init__(self, p, alpha=1.0, *, u):
        u = _f64(u)
        _check_lp(p, alpha)
        if np.any(u < 0):
            raise ValueError("vector u must have nonnegative entries")
        self.lam, self.lo, self.hi, self.u = float(alpha), float(p), 0.0, u


class NormLpPowerNonneg:
    """src/proxoperators/normLpNonneg.jl:12-90: alpha * sum_i x_i^p, 0 < p < 1, on x >= 0."""
    g_kind = K.G_NORMLP_NONNEG

    def __init__(self, p, alpha=1.0):
        _check_lp(p, alpha)
        self.lam, self.lo, self.hi = float(alpha), float(p), 0.0


class NonsmoothCostObstacleL1:
    """demo/obstacle.jl:64-81."""
    g_kind = K.G_OBSTACLE_L1

    def __init__(self, N):
        self.N = int(N)


class NonsmoothCostObstacleL2:
    """demo/obstacle.jl:49-62: max(0, x) on all three blocks, then the complementarity zeroing; value 0."""
    g_kind = K.G_OBSTACLE_L2

    def __init__(self, N):
        self.N = int(N)


class NonsmoothCostObstacleL2Red:
    """demo/obstacle.jl:121-126: max(0, x), value 0 (the same prox as demo/mpvca.jl's NonsmoothCostMPVCA)."""
    g_kind = K.G_NONNEG

    def __init__(self, N):
        self.N = int(N)


class NonsmoothCostObstacleRedL1:
    """demo/obstacle.jl:83-91: soft threshold + max(0, .) on block 1, identity on block 2 (nx = 2N)."""
    g_kind = K.G_OBSTACLE_RED_L1

    def __init__(self, N):
        self.N = int(N)


# ------------------------------------------------------------------ constraint functions c
class IdentityFunction:
    """test/definitions/identityFunction.jl:3-13."""
    c_kind = K.C_IDENTITY


class ConstraintBasisPursuit:
    """demo/basispursuit.jl:38-49: cx = A*x - b, jtv = A'*v.  ``A`` may be C-contiguous
    (row-major) or Fortran-contiguous (Julia's layout); or pass ``gen_seed`` to have the library
    generate A on the device."""
    c_kind = K.C_DENSE_AFFINE

    def __init__(self, A, b, *, gen_seed=None, shape=None, row_begin=0, ny_global=None):
        self.b = _f64(b)
        self.gen_seed = gen_seed
        self.row_begin = int(row_begin)
        self.ny_global = ny_global
        if A is None:
            if gen_seed is None or shape is None:
                raise ValueError("ConstraintBasisPursuit: need A, or gen_seed and shape")
            self.A, self.shape, self.layout = None, tuple(shape), K.ROW_MAJOR
        else:
            A = np.asarray(A, dtype=np.float64)
            if A.flags.c_contiguous:
                self.A, self.layout = A, K.ROW_MAJOR
            elif A.flags.f_contiguous:
                self.A, self.layout = A, K.COL_MAJOR
            else:
                self.A, self.layout = np.ascontiguousarray(A), K.ROW_MAJOR
            self.shape = A.shape
        if self.shape[0] != self.b.shape[0]:
            raise ValueError("ConstraintBasisPursuit: size mismatch")


class ConstraintP(ConstraintBasisPursuit):
    """demo/portfolio.jl:46-58 ConstraintP(mu): c(x) = [mu'x; e'x], jtv = mu*v[1] + v[2] -- the dense affine
    provider with the two rows [mu'; e'] and b = 0."""

    def __init__(self, mu):
        mu = _f64(mu).reshape(-1)
        ConstraintBasisPursuit.__init__(self, np.stack([mu, np.ones_like(mu)]), np.zeros(2))
        self.mu = mu


class ConstraintObstacle:
    """demo/obstacle.jl:92-113; ``grid=(gy, gx)`` selects the 2-D 5-point extension (config C4).
    ``slab=(rank, nranks)``: this object describes ONE rank's slab of a grid split by rows over ranks
    (N and grid are local: gy local grid rows; every rank the same) -- SURVEY.md §8e."""
    c_kind = K.C_OBSTACLE

    def __init__(self, N, grid=None, slab=None):
        self.N = int(N)
        self.grid = grid
        self.slab = slab
        if slab is not None and (grid is None or int(grid[0]) * int(grid[1]) != self.N):
            raise ValueError("slab decomposition needs the local 2-D grid: grid=(gy_local, gx), N = gy_local*gx")


class ConstraintObstacleRed:
    """demo/obstacle.jl:128-146: c(x) = [x1 + A x2; x2] (nx = ny = 2N); used with D = SetObstacleRed = CC pairs (i, i+N)."""
    c_kind = K.C_OBSTACLE_RED

    def __init__(self, N, grid=None):
        self.N = int(N)
        self.grid = grid
        self.slab = None


# ------------------------------------------------------------------ closed sets D
class ZeroSet:
    """src/projections/zeroSet.jl (also demo/obstacle.jl SetObstacle)."""
    d_kind = K.D_ZERO


class FreeSet:
    """src/projections/freeSet.jl."""
    d_kind = K.D_FREE


class IndicatorSet:
    """src/projections/indicatorSet.jl: ClosedSet(f) = IndicatorSet(f) (src/Bazinga.jl:18).
    Only IndBox is in the catalogue."""
    d_kind = K.D_BOX

    def __init__(self, f):
        if not isinstance(f, IndBox):
            raise K.BazingaError(K.BZ_ERR_UNSUPPORTED, "IndicatorSet: only IndBox is supported on the GPU path")
        self.lo, self.hi = f.lo, f.hi


class SetP:
    """demo/portfolio.jl:60-69 SetP(rho) = [rho, inf) x {1}: proj! copies cx, then z[1] = max(rho, cx[1]), z[2] = 1."""
    d_kind = K.D_PORTFOLIO

    def __init__(self, rho):
        self.lo, self.hi = float(rho), 1.0
        self.rho = float(rho)


def ClosedSet(f):
    return IndicatorSet(f)


class PairSet:
    """Pairs projected by project_onto_{EITHEROR,XOR,VC,CC}_set! (src/projections/orConstraints.jl,
    vanishingConstraints.jl, complementarityConstraints.jl), adjacent (stride 1) as in SetEOR /
    SetXOR / SetMPVCA, or (i, i+N) as in SetObstacleRed."""
    _KINDS = {"eitheror": K.D_EITHEROR, "xor": K.D_XOR, "vc": K.D_VC, "cc": K.D_CC}

    def __init__(self, kind, stride=1):
        self.d_kind = self._KINDS[kind]
        self.stride = int(stride)


def SetEOR():
    return PairSet("eitheror")


def SetXOR():
    return PairSet("xor")


def SetMPVCA():
    return PairSet("vc")
