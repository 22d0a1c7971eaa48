"""CPU ORACLE (test infrastructure -- NOT a product path).

A plain numpy restatement of the PANOC+/ALPS hot path of aldma/Bazinga.jl.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this module; the product (``bazinga.jl_b200``) never does.

What is restated, and from where (all file:line into the reference tree):

* outer augmented-Lagrangian loop ........ src/algorithms/alps.jl:7-117
* AugLagFun value / gradient / update .... src/utilities/auglagfun.jl:32-101
* NonsmoothCostFun ....................... src/utilities/nonsmoothcostfun.jl:11-22
* safeguards / default penalty ........... src/utilities/safeguards.jl:2-18
* projections ............................ src/projections/{zeroSet,freeSet,indicatorSet,
                                           orConstraints,vanishingConstraints,
                                           complementarityConstraints}.jl
* prox operators ......................... src/proxoperators/{zero,normL1Box,normL1Nonneg,
                                           normL0Box}.jl
* demo oracles ........................... demo/{rosenbrock,basispursuit,eitheror,mpvca,obstacle}.jl
* test fixture ........................... test/definitions/identityFunction.jl

The inner solver is NOT in the reference tree: ``ProximalAlgorithms.PANOCplus`` with
``LBFGS(5)`` directions (Project.toml:9, unpinned, no Manifest; call sites alps.jl:5,64,66).
It is restated here from its published algorithm (De Marchi & Themelis, "Proximal gradient
algorithms under local Lipschitz gradient continuity", PANOC+) and the 0.5.x-series behaviour
written down in SURVEY.md Appendix A.  ProximalOperators' NormL1 / NormL0 / IndBox / IndFree /
Zero / Quadratic / LeastSquares (used by the reference's demos and tests, also third-party) follow
SURVEY.md Appendix B.

PARITY PIN STATUS: Julia is not available, so the reference cannot be executed.  The oracle is
pinned against everything the reference's own tests hold for this path (tests/test_oracle_pins.py):
the golden vector and iteration bounds of test/problems/test_verbose.jl:29,42-44 and the
fixed-point-residual inequalities of test/problems/test_nonconvex_qp.jl:35,50,89,104.  Per-kernel
values and exact iteration counts are NOT pinned by the reference ("parity unpinned" at kernel
granularity) -- see DESIGN.md.
"""
from __future__ import annotations

import math
import time
from dataclasses import dataclass, field

import numpy as np

EPS = float(np.finfo(np.float64).eps)


# ----------------------------------------------------------------------------------------------
# 2-element set projections (src/projections/*.jl).  Scalar versions, exact branch order.
# ----------------------------------------------------------------------------------------------
def project_onto_EITHEROR_set(x1, x2):
    """orConstraints.jl:7-17: z = x; if both < 0, zero the larger one (tie -> second)."""
    z1, z2 = x1, x2
    if x1 < 0 and x2 < 0:
        if x1 > x2:
            z1 = 0.0
        else:
            z2 = 0.0
    return z1, z2


def project_onto_XOR_set(x1, x2):
    """orConstraints.jl:24-36."""
    z1, z2 = x1, x2
    if x1 * x2 > 0:
        if x1 > x2:
            z1 = max(0.0, x1)
            z2 = min(0.0, x2)
        else:
            z1 = min(0.0, x1)
            z2 = max(0.0, x2)
    return z1, z2


def project_onto_VC_set(x1, x2):
    """vanishingConstraints.jl:27-46 (tie x1+x2 == 0 keeps z2)."""
    z1, z2 = 0.0, 0.0
    if x1 <= 0:
        z2 = x2
    else:
        if x2 >= 0:
            z1, z2 = x1, x2
        else:
            if x1 + x2 > 0:
                z1 = x1
            elif x1 + x2 < 0:
                z2 = x2
            else:
                z2 = x2
    return z1, z2


def project_onto_CC_set(x1, x2):
    """complementarityConstraints.jl:8-20."""
    if x1 > 0 and x2 > 0:
        z1, z2 = x1, x2
        if x2 > x1:
            z1 = 0.0
        else:
            z2 = 0.0
    else:
        z1, z2 = max(0.0, x1), max(0.0, x2)
    return z1, z2


_PAIR_PROJ = {
    "eitheror": project_onto_EITHEROR_set,
    "xor": project_onto_XOR_set,
    "vc": project_onto_VC_set,
    "cc": project_onto_CC_set,
}


def _pair_proj_vec(kind, a, b):
    """Vectorised forms of the pair projections (same branch order, elementwise)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if kind == "eitheror":
        both = (a < 0) & (b < 0)
        z1 = np.where(both & (a > b), 0.0, a)
        z2 = np.where(both & ~(a > b), 0.0, b)
        return z1, z2
    if kind == "xor":
        same = (a * b) > 0
        gt = a > b
        z1 = np.where(same, np.where(gt, np.maximum(0.0, a), np.minimum(0.0, a)), a)
        z2 = np.where(same, np.where(gt, np.minimum(0.0, b), np.maximum(0.0, b)), b)
        return z1, z2
    if kind == "vc":
        le = a <= 0
        ge2 = b >= 0
        ssum = a + b
        # x1<=0: (0,x2); x1>0,x2>=0: (x1,x2); x1>0,x2<0: sum>0 -> (x1,0) else (0,x2)
        z1 = np.where(le, 0.0, np.where(ge2, a, np.where(ssum > 0, a, 0.0)))
        z2 = np.where(le, b, np.where(ge2, b, np.where(ssum > 0, 0.0, b)))
        return z1, z2
    if kind == "cc":
        pos = (a > 0) & (b > 0)
        z1 = np.where(pos, np.where(b > a, 0.0, a), np.maximum(0.0, a))
        z2 = np.where(pos, np.where(b > a, b, 0.0), np.maximum(0.0, b))
        return z1, z2
    raise ValueError(kind)


# ----------------------------------------------------------------------------------------------
# Closed sets D:  proj!(z, D, v)
# ----------------------------------------------------------------------------------------------
class ZeroSet:
    """src/projections/zeroSet.jl:17-20."""

    def proj(self, z, v):
        z[:] = 0.0


class FreeSet:
    """src/projections/freeSet.jl:17-20."""

    def proj(self, z, v):
        z[:] = v


class BoxSet:
    """ClosedSet(IndBox(lo,hi)) == IndicatorSet: proj! -> prox!(z, f.f, x) (indicatorSet.jl:8-11),
    i.e. the clamp of ProximalOperators.IndBox."""

    def __init__(self, lo, hi):
        self.lo, self.hi = lo, hi

    def proj(self, z, v):
        z[:] = np.minimum(self.hi, np.maximum(self.lo, v))


class SetPortfolio:
    """demo/portfolio.jl:60-69 SetP = [rho, inf) x {1}: z .= cx; z[1] = max(rho, cx[1]); z[2] = 1."""

    def __init__(self, rho):
        self.rho = float(rho)

    def proj(self, z, v):
        z[:] = v
        z[0] = max(self.rho, v[0])
        z[1] = 1.0


class PairSet:
    """Adjacent pairs (1:2),(3:4),... projected by one of the 2-element kernels
    (demo/eitheror.jl:122-126 SetEOR, :128-132 SetXOR; demo/mpvca.jl:146-150 SetMPVCA;
    demo/rosenbrock.jl:77-80 SetRosenbrock).  ``stride`` = 1 -> adjacent pairs (2k, 2k+1);
    ``stride`` = N -> pairs (i, i+N) (demo/obstacle.jl:150-156 SetObstacleRed)."""

    def __init__(self, kind, stride=1):
        self.kind, self.stride = kind, stride

    def proj(self, z, v):
        if self.stride == 1:
            z1, z2 = _pair_proj_vec(self.kind, v[0::2], v[1::2])
            z[0::2] = z1
            z[1::2] = z2
        else:
            N = self.stride
            z1, z2 = _pair_proj_vec(self.kind, v[:N], v[N : 2 * N])
            z[:N] = z1
            z[N : 2 * N] = z2


# ----------------------------------------------------------------------------------------------
# Nonsmooth terms g:  prox!(z, g, x, gamma) -> g(z)
# ----------------------------------------------------------------------------------------------
class GZero:
    """proxoperators/zero.jl:22-25; also ProximalOperators.IndFree/Zero and demo NonsmoothCostOR
    (demo/eitheror.jl:71-74)."""

    def prox(self, z, x, gamma):
        z[:] = x
        return 0.0


class GNormL1:
    """ProximalOperators.NormL1(lambda) (Appendix B): soft threshold, returns lambda*||z||_1."""

    def __init__(self, lam=1.0):
        self.lam = float(lam)

    def prox(self, z, x, gamma):
        gl = gamma * self.lam
        z[:] = np.sign(x) * np.maximum(np.abs(x) - gl, 0.0)
        return self.lam * float(np.sum(np.abs(z)))


class GNormL0:
    """ProximalOperators.NormL0(lambda) (Appendix B): hard threshold at sqrt(2*gamma*lambda)."""

    def __init__(self, lam=1.0):
        self.lam = float(lam)

    def prox(self, z, x, gamma):
        thr = math.sqrt(2.0 * gamma * self.lam)
        keep = np.abs(x) > thr
        z[:] = np.where(keep, x, 0.0)
        return self.lam * float(np.count_nonzero(z))


class GIndBox:
    """ProximalOperators.IndBox(lo,hi) (Appendix B): clamp, returns 0."""

    def __init__(self, lo, hi):
        self.lo, self.hi = lo, hi

    def prox(self, z, x, gamma):
        z[:] = np.minimum(self.hi, np.maximum(self.lo, x))
        return 0.0


class GNormL1Box:
    """proxoperators/normL1Box.jl:30-39: y = max(0, min(x - gamma*lambda, u)); lambda*sum(y)."""

    def __init__(self, lam, u):
        if lam < 0:
            raise ValueError("parameter λ must be nonnegative")
        u = np.asarray(u, dtype=np.float64)
        if np.any(u < 0):
            raise ValueError("vector u must have nonnegative entries")
        self.lam, self.u = float(lam), u

    def prox(self, z, x, gamma):
        gl = gamma * self.lam
        z[:] = np.maximum(0.0, np.minimum(x - gl, self.u))
        return self.lam * float(np.sum(z))


class GNormL1Nonneg:
    """proxoperators/normL1Nonneg.jl:29-42."""

    def __init__(self, lam=1.0):
        if lam < 0:
            raise ValueError("λ must be nonnegative")
        self.lam = float(lam)

    def prox(self, z, x, gamma):
        gl = gamma * self.lam
        z[:] = np.where(x >= gl, x - gl, 0.0)
        return self.lam * float(np.sum(z))


class GNormL0Box:
    """proxoperators/normL0Box.jl:33-58 -- quirks kept on purpose: threshold sqrt(gamma*lambda)
    (not sqrt(2*gamma*lambda)), and the over-the-box branch returns x[i], not u[i]."""

    def __init__(self, lam, u):
        if lam < 0:
            raise ValueError("parameter λ must be nonnegative")
        u = np.asarray(u, dtype=np.float64)
        if np.any(u < 0):
            raise ValueError("vector u must have nonnegative entries")
        self.lam, self.u = float(lam), u

    def prox(self, z, x, gamma):
        gl2 = gamma * self.lam
        u = self.u
        big = x > math.sqrt(gl2)
        over = x > u
        keep_over = x * x > gl2 + (u - x) * (u - x)
        keep = (u != 0) & big & (~over | keep_over)
        z[:] = np.where(keep, x, 0.0)
        return self.lam * float(np.count_nonzero(keep))


def _lp_subproblem(x, p, a, gamma, u=None, perturb=0.1):
    """src/proxoperators/normLpBox.jl:46-97 (u given, perturbation 0.1) / normLpNonneg.jl:43-85 (u None,
    perturbation 1): min a z^p + (z-x)^2/(2 gamma) over 0 <= z (<= u); Newton from the right of the inflection."""
    alpha = a * gamma
    if x <= 0 or (u is not None and u == 0):
        return 0.0
    zbar = (1 / (alpha * p * (1 - p))) ** (1 / (p - 2))
    psi_zbar = zbar + alpha * p * zbar ** (p - 1)
    if psi_zbar >= x:
        return 0.0
    z = zbar + perturb
    it = 0
    while it < 1000:
        dphi = z - x + alpha * p * z ** (p - 1)
        if abs(dphi) <= 1e-12:
            break
        it += 1
        ddphi = 1 + alpha * p * (p - 1) * z ** (p - 2)
        z -= dphi / ddphi
    phi_0 = 0.5 * x**2
    phi_z = 0.5 * (z - x) ** 2 + alpha * z**p
    if phi_0 <= phi_z:
        return 0.0
    if u is not None and z > u:
        phi_u = 0.5 * (u - x) ** 2 + alpha * u**p
        return u if phi_u < phi_0 else 0.0
    return z


class GNormLpBox:
    """src/proxoperators/normLpBox.jl:11-44."""

    def __init__(self, p, alpha, u):
        if p <= 0:
            raise ValueError("p must be positive")
        if p >= 1:
            raise ValueError("p must be smaller than one")
        if alpha < 0:
            raise ValueError("alpha must be nonnegative")
        u = np.asarray(u, dtype=np.float64)
        if np.any(u < 0):
            raise ValueError("vector u must have nonnegative entries")
        self.p, self.alpha, self.u = float(p), float(alpha), u

    def prox(self, z, x, gamma):
        for i in range(x.shape[0]):
            z[i] = _lp_subproblem(float(x[i]), self.p, self.alpha, gamma, float(self.u[i]), 0.1)
        return self.alpha * float(np.sum(z**self.p))


class GNormLpNonneg:
    """src/proxoperators/normLpNonneg.jl:12-41."""

    def __init__(self, p, alpha=1.0):
        if p <= 0:
            raise ValueError("p must be positive")
        if p >= 1:
            raise ValueError("p must be smaller than one")
        if alpha < 0:
            raise ValueError("alpha must be nonnegative")
        self.p, self.alpha = float(p), float(alpha)

    def prox(self, z, x, gamma):
        for i in range(x.shape[0]):
            z[i] = _lp_subproblem(float(x[i]), self.p, self.alpha, gamma, None, 1.0)
        return self.alpha * float(np.sum(z**self.p))


def _schatten_subproblem(x, p, a, gamma):
    """src/proxoperators/schattenNormLp.jl:60-97 solve_lp_quasi_norm_subproblem: the scalar problem of normLpBox.jl without
    the upper bound, tolerance 1e-14, perturbation 0.1."""
    alpha = a * gamma
    if x <= 0:
        return 0.0
    zbar = (1 / (alpha * p * (1 - p))) ** (1 / (p - 2))
    psi_zbar = zbar + alpha * p * zbar ** (p - 1)
    if psi_zbar >= x:
        return 0.0
    z = zbar + 0.1
    it = 0
    while it < 1000:
        dphi = z - x + alpha * p * z ** (p - 1)
        if abs(dphi) <= 1e-14:
            break
        it += 1
        ddphi = 1 + alpha * p * (p - 1) * z ** (p - 2)
        z -= dphi / ddphi
    phi_0 = 0.5 * x**2
    phi_z = 0.5 * (z - x) ** 2 + alpha * z**p
    return 0.0 if phi_0 <= phi_z else z


def check_and_reshape_as_matrix(x):
    """src/proxoperators/rank.jl:36-44 (column-major reshape of a vector of n^2 entries)"""
    x = np.asarray(x, dtype=np.float64)
    n = int(round(np.sqrt(x.size)))
    if n * n != x.size:
        raise ValueError("do not know how to reshape passed vector x")
    return x.reshape((n, n), order="F")


class GRank:
    """src/proxoperators/rank.jl:11-33 (LAPACK through numpy.linalg.svd, as Julia's `svd`)."""

    def __init__(self, lam=1.0):
        if lam < 0:
            raise ValueError("parameter λ must be nonnegative")
        self.lam = float(lam)

    def __call__(self, X):
        return self.lam * int(np.linalg.matrix_rank(np.asarray(X, dtype=np.float64)))   # rank.jl:24-26

    def prox(self, Y, X, gamma):
        U, S, Vt = np.linalg.svd(X, full_matrices=False)          # rank.jl:29
        over = S > np.sqrt(2 * gamma * self.lam)                  # :30
        Y[...] = U @ (np.diag(S * over) @ Vt)                     # :31
        return self.lam * float(np.sum(over))                     # :32


class GNuclearNorm:
    """src/proxoperators/nuclearNorm.jl:11-34."""

    def __init__(self, lam=1.0):
        if lam < 0:
            raise ValueError("parameter λ must be nonnegative")
        self.lam = float(lam)

    def __call__(self, X):
        return self.lam * float(np.sum(np.linalg.svd(X, compute_uv=False)))   # nuclearNorm.jl:24-27

    def prox(self, Y, X, gamma):
        U, S, Vt = np.linalg.svd(X, full_matrices=False)          # nuclearNorm.jl:30
        S = np.maximum(0.0, S - self.lam * gamma)                 # :31
        Y[...] = U @ (np.diag(S) @ Vt)                            # :32
        return self.lam * float(np.sum(S))                        # :33


class GSchattenNormLpPower:
    """src/proxoperators/schattenNormLp.jl:14-42."""

    def __init__(self, p, alpha=1.0):
        if p <= 0:
            raise ValueError("p must be positive")
        if p >= 1:
            raise ValueError("p must be smaller than one")
        if alpha < 0:
            raise ValueError("alpha must be nonnegative")
        self.p, self.alpha = float(p), float(alpha)

    def __call__(self, X):
        return self.alpha * float(np.sum(np.linalg.svd(X, compute_uv=False) ** self.p))   # :32-34

    def prox(self, Y, X, gamma):
        U, S, Vt = np.linalg.svd(X, full_matrices=False)          # schattenNormLp.jl:37-38
        S = np.array([_schatten_subproblem(float(v), self.p, self.alpha, gamma) for v in S])   # :39
        Y[...] = U @ (np.diag(S) @ Vt)                            # :40
        return self.alpha * float(np.sum(S**self.p))              # :41


class GRosenbrock:
    """demo/rosenbrock.jl:55-64: soft threshold on x[1] only."""

    def __init__(self, lam):
        self.lam = float(lam)

    def prox(self, z, x, gamma):
        gl = gamma * self.lam
        if abs(x[0]) <= gl:
            z[0] = 0.0
        else:
            z[0] = math.copysign(1.0, x[0]) * (abs(x[0]) - gl)
        z[1] = x[1]
        return self.lam * abs(z[0])


class GNonneg:
    """demo/mpvca.jl:90-93 NonsmoothCostMPVCA: z = max(0, x), returns 0."""

    def prox(self, z, x, gamma):
        z[:] = np.maximum(0.0, x)
        return 0.0


class GObstacleL1:
    """demo/obstacle.jl:67-81 NonsmoothCostObstacleL1."""

    def __init__(self, N):
        self.N = N

    def prox(self, y, x, gamma):
        N = self.N
        y[:N] = np.maximum(0.0, x[:N] - gamma)
        y[N:] = np.maximum(0.0, x[N:])
        a = y[N : 2 * N].copy()
        b = y[2 * N : 3 * N].copy()
        gt = a > b
        # a > b: zero block 3; a < b: zero block 2; tie: zero block 2
        y[2 * N : 3 * N] = np.where(gt, 0.0, b)
        y[N : 2 * N] = np.where(gt, a, 0.0)
        return float(np.sum(y[:N]))


class GObstacleL2:
    """demo/obstacle.jl:49-62 NonsmoothCostObstacleL2: y = max(0, x); per i: block 2 > block 3 zeroes block 3, else
    (tie included) zeroes block 2; returns 0."""

    def __init__(self, N):
        self.N = N

    def prox(self, y, x, gamma):
        N = self.N
        y[:] = np.maximum(0.0, x)
        a = y[N : 2 * N].copy()
        b = y[2 * N : 3 * N].copy()
        gt = a > b
        y[2 * N : 3 * N] = np.where(gt, 0.0, b)
        y[N : 2 * N] = np.where(gt, a, 0.0)
        return 0.0


# ----------------------------------------------------------------------------------------------
# Smooth costs f:  f(x), gradient!(dfx, f, x) -> f(x)
# ----------------------------------------------------------------------------------------------
class FZero:
    """ProximalOperators.Zero / proxoperators/zero.jl:12-20."""

    def value(self, x):
        return 0.0

    def gradient(self, dfx, x):
        dfx[:] = 0.0
        return 0.0


class FLeastSquares:
    """ProximalOperators.LeastSquares(A,b): 0.5||Ax-b||^2 (test_verbose.jl:22)."""

    def __init__(self, A, b):
        self.A, self.b = np.asarray(A, dtype=np.float64), np.asarray(b, dtype=np.float64)

    def value(self, x):
        r = self.A @ x - self.b
        return 0.5 * float(r @ r)

    def gradient(self, dfx, x):
        r = self.A @ x - self.b
        dfx[:] = self.A.T @ r
        return 0.5 * float(r @ r)


class FQuadratic:
    """ProximalOperators.Quadratic(Q,q): 0.5 x'Qx + q'x (test_nonconvex_qp.jl:14)."""

    def __init__(self, Q, q):
        self.Q, self.q = np.asarray(Q, dtype=np.float64), np.asarray(q, dtype=np.float64)

    def value(self, x):
        return 0.5 * float(x @ (self.Q @ x)) + float(self.q @ x)

    def gradient(self, dfx, x):
        Qx = self.Q @ x
        dfx[:] = Qx + self.q
        return 0.5 * float(x @ Qx) + float(self.q @ x)


class FRosenbrock:
    """demo/rosenbrock.jl:39-50."""

    def __init__(self, w):
        self.w = float(w)

    def value(self, x):
        return self.w * (x[1] + 1 - (x[0] + 1) ** 2) ** 2

    def gradient(self, dfx, x):
        tmp = x[1] + 1 - (x[0] + 1) ** 2
        dfx[0] = -4 * self.w * tmp * (x[0] + 1)
        dfx[1] = 2 * self.w * tmp
        return self.w * tmp**2


class FSqDist2:
    """demo/eitheror.jl:58-68 SmoothCostOR: ||x[1:2] - c||^2 (norm then square)."""

    def __init__(self, c):
        self.c = np.asarray(c, dtype=np.float64)

    def value(self, x):
        d0, d1 = x[0] - self.c[0], x[1] - self.c[1]
        return math.sqrt(d0 * d0 + d1 * d1) ** 2

    def gradient(self, dfx, x):
        dfx[:] = 0.0
        dfx[0] = 2 * (x[0] - self.c[0])
        dfx[1] = 2 * (x[1] - self.c[1])
        return self.value(x)


class FLinear2:
    """demo/mpvca.jl:77-87 SmoothCostMPVCA: dot(c, x[1:2])."""

    def __init__(self, c):
        self.c = np.asarray(c, dtype=np.float64)

    def value(self, x):
        return self.c[0] * x[0] + self.c[1] * x[1]

    def gradient(self, dfx, x):
        dfx[:] = 0.0
        dfx[0] = self.c[0]
        dfx[1] = self.c[1]
        return self.value(x)


class FObstacleL1:
    """demo/obstacle.jl:37-47 SmoothCostObstacleL1."""

    def __init__(self, N):
        self.N = N

    def value(self, x):
        N = self.N
        v = x[N : 2 * N]
        return 0.5 * float(np.sum(v**2)) - float(np.sum(v))

    def gradient(self, dfx, x):
        N = self.N
        dfx[:] = 0.0
        dfx[N : 2 * N] = x[N : 2 * N] - 1.0
        return self.value(x)


class FObstacleL2:
    """demo/obstacle.jl:21-35 SmoothCostObstacleL2: 0.5*norm(x[1:2N])^2 - sum(x[N+1:2N]); dfx = [x1; x2 - 1; 0]."""

    def __init__(self, N):
        self.N = N

    def value(self, x):
        N = self.N
        return 0.5 * float(np.linalg.norm(x[: 2 * N])) ** 2 - float(np.sum(x[N : 2 * N]))

    def gradient(self, dfx, x):
        N = self.N
        dfx[:] = 0.0
        dfx[:N] = x[:N]
        dfx[N : 2 * N] = x[N : 2 * N] - 1.0
        return self.value(x)


# ----------------------------------------------------------------------------------------------
# Constraint functions c:  eval!(cx, c, x), jtprod!(jtv, c, x, v)
# ----------------------------------------------------------------------------------------------
class CIdentity:
    """test/definitions/identityFunction.jl:3-13."""

    def eval(self, cx, x):
        cx[:] = x

    def jtprod(self, jtv, x, v):
        jtv[:] = v


class CDenseAffine:
    """demo/basispursuit.jl:38-49: cx = A*x - b, jtv = A'*v."""

    def __init__(self, A, b):
        self.A, self.b = A, b

    def eval(self, cx, x):
        cx[:] = self.A @ x - self.b

    def jtprod(self, jtv, x, v):
        jtv[:] = self.A.T @ v


class CPortfolio:
    """demo/portfolio.jl:46-58 ConstraintP: cx = [dot(mu, x); sum(x)], jtv = mu*v[1] + v[2]."""

    def __init__(self, mu):
        self.mu = np.asarray(mu, dtype=np.float64)

    def eval(self, cx, x):
        cx[0] = float(self.mu @ x)
        cx[1] = float(np.sum(x))

    def jtprod(self, jtv, x, v):
        jtv[:] = self.mu * v[0] + v[1]


class CRosenbrock:
    """demo/rosenbrock.jl:66-74."""

    def eval(self, cx, x):
        cx[0] = -x[0] - x[1]
        cx[1] = x[1] - x[0]

    def jtprod(self, jtv, x, v):
        jtv[0] = -v[0] - v[1]
        jtv[1] = v[1] - v[0]


class CEitherOr:
    """demo/eitheror.jl:92-104 ConstraintOR."""

    def eval(self, cx, x):
        cx[0] = 2 * x[1] - x[0] - 4
        cx[1] = 2 - x[0]
        cx[2] = 4 * x[1] - x[0] ** 2
        cx[3] = 10 - (x[0] - 3) ** 2 - (x[1] - 1) ** 2

    def jtprod(self, jtv, x, v):
        jtv[0] = -v[0] - v[1] - 2 * x[0] * v[2] + 2 * (3 - x[0]) * v[3]
        jtv[1] = 2 * v[0] + 4 * v[2] + 2 * (1 - x[1]) * v[3]


class CMpvca:
    """demo/mpvca.jl:110-118 ConstraintMPVCA."""

    def eval(self, cx, x):
        cx[0] = x[0]
        cx[1] = x[0] + x[1] - 5.0 * math.sqrt(2.0)
        cx[2] = x[1]
        cx[3] = x[0] + x[1] - 5.0

    def jtprod(self, jtv, x, v):
        jtv[0] = v[0] + v[1] + v[3]
        jtv[1] = v[1] + v[2] + v[3]


def _stencil_apply(v, grid):
    """A*v with A = SymTridiagonal(2,-1) (demo/obstacle.jl:97-103) when grid is None, or the 2-D
    5-point extension (4 on the diagonal, -1 neighbours, Dirichlet) on a (ny, nx) grid stored
    row-major -- restricted to one grid line it reproduces the 1-D operator's off-diagonals
    (config C4 is an extrapolation of the demo)."""
    if grid is None:
        out = 2.0 * v
        out[1:] -= v[:-1]
        out[:-1] -= v[1:]
        return out
    gy, gx = grid
    V = v.reshape(gy, gx)
    out = 4.0 * V
    out[:, 1:] -= V[:, :-1]
    out[:, :-1] -= V[:, 1:]
    out[1:, :] -= V[:-1, :]
    out[:-1, :] -= V[1:, :]
    return out.reshape(-1)


class CObstacle:
    """demo/obstacle.jl:92-113 ConstraintObstacle: cx = x1 + A x2 - x3."""

    def __init__(self, N, grid=None):
        self.N, self.grid = N, grid

    def eval(self, cx, x):
        N = self.N
        cx[:] = x[:N] + _stencil_apply(x[N : 2 * N], self.grid) - x[2 * N : 3 * N]

    def jtprod(self, jtv, x, v):
        N = self.N
        jtv[:N] = v
        jtv[N : 2 * N] = _stencil_apply(v, self.grid)
        jtv[2 * N : 3 * N] = -v


# ----------------------------------------------------------------------------------------------
# AugLagFun / NonsmoothCostFun (src/utilities)
# ----------------------------------------------------------------------------------------------
class CObstacleRed:
    """demo/obstacle.jl:128-146 ConstraintObstacleRed: cx = [x1 + A x2; x2]."""

    def __init__(self, N, grid=None):
        self.N, self.grid = N, grid

    def eval(self, cx, x):
        N = self.N
        cx[:N] = x[:N] + _stencil_apply(x[N : 2 * N], self.grid)
        cx[N : 2 * N] = x[N : 2 * N]

    def jtprod(self, jtv, x, v):
        N = self.N
        jtv[:N] = v[:N]
        jtv[N : 2 * N] = _stencil_apply(v[:N], self.grid) + v[N : 2 * N]


class GObstacleRedL1:
    """demo/obstacle.jl:83-91 NonsmoothCostObstacleRedL1."""

    def __init__(self, N):
        self.N = N

    def prox(self, y, x, gamma):
        N = self.N
        y[:N] = np.maximum(0.0, x[:N] - gamma)
        y[N : 2 * N] = x[N : 2 * N]
        return float(np.sum(y[:N]))


def problem_obstacle_red(N, grid=None):
    """demo/obstacle.jl:233-239 (obstacle_l1red): nx = ny = 2N, D = SetObstacleRed (CC pairs (i, i+N))."""
    return dict(f=FObstacleL1(N), g=GObstacleRedL1(N), c=CObstacleRed(N, grid), D=PairSet("cc", stride=N),
                nx=2 * N, ny=2 * N,
                kw=dict(subsolver_maxit=1_000_000,
                        subsolver=PanocOptions(minimum_gamma=EPS, maxit=1_000_000)))


def problem_obstacle_l2(N, grid=None, red=False):
    """demo/obstacle.jl:212-225: 'obstacle_l2' (nx = 3N, ny = N, D = {0}) and 'obstacle_l2red' (nx = ny = 2N, g = max(0, x),
    D = SetObstacleRed)."""
    kw = dict(subsolver_maxit=1_000_000, subsolver=PanocOptions(minimum_gamma=EPS, maxit=1_000_000))
    if red:
        return dict(f=FObstacleL2(N), g=GNonneg(), c=CObstacleRed(N, grid), D=PairSet("cc", stride=N), nx=2 * N, ny=2 * N, kw=kw)
    return dict(f=FObstacleL2(N), g=GObstacleL2(N), c=CObstacle(N, grid), D=ZeroSet(), nx=3 * N, ny=N, kw=kw)


class AugLagFun:
    """src/utilities/auglagfun.jl:11-101."""

    def __init__(self, f, c, D, mu, y, x):
        if np.any(mu <= 0):
            raise ValueError("parameters `mu` must be positive")
        self.f, self.c, self.D = f, c, D
        self.mu, self.y = mu, y  # aliases, as in the reference (auglagfun.jl:38-52)
        self.muy = mu * y
        self.musqy = 0.5 * float(np.sum(self.muy * y))
        self.cx = np.empty_like(y)
        self.s = np.empty_like(y)
        self.yupd = np.empty_like(y)
        self.fx = 0.0
        self.dfx = np.empty_like(x)
        self.jtv = np.empty_like(x)
        self.n_grad = 0

    def update(self, mu, y):  # AugLagUpdate! auglagfun.jl:91-101
        if np.any(mu <= 0):
            raise ValueError("parameters `mu` must be positive")
        self.mu[:] = mu
        self.y[:] = y
        self.muy[:] = self.mu * self.y
        self.musqy = 0.5 * float(np.sum(self.muy * self.y))

    def _common(self, x):
        self.c.eval(self.cx, x)
        self.yupd[:] = self.cx + self.muy
        self.D.proj(self.s, self.yupd)
        self.yupd -= self.s
        lx = 0.5 * float(np.sum(self.yupd**2 / self.mu))
        self.yupd /= self.mu
        return lx

    def value(self, x):  # auglagfun.jl:58-69
        lx = self._common(x)
        self.fx = self.f.value(x)
        lx += self.fx
        lx -= self.musqy
        return lx

    def gradient(self, dlx, x):  # auglagfun.jl:73-86
        self.n_grad += 1
        lx = self._common(x)
        self.fx = self.f.gradient(self.dfx, x)
        lx += self.fx
        lx -= self.musqy
        self.c.jtprod(self.jtv, x, self.yupd)
        dlx[:] = self.dfx + self.jtv
        return lx


class NonsmoothCostFun:
    """src/utilities/nonsmoothcostfun.jl:1-22."""

    def __init__(self, g):
        self.g, self.gamma, self.gz = g, 0.0, 0.0
        self.n_prox = 0

    def prox(self, z, x, gamma):
        self.n_prox += 1
        gz = self.g.prox(z, x, gamma)
        self.gamma, self.gz = gamma, gz
        return gz


def default_dual_safeguard(y, cx=None):
    """src/utilities/safeguards.jl:2-10."""
    y[:] = np.maximum(-1e20, np.minimum(y, 1e20))


def default_penalty_parameter(mu, cx, proj_cx, objx):
    """src/utilities/safeguards.jl:13-18."""
    mu[:] = np.maximum(1.0, 0.5 * (cx - proj_cx) ** 2) / max(1.0, objx)
    mu *= 0.1
    mu[:] = np.maximum(1e-8, np.minimum(mu, 1e8))


# ----------------------------------------------------------------------------------------------
# L-BFGS operator and PANOCplus (third-party ProximalAlgorithms.jl)
# ----------------------------------------------------------------------------------------------
def _dot(a, b):
    return float(np.dot(a, b))


def _norm(a):
    return math.sqrt(_dot(a, a))


class LBFGS:
    """Appendix A 'L-BFGS operator' (ring buffers, 1-based curridx kept to mirror upstream)."""

    def __init__(self, n, M=5):
        self.M = M
        self.s_M = [np.zeros(n) for _ in range(M)]
        self.y_M = [np.zeros(n) for _ in range(M)]
        self.ys_M = [0.0] * M
        self.alphas = [0.0] * M
        self.currmem = 0
        self.curridx = 0
        self.H = 1.0

    def reset(self):
        self.currmem, self.curridx, self.H = 0, 0, 1.0

    def update(self, s, yv):
        if self.M == 0:
            return
        ys = _dot(s, yv)
        if ys > 0:
            self.curridx += 1
            if self.curridx > self.M:
                self.curridx = 1
            self.currmem = min(self.currmem + 1, self.M)
            i = self.curridx - 1
            self.ys_M[i] = ys
            self.s_M[i][:] = s
            self.y_M[i][:] = yv
            self.H = ys / _dot(yv, yv)

    def mul(self, v):
        d = v.copy()
        idx = self.curridx
        for _ in range(self.currmem):
            i = idx - 1
            self.alphas[i] = _dot(self.s_M[i], d) / self.ys_M[i]
            d -= self.alphas[i] * self.y_M[i]
            idx -= 1
            if idx == 0:
                idx = self.M
        d *= self.H
        for _ in range(self.currmem):
            idx += 1
            if idx > self.M:
                idx = 1
            i = idx - 1
            beta = _dot(self.y_M[i], d) / self.ys_M[i]
            d += (self.alphas[i] - beta) * self.s_M[i]
        return d


def _f_model(fx, grad, res, L):
    """fb_tools: f_x - <grad,res> + (L/2)*norm(res)^2  (norm then square, as upstream)."""
    return fx - _dot(grad, res) + (L / 2) * _norm(res) ** 2


@dataclass
class PanocOptions:
    """Appendix A.1 defaults; demo closures override maxit / minimum_gamma / memory."""

    alpha: float = 0.95
    beta: float = 0.5
    minimum_gamma: float = 1e-7
    max_backtracks: int = 20
    memory: int = 5
    maxit: int = 1000
    tol: float = 1e-8
    trace: list | None = None  # if a list, per-iteration dicts are appended (golden dumps)


def panocplus(f: AugLagFun, g: NonsmoothCostFun, x0, opt: PanocOptions):
    """PANOCplus with adaptive gamma, A = I.
    Returns (z, k) with k counting the initial state as 1."""
    n = x0.size
    a, minimum_gamma = opt.alpha, opt.minimum_gamma
    x = x0.copy()
    grad_x = np.empty(n)
    f_x = f.gradient(grad_x, x)
    # lower_bound_smoothness_constant (A.4): one extra AL-gradient at x .+ 1
    tmpg = np.empty(n)
    f.gradient(tmpg, x + 1.0)
    Lest = _norm(tmpg - grad_x) / math.sqrt(n)
    with np.errstate(divide="ignore"):
        gamma = float(np.float64(a) / np.float64(Lest))
    y = x - gamma * grad_x
    z = np.empty(n)
    g_z = g.prox(z, y, gamma)
    res = x - z
    H = LBFGS(n, opt.memory)
    grad_z = np.empty(n)
    # backtrack_stepsize! (A.4)
    f_up = _f_model(f_x, grad_x, res, a / gamma)
    f_z = f.gradient(grad_z, z)
    while f_z > f_up + 10 * EPS * (1 + abs(f_z)) and gamma >= minimum_gamma:
        gamma /= 2
        y[:] = x - gamma * grad_x
        g_z = g.prox(z, y, gamma)
        res[:] = x - z
        f_up = _f_model(f_x, grad_x, res, a / gamma)
        f_z = f.gradient(grad_z, z)
    x_prev = np.empty(n)
    res_prev = np.empty(n)
    x_d = np.empty(n)
    z_curr = np.empty(n)
    tau = 1.0
    k = 1
    while True:
        stopnorm = float(np.max(np.abs(res / gamma - grad_x + grad_z))) if n else 0.0
        if opt.trace is not None:
            opt.trace.append(
                dict(k=k, gamma=gamma, tau=tau, f_x=f_x, f_z=f_z, g_z=g_z, stopnorm=stopnorm,
                     x=x.copy(), z=z.copy(), res=res.copy(), grad_x=grad_x.copy(),
                     grad_z=grad_z.copy())
            )
        if k >= opt.maxit or stopnorm <= opt.tol:
            return z, k
        # ---- one iteration (A.5)
        x_prev[:] = x
        res_prev[:] = res
        FBE_x = _f_model(f_x, grad_x, res, a / gamma) + g_z
        d = H.mul(-res)
        x_d[:] = x + d
        f_x = f.gradient(grad_x, x_d)
        x[:] = x_d
        z_curr[:] = z
        tau = 1.0
        sigma = opt.beta * (0.5 / gamma) * (1 - a)
        thresh = FBE_x - sigma * _norm(res) ** 2 + 10 * EPS * (1 + abs(FBE_x))
        for kb in range(1, opt.max_backtracks + 1):
            y[:] = x - gamma * grad_x
            g_z = g.prox(z, y, gamma)
            res[:] = x - z
            f_up = _f_model(f_x, grad_x, res, a / gamma)
            f_z = f.gradient(grad_z, z)
            if f_z > f_up + 10 * EPS * (1 + abs(f_z)) and gamma >= minimum_gamma:
                gamma *= 0.5
                H.reset()
                continue
            if f_up + g_z <= thresh or kb >= opt.max_backtracks:
                break
            tau = 0.0 if kb >= opt.max_backtracks - 1 else tau / 2
            x[:] = tau * x_d + (1 - tau) * z_curr
            f_x = f.gradient(grad_x, x)
        H.update(x - x_prev, res - res_prev)
        k += 1


# ----------------------------------------------------------------------------------------------
# ALPS outer loop (src/algorithms/alps.jl:7-117)
# ----------------------------------------------------------------------------------------------
@dataclass
class AlpsResult:
    x: np.ndarray
    y: np.ndarray
    tot_it: int
    tot_inner_it: int
    elapsed_time: float
    status: str
    inner_tol: float
    norm_res_prim: float | None
    s: np.ndarray
    mu: np.ndarray
    objective: float = float("nan")
    n_grad: int = 0
    n_prox: int = 0
    sub_its: list = field(default_factory=list)

    def as_tuple(self):
        return (self.x, self.y, self.tot_it, self.tot_inner_it, self.elapsed_time, self.status,
                self.inner_tol, self.norm_res_prim, self.s, self.mu)


def alps(f, g, c, D, x0, y0, *, tol=1e-6, tol_prim=None, tol_dual=None, inner_tol=None,
         maxit=100, theta_penalty=0.8, kappa_penalty=0.5, kappa_tol=0.1,
         subsolver=None, subsolver_maxit=1_000_000_000, max_inner_total=None):
    """alps.jl:7-117.  ``subsolver`` is a PanocOptions template (its ``tol`` is overwritten every
    outer iteration by ``inner_tol``, alps.jl:64).  ``max_inner_total`` (oracle-only, for bounded
    CPU-baseline timing) stops after that many inner iterations in total."""
    start = time.time()
    tol_prim = tol if tol_prim is None else tol_prim
    tol_dual = tol if tol_dual is None else tol_dual
    inner_tol = float(np.cbrt(tol_dual)) if inner_tol is None else inner_tol
    sub = subsolver if subsolver is not None else PanocOptions()
    x0 = np.asarray(x0, dtype=np.float64)
    y0 = np.asarray(y0, dtype=np.float64)
    x = np.empty_like(x0)
    y = np.empty_like(y0)
    cx = np.empty_like(y0)
    s = np.empty_like(y0)
    mu = np.empty_like(y0)
    gFun = NonsmoothCostFun(g)
    gFun.prox(x, x0, EPS)
    objx = f.value(x) + gFun.gz
    c.eval(cx, x)
    D.proj(s, cx)
    default_penalty_parameter(mu, cx, s, objx)
    y[:] = y0
    norm_res_prim = None
    norm_res_prim_old = None
    alFun = AugLagFun(f, c, D, mu, y, x)
    tot_it = 0
    tot_inner_it = 0
    solved = False
    tired = tot_it >= maxit
    broken = math.isnan(objx)
    sub_its = []
    can_stop = solved or tired or broken
    while not can_stop:
        tot_it += 1
        default_dual_safeguard(y, cx)
        opt = PanocOptions(alpha=sub.alpha, beta=sub.beta, minimum_gamma=sub.minimum_gamma,
                           max_backtracks=sub.max_backtracks, memory=sub.memory,
                           maxit=sub.maxit, tol=inner_tol, trace=sub.trace)
        if max_inner_total is not None:
            opt.maxit = max(1, min(opt.maxit, max_inner_total - tot_inner_it))
        alFun.update(mu, y)
        sub_sol, sub_it = panocplus(alFun, gFun, x, opt)
        x[:] = sub_sol
        objx = alFun.fx + gFun.gz
        tot_inner_it += sub_it
        sub_its.append(sub_it)
        sub_solved = sub_it < subsolver_maxit
        c.eval(cx, x)
        y[:] = cx + alFun.muy
        D.proj(s, y)
        y -= s
        y /= mu
        norm_res_prim_old = norm_res_prim
        norm_res_prim = float(np.max(np.abs(cx - s))) if cx.size else 0.0
        solved = (inner_tol <= tol_dual and sub_solved) and (norm_res_prim <= tol_prim)
        tired = tot_it >= maxit
        broken = math.isnan(objx)
        can_stop = solved or tired or broken
        if max_inner_total is not None and tot_inner_it >= max_inner_total:
            can_stop = True
        if not can_stop:
            if norm_res_prim_old is None:
                pass
            elif norm_res_prim > max(theta_penalty * norm_res_prim_old, tol_prim):
                mu *= kappa_penalty
            inner_tol = max(kappa_tol * inner_tol, tol_dual)
    elapsed = time.time() - start
    status = "first_order" if solved else "max_iter" if tired else "exception" if broken else "unknown"
    return AlpsResult(x, y, tot_it, tot_inner_it, elapsed, status, inner_tol, norm_res_prim, s, mu,
                      objective=objx, n_grad=alFun.n_grad, n_prox=gFun.n_prox, sub_its=sub_its)


# ----------------------------------------------------------------------------------------------
# ALS: the augmented-Lagrangian method on the stacked variable [x; s] with explicit slacks
# ----------------------------------------------------------------------------------------------
class AugLagFunSlack:
    """src/utilities/auglagfunslack.jl:15-126: al_f(x, s) = f(x) + 1/(2 mu) ||c(x) + mu y - s||^2 - (mu/2)||y||^2."""

    def __init__(self, f, c, mu, y, x):
        if np.any(mu <= 0):
            raise ValueError("parameters `mu` must be positive")          # auglagfunslack.jl:33-35
        self.f, self.c = f, c
        self.nx, self.ny = x.shape[0], y.shape[0]
        self.mu, self.y = mu, y
        self.muy = mu * y
        self.musqy = 0.5 * float(np.sum(self.muy * y))
        self.cx = np.empty_like(y)
        self.yupd = np.empty_like(y)
        self.dfx = np.empty_like(x)
        self.jtv = np.empty_like(x)
        self.fx = 0.0
        self.n_grad = 0

    def update(self, mu, y):  # AugLagUpdate! auglagfunslack.jl:114-126
        if np.any(mu <= 0):
            raise ValueError("parameters `mu` must be positive")
        if y.shape[0] != self.ny:
            raise ValueError("wrong length of passed argument y")
        self.mu[:] = mu
        self.y[:] = y
        self.muy[:] = self.mu * self.y
        self.musqy = 0.5 * float(np.sum(self.muy * self.y))

    def value(self, xs):  # auglagfunslack.jl:58-72
        if xs.shape[0] != self.nx + self.ny:
            raise ValueError("wrong length of passed argument xs")
        x, s = xs[: self.nx], xs[self.nx :]
        self.fx = self.f.value(x)
        self.c.eval(self.cx, x)
        self.yupd[:] = self.cx + self.muy - s
        Fxs = 0.5 * float(np.sum(self.yupd**2 / self.mu))
        Fxs += self.fx
        Fxs -= self.musqy
        self.yupd[:] = self.y + (self.cx - s) / self.mu
        return Fxs

    def gradient(self, dFxs, xs):  # auglagfunslack.jl:77-97
        if xs.shape[0] != self.nx + self.ny or dFxs.shape[0] != self.nx + self.ny:
            raise ValueError("wrong length of passed argument xs")
        self.n_grad += 1
        x, s = xs[: self.nx], xs[self.nx :]
        self.fx = self.f.gradient(self.dfx, x)
        self.c.eval(self.cx, x)
        Fxs = 0.5 * float(np.sum((self.cx + self.muy - s) ** 2 / self.mu))
        Fxs += self.fx
        Fxs -= self.musqy
        self.yupd[:] = self.y + (self.cx - s) / self.mu
        self.c.jtprod(self.jtv, x, self.yupd)
        dFxs[: self.nx] = self.dfx + self.jtv
        dFxs[self.nx :] = -self.yupd
        return Fxs


class NonsmoothCostFunSlack:
    """src/utilities/auglagfunslack.jl:130-164: prox of g on x, projection onto D on s."""

    def __init__(self, g, D, nx, ny):
        self.g, self.D, self.nx, self.ny = g, D, nx, ny
        self.gamma, self.gz = 0.0, 0.0
        self.n_prox = 0

    def prox(self, z, xs, gamma):
        if xs.shape[0] != self.nx + self.ny or z.shape[0] != self.nx + self.ny:
            raise ValueError("wrong length of passed argument xs")
        self.n_prox += 1
        self.gamma = gamma
        zx = np.empty(self.nx)
        gz = self.g.prox(zx, xs[: self.nx].copy(), gamma)
        z[: self.nx] = zx
        self.gz = gz
        zs = np.empty(self.ny)
        self.D.proj(zs, xs[self.nx :].copy())
        z[self.nx :] = zs
        return gz


def als(f, g, c, D, x0, y0, *, tol=1e-6, tol_prim=None, tol_dual=None, inner_tol=None,
        maxit=100, theta_penalty=0.8, kappa_penalty=0.5, kappa_tol=0.1,
        subsolver=None, subsolver_maxit=1_000_000_000):
    """src/algorithms/als.jl:7-120 (same keywords and return tuple as alps)."""
    start = time.time()
    tol_prim = tol if tol_prim is None else tol_prim
    tol_dual = tol if tol_dual is None else tol_dual
    inner_tol = float(np.cbrt(tol_dual)) if inner_tol is None else inner_tol
    sub = subsolver if subsolver is not None else PanocOptions()
    x0 = np.asarray(x0, dtype=np.float64)
    y0 = np.asarray(y0, dtype=np.float64)
    nx, ny = x0.shape[0], y0.shape[0]
    x = np.empty_like(x0)
    s = np.empty_like(y0)
    xSlack = np.zeros(nx + ny)
    y = np.empty_like(y0)
    cx = np.empty_like(y0)
    mu = np.empty_like(y0)
    gFun = NonsmoothCostFun(g)
    gFun.prox(x, x0, EPS)
    objx = f.value(x) + gFun.gz
    c.eval(cx, x)
    D.proj(s, cx)
    default_penalty_parameter(mu, cx, s, objx)
    y[:] = y0
    norm_res_prim = None
    norm_res_prim_old = None
    fSlack = AugLagFunSlack(f, c, mu, y, x)
    gSlack = NonsmoothCostFunSlack(g, D, nx, ny)
    tot_it = 0
    tot_inner_it = 0
    solved = False
    tired = tot_it >= maxit
    broken = math.isnan(objx)
    sub_its = []
    can_stop = solved or tired or broken
    while not can_stop:
        tot_it += 1
        default_dual_safeguard(y, cx)
        opt = PanocOptions(alpha=sub.alpha, beta=sub.beta, minimum_gamma=sub.minimum_gamma,
                           max_backtracks=sub.max_backtracks, memory=sub.memory,
                           maxit=sub.maxit, tol=inner_tol, trace=sub.trace)
        fSlack.update(mu, y)
        xSlack[:nx] = x
        xSlack[nx:] = s
        sub_sol, sub_it = panocplus(fSlack, gSlack, xSlack, opt)
        xSlack[:] = sub_sol
        x[:] = xSlack[:nx]
        s[:] = xSlack[nx:]
        objx = f.value(x) + gSlack.gz
        tot_inner_it += sub_it
        sub_its.append(sub_it)
        sub_solved = sub_it < subsolver_maxit
        c.eval(cx, x)
        y[:] = y + (cx - s) / mu
        norm_res_prim_old = norm_res_prim
        norm_res_prim = float(np.max(np.abs(cx - s))) if cx.size else 0.0
        solved = (inner_tol <= tol_dual and sub_solved) and (norm_res_prim <= tol_prim)
        tired = tot_it >= maxit
        broken = math.isnan(objx)
        can_stop = solved or tired or broken
        if not can_stop:
            if norm_res_prim_old is None:
                pass
            elif norm_res_prim > max(theta_penalty * norm_res_prim_old, tol_prim):
                mu *= kappa_penalty
            inner_tol = max(kappa_tol * inner_tol, tol_dual)
    elapsed = time.time() - start
    status = "first_order" if solved else "max_iter" if tired else "exception" if broken else "unknown"
    return AlpsResult(x, y, tot_it, tot_inner_it, elapsed, status, inner_tol, norm_res_prim, s, mu,
                      objective=objx, n_grad=fSlack.n_grad, n_prox=gSlack.n_prox, sub_its=sub_its)


# ----------------------------------------------------------------------------------------------
# Counter-based synthetic data (shared bit-for-bit with the CUDA generator, csrc/bz_rng.cuh)
# ----------------------------------------------------------------------------------------------
_PHILOX_M0 = np.uint64(0xD2511F53)
_PHILOX_M1 = np.uint64(0xCD9E8D57)
_PHILOX_W0 = 0x9E3779B9
_PHILOX_W1 = 0xBB67AE85
_MASK32 = np.uint64(0xFFFFFFFF)
# sqrt(3) * 2^-32 as one double literal: the only rounding in the generator is one multiply.
BZ_RNG_SCALE = float.fromhex("0x1.bb67ae8584caap-32")


def philox4x32_10(ctr_lo, ctr_hi, key0, key1):
    """Philox-4x32-10 on uint64 numpy arrays (counter = (lo32, hi32, 0, 0)); returns 4 uint32
    streams as uint64 arrays."""
    c0 = ctr_lo.astype(np.uint64) & _MASK32
    c1 = ctr_hi.astype(np.uint64) & _MASK32
    c2 = np.zeros_like(c0)
    c3 = np.zeros_like(c0)
    k0, k1 = int(key0) & 0xFFFFFFFF, int(key1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = _PHILOX_M0 * c0
        p1 = _PHILOX_M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & _MASK32
        hi1, lo1 = p1 >> np.uint64(32), p1 & _MASK32
        n0 = hi1 ^ c1 ^ np.uint64(k0)
        n1 = lo1
        n2 = hi0 ^ c3 ^ np.uint64(k1)
        n3 = lo0
        c0, c1, c2, c3 = n0, n1, n2, n3
        k0 = (k0 + _PHILOX_W0) & 0xFFFFFFFF
        k1 = (k1 + _PHILOX_W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def bz_randn_like(seed, start, count, stream=0):
    """Zero-mean unit-variance iid entries for element indices [start, start+count):
    (u0+u1+u2+u3 - 2*(2^32-1)) * sqrt(3)*2^-32  (Irwin-Hall(4), exact integer sum, one multiply),
    so host and device agree bit-for-bit.  Not Gaussian in the tails; it stands in for
    demo/basispursuit.jl:57 ``randn`` as synthetic data."""
    idx = np.arange(start, start + count, dtype=np.uint64)
    u0, u1, u2, u3 = philox4x32_10(idx & _MASK32, idx >> np.uint64(32), seed, stream)
    ssum = (u0 + u1 + u2 + u3).astype(np.int64) - np.int64(2 * 0xFFFFFFFF)
    return ssum.astype(np.float64) * BZ_RNG_SCALE


def bz_sparse_randn(seed, count, density=0.1):
    """Bernoulli(density) support times the entries above (stands in for ``sprandn``,
    demo/basispursuit.jl:58): element i is nonzero iff u0(stream 1) < density*2^32."""
    idx = np.arange(0, count, dtype=np.uint64)
    u0, _, _, _ = philox4x32_10(idx & _MASK32, idx >> np.uint64(32), seed, 1)
    keep = u0 < np.uint64(int(density * 4294967296.0))
    vals = bz_randn_like(seed, 0, count, stream=2)
    return np.where(keep, vals, 0.0)


def make_basis_pursuit(m, n, seed=0, density=0.1):
    """Synthetic basis pursuit instance of SURVEY §8(d) C2/C5: A (row-major m x n), xtrue, b=A xtrue."""
    A = bz_randn_like(seed, 0, m * n).reshape(m, n)
    xtrue = bz_sparse_randn(seed + 1, n, density)
    b = A @ xtrue
    return A, xtrue, b


# ----------------------------------------------------------------------------------------------
# Named problem builders (the oracle catalogue of SURVEY §8 a11-a14 + the reference's tests)
# ----------------------------------------------------------------------------------------------
def problem_rosenbrock():
    """demo/rosenbrock.jl:97-136: w=10, lambda=1; tol=1e-8, inner_tol=1, gamma_min=1e-32."""
    return dict(f=FRosenbrock(10.0), g=GRosenbrock(1.0), c=CRosenbrock(), D=PairSet("eitheror"),
                nx=2, ny=2,
                kw=dict(tol=1e-8, inner_tol=1.0, subsolver_maxit=1_000_000_000,
                        subsolver=PanocOptions(minimum_gamma=1e-32, maxit=1_000_000_000)))


def problem_eitheror(kind="eitheror"):
    """demo/eitheror.jl:133-209 ('eitheror' or 'xor')."""
    return dict(f=FSqDist2([8.0, -3.0]), g=GZero(), c=CEitherOr(), D=PairSet(kind), nx=2, ny=4,
                kw=dict(tol=1e-8, inner_tol=1.0, subsolver_maxit=1000,
                        subsolver=PanocOptions(minimum_gamma=EPS, maxit=1000)))


def problem_mpvca():
    """demo/mpvca.jl:155-218."""
    return dict(f=FLinear2([4.0, 2.0]), g=GNonneg(), c=CMpvca(), D=PairSet("vc"), nx=2, ny=4,
                kw=dict(tol=1e-8, inner_tol=1.0, subsolver_maxit=1000,
                        subsolver=PanocOptions(minimum_gamma=EPS, maxit=1000)))


class CEitherOrFullSlack(CEitherOr):
    """demo/eitheror.jl:106-119 ConstraintORfullslack: x = [x1, x2, s1..s4], c(x) = c_OR(x1, x2) - s."""

    def eval(self, cx, x):
        CEitherOr.eval(self, cx, x)
        cx -= x[2:6]

    def jtprod(self, jtv, x, v):
        CEitherOr.jtprod(self, jtv, x, v)
        jtv[2:6] = -v[0:4]


class CMpvcaSlack:
    """demo/mpvca.jl:120-128 ConstraintMPVCAslack: x = [x1, x2, s1, s2]."""

    def eval(self, cx, x):
        cx[0] = x[0] + x[1] - 5.0 * math.sqrt(2.0) - x[2]
        cx[1] = x[0] + x[1] - 5.0 - x[3]

    def jtprod(self, jtv, x, v):
        jtv[0] = v[0] + v[1]
        jtv[1] = v[0] + v[1]
        jtv[2] = -v[0]
        jtv[3] = -v[1]


class CMpvcaFullSlack(CMpvca):
    """demo/mpvca.jl:130-143 ConstraintMPVCAfullslack: x = [x1, x2, s1..s4]."""

    def eval(self, cx, x):
        cx[0] = x[0] - x[2]
        cx[1] = x[0] + x[1] - 5.0 * math.sqrt(2.0) - x[3]
        cx[2] = x[1] - x[4]
        cx[3] = x[0] + x[1] - 5.0 - x[5]

    def jtprod(self, jtv, x, v):
        CMpvca.jtprod(self, jtv, x, v)
        jtv[2:6] = -v[0:4]


class GPairSlack:
    """The nonsmooth terms of the demos' slack variants: the identity or max(0, .) on (x1, x2) and a 2-element
    projection on pairs of slacks -- demo/eitheror.jl:76-90 (EITHEROR / XOR on (3,4), (5,6)), demo/mpvca.jl:95-100
    (VC on (1,3), (2,4)) and :102-108 (max(0,.) on 1:2, VC on (3,4), (5,6)); all return 0."""

    def __init__(self, kind, pairs, nonneg_head=False):
        self.kind, self.pairs, self.nonneg_head = kind, pairs, nonneg_head

    def prox(self, z, x, gamma):
        z[:] = x
        if self.nonneg_head:
            z[0:2] = np.maximum(0.0, x[0:2])
        fn = {"eitheror": project_onto_EITHEROR_set, "xor": project_onto_XOR_set, "vc": project_onto_VC_set}[self.kind]
        for i, j in self.pairs:
            z[i], z[j] = fn(float(x[i]), float(x[j]))
        return 0.0


def problem_eitheror_fullslack(kind="eitheror"):
    """demo/eitheror.jl:147-165 ('eitheror_fullslack' / 'xor_fullslack'): nx = 6, ny = 4, D = ZeroSet."""
    return dict(f=FSqDist2([8.0, -3.0]), g=GPairSlack(kind, [(2, 3), (4, 5)]), c=CEitherOrFullSlack(), D=ZeroSet(), nx=6, ny=4,
                kw=dict(tol=1e-8, inner_tol=1.0, subsolver_maxit=1000,
                        subsolver=PanocOptions(minimum_gamma=EPS, maxit=1000)))


def problem_mpvca_slack(full=False):
    """demo/mpvca.jl:165-176 ('mpvca_slack': nx = 4, ny = 2; 'mpvca_fullslack': nx = 6, ny = 4), D = ZeroSet."""
    if full:
        g, c, nx, ny = GPairSlack("vc", [(2, 3), (4, 5)], nonneg_head=True), CMpvcaFullSlack(), 6, 4
    else:
        g, c, nx, ny = GPairSlack("vc", [(0, 2), (1, 3)]), CMpvcaSlack(), 4, 2
    return dict(f=FLinear2([4.0, 2.0]), g=g, c=c, D=ZeroSet(), nx=nx, ny=ny,
                kw=dict(tol=1e-8, inner_tol=1.0, subsolver_maxit=1000,
                        subsolver=PanocOptions(minimum_gamma=EPS, maxit=1000)))


def slack_start(name, X):
    """The demos' starting points for the slack variants: [x; c(x)] (demo/eitheror.jl:241-250, demo/mpvca.jl:244-259)."""
    X = np.asarray(X, dtype=np.float64).reshape(-1, 2)
    x1, x2 = X[:, 0], X[:, 1]
    if name in ("eitheror_fullslack", "xor_fullslack"):
        return np.stack([x1, x2, 2 * x2 - x1 - 4, 2 - x1, 4 * x2 - x1**2, 10 - (x1 - 3) ** 2 - (x2 - 1) ** 2], axis=1)
    if name == "mpvca_slack":
        return np.stack([x1, x2, x1 + x2 - 5 * math.sqrt(2), x1 + x2 - 5], axis=1)
    if name == "mpvca_fullslack":
        return np.stack([x1, x2, x1, x1 + x2 - 5 * math.sqrt(2), x2, x1 + x2 - 5], axis=1)
    return X


def make_portfolio(n, seed=0, factors=None):
    """Synthetic stand-in for the demo's `portfolio_data/dim200/*/{Q,rho,mu,ub}.csv` (demo/portfolio.jl:71-90; the data
    files are not part of the reference repository): a factor-model covariance Q = F'F/k + diag(d) (symmetric positive
    definite), expected returns mu in [0.02, 0.12], a return floor rho at the median of mu and holding caps ub in
    [0.05, 0.3] (so at least four assets are held).  Entries from the counter-based generator above: the same bits on
    every machine."""
    k = factors or max(4, n // 10)
    F = bz_randn_like(seed, 0, k * n).reshape(k, n)
    d = 0.05 + 0.1 * np.abs(bz_randn_like(seed, 0, n, stream=3))
    Q = (F.T @ F) / k + np.diag(d)
    Q = 0.5 * (Q + Q.T)
    u = np.abs(bz_randn_like(seed, 0, n, stream=4))
    mu = 0.02 + 0.1 * (u / (1.0 + u))
    rho = float(np.median(mu))
    w = np.abs(bz_randn_like(seed, 0, n, stream=5))
    ub = 0.05 + 0.25 * (w / (1.0 + w))
    return Q, rho, mu, ub


def problem_portfolio(Q, rho, mu, ub, g="l1", a=100.0, p=0.5):
    """demo/portfolio.jl:117-166: f = SmoothCostP(Q), c = ConstraintP(mu), D = SetP(rho) and g = NormL1Box(a, u = ub)
    ('l1'), NormLpPowerBox(0.5, a, u = ub) ('lp') or NormL0Box(a, u = ub) ('l0'); solved with `als`, x0 = ones(n),
    y0 = zeros(2), PANOCplus(LBFGS(5), maxit = 1000, minimum_gamma = 1e-32)."""
    n = len(mu)
    gg = {"l1": lambda: GNormL1Box(a, ub), "lp": lambda: GNormLpBox(p, a, ub), "l0": lambda: GNormL0Box(a, ub)}[g]()
    return dict(f=FQuadratic(Q, np.zeros(n)), g=gg, c=CPortfolio(mu), D=SetPortfolio(rho), nx=n, ny=2,
                kw=dict(subsolver=PanocOptions(minimum_gamma=1e-32, maxit=1000)))


def problem_basis_pursuit(A, b, g="l1"):
    """demo/basispursuit.jl:62-89."""
    return dict(f=FZero(), g=GNormL1(1.0) if g == "l1" else GNormL0(1.0), c=CDenseAffine(A, b),
                D=ZeroSet(), nx=A.shape[1], ny=A.shape[0],
                kw=dict(subsolver_maxit=100_000,
                        subsolver=PanocOptions(minimum_gamma=EPS, maxit=100_000)))


def problem_obstacle(N, grid=None):
    """demo/obstacle.jl:226-232 (obstacle_l1); grid=(gy,gx) selects the 2-D extension (C4)."""
    return dict(f=FObstacleL1(N), g=GObstacleL1(N), c=CObstacle(N, grid), D=ZeroSet(),
                nx=3 * N, ny=N,
                kw=dict(subsolver_maxit=1_000_000,
                        subsolver=PanocOptions(minimum_gamma=EPS, maxit=1_000_000)))


def solve(problem, x0, y0=None, **overrides):
    kw = dict(problem["kw"])
    kw.update(overrides)
    if y0 is None:
        y0 = np.zeros(problem["ny"])
    return alps(problem["f"], problem["g"], problem["c"], problem["D"], x0, y0, **kw)
