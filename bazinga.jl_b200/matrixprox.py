"""Host mirror of the reference's matrix prox operators (src/proxoperators/rank.jl, nuclearNorm.jl, schattenNormLp.jl):
``Rank(lam)``, ``NuclearNorm(lam)``, ``SchattenNormLpPower(p, alpha=1)`` with the reference's constructor errors, the call
``f(X)`` and ``prox(f, X, gamma) -> (Y, f(Y))`` for ``prox!(Y, f, X, gamma)``.  The factorisation runs on the device
(bz_prox_matrix, csrc/bz_matprox.cu: one-sided Jacobi, one CTA per matrix); a leading axis is a batch of matrices.  Vector
input is reshaped like ``check_and_reshape_as_matrix`` (rank.jl:36-44: column-major, must be a perfect square).  There is
no CPU path."""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import _capi as K


def check_and_reshape_as_matrix(x):
    """rank.jl:36-44"""
    x = np.asarray(x, dtype=np.float64)
    n = math.isqrt(x.size)
    if n * n != x.size:
        raise ValueError("do not know how to reshape passed vector x")
    return x.reshape((n, n), order="F")


def _device_prox(kind, lam, p, gamma, X, ctx, want_sigma=False):
    """X: (..., rows, cols).  Returns Y of the same shape, the values, and the singular values if asked for."""
    from .solver import default_context
    ctx = ctx or default_context()
    X = np.asarray(X, dtype=np.float64)
    if X.ndim < 2:
        raise ValueError("matrix input expected")
    rows, cols = X.shape[-2:]
    B = int(np.prod(X.shape[:-2], dtype=np.int64)) if X.ndim > 2 else 1
    # the kernel wants contiguous columns; a C-ordered (rows, cols) matrix is the column-major (cols, rows) one, and the prox
    # commutes with the transpose -- so the buffer goes down as it is with the two dimensions swapped
    Xc = np.ascontiguousarray(X).reshape(B, rows, cols)
    Y = np.empty_like(Xc)
    gz = np.empty(B)
    sig = np.empty((B, rows)) if want_sigma else None
    lib = K.load_library()
    P = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None
    K.check(lib.bz_prox_matrix(ctx.handle, kind, float(lam), float(p), float(gamma), cols, rows, B, P(Xc), P(Y), P(gz),
                               P(sig), None), ctx.handle)
    Y = Y.reshape(X.shape)
    gz = gz.reshape(X.shape[:-2]) if X.ndim > 2 else float(gz[0])
    if want_sigma:
        k = min(rows, cols)
        sig = sig[:, :k].reshape(X.shape[:-2] + (k,)) if X.ndim > 2 else sig[0, :k]
    return Y, gz, sig


class _MatrixFunction:
    kind = None
    lam = 1.0
    p = 0.5

    def _matrix(self, X):
        X = np.asarray(X, dtype=np.float64)
        return (check_and_reshape_as_matrix(X), True) if X.ndim == 1 else (X, False)

    def svdvals(self, X, ctx=None):
        """singular values of X (descending), from the same device factorisation the prox uses"""
        return _device_prox(K.MATPROX_NUCLEAR, 0.0, 0.5, 0.0, X, ctx, want_sigma=True)[2]

    def prox(self, X, gamma, ctx=None):
        """``prox!(Y, f, X, gamma)`` -> (Y, value); vector in, vector out (rank.jl:52-58)"""
        M, vec = self._matrix(X)
        Y, gz, _ = _device_prox(self.kind, self.lam, self.p, gamma, M, ctx)
        return (Y.reshape(-1, order="F") if vec else Y), gz


class Rank(_MatrixFunction):
    """src/proxoperators/rank.jl:11-33: lambda * rank(X); prox = hard threshold of the singular values at sqrt(2 gamma lambda)"""
    kind = K.MATPROX_RANK

    def __init__(self, lam=1.0):
        if lam < 0:
            raise ValueError("parameter λ must be nonnegative")   # rank.jl:14-15
        self.lam = float(lam)

    def __call__(self, X, ctx=None):
        M, _ = self._matrix(X)
        s = self.svdvals(M, ctx)
        # LinearAlgebra.rank: singular values above min(m, n) * eps * sigma_max
        tol = min(M.shape[-2:]) * np.finfo(float).eps * (s[..., :1] if s.shape[-1] else 0.0)
        return self.lam * np.sum(s > tol, axis=-1)


class NuclearNorm(_MatrixFunction):
    """src/proxoperators/nuclearNorm.jl:11-34: lambda * sum of singular values; prox = soft threshold by lambda * gamma"""
    kind = K.MATPROX_NUCLEAR

    def __init__(self, lam=1.0):
        if lam < 0:
            raise ValueError("parameter λ must be nonnegative")   # nuclearNorm.jl:14-15
        self.lam = float(lam)

    def __call__(self, X, ctx=None):
        M, _ = self._matrix(X)
        return self.lam * np.sum(self.svdvals(M, ctx), axis=-1)


class SchattenNormLpPower(_MatrixFunction):
    """src/proxoperators/schattenNormLp.jl:14-42: alpha * sum sigma_i^p, 0 < p < 1; prox = the scalar Lp subproblem (:60-97)
    on every singular value"""
    kind = K.MATPROX_SCHATTEN_LP

    def __init__(self, p, alpha=1.0):
        if p <= 0:
            raise ValueError("p must be positive")               # schattenNormLp.jl:18-19
        if p >= 1:
            raise ValueError("p must be smaller than one")       # :20-21
        if alpha < 0:
            raise ValueError("alpha must be nonnegative")        # :23-24
        self.p, self.lam = float(p), float(alpha)

    def __call__(self, X, ctx=None):
        M, _ = self._matrix(X)
        return self.lam * np.sum(self.svdvals(M, ctx) ** self.p, axis=-1)


def prox(f, X, gamma, ctx=None):
    """``prox!(Y, f, X, gamma)`` of Bazinga.jl for the matrix functions: returns (Y, f(Y))"""
    return f.prox(X, gamma, ctx=ctx)
