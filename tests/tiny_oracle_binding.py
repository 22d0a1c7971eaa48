"""ctypes binding of oracle/tiny_oracle.c (test infrastructure).  Builds the shared object on first
use with gcc -O2 -ffp-contract=off (no FMA contraction, see the file header)."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "oracle", "tiny_oracle.c")
SO = os.path.join(ROOT, "oracle", "_build", "libtiny_oracle.so")


def build_tiny_oracle(force=False):
    if force or not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(SRC):
        os.makedirs(os.path.dirname(SO), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-fopenmp", "-o", SO, SRC, "-lm"])
    return SO


class TinyOpts(C.Structure):
    _fields_ = [("tol", C.c_double), ("tol_prim", C.c_double), ("tol_dual", C.c_double), ("inner_tol", C.c_double),
                ("maxit", C.c_int), ("theta_penalty", C.c_double), ("kappa_penalty", C.c_double),
                ("kappa_tol", C.c_double), ("subsolver_maxit", C.c_longlong), ("alpha", C.c_double),
                ("beta", C.c_double), ("minimum_gamma", C.c_double), ("max_backtracks", C.c_int),
                ("lbfgs_memory", C.c_int), ("sub_maxit", C.c_longlong)]


def tiny_batch(kind, X0, *, min_gamma, sub_maxit, subsolver_maxit, tol=1e-8, inner_tol=1.0, maxit=100, memory=5,
               solver="alps"):
    """solver = "alps" (src/algorithms/alps.jl) or "als" (src/algorithms/als.jl)"""
    lib = C.CDLL(build_tiny_oracle())
    nx = 2 if kind <= 3 else (4 if kind == 6 else 6)      # kinds 4..7: the demos' slack variants
    X0 = np.ascontiguousarray(X0, dtype=np.float64).reshape(-1, nx)
    B = X0.shape[0]
    ny = 2 if kind in (0, 6) else 4
    o = TinyOpts(tol, -1.0, -1.0, inner_tol, maxit, 0.8, 0.5, 0.1, subsolver_maxit, 0.95, 0.5, min_gamma, 20, memory, sub_maxit)
    Y0 = np.zeros((B, ny))
    x, y = np.zeros((B, nx)), np.zeros((B, ny))
    it, st = np.zeros(B, np.int32), np.zeros(B, np.int32)
    inn, nrp = np.zeros(B, np.int64), np.zeros(B)
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    fn = lib.tiny_als_batch if solver == "als" else lib.tiny_alps_batch
    fn(C.c_int(kind), C.c_longlong(B), P(X0), P(Y0), C.byref(o), P(x), P(y), P(it), P(inn), P(st), P(nrp))
    return dict(x=x, y=y, iters=it, sub_iters=inn, status=st, norm_res_prim=nrp)


def oracle_randn_fill(seed, start, count, stream=0):
    lib = C.CDLL(build_tiny_oracle())
    out = np.empty(int(count))
    lib.oracle_randn_fill(C.c_uint64(seed), C.c_uint64(stream), C.c_uint64(start), C.c_longlong(int(count)),
                          out.ctypes.data_as(C.c_void_p))
    return out
