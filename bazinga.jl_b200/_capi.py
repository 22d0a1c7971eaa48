"""ctypes binding of include/bazinga_b200.h -- the same symbols a Julia ``ccall`` binds
(see INTEGRATION.md and julia/BazingaB200.jl).  No compute happens in Python and there is no
fallback: if the shared library is missing, or there is no CUDA device, calls raise."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libbazinga_b200.so")

BZ_ABI_VERSION = 5

# enums (include/bazinga_b200.h)
F_ZERO, F_LEAST_SQUARES, F_QUADRATIC, F_ROSENBROCK, F_SQDIST2, F_LINEAR2, F_OBSTACLE_L1, F_OBSTACLE_L2 = range(8)
(G_ZERO, G_NORM_L1, G_NORM_L0, G_IND_BOX, G_NORML1_BOX, G_NORML1_NONNEG, G_NORML0_BOX, G_ROSENBROCK,
 G_NONNEG, G_OBSTACLE_L1, G_NORMLP_BOX, G_NORMLP_NONNEG, G_OBSTACLE_RED_L1, G_OBSTACLE_L2) = range(14)
C_IDENTITY, C_DENSE_AFFINE, C_ROSENBROCK, C_EITHEROR, C_MPVCA, C_OBSTACLE, C_OBSTACLE_RED = range(7)
D_ZERO, D_FREE, D_BOX, D_EITHEROR, D_XOR, D_VC, D_CC, D_PORTFOLIO = range(8)
ROW_MAJOR, COL_MAJOR = 0, 1
MATPROX_RANK, MATPROX_NUCLEAR, MATPROX_SCHATTEN_LP = range(3)

BZ_OK, BZ_ERR_ARG, BZ_ERR_CUDA, BZ_ERR_UNSUPPORTED, BZ_ERR_COMM, BZ_ERR_NONPOSITIVE_MU = 0, -1, -2, -3, -4, -5
STATUS_SYMBOLS = ("first_order", "max_iter", "exception", "unknown")   # alps.jl:105-113

_dp = C.POINTER(C.c_double)


class ProblemDesc(C.Structure):
    _fields_ = [
        ("f_kind", C.c_int32), ("g_kind", C.c_int32), ("c_kind", C.c_int32), ("d_kind", C.c_int32),
        ("nx", C.c_int64), ("ny", C.c_int64),
        ("f_A", C.c_void_p), ("f_b", C.c_void_p), ("f_m", C.c_int64), ("f_layout", C.c_int32),
        ("f_p", C.c_double * 4),
        ("g_lambda", C.c_double), ("g_lo", C.c_double), ("g_hi", C.c_double), ("g_u", C.c_void_p),
        ("c_A", C.c_void_p), ("c_b", C.c_void_p), ("c_layout", C.c_int32),
        ("c_A_is_device", C.c_int32), ("c_gen", C.c_int32), ("c_gen_seed", C.c_uint64),
        ("grid_y", C.c_int64), ("grid_x", C.c_int64),
        ("d_lo", C.c_double), ("d_hi", C.c_double), ("d_pair_stride", C.c_int64),
        ("row_begin", C.c_int64), ("ny_local", C.c_int64),
        ("slab_nranks", C.c_int32), ("slab_rank", C.c_int32), ("nx_global", C.c_int64),
        ("slack_formulation", C.c_int32),
    ]


class Options(C.Structure):
    _fields_ = [
        ("tol", C.c_double), ("tol_prim", C.c_double), ("tol_dual", C.c_double),
        ("inner_tol", C.c_double), ("maxit", C.c_int32),
        ("theta_penalty", C.c_double), ("kappa_penalty", C.c_double), ("kappa_tol", C.c_double),
        ("subsolver_maxit", C.c_int64),
        ("alpha", C.c_double), ("beta", C.c_double), ("minimum_gamma", C.c_double),
        ("max_backtracks", C.c_int32), ("lbfgs_memory", C.c_int32),
        ("sub_maxit", C.c_int64), ("max_inner_total", C.c_int64), ("verbose", C.c_int32),
        ("dual_safeguard", C.c_void_p), ("dual_safeguard_user", C.c_void_p),
    ]


class Result(C.Structure):
    _fields_ = [
        ("tot_it", C.c_int32), ("tot_inner_it", C.c_int64), ("elapsed_time", C.c_double),
        ("status", C.c_int32), ("inner_tol", C.c_double), ("norm_res_prim", C.c_double),
        ("objective", C.c_double), ("n_grad", C.c_int64), ("n_prox", C.c_int64),
        ("n_launches", C.c_int64), ("dominant_kernel_ms", C.c_double),
        ("dominant_kernel_launches", C.c_int64),
        ("kind_ms", C.c_double * 4), ("kind_launches", C.c_int64 * 4),
        ("mu0_min", C.c_double), ("mu0_max", C.c_double), ("inner_tol0", C.c_double),
        ("device_ms", C.c_double),
    ]


# every symbol include/bazinga_b200.h declares: name -> (restype, argtypes)
_vp = C.c_void_p
SYMBOLS = {
    "bz_abi_version": (C.c_int, []),
    "bz_ctx_create": (C.c_int, [C.c_int, C.POINTER(_vp)]),
    "bz_ctx_destroy": (C.c_int, [_vp]),
    "bz_last_error": (C.c_char_p, [_vp]),
    "bz_ctx_set_profiling": (C.c_int, [_vp, C.c_int]),
    "bz_ctx_sync": (C.c_int, [_vp]),
    "bz_ctx_timer_start": (C.c_int, [_vp]),
    "bz_ctx_timer_stop": (C.c_int, [_vp, _dp]),
    "bz_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "bz_default_options": (None, [C.POINTER(Options)]),
    "bz_problem_create": (C.c_int, [_vp, C.POINTER(ProblemDesc), C.POINTER(_vp)]),
    "bz_problem_destroy": (C.c_int, [_vp]),
    "bz_problem_upload_dense": (C.c_int, [_vp, _vp, _vp]),
    "bz_problem_xr_handle": (C.c_int, [_vp, _vp]),
    "bz_problem_xr_attach": (C.c_int, [_vp, _vp]),
    "bz_alps_solve": (C.c_int, [_vp, _vp, _vp, C.POINTER(Options), _vp, _vp, _vp, _vp, C.POINTER(Result)]),
    "bz_als_solve": (C.c_int, [_vp, _vp, _vp, C.POINTER(Options), _vp, _vp, _vp, _vp, C.POINTER(Result)]),
    "bz_panocplus_solve": (C.c_int, [_vp, _vp, _vp, _vp, C.POINTER(Options), C.c_double, _vp,
                                     C.POINTER(C.c_int64), _dp, _dp]),
    "bz_auglag_state": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _dp]),
    "bz_panoc_state": (C.c_int, [_vp, _vp]),
    "bz_auglag_update": (C.c_int, [_vp, _vp, _vp]),
    "bz_auglag_grad": (C.c_int, [_vp, _vp, _vp, _dp, _dp, _vp, _vp, _vp]),
    "bz_prox": (C.c_int, [_vp, _vp, C.c_double, _vp, _dp]),
    "bz_proj": (C.c_int, [_vp, _vp, _vp]),
    "bz_fb_step": (C.c_int, [_vp, _vp, _vp, C.c_double, _vp, _vp, _dp, _dp, _dp]),
    "bz_obstacle_fused_step": (C.c_int, [_vp, _vp, C.c_double, _vp, _vp, _vp, _vp, _vp]),
    "bz_lbfgs_reset": (C.c_int, [_vp, C.c_int]),
    "bz_lbfgs_update": (C.c_int, [_vp, _vp, _vp]),
    "bz_lbfgs_apply": (C.c_int, [_vp, _vp, _vp]),
    "bz_multiplier_update": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _dp]),
    "bz_default_penalty": (C.c_int, [_vp, _vp, _vp, C.c_double, _vp]),
    "bz_randn_fill": (C.c_int, [_vp, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int64, _vp]),
    "bz_batch_solve": (C.c_int, [_vp, C.c_int, C.c_int64, _vp, _vp, C.POINTER(Options), _vp, _vp, _vp,
                                 _vp, _vp, _vp, _dp]),
    "bz_batch_solve_als": (C.c_int, [_vp, C.c_int, C.c_int64, _vp, _vp, C.POINTER(Options), _vp, _vp, _vp,
                                     _vp, _vp, _vp, _dp]),
    "bz_fp64_fma_peak": (C.c_int, [_vp, _dp]),
    "bz_prox_matrix": (C.c_int, [_vp, C.c_int, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int64, _vp, _vp, _vp, _vp, _vp]),
    "bz_bdense_create": (C.c_int, [_vp, C.c_int64, C.c_int64, _vp, C.c_int32, C.c_int32, C.c_uint64, C.POINTER(_vp)]),
    "bz_bdense_destroy": (C.c_int, [_vp]),
    "bz_bdense_auglag_grad": (C.c_int, [_vp, C.c_int32, _vp, _vp, _vp, _vp, C.c_int32, C.c_double, C.c_double, _vp, _vp, _vp,
                                        _vp, C.c_int32, _dp]),
    "bz_comm_unique_id": (C.c_int, [_vp]),
    "bz_comm_init": (C.c_int, [_vp, _vp, C.c_int, C.c_int]),
    "bz_comm_destroy": (C.c_int, [_vp]),
    "bz_comm_bench_allreduce": (C.c_int, [_vp, C.c_int64, C.c_int, _dp]),
}

# alps.jl:23,62: dual_safeguard(y, cx) as a host callback (see bz_options.dual_safeguard)
DUAL_SAFEGUARD_FN = C.CFUNCTYPE(None, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int64, C.c_void_p)


class BazingaError(RuntimeError):
    """Mirrors Julia's ``error(msg)`` raised by the reference's guards; carries the bz_status."""

    def __init__(self, code, msg):
        super().__init__(f"[bz_status {code}] {msg}")
        self.code = code
        self.msg = msg


_lib = None


def load_library():
    """Load libbazinga_b200.so (built in-tree by ``make -C bazinga.jl_b200`` /
    ``__graft_entry__.build()``).  Raises if it is missing -- there is no other implementation."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise BazingaError(BZ_ERR_CUDA, f"{LIB_PATH} not found: build the CUDA library first "
                                        "(python -c 'import __graft_entry__ as g; g.build()'); "
                                        "there is no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SYMBOLS.items():
        fn = getattr(lib, name)   # AttributeError if the header and the library ever disagree
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.bz_abi_version() != BZ_ABI_VERSION:
        raise BazingaError(BZ_ERR_ARG, "ABI version mismatch between _capi.py and the library")
    _lib = lib
    return lib


def check(rc, ctx=None):
    if rc != BZ_OK:
        lib = load_library()
        msg = lib.bz_last_error(ctx)
        raise BazingaError(rc, (msg or b"").decode("utf-8", "replace"))
