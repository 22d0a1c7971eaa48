"""Host-side mirror of the reference's plugin/operator interface for the PANOC+/ALPS hot path
(Julia is not available in this image, so the mirror is Python over the same C ABI a Julia
``ccall`` would bind; see julia/BazingaB200.jl and INTEGRATION.md).

    alps(f, g, c, D, x0, y0; tol, tol_prim, tol_dual, inner_tol, maxit, theta_penalty,
         kappa_penalty, kappa_tol, verbose, subsolver, subsolver_maxit)
        -> (x, y, tot_it, tot_inner_it, elapsed_time, status, inner_tol, norm_res_prim, s, mu)

is src/algorithms/alps.jl:7-117 with the same names, argument meaning and error behaviour.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi as K
from . import oracles as O


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


class Context:
    """One CUDA device + stream (bz_ctx).  Not thread-safe, like the reference."""

    def __init__(self, device=0):
        self.lib = K.load_library()
        self.handle = C.c_void_p()
        K.check(self.lib.bz_ctx_create(int(device), C.byref(self.handle)))
        self.device = device

    def set_profiling(self, on=True):
        K.check(self.lib.bz_ctx_set_profiling(self.handle, 1 if on else 0), self.handle)

    def sync(self):
        K.check(self.lib.bz_ctx_sync(self.handle), self.handle)

    def timer_start(self):
        K.check(self.lib.bz_ctx_timer_start(self.handle), self.handle)

    def timer_stop(self):
        ms = C.c_double()
        K.check(self.lib.bz_ctx_timer_stop(self.handle, C.byref(ms)), self.handle)
        return ms.value

    def comm_init(self, unique_id: bytes, nranks: int, rank: int):
        buf = C.create_string_buffer(unique_id, 128)
        K.check(self.lib.bz_comm_init(self.handle, buf, nranks, rank), self.handle)

    def comm_bench_allreduce(self, count: int, reps: int = 50) -> float:
        """Average device time (us) of one allreduce(sum) of ``count`` doubles on the library's NCCL communicator."""
        us = C.c_double()
        K.check(self.lib.bz_comm_bench_allreduce(self.handle, int(count), int(reps), C.byref(us)), self.handle)
        return us.value

    def fp64_fma_peak(self) -> float:
        """fp64 FMA issue peak of the device in TFLOP/s (microbenchmark; the batched path's secondary roofline)."""
        v = C.c_double()
        K.check(self.lib.bz_fp64_fma_peak(self.handle, C.byref(v)), self.handle)
        return v.value

    def randn_fill(self, seed, start, count, stream=0):
        out = np.empty(int(count), dtype=np.float64)
        K.check(self.lib.bz_randn_fill(self.handle, seed, stream, start, int(count), _ptr(out)), self.handle)
        return out

    def close(self):
        if self.handle:
            self.lib.bz_ctx_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default_ctx = None


def default_context():
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(0)
    return _default_ctx


def comm_unique_id() -> bytes:
    lib = K.load_library()
    buf = C.create_string_buffer(128)
    K.check(lib.bz_comm_unique_id(buf))
    return buf.raw


def row_shard(ny: int, nranks: int, rank: int):
    """Row block of rank `rank` when a dense-affine instance is row-sharded over `nranks` GPUs
    (SURVEY.md §8e): contiguous blocks, sizes differing by at most one row, even-sized except possibly
    the last so that adjacent pairs never straddle shards.  Returns (row_begin, ny_local)."""
    if not (0 <= rank < nranks) or ny < nranks:
        raise ValueError("row_shard: need 0 <= rank < nranks <= ny")
    base = (ny // nranks) & ~1 if ny // nranks >= 2 else ny // nranks
    begin = rank * base
    local = base if rank < nranks - 1 else ny - begin
    return begin, local


class LBFGS:
    """ProximalAlgorithms.LBFGS(memory) (demo/basispursuit.jl:69)."""

    def __init__(self, memory=5):
        self.memory = int(memory)


class NoAcceleration(LBFGS):
    """ProximalAlgorithms.NoAcceleration() == L-BFGS with memory 0 (d = -res)."""

    def __init__(self):
        super().__init__(0)


class PANOCplus:
    """ProximalAlgorithms.PANOCplus(; directions, maxit, tol, minimum_gamma, ...) -- the ``subsolver``
    plugin slot of alps (alps.jl:24,64,66).  Calling ``subsolver(tol=..., verbose=...)`` returns a
    configured copy, exactly how alps.jl:64 uses it; calling the copy with (f=alFun, g=gFun, x0=x)
    runs the GPU inner loop (see Problem.panocplus)."""

    def __init__(self, *, directions=None, maxit=1000, tol=1e-8, minimum_gamma=1e-7, alpha=0.95,
                 beta=0.5, max_backtracks=20, verbose=False, freq=10):
        self.directions = directions if directions is not None else LBFGS(5)
        self.maxit, self.tol, self.minimum_gamma = int(maxit), float(tol), float(minimum_gamma)
        self.alpha, self.beta, self.max_backtracks = float(alpha), float(beta), int(max_backtracks)
        self.verbose, self.freq = verbose, freq

    def __call__(self, **kw):
        cfg = dict(directions=self.directions, maxit=self.maxit, tol=self.tol,
                   minimum_gamma=self.minimum_gamma, alpha=self.alpha, beta=self.beta,
                   max_backtracks=self.max_backtracks, verbose=self.verbose, freq=self.freq)
        cfg.update(kw)
        return PANOCplus(**cfg)


default_subsolver = PANOCplus   # alps.jl:5


def _julia_float(v: float) -> str:
    """Julia's `print(::Float64)`: shortest round-trip digits, always with a fractional part or an `e` exponent."""
    if v != v:
        return "NaN"
    if v in (float("inf"), float("-inf")):
        return "Inf" if v > 0 else "-Inf"
    r = repr(float(v))
    if "e" in r:
        m, e = r.split("e")
        if "." not in m:
            m += ".0"
        return f"{m}e{int(e)}"
    return r


def _wrap_dual_safeguard(fn):
    """alps.jl:23,62 ``dual_safeguard(y, cx)`` (in-place on y) -> the C-ABI host callback of bz_options."""
    def tramp(yp, cxp, m, _user):
        y = np.ctypeslib.as_array(yp, shape=(m,))
        cx = np.ctypeslib.as_array(cxp, shape=(m,))
        fn(y, cx)
    return K.DUAL_SAFEGUARD_FN(tramp)


def _options(*, tol, tol_prim, tol_dual, inner_tol, maxit, theta_penalty, kappa_penalty, kappa_tol,
             subsolver_maxit, sub: PANOCplus, max_inner_total=0, verbose=False, dual_safeguard=None):
    lib = K.load_library()
    o = K.Options()
    lib.bz_default_options(C.byref(o))
    o.tol = tol
    o.tol_prim = -1.0 if tol_prim is None else tol_prim
    o.tol_dual = -1.0 if tol_dual is None else tol_dual
    o.inner_tol = -1.0 if inner_tol is None else inner_tol
    o.maxit = maxit
    o.theta_penalty, o.kappa_penalty, o.kappa_tol = theta_penalty, kappa_penalty, kappa_tol
    o.subsolver_maxit = int(subsolver_maxit)
    o.alpha, o.beta, o.minimum_gamma = sub.alpha, sub.beta, sub.minimum_gamma
    o.max_backtracks, o.lbfgs_memory = sub.max_backtracks, sub.directions.memory
    o.sub_maxit = sub.maxit
    o.max_inner_total = int(max_inner_total)
    o.verbose = 1 if verbose else 0
    if dual_safeguard is not None:
        cb = _wrap_dual_safeguard(dual_safeguard)
        o._keep_cb = cb                       # the callback object must outlive the solve
        o.dual_safeguard = C.cast(cb, C.c_void_p)
    return o


class Problem:
    """The four oracles on the device (bz_problem): the GPU counterpart of the Julia oracle structs
    plus ``AugLagFun`` / ``NonsmoothCostFun`` (src/utilities/auglagfun.jl, nonsmoothcostfun.jl)."""

    def __init__(self, f, g, c, D, nx=None, ny=None, ctx: Context | None = None, slack=False):
        self.handle = C.c_void_p()
        for name, obj, attr in (("f", f, "f_kind"), ("g", g, "g_kind"), ("c", c, "c_kind"), ("D", D, "d_kind")):
            if not hasattr(obj, attr):
                raise K.BazingaError(K.BZ_ERR_UNSUPPORTED,
                                     f"{name} = {type(obj).__name__} is not in the GPU oracle catalogue "
                                     "(no CPU fallback)")
        self.ctx = ctx or default_context()
        self.lib = self.ctx.lib
        self.f, self.g, self.c, self.D = f, g, c, D
        d = K.ProblemDesc()
        self._keep = []
        d.f_kind, d.g_kind, d.c_kind, d.d_kind = f.f_kind, g.g_kind, c.c_kind, D.d_kind
        # ---- c
        if isinstance(c, O.ConstraintBasisPursuit):
            ny_local, nx_ = c.shape
            nx = nx_ if nx is None else nx
            ny = (c.ny_global or ny_local) if ny is None else ny
            d.c_A = _ptr(c.A)
            d.c_b = _ptr(c.b)
            d.c_layout = c.layout
            if c.gen_seed is not None and c.A is None:
                d.c_gen, d.c_gen_seed = 1, int(c.gen_seed)
            if c.ny_global is not None and c.ny_global != ny_local:
                d.row_begin, d.ny_local = c.row_begin, ny_local
            self._keep += [c.A, c.b]
        elif isinstance(c, (O.ConstraintObstacle, O.ConstraintObstacleRed)):
            red = isinstance(c, O.ConstraintObstacleRed)
            nx = (2 if red else 3) * c.N if nx is None else nx
            ny = (2 if red else 1) * c.N if ny is None else ny
            if c.grid is not None:
                d.grid_y, d.grid_x = int(c.grid[0]), int(c.grid[1])
            if c.slab is not None:
                d.slab_rank, d.slab_nranks = int(c.slab[0]), int(c.slab[1])
                d.nx_global = 3 * c.N * int(c.slab[1])
        # ---- f
        if isinstance(f, O.LeastSquares):
            d.f_A, d.f_b, d.f_m = _ptr(f.A), _ptr(f.b), f.A.shape[0]
            d.f_layout = K.ROW_MAJOR
            nx = f.A.shape[1] if nx is None else nx
            self._keep += [f.A, f.b]
        elif isinstance(f, O.Quadratic):
            d.f_A, d.f_b, d.f_m = _ptr(f.Q), _ptr(f.q), f.Q.shape[0]
            d.f_layout = K.ROW_MAJOR
            nx = f.Q.shape[0] if nx is None else nx
            self._keep += [f.Q, f.q]
        if nx is None:
            raise ValueError("nx could not be inferred; pass nx=")
        if ny is None:
            ny = nx if isinstance(c, O.IdentityFunction) else None
        if ny is None:
            raise ValueError("ny could not be inferred; pass ny=")
        d.nx, d.ny = int(nx), int(ny)
        # ---- g
        d.g_lambda = getattr(g, "lam", 1.0)
        d.g_lo, d.g_hi = getattr(g, "lo", 0.0), getattr(g, "hi", 0.0)
        if hasattr(g, "u"):
            if g.u.shape[0] != nx:
                raise ValueError("g.u must have nx entries")
            d.g_u = _ptr(g.u)
            self._keep.append(g.u)
        # ---- D
        d.d_lo, d.d_hi = getattr(D, "lo", 0.0), getattr(D, "hi", 0.0)
        d.d_pair_stride = getattr(D, "stride", 1)
        d.slack_formulation = 1 if slack else 0     # als: PANOC+ on [x; s] (src/algorithms/als.jl)
        self.slack = bool(slack)
        self.nx, self.ny = int(nx), int(ny)
        self.ny_local = int(d.ny_local) if d.ny_local > 0 else self.ny
        self.row_begin = int(d.row_begin)
        self.desc = d
        self.handle = C.c_void_p()
        K.check(self.lib.bz_problem_create(self.ctx.handle, C.byref(d), C.byref(self.handle)), self.ctx.handle)

    # ---- whole solves
    def alps(self, x0, y0, options: K.Options):
        x0, y0 = _f64(x0), _f64(y0)
        if x0.shape[0] != self.nx or y0.shape[0] != self.ny:
            raise ValueError("x0 / y0 have the wrong length")
        x = np.empty(self.nx)
        y, s, mu = np.empty(self.ny_local), np.empty(self.ny_local), np.empty(self.ny_local)
        res = K.Result()
        K.check(self.lib.bz_alps_solve(self.handle, _ptr(x0), _ptr(y0), C.byref(options), _ptr(x), _ptr(y),
                                       _ptr(s), _ptr(mu), C.byref(res)), self.ctx.handle)
        return x, y, s, mu, res

    def als(self, x0, y0, options: K.Options):
        """``als`` (src/algorithms/als.jl:7-120) on a problem created with ``slack=True``."""
        x0, y0 = _f64(x0), _f64(y0)
        if x0.shape[0] != self.nx or y0.shape[0] != self.ny:
            raise ValueError("x0 / y0 have the wrong length")
        x = np.empty(self.nx)
        y, s, mu = np.empty(self.ny_local), np.empty(self.ny_local), np.empty(self.ny_local)
        res = K.Result()
        K.check(self.lib.bz_als_solve(self.handle, _ptr(x0), _ptr(y0), C.byref(options), _ptr(x), _ptr(y),
                                      _ptr(s), _ptr(mu), C.byref(res)), self.ctx.handle)
        return x, y, s, mu, res

    # ---- slab-decomposed obstacle grid (one process per GPU): export / attach the peer-mapped slabs
    def xr_handle(self) -> bytes:
        buf = C.create_string_buffer(64)
        K.check(self.lib.bz_problem_xr_handle(self.handle, buf), self.ctx.handle)
        return buf.raw

    def xr_attach(self, handles):
        """``handles``: the 64-byte handles of all ranks in rank order (e.g. from dist.all_gather_object)."""
        blob = b"".join(handles)
        K.check(self.lib.bz_problem_xr_attach(self.handle, C.c_char_p(blob)), self.ctx.handle)

    def upload_dense(self, A, b):
        """Replace the dense constraint data; ``A=None`` keeps the device-resident matrix and replaces b only."""
        b = _f64(b)
        if A is None:
            K.check(self.lib.bz_problem_upload_dense(self.handle, None, _ptr(b)), self.ctx.handle)
            return
        # the library reads A in the layout the problem was created with: normalise, never upload a strided view
        A = np.asarray(A, dtype=np.float64)
        if A.shape != (self.ny_local, self.nx):
            raise ValueError("upload_dense: A has the wrong shape")
        A = np.asfortranarray(A) if self.desc.c_layout == K.COL_MAJOR else np.ascontiguousarray(A)
        K.check(self.lib.bz_problem_upload_dense(self.handle, _ptr(A), _ptr(b)), self.ctx.handle)

    def panocplus(self, mu, y, x0, sub: PANOCplus):
        """The subsolver slot: ``sub_sol, sub_it = sub_solver(f = alFun, g = gFun, x0 = x)``
        (alps.jl:66); also returns alFun.fx and gFun.gz at the solution (alps.jl:68)."""
        o = _options(tol=1e-6, tol_prim=None, tol_dual=None, inner_tol=None, maxit=100, theta_penalty=0.8,
                     kappa_penalty=0.5, kappa_tol=0.1, subsolver_maxit=10**9, sub=sub)
        mu, y, x0 = _f64(mu), _f64(y), _f64(x0)
        x = np.empty(self.nx)
        it, fx, gz = C.c_int64(), C.c_double(), C.c_double()
        K.check(self.lib.bz_panocplus_solve(self.handle, _ptr(mu), _ptr(y), _ptr(x0), C.byref(o), sub.tol,
                                            _ptr(x), C.byref(it), C.byref(fx), C.byref(gz)), self.ctx.handle)
        return x, it.value, fx.value, gz.value

    def panoc_state(self):
        """Scalars of the PANOCplus iterate after the last inner solve: dict(gamma, tau, f_x, f_z, g_z, stopnorm, k, H)."""
        out = np.empty(8)
        K.check(self.lib.bz_panoc_state(self.handle, _ptr(out)), self.ctx.handle)
        return dict(gamma=out[0], tau=out[1], f_x=out[2], f_z=out[3], g_z=out[4], stopnorm=out[5], k=int(out[6]), H=out[7])

    # ---- per-kernel entry points (names follow the reference's functions)
    def auglag_update(self, mu, y):
        mu, y = _f64(mu), _f64(y)
        K.check(self.lib.bz_auglag_update(self.handle, _ptr(mu), _ptr(y)), self.ctx.handle)

    def gradient(self, x):
        """lx = gradient!(dlx, alFun, x) -> dict(dlx, lx, fx, cx, s, yupd) (auglagfun.jl:73-86)."""
        x = _f64(x)
        nv = self.nx + self.ny if getattr(self, "slack", False) else self.nx    # slack: gradient! of AugLagFunSlack on [x; s]
        if x.shape[0] != nv:
            raise ValueError("x has the wrong length")
        dlx = np.empty(nv)
        cx, s, yupd = (np.empty(self.ny_local) for _ in range(3))
        lx, fx = C.c_double(), C.c_double()
        K.check(self.lib.bz_auglag_grad(self.handle, _ptr(x), _ptr(dlx), C.byref(lx), C.byref(fx), _ptr(cx),
                                        _ptr(s), _ptr(yupd)), self.ctx.handle)
        return dict(dlx=dlx, lx=lx.value, fx=fx.value, cx=cx, s=s, yupd=yupd)

    def auglag_state(self):
        cx, s, yupd, muy = (np.empty(self.ny_local) for _ in range(4))
        musqy = C.c_double()
        K.check(self.lib.bz_auglag_state(self.handle, _ptr(cx), _ptr(s), _ptr(yupd), _ptr(muy),
                                         C.byref(musqy)), self.ctx.handle)
        return dict(cx=cx, s=s, yupd=yupd, muy=muy, musqy=musqy.value)

    def prox(self, x, gamma):
        """(z, gz) with gz = prox!(z, g, x, gamma)."""
        x = _f64(x)
        z = np.empty(self.nx)
        gz = C.c_double()
        K.check(self.lib.bz_prox(self.handle, _ptr(x), float(gamma), _ptr(z), C.byref(gz)), self.ctx.handle)
        return z, gz.value

    def proj(self, v):
        v = _f64(v)
        z = np.empty(self.ny_local)
        K.check(self.lib.bz_proj(self.handle, _ptr(v), _ptr(z)), self.ctx.handle)
        return z

    def fb_step(self, x, grad, gamma):
        x, grad = _f64(x), _f64(grad)
        z, res = np.empty(self.nx), np.empty(self.nx)
        gz, dot, sq = C.c_double(), C.c_double(), C.c_double()
        K.check(self.lib.bz_fb_step(self.handle, _ptr(x), _ptr(grad), float(gamma), _ptr(z), _ptr(res),
                                    C.byref(gz), C.byref(dot), C.byref(sq)), self.ctx.handle)
        return dict(z=z, res=res, gz=gz.value, dot_grad_res=dot.value, res_sq=sq.value)

    def obstacle_fused_step(self, x, gamma):
        """One fused forward-backward evaluation (csrc/bz_obfused.cu): dict(z, res, grad_x, grad_z, scalars[9])."""
        x = _f64(x)
        z, res, gx, gz = (np.empty(self.nx) for _ in range(4))
        sc = np.empty(9)
        K.check(self.lib.bz_obstacle_fused_step(self.handle, _ptr(x), float(gamma), _ptr(z), _ptr(res), _ptr(gx), _ptr(gz),
                                                _ptr(sc)), self.ctx.handle)
        return dict(z=z, res=res, grad_x=gx, grad_z=gz, scalars=sc)

    def lbfgs_reset(self, memory=5):
        K.check(self.lib.bz_lbfgs_reset(self.handle, int(memory)), self.ctx.handle)

    def lbfgs_update(self, s, y):
        s, y = _f64(s), _f64(y)
        K.check(self.lib.bz_lbfgs_update(self.handle, _ptr(s), _ptr(y)), self.ctx.handle)

    def lbfgs_apply(self, v):
        v = _f64(v)
        d = np.empty(self.nx)
        K.check(self.lib.bz_lbfgs_apply(self.handle, _ptr(v), _ptr(d)), self.ctx.handle)
        return d

    def multiplier_update(self, cx, mu, y):
        cx, mu, y = _f64(cx), _f64(mu), _f64(y)
        yo, so = np.empty(self.ny_local), np.empty(self.ny_local)
        nrp = C.c_double()
        K.check(self.lib.bz_multiplier_update(self.handle, _ptr(cx), _ptr(mu), _ptr(y), _ptr(yo), _ptr(so),
                                              C.byref(nrp)), self.ctx.handle)
        return yo, so, nrp.value

    def default_penalty(self, cx, proj_cx, objx):
        cx, pc = _f64(cx), _f64(proj_cx)
        mu = np.empty(self.ny_local)
        K.check(self.lib.bz_default_penalty(self.handle, _ptr(cx), _ptr(pc), float(objx), _ptr(mu)),
                self.ctx.handle)
        return mu

    def close(self):
        if self.handle:
            self.lib.bz_problem_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class SharedMatrixBatch:
    """Batched shared-A mode (csrc/bz_bdense.cu): up to 8 dense-affine instances c_k(x) = A x - b_k that share A
    (demo/basispursuit.jl:38-49 over a batch).  ``gradient`` evaluates gradient!(dlx_k, alFun_k, x_k) (auglagfun.jl:73-86,
    f = Zero) for the whole batch with A read once per pass, on the fp64 tensor cores."""

    def __init__(self, A=None, *, shape=None, gen_seed=None, ctx: Context | None = None):
        self.ctx = ctx or default_context()
        self.lib = self.ctx.lib
        self.handle = C.c_void_p()
        if A is not None:
            A = _f64(A)
            self.m, self.n = A.shape
            K.check(self.lib.bz_bdense_create(self.ctx.handle, self.m, self.n, _ptr(A), K.ROW_MAJOR, 0, 0, C.byref(self.handle)),
                    self.ctx.handle)
        else:
            self.m, self.n = shape
            K.check(self.lib.bz_bdense_create(self.ctx.handle, self.m, self.n, None, K.ROW_MAJOR, 1, int(gen_seed), C.byref(self.handle)),
                    self.ctx.handle)

    def gradient(self, X, b, mu, y, D=None, reps=1):
        """X: batch x n; b, mu, y: batch x m.  Returns dict(dlx, lx, cx, yupd, us) with the instance-major layouts of the inputs."""
        X, b, mu, y = _f64(X), _f64(b), _f64(mu), _f64(y)
        B = X.shape[0]
        if X.shape != (B, self.n) or b.shape != (B, self.m) or mu.shape != (B, self.m) or y.shape != (B, self.m):
            raise ValueError("SharedMatrixBatch.gradient: shapes must be (batch, n) and (batch, m)")
        D = D or O.ZeroSet()
        dlx, cx, yupd, lx = np.empty((B, self.n)), np.empty((B, self.m)), np.empty((B, self.m)), np.empty(B)
        us = C.c_double()
        K.check(self.lib.bz_bdense_auglag_grad(self.handle, B, _ptr(X), _ptr(b), _ptr(mu), _ptr(y), D.d_kind,
                                               getattr(D, "lo", 0.0), getattr(D, "hi", 0.0), _ptr(dlx), _ptr(lx), _ptr(cx),
                                               _ptr(yupd), int(reps), C.byref(us)), self.ctx.handle)
        return dict(dlx=dlx, lx=lx, cx=cx, yupd=yupd, us=us.value)

    def close(self):
        if self.handle:
            self.lib.bz_bdense_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def alps(f, g, c, D, x0, y0, *, tol=1e-6, tol_prim=None, tol_dual=None, inner_tol=None, maxit=100,
         theta_penalty=0.8, kappa_penalty=0.5, kappa_tol=0.1, verbose=False, dual_safeguard=None,
         subsolver=default_subsolver, subsolver_maxit=1_000_000_000, ctx=None, problem=None,
         return_result=False, max_inner_total=0):
    """Drop-in for ``Bazinga.alps`` (src/algorithms/alps.jl:7-26).  Returns the reference's 10-tuple
    ``(x, y, tot_it, tot_inner_it, elapsed_time, status, inner_tol, norm_res_prim, s, mu)`` with
    ``status`` in {"first_order", "max_iter", "exception", "unknown"} (the Julia Symbols).
    x0 and y0 are never modified (test/problems/test_nonconvex_qp.jl:36).  ``dual_safeguard(y, cx)`` (alps.jl:23,62)
    may be a Python callable that edits ``y`` in place: it runs on the host once per outer iteration; ``None`` keeps
    ``default_dual_safeguard!`` on the device."""
    sub = subsolver(tol=1.0, verbose=verbose)   # alps.jl:64: constructor call with two kwargs only
    if not isinstance(sub, PANOCplus):
        raise K.BazingaError(K.BZ_ERR_UNSUPPORTED, "subsolver must be (a closure over) PANOCplus")
    x0 = np.asarray(x0, dtype=np.float64)
    y0 = np.asarray(y0, dtype=np.float64)
    own = problem is None
    p = problem or Problem(f, g, c, D, nx=x0.shape[0], ny=None if isinstance(c, O.ConstraintBasisPursuit) else y0.shape[0], ctx=ctx)
    try:
        o = _options(tol=tol, tol_prim=tol_prim, tol_dual=tol_dual, inner_tol=inner_tol, maxit=maxit,
                     theta_penalty=theta_penalty, kappa_penalty=kappa_penalty, kappa_tol=kappa_tol,
                     subsolver_maxit=subsolver_maxit, sub=sub, max_inner_total=max_inner_total,
                     verbose=verbose, dual_safeguard=dual_safeguard)
        x, y, s, mu, res = p.alps(x0, y0, o)
    finally:
        if own:
            p.close()
    if verbose:      # the two @info lines of alps.jl:52-55 (Julia's logger writes them to stderr)
        import sys
        print(f"[ Info: initial penalty parameters μ ∈ [{_julia_float(res.mu0_min)}, {_julia_float(res.mu0_max)}]", file=sys.stderr)
        print(f"[ Info: initial inner tolerance {_julia_float(res.inner_tol0)}", file=sys.stderr)
    out = (x, y, res.tot_it, res.tot_inner_it, res.elapsed_time, K.STATUS_SYMBOLS[res.status], res.inner_tol,
           res.norm_res_prim, s, mu)
    return (out, res) if return_result else out


def als(f, g, c, D, x0, y0, *, tol=1e-6, tol_prim=None, tol_dual=None, inner_tol=None, maxit=100,
         theta_penalty=0.8, kappa_penalty=0.5, kappa_tol=0.1, verbose=False, dual_safeguard=None,
         subsolver=default_subsolver, subsolver_maxit=1_000_000_000, ctx=None, problem=None,
         return_result=False, max_inner_total=0):
    """Drop-in for ``Bazinga.als`` (src/algorithms/als.jl:7-26): the augmented-Lagrangian method with explicit slack
    variables (PANOCplus on [x; s] with AugLagFunSlack / NonsmoothCostFunSlack).  Returns the reference's 10-tuple
    ``(x, y, tot_it, tot_inner_it, elapsed_time, status, inner_tol, norm_res_prim, s, mu)`` with
    ``status`` in {"first_order", "max_iter", "exception", "unknown"} (the Julia Symbols).
    x0 and y0 are never modified (test/problems/test_nonconvex_qp.jl:36).  ``dual_safeguard(y, cx)`` (alps.jl:23,62)
    may be a Python callable that edits ``y`` in place: it runs on the host once per outer iteration; ``None`` keeps
    ``default_dual_safeguard!`` on the device."""
    sub = subsolver(tol=1.0, verbose=verbose)   # alps.jl:64: constructor call with two kwargs only
    if not isinstance(sub, PANOCplus):
        raise K.BazingaError(K.BZ_ERR_UNSUPPORTED, "subsolver must be (a closure over) PANOCplus")
    x0 = np.asarray(x0, dtype=np.float64)
    y0 = np.asarray(y0, dtype=np.float64)
    own = problem is None
    p = problem or Problem(f, g, c, D, nx=x0.shape[0], ny=None if isinstance(c, O.ConstraintBasisPursuit) else y0.shape[0], ctx=ctx,
                           slack=True)
    try:
        o = _options(tol=tol, tol_prim=tol_prim, tol_dual=tol_dual, inner_tol=inner_tol, maxit=maxit,
                     theta_penalty=theta_penalty, kappa_penalty=kappa_penalty, kappa_tol=kappa_tol,
                     subsolver_maxit=subsolver_maxit, sub=sub, max_inner_total=max_inner_total,
                     verbose=verbose, dual_safeguard=dual_safeguard)
        x, y, s, mu, res = p.als(x0, y0, o)
    finally:
        if own:
            p.close()
    if verbose:      # the two @info lines of alps.jl:52-55 (Julia's logger writes them to stderr)
        import sys
        print(f"[ Info: initial penalty parameters μ ∈ [{_julia_float(res.mu0_min)}, {_julia_float(res.mu0_max)}]", file=sys.stderr)
        print(f"[ Info: initial inner tolerance {_julia_float(res.inner_tol0)}", file=sys.stderr)
    out = (x, y, res.tot_it, res.tot_inner_it, res.elapsed_time, K.STATUS_SYMBOLS[res.status], res.inner_tol,
           res.norm_res_prim, s, mu)
    return (out, res) if return_result else out


BATCH_KINDS = {"rosenbrock": 0, "eitheror": 1, "xor": 2, "mpvca": 3, "eitheror_fullslack": 4, "xor_fullslack": 5,
               "mpvca_slack": 6, "mpvca_fullslack": 7}
BATCH_DIMS = {0: (2, 2), 1: (2, 4), 2: (2, 4), 3: (2, 4), 4: (6, 4), 5: (6, 4), 6: (4, 2), 7: (6, 4)}   # kind -> (nx, ny)


def als_batch(kind, X0, Y0=None, **kw):
    """``als`` (src/algorithms/als.jl:7-120) over a batch of starting points; see alps_batch."""
    return alps_batch(kind, X0, Y0, solver="als", **kw)


def alps_batch(kind, X0, Y0=None, *, tol=1e-8, inner_tol=1.0, maxit=100, theta_penalty=0.8,
               kappa_penalty=0.5, kappa_tol=0.1, subsolver=None, subsolver_maxit=1000, ctx=None, solver="alps"):
    """The demos' ``for i = 1:ntests`` loop over starting points (demo/eitheror.jl:238-271,
    demo/mpvca.jl:241-281, demo/rosenbrock.jl:154-177) as one launch: one alps solve per row of X0.
    Returns dict(x, y, iters, sub_iters, status, norm_res_prim, kernel_ms) -- the columns the demos
    write to CSV (id, xfinal, iters, sub_iters, is_solved).  ``solver="als"`` runs ``Bazinga.als``
    (src/algorithms/als.jl) instead, as the demos do beside alps (demo/mpvca.jl:404-418)."""
    if solver not in ("alps", "als"):
        raise ValueError("solver must be 'alps' or 'als'")
    ctx = ctx or default_context()
    lib = ctx.lib
    k = BATCH_KINDS[kind] if isinstance(kind, str) else int(kind)
    nx, ny = BATCH_DIMS[k]
    X0 = _f64(X0).reshape(-1, nx)
    B = X0.shape[0]
    Y0 = np.zeros((B, ny)) if Y0 is None else _f64(Y0).reshape(B, ny)
    sub = subsolver or PANOCplus(minimum_gamma=float(np.finfo(np.float64).eps), maxit=1000)
    o = _options(tol=tol, tol_prim=None, tol_dual=None, inner_tol=inner_tol, maxit=maxit,
                 theta_penalty=theta_penalty, kappa_penalty=kappa_penalty, kappa_tol=kappa_tol,
                 subsolver_maxit=subsolver_maxit, sub=sub)
    x, y = np.empty((B, nx)), np.empty((B, ny))
    it, st = np.empty(B, dtype=np.int32), np.empty(B, dtype=np.int32)
    inn, nrp = np.empty(B, dtype=np.int64), np.empty(B)
    ms = C.c_double()
    fn = lib.bz_batch_solve_als if solver == "als" else lib.bz_batch_solve
    K.check(fn(ctx.handle, k, B, _ptr(X0), _ptr(Y0), C.byref(o), _ptr(x), _ptr(y), _ptr(it),
               _ptr(inn), _ptr(st), _ptr(nrp), C.byref(ms)), ctx.handle)
    return dict(x=x, y=y, iters=it, sub_iters=inn, status=st, norm_res_prim=nrp, kernel_ms=ms.value)
