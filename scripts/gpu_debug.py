"""Scratch diagnostics run on the GPU box: compares the CUDA path with the oracle and prints numbers."""
import os, sys, time, traceback
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import bazinga_oracle as bo
import bazinga_b200 as bz

def section(name):
    print("\n==== " + name, flush=True)

def run(name, fn):
    section(name)
    try:
        fn()
    except Exception:
        traceback.print_exc()

def t_rng():
    ctx = bz.default_context()
    a = ctx.randn_fill(0, 12345, 100000)
    b = bo.bz_randn_like(0, 12345, 100000)
    print("rng bit-exact:", np.array_equal(a, b), a[:3], b[:3])

def t_kernels_bp():
    m, n = 64, 256
    A, xt, b = bo.make_basis_pursuit(m, n)
    p = bz.Problem(bz.Zero(), bz.NormL1(1.0), bz.ConstraintBasisPursuit(A, b), bz.ZeroSet())
    rng = np.random.default_rng(0)
    mu = rng.uniform(0.1, 2.0, m); y = rng.standard_normal(m); x = rng.standard_normal(n)
    p.auglag_update(mu, y)
    g = p.gradient(x)
    al = bo.AugLagFun(bo.FZero(), bo.CDenseAffine(A, b), bo.ZeroSet(), mu.copy(), y.copy(), x)
    dl = np.empty(n); lx = al.gradient(dl, x)
    sc = np.linalg.norm(A, 2) * np.linalg.norm(x) + np.linalg.norm(b)
    print("grad err (normwise)", np.abs(g["dlx"] - dl).max() / (np.linalg.norm(A, 2) * np.linalg.norm(al.yupd)),
          "lx rel", abs(g["lx"] - lx) / abs(lx), "cx err", np.abs(g["cx"] - al.cx).max() / sc,
          "yupd", np.abs(g["yupd"] - al.yupd).max())
    z, gz = p.prox(x, 0.3)
    zz = np.empty(n); gzz = bo.GNormL1(1.0).prox(zz, x, 0.3)
    print("prox", np.abs(z - zz).max(), gz - gzz)
    fb = p.fb_step(x, dl, 0.01)
    yy = x - 0.01 * dl; gzz = bo.GNormL1(1.0).prox(zz, yy, 0.01); res = x - zz
    print("fb", np.abs(fb["z"] - zz).max(), np.abs(fb["res"] - res).max(), fb["gz"] - gzz, fb["dot_grad_res"] - dl @ res, fb["res_sq"] - res @ res)
    # lbfgs
    p.lbfgs_reset(5); L = bo.LBFGS(n, 5)
    for k in range(8):
        s = rng.standard_normal(n); yv = s * rng.uniform(0.5, 1.5, n) + 0.01 * rng.standard_normal(n)
        p.lbfgs_update(s, yv); L.update(s, yv)
        v = rng.standard_normal(n)
        d = p.lbfgs_apply(v); dd = L.mul(v)
        print("  lbfgs", k, np.abs(d - dd).max() / np.abs(dd).max())
    p.close()

def cmp_alps(name, P, kw_gpu, r_or, x0):
    t = time.time()
    out, res = bz.alps(*P, x0, np.zeros(len(r_or.y)), return_result=True, **kw_gpu)
    dt = time.time() - t
    x = out[0]
    print(f"{name}: GPU it={out[2]} inner={out[3]} status={out[5]} nrp={out[7]:.3e} obj={res.objective:.9e} ngrad={res.n_grad} nprox={res.n_prox} launches={res.n_launches} t={dt:.3f}s")
    print(f"{name}: ORA it={r_or.tot_it} inner={r_or.tot_inner_it} status={r_or.status} nrp={r_or.norm_res_prim:.3e} obj={r_or.objective:.9e} ngrad={r_or.n_grad} nprox={r_or.n_prox}")
    print(f"{name}: |x-x_or|inf={np.abs(x - r_or.x).max():.3e} |y-y_or|inf={np.abs(out[1] - r_or.y).max():.3e} rel={np.abs(out[1] - r_or.y).max() / max(1e-300, np.abs(r_or.y).max()):.3e}")
    return out, res

def t_lasso():
    A = np.array([[1, -2, 3, -4, 5], [2, -1, 0, -1, 3], [-1, 0, 4, -3, 2], [-1, -1, -1, 1, 3]], float); b = np.array([1, 2, 3, 4.])
    lam = 0.1 * np.max(np.abs(A.T @ b))
    r = bo.alps(bo.FLeastSquares(A, b), bo.GNormL1(lam), bo.CIdentity(), bo.FreeSet(), np.zeros(5), np.zeros(5))
    out, _ = cmp_alps("lasso", (bz.LeastSquares(A, b), bz.NormL1(lam), bz.IdentityFunction(), bz.FreeSet()), {}, r, np.zeros(5))
    xs = np.array([-3.877278911564627e-01, 0, 0, 2.174149659863943e-02, 6.168435374149660e-01])
    print("lasso err vs golden", np.abs(out[0] - xs).max())

def t_qp():
    Q = np.diag([-0.5, 1.0]); q = np.array([0.3, 0.5])
    for gname, go, gg in (("box", bo.GIndBox(-1, 1), bz.IndBox(-1, 1)), ("free", bo.GZero(), bz.IndFree())):
        r = bo.alps(bo.FQuadratic(Q, q), go, bo.CIdentity(), bo.BoxSet(-1, 1), np.zeros(2), np.zeros(2))
        cmp_alps("qp-" + gname, (bz.Quadratic(Q, q), gg, bz.IdentityFunction(), bz.ClosedSet(bz.IndBox(-1, 1))), {}, r, np.zeros(2))

def t_bp(m, n):
    A, xt, b = bo.make_basis_pursuit(m, n)
    prob = bo.problem_basis_pursuit(A, b)
    t = time.time(); r = bo.solve(prob, np.zeros(n)); print("oracle time", time.time() - t)
    sub = bz.PANOCplus(directions=bz.LBFGS(5), maxit=100_000, minimum_gamma=bo.EPS)
    cmp_alps(f"bp{m}x{n}", (bz.Zero(), bz.NormL1(1.0), bz.ConstraintBasisPursuit(A, b), bz.ZeroSet()),
             dict(subsolver=sub, subsolver_maxit=100_000), r, np.zeros(n))

def t_obstacle():
    N = 16
    prob = bo.problem_obstacle(N)
    r = bo.solve(prob, 2 * np.ones(3 * N), tol=1e-3)
    sub = bz.PANOCplus(directions=bz.LBFGS(5), maxit=1_000_000, minimum_gamma=bo.EPS)
    cmp_alps("obstacle16", (bz.SmoothCostObstacleL1(N), bz.NonsmoothCostObstacleL1(N), bz.ConstraintObstacle(N), bz.ZeroSet()),
             dict(subsolver=sub, subsolver_maxit=1_000_000, tol=1e-3), r, 2 * np.ones(3 * N))

def t_batch():
    import ctypes
    v = np.arange(-4, 8 + 1e-9, 0.25); g = np.array([(i, j) for j in v for i in v])
    t = time.time(); out = bz.alps_batch("eitheror", g); dt = time.time() - t
    print("eor batch", len(g), "solved", (out["status"] == 0).sum(), "inner med/max", np.median(out["sub_iters"]), out["sub_iters"].max(), "kernel_ms", out["kernel_ms"], "wall", dt)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from tiny_oracle_binding import tiny_batch
    ref = tiny_batch(1, g, min_gamma=bo.EPS, sub_maxit=1000, subsolver_maxit=1000)
    print("  bit-exact x:", np.array_equal(out["x"], ref["x"]), "iters equal:", np.array_equal(out["sub_iters"], ref["sub_iters"]), "maxdx", np.abs(out["x"] - ref["x"]).max())

if __name__ == "__main__":
    run("rng", t_rng)
    run("kernels_bp", t_kernels_bp)
    run("lasso", t_lasso)
    run("qp", t_qp)
    run("bp 20x100", lambda: t_bp(20, 100))
    run("bp 256x1024", lambda: t_bp(256, 1024))
    run("obstacle", t_obstacle)
    run("batch", t_batch)
