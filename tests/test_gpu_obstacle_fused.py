"""The fused forward-backward kernel of the obstacle problems (csrc/bz_obfused.cu; BASELINE configs[3]) against the numpy
oracle: grad_x = gradient!(., alFun, x) (auglagfun.jl:73-86 around demo/obstacle.jl:97-113), z = prox_g(x - gamma grad_x)
(demo/obstacle.jl:67-81), res, grad_z = gradient!(., alFun, z) -- elementwise outputs bit for bit, the nine scalars to
1e-12 -- and whole solves against the unfused kernel cycle (BZ_OBSTACLE_FUSED=0 is the A/B switch)."""
import numpy as np
import pytest

import bazinga_b200 as bz
import bazinga_oracle as bo

pytestmark = pytest.mark.gpu
KTOL = 1e-12


def _problem(ctx, N, grid, l2):
    f = (bz.SmoothCostObstacleL2 if l2 else bz.SmoothCostObstacleL1)(N)
    g = (bz.NonsmoothCostObstacleL2 if l2 else bz.NonsmoothCostObstacleL1)(N)
    return bz.Problem(f, g, bz.ConstraintObstacle(N, grid), bz.ZeroSet(), ctx=ctx)


def _oracle(N, grid, l2):
    P = bo.problem_obstacle_l2(N, grid) if l2 else bo.problem_obstacle(N, grid)
    return P["f"], P["g"], P["c"]


@pytest.mark.parametrize("grid", [None, (1, 50), (7, 5), (24, 40), (37, 53), (130, 100), (64, 300)])
@pytest.mark.parametrize("l2", [False, True])
def test_fused_step_matches_oracle(gpu_ctx, grid, l2):
    N = 100 if grid is None else grid[0] * grid[1]
    rng = np.random.default_rng(N + (7 if l2 else 0))
    x, mu, y = rng.standard_normal(3 * N), rng.uniform(0.1, 2.0, N), rng.standard_normal(N)
    x[N + 3] = x[2 * N + 3] = 0.75                       # the set-valued tie of the complementarity zeroing
    p = _problem(gpu_ctx, N, grid, l2)
    p.auglag_update(mu, y)
    fo, go, co = _oracle(N, grid, l2)
    al = bo.AugLagFun(fo, co, bo.ZeroSet(), mu.copy(), y.copy(), x)
    for gamma in (0.37, 1e-3, 4.0):
        out = p.obstacle_fused_step(x, gamma)
        gx = np.empty(3 * N)
        lx = al.gradient(gx, x)
        fx = al.fx
        z = np.empty(3 * N)
        gz_val = go.prox(z, x - gamma * gx, gamma)
        res = x - z
        gzv = np.empty(3 * N)
        lz = al.gradient(gzv, z)
        fz = al.fx
        assert np.array_equal(out["grad_x"], gx)
        assert np.array_equal(out["z"], z) and np.array_equal(out["res"], res)
        assert np.array_equal(out["grad_z"], gzv)
        sc = out["scalars"]
        big = np.sum(al.yupd**2 * mu) + abs(lx) + abs(lz)
        assert abs(sc[0] - lx) <= KTOL * big and abs(sc[1] - fx) <= KTOL * max(1.0, np.sum(x * x))
        assert abs(sc[2] - gx @ res) <= KTOL * np.linalg.norm(gx) * np.linalg.norm(res)
        assert abs(sc[3] - res @ res) <= KTOL * (res @ res)
        assert abs(sc[4] - gz_val) <= KTOL * max(1.0, abs(gz_val))
        assert abs(sc[5] - lz) <= KTOL * big and abs(sc[6] - fz) <= KTOL * max(1.0, np.sum(z * z))
        assert sc[7] == np.max(np.abs(res / gamma - gx + gzv))
        assert sc[8] == np.max(np.abs(res / (gamma * 0.5) - gx + gzv))
    p.close()


def test_fused_step_is_deterministic_and_matches_the_unfused_kernels(gpu_ctx):
    gy, gx_ = 96, 200
    N = gy * gx_
    rng = np.random.default_rng(4)
    x, mu, y = rng.standard_normal(3 * N), rng.uniform(0.1, 2.0, N), rng.standard_normal(N)
    p = _problem(gpu_ctx, N, (gy, gx_), False)
    p.auglag_update(mu, y)
    a = p.obstacle_fused_step(x, 0.2)
    b = p.obstacle_fused_step(x, 0.2)
    for k in ("z", "res", "grad_x", "grad_z", "scalars"):
        assert np.array_equal(a[k], b[k])
    g = p.gradient(x)                                     # k_obstacle_c + k_obstacle_jt
    assert np.array_equal(g["dlx"], a["grad_x"])
    fb = p.fb_step(x, g["dlx"], 0.2)                      # k_fstep
    assert np.array_equal(fb["z"], a["z"]) and np.array_equal(fb["res"], a["res"])
    assert np.array_equal(p.gradient(a["z"])["dlx"], a["grad_z"])
    p.close()


@pytest.mark.parametrize("grid,l2", [(None, False), ((12, 20), False), ((12, 20), True)])
def test_fused_solve_follows_the_oracle(gpu_ctx, grid, l2):
    """the first inner iterations of alps on the obstacle problem: same counts and iterates as the numpy oracle, and the
    side-effect state (cx, s, yupd) at the returned point"""
    N = 32 if grid is None else grid[0] * grid[1]
    p = _problem(gpu_ctx, N, grid, l2)
    from bazinga_b200.solver import _options
    sub = bz.PANOCplus(directions=bz.LBFGS(5), maxit=1_000_000, minimum_gamma=bo.EPS)
    P = bo.problem_obstacle_l2(N, grid) if l2 else bo.problem_obstacle(N, grid)
    for cap in (1, 4, 12, 40):
        o = _options(tol=1e-4, tol_prim=None, tol_dual=None, inner_tol=None, maxit=100, theta_penalty=0.8, kappa_penalty=0.5,
                     kappa_tol=0.1, subsolver_maxit=1_000_000, sub=sub, max_inner_total=cap)
        x, y, s, mu, res = p.alps(2.0 * np.ones(3 * N), np.zeros(N), o)
        r = bo.solve(P, 2.0 * np.ones(3 * N), tol=1e-4, max_inner_total=cap)
        assert (res.tot_it, res.tot_inner_it) == (r.tot_it, r.tot_inner_it)
        assert res.n_grad == r.n_grad
        sc = np.max(np.abs(r.x))
        tol = 1e-12 if cap <= 4 else (1e-8 if cap <= 12 else 1e-5)      # later iterates: the summation orders of the two codes differ
        assert np.max(np.abs(x - r.x)) <= tol * sc
        assert np.max(np.abs(y - r.y)) <= tol * max(1.0, np.max(np.abs(r.y)))
        assert abs(res.objective - r.objective) <= tol * max(1.0, abs(r.objective))
        assert abs(res.norm_res_prim - r.norm_res_prim) <= tol * max(1.0, r.norm_res_prim)
    p.close()
