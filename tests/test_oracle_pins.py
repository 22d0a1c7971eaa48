"""Pins the CPU oracle (oracle/) against everything the reference's own tests hold for this path
(SURVEY.md §8c): the golden vector + iteration bounds of test/problems/test_verbose.jl and the
fixed-point residual inequalities of test/problems/test_nonconvex_qp.jl, plus internal consistency of
the three restatements (numpy oracle, C tiny oracle, counter-based generator)."""
import json
import os

import numpy as np
import pytest

import bazinga_oracle as bo
from tiny_oracle_binding import oracle_randn_fill, tiny_batch

HERE = os.path.dirname(os.path.abspath(__file__))


def test_reference_test_verbose_golden_vector():
    # test/problems/test_verbose.jl:7-44
    A = np.array([[1, -2, 3, -4, 5], [2, -1, 0, -1, 3], [-1, 0, 4, -3, 2], [-1, -1, -1, 1, 3]], float)
    b = np.array([1.0, 2.0, 3.0, 4.0])
    lam = 0.1 * np.max(np.abs(A.T @ b))
    x_star = np.array([-3.877278911564627e-01, 0, 0, 2.174149659863943e-02, 6.168435374149660e-01])
    x0, y0 = np.zeros(5), np.zeros(5)
    r = bo.alps(bo.FLeastSquares(A, b), bo.GNormL1(lam), bo.CIdentity(), bo.FreeSet(), x0, y0)
    assert r.x.dtype == np.float64
    assert np.max(np.abs(r.x - x_star)) <= 1e-4       # :42
    assert r.tot_it < 10 and r.tot_inner_it < 50       # :43-44
    assert r.status == "first_order"
    assert not x0.any()


@pytest.mark.parametrize("gname", ["box", "free"])
def test_reference_nonconvex_qp_tiny(gname):
    # test/problems/test_nonconvex_qp.jl:8-52
    Q, q = np.diag([-0.5, 1.0]), np.array([0.3, 0.5])
    g = bo.GIndBox(-1.0, 1.0) if gname == "box" else bo.GZero()
    gamma = 0.95 / 1.0
    x0 = np.zeros(2)
    r = bo.alps(bo.FQuadratic(Q, q), g, bo.CIdentity(), bo.BoxSet(-1.0, 1.0), x0, np.zeros(2))
    z = np.clip(r.x - gamma * (Q @ r.x + q), -1.0, 1.0)
    assert np.max(np.abs(r.x - z)) / gamma <= 1e-4    # :35,50
    assert not x0.any()                                # :36,51


@pytest.mark.parametrize("k", [1, 2, 3, 4, 5])
@pytest.mark.parametrize("gname", ["box", "free"])
def test_reference_nonconvex_qp_small(k, gname):
    # test/problems/test_nonconvex_qp.jl:55-107 (Julia's Random.seed!(k) stream cannot be regenerated
    # without Julia: the assertion, not the instance, is what is portable -- SURVEY.md §8c)
    rng = np.random.default_rng(k)
    n = 100
    U, _ = np.linalg.qr(rng.standard_normal((n, n)))
    ev = 2.0 * rng.random(n) - 1.0
    Q = U @ np.diag(ev) @ U.T
    Q = 0.5 * (Q + Q.T)
    q = rng.standard_normal(n)
    g = bo.GIndBox(-1.0, 1.0) if gname == "box" else bo.GZero()
    gamma = 0.95 / np.max(np.abs(ev))
    x0 = np.zeros(n)
    r = bo.alps(bo.FQuadratic(Q, q), g, bo.CIdentity(), bo.BoxSet(-1.0, 1.0), x0, np.zeros(n))
    z = np.clip(r.x - gamma * (Q @ r.x + q), -1.0, 1.0)
    assert np.max(np.abs(r.x - z)) / gamma <= 1e-4
    assert r.status == "first_order"


def test_outer_count_floor():
    # 0.1*inner_tol never lands exactly on tol_dual -> >= 6 outer iterations at tol=1e-6
    r = bo.solve(bo.problem_mpvca(), np.array([1.0, 1.0]))
    assert r.tot_it >= 10   # tol=1e-8, inner_tol=1.0


def test_pair_projections_branch_order():
    # src/projections/*.jl, incl. the set-valued ties the reference resolves in a fixed way
    assert bo.project_onto_EITHEROR_set(-1.0, -1.0) == (-1.0, 0.0)      # tie -> second entry zeroed
    assert bo.project_onto_EITHEROR_set(-1.0, -2.0) == (0.0, -2.0)
    assert bo.project_onto_EITHEROR_set(1.0, -2.0) == (1.0, -2.0)
    assert bo.project_onto_XOR_set(2.0, 1.0) == (2.0, 0.0)
    assert bo.project_onto_XOR_set(-2.0, -1.0) == (-2.0, 0.0)
    assert bo.project_onto_XOR_set(1.0, 1.0) == (0.0, 1.0)
    assert bo.project_onto_VC_set(1.0, -1.0) == (0.0, -1.0)             # x1 + x2 == 0 keeps z2
    assert bo.project_onto_VC_set(2.0, -1.0) == (2.0, 0.0)
    assert bo.project_onto_VC_set(-1.0, -3.0) == (0.0, -3.0)
    assert bo.project_onto_CC_set(1.0, 1.0) == (1.0, 0.0)               # tie -> second zeroed
    assert bo.project_onto_CC_set(1.0, 2.0) == (0.0, 2.0)
    assert bo.project_onto_CC_set(-1.0, 2.0) == (0.0, 2.0)
    # vectorised forms agree with the scalar ones on a grid that contains every tie
    v = np.arange(-2.0, 2.01, 0.5)
    a, b = np.meshgrid(v, v)
    a, b = a.ravel(), b.ravel()
    for kind, fn in bo._PAIR_PROJ.items():
        z1, z2 = bo._pair_proj_vec(kind, a, b)
        ref = np.array([fn(float(p), float(q)) for p, q in zip(a, b)])
        assert np.array_equal(z1, ref[:, 0]) and np.array_equal(z2, ref[:, 1]), kind


def test_prox_quirks_kept():
    # normL0Box.jl:35,40-44: threshold sqrt(gamma*lambda) and the over-the-box branch returns x, not u
    g = bo.GNormL0Box(1.0, np.array([1.0, 1.0, 0.0, 5.0]))
    z = np.empty(4)
    val = g.prox(z, np.array([3.0, 0.5, 9.0, 1.5]), 1.0)
    assert z.tolist() == [3.0, 0.0, 0.0, 1.5] and val == 2.0
    with pytest.raises(ValueError):
        bo.GNormL1Box(-1.0, np.ones(2))
    with pytest.raises(ValueError):
        bo.GNormL0Box(1.0, -np.ones(2))
    with pytest.raises(ValueError):
        bo.AugLagFun(bo.FZero(), bo.CIdentity(), bo.FreeSet(), np.array([1.0, 0.0]), np.zeros(2), np.zeros(2))


def test_philox_known_answer_and_generators_agree():
    # Random123 known-answer vectors for Philox-4x32-10
    z = np.zeros(1, dtype=np.uint64)
    out = [int(v[0]) for v in bo.philox4x32_10(z, z, 0, 0)]
    assert out == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    a = oracle_randn_fill(7, (1 << 33) + 5, 4096)
    b = bo.bz_randn_like(7, (1 << 33) + 5, 4096)
    assert np.array_equal(a, b)
    big = oracle_randn_fill(0, 0, 1 << 20)
    assert abs(big.mean()) < 5e-3 and abs(big.std() - 1.0) < 5e-3


@pytest.mark.parametrize("kind,name", [(0, "rosenbrock"), (1, "eitheror"), (3, "mpvca")])
def test_tiny_c_oracle_matches_numpy_oracle(kind, name):
    prob = {0: bo.problem_rosenbrock, 1: bo.problem_eitheror, 3: bo.problem_mpvca}[kind]()
    mg = 1e-32 if kind == 0 else bo.EPS
    sm = 10**9 if kind == 0 else 1000
    pts = np.array([(-3.0, 2.5), (0.25, -1.75), (4.0, 4.0), (-1.0, 6.5), (2.0, -2.0), (7.5, 0.5)])
    ref = tiny_batch(kind, pts, min_gamma=mg, sub_maxit=sm, subsolver_maxit=sm)
    agree = 0
    for i, x0 in enumerate(pts):
        r = bo.solve(prob, x0.copy())
        assert np.max(np.abs(r.x - ref["x"][i])) <= 1e-6, (name, i)
        assert (r.status == "first_order") == (ref["status"][i] == 0)
        agree += int(r.tot_inner_it == ref["sub_iters"][i])
    assert agree >= 4   # summation order differs (BLAS dot vs left-to-right); counts nearly always agree


def test_demo_grid_expectations():
    # either-or 2400/2401 solved, MPVC 2601/2601 (demo grids, column-major order)
    v = np.arange(-4.0, 8.0 + 1e-9, 0.25)
    g = np.array([(i, j) for j in v for i in v])
    r = tiny_batch(1, g, min_gamma=bo.EPS, sub_maxit=1000, subsolver_maxit=1000)
    assert (r["status"] == 0).sum() == 2400
    glob = np.sum(np.max(np.abs(r["x"] - np.array([2.0, -2.0])), axis=1) < 1e-4)
    assert glob > 1800
    v = np.arange(-5.0, 20.0 + 1e-9, 0.5)
    g = np.array([(i, j) for j in v for i in v])
    r = tiny_batch(3, g, min_gamma=bo.EPS, sub_maxit=1000, subsolver_maxit=1000)
    assert (r["status"] == 0).sum() == 2601
    assert np.median(r["sub_iters"]) == 10


def test_golden_fixtures_frozen():
    # tests/golden/*.json were written by tests/golden/make_golden.py from this oracle; a change in
    # the oracle's arithmetic shows up here before it can silently move the GPU parity target
    with open(os.path.join(HERE, "golden", "oracle_golden.json")) as fh:
        gold = json.load(fh)
    A, _, b = bo.make_basis_pursuit(20, 100)
    r = bo.solve(bo.problem_basis_pursuit(A, b), np.zeros(100))
    assert r.tot_it == gold["bp20x100"]["tot_it"]
    assert abs(r.tot_inner_it - gold["bp20x100"]["tot_inner_it"]) <= 0.05 * gold["bp20x100"]["tot_inner_it"]
    assert np.max(np.abs(r.x - np.array(gold["bp20x100"]["x"]))) <= 1e-5
    assert abs(r.objective - gold["bp20x100"]["objective"]) <= 1e-6 * abs(r.objective)
    al = bo.AugLagFun(bo.FZero(), bo.CDenseAffine(A, b), bo.ZeroSet(), np.array(gold["grad"]["mu"]),
                      np.array(gold["grad"]["y"]), np.zeros(100))
    dl = np.empty(100)
    lx = al.gradient(dl, np.array(gold["grad"]["x"]))
    assert abs(lx - gold["grad"]["lx"]) <= 1e-12 * abs(lx)
    assert np.max(np.abs(dl - np.array(gold["grad"]["dlx"]))) <= 1e-11 * np.max(np.abs(dl))


def test_als_c_oracle_matches_numpy_restatement():
    """`als` (src/algorithms/als.jl:7-120 with src/utilities/auglagfunslack.jl) has no test in the reference; the two
    independent restatements (numpy, following the Julia source statement by statement; plain C, the batched
    kernel's bit-exact oracle) must agree on the demo problems.  These problems are nonconvex with a handful of
    stationary points, and which one a start ends in -- like the iteration counts -- moves with the summation order
    of the 6-element dot products (numpy's dot vs the C loop): most starts must agree point for point, all must end
    in the same solution set, and als must end in the solution set of alps."""
    from tiny_oracle_binding import tiny_batch
    v = np.arange(-4.0, 8.0 + 1e-9, 2.0)
    X0 = np.array([(i, j) for j in v for i in v])
    for kind, pr in ((1, bo.problem_eitheror()), (3, bo.problem_mpvca())):
        kw = dict(pr["kw"])
        ra = tiny_batch(kind, X0, min_gamma=bo.EPS, sub_maxit=1000, subsolver_maxit=1000, solver="als")
        rp = tiny_batch(kind, X0, min_gamma=bo.EPS, sub_maxit=1000, subsolver_maxit=1000, solver="alps")
        assert (ra["status"] == 0).all()
        xs_np, tot_np, same = [], 0, 0
        for i, x0 in enumerate(X0):
            o = bo.als(pr["f"], pr["g"], pr["c"], pr["D"], x0.copy(), np.zeros(pr["ny"]), **kw)
            assert o.status == "first_order"
            xs_np.append(o.x.copy())
            tot_np += o.tot_it
            same += np.max(np.abs(o.x - ra["x"][i])) <= 1e-6
        xs_np = np.array(xs_np)
        assert same >= 0.8 * len(X0)
        assert abs(tot_np - int(ra["iters"].sum())) <= 0.2 * tot_np
        for x in xs_np:
            assert np.min(np.max(np.abs(ra["x"] - x), axis=1)) <= 1e-5
        for x in ra["x"]:
            assert np.min(np.max(np.abs(xs_np - x), axis=1)) <= 1e-5
            assert np.min(np.max(np.abs(rp["x"] - x), axis=1)) <= 1e-5


def test_slack_variant_oracles_agree():
    """The demos' slack formulations (kinds 4..7 of the batched path: demo/eitheror.jl:76-119, demo/mpvca.jl:95-143): the
    plain-C oracle and the numpy restatement (separately written problem classes) end in the same points."""
    from tiny_oracle_binding import tiny_batch
    G = np.array([(1.0, 2.0), (-3.0, 5.0)])
    for kind, name, pr in ((4, "eitheror_fullslack", bo.problem_eitheror_fullslack()), (5, "xor_fullslack", bo.problem_eitheror_fullslack("xor")),
                           (6, "mpvca_slack", bo.problem_mpvca_slack()), (7, "mpvca_fullslack", bo.problem_mpvca_slack(True))):
        X0 = bo.slack_start(name, G)
        assert X0.shape == (2, pr["nx"])
        r = tiny_batch(kind, X0, min_gamma=bo.EPS, sub_maxit=1000, subsolver_maxit=1000)
        assert (r["status"] == 0).all()
        for i, x0 in enumerate(X0):
            o = bo.alps(pr["f"], pr["g"], pr["c"], pr["D"], x0.copy(), np.zeros(pr["ny"]), **pr["kw"])
            assert o.status == "first_order" and np.max(np.abs(o.x - r["x"][i])) <= 1e-6
            cx = np.empty(pr["ny"])
            pr["c"].eval(cx, r["x"][i])
            assert np.max(np.abs(cx)) <= 1e-7            # D = {0}: the slacks equal the constraint values


def test_portfolio_oracles_and_als_solves():
    """demo/portfolio.jl: SetP projection (:60-69), ConstraintP eval!/jtprod! (:46-58), SmoothCostP (:33-44), and the
    demo's solves through als on a synthetic data set (the reference's CSV data is not in its repository): every
    solve ends :first_order at a feasible point; L0 / Lp solutions are sparse, warm-started L0 keeps the support small."""
    D = bo.SetPortfolio(0.3)
    z = np.empty(2)
    D.proj(z, np.array([0.1, 7.0]))
    assert z.tolist() == [0.3, 1.0]
    D.proj(z, np.array([0.9, -2.0]))
    assert z.tolist() == [0.9, 1.0]
    z2 = np.empty(2)
    D.proj(z2, z)
    assert np.array_equal(z, z2)                                   # idempotent
    rng = np.random.default_rng(0)
    n = 60
    Q, rho, mu, ub = bo.make_portfolio(n, seed=5)
    assert np.allclose(Q, Q.T) and np.linalg.eigvalsh(Q)[0] > 0 and ub.sum() > 1 and mu.min() < rho < mu.max()
    c = bo.CPortfolio(mu)
    x, v = rng.standard_normal(n), rng.standard_normal(2)
    cx, jtv = np.empty(2), np.empty(n)
    c.eval(cx, x)
    c.jtprod(jtv, x, v)
    assert abs(cx @ v - x @ jtv) <= 1e-12 * (1 + abs(cx @ v))      # <c(x), v> = <x, J'v> for the linear c
    f = bo.FQuadratic(Q, np.zeros(n))
    d = np.empty(n)
    assert abs(f.gradient(d, x) - 0.5 * x @ Q @ x) <= 1e-12 * abs(x @ Q @ x) and np.allclose(d, Q @ x)
    res = {}
    for g in ("l0", "lp", "l1"):
        P = bo.problem_portfolio(Q, rho, mu, ub, g)
        with np.errstate(all="ignore"):
            r = bo.als(P["f"], P["g"], P["c"], P["D"], np.ones(n), np.zeros(2), **P["kw"])
        res[g] = r
        assert r.status == "first_order" and r.norm_res_prim <= 1e-6
        assert abs(r.x.sum() - 1) <= 1e-5 and mu @ r.x >= rho - 1e-5
        assert np.all(r.x >= 0)
        if g != "l0":     # normL0Box.jl:41-46 keeps x[i] (not u[i]) in its over-the-box branch: L0 iterates may exceed u
            assert np.all(r.x <= ub + 1e-12)
    assert (res["l0"].x > 0).sum() < n // 2 and (res["lp"].x > 0).sum() < n // 2
    P = bo.problem_portfolio(Q, rho, mu, ub, "l0")
    with np.errstate(all="ignore"):
        r = bo.als(P["f"], P["g"], P["c"], P["D"], res["lp"].x, res["lp"].y, **P["kw"])     # Lp + L0, portfolio.jl:170-173
    assert r.status == "first_order" and (r.x > 0).sum() <= (res["lp"].x > 0).sum()


def test_obstacle_l2_oracles():
    """demo/obstacle.jl:21-35 (SmoothCostObstacleL2), :49-62 (NonsmoothCostObstacleL2): gradient matches finite
    differences of the value, the prox is max(0, .) followed by the complementarity zeroing (tie zeroes block 2) and
    returns 0; a small obstacle_l2 solve ends first-order with complementary blocks 2 and 3."""
    N = 7
    rng = np.random.default_rng(3)
    f = bo.FObstacleL2(N)
    x = rng.standard_normal(3 * N)
    d = np.empty(3 * N)
    fx = f.gradient(d, x)
    assert abs(fx - (0.5 * np.sum(x[:2 * N] ** 2) - np.sum(x[N:2 * N]))) <= 1e-12 * (1 + abs(fx))
    for j in (0, N + 2, 2 * N + 1):
        e = np.zeros(3 * N); e[j] = 1e-6
        assert abs((f.value(x + e) - f.value(x - e)) / 2e-6 - d[j]) <= 1e-6
    g = bo.GObstacleL2(N)
    xin = np.concatenate([rng.standard_normal(N), [1.0, 2.0, -1.0, 0.5, 0.5, 3.0, -2.0], [2.0, 1.0, 4.0, 0.5, -1.0, -3.0, -5.0]])
    z = np.empty(3 * N)
    assert g.prox(z, xin, 0.7) == 0.0
    assert np.array_equal(z[:N], np.maximum(0, xin[:N]))
    assert z[N:2 * N].tolist() == [0.0, 2.0, 0.0, 0.0, 0.5, 3.0, 0.0] and z[2 * N:].tolist() == [2.0, 0.0, 4.0, 0.5, 0.0, 0.0, 0.0]
    P = bo.problem_obstacle_l2(8)
    r = bo.alps(P["f"], P["g"], P["c"], P["D"], 2.0 * np.ones(24), np.zeros(8), tol=1e-4, **P["kw"])
    assert r.status == "first_order" and r.norm_res_prim <= 1e-4
    assert np.all((r.x[8:16] == 0) | (r.x[16:] == 0)) and np.all(r.x >= 0)


def test_golden_fixtures_widened_rows_frozen():
    """tests/golden/oracle_golden_widen.json (written by make_golden_widen.py): the synthetic portfolio data set, an Lp
    portfolio solve, the obstacle_l2 AL-gradient / prox and an Lp prox sweep, recomputed and compared."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden_widen", os.path.join(HERE, "golden", "make_golden_widen.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    with open(os.path.join(HERE, "golden", "oracle_golden_widen.json")) as fh:
        gold = json.load(fh)
    now = mod.build()
    gp, npf = gold["portfolio40"], now["portfolio40"]
    for k in ("Q_row0", "mu", "ub"):
        assert np.array_equal(np.array(gp[k]), np.array(npf[k])), k      # counter-based generator: the same bits everywhere
    assert gp["rho"] == npf["rho"] and abs(gp["Q_trace"] - npf["Q_trace"]) <= 1e-12 * gp["Q_trace"]
    assert gp["lp"]["status"] == npf["lp"]["status"] == "first_order" and gp["lp"]["tot_it"] == npf["lp"]["tot_it"]
    assert abs(gp["lp"]["tot_inner_it"] - npf["lp"]["tot_inner_it"]) <= 0.1 * gp["lp"]["tot_inner_it"] + 5
    assert gp["lp"]["nnz"] == npf["lp"]["nnz"] and abs(gp["lp"]["objective_q"] - npf["lp"]["objective_q"]) <= 1e-6 * gp["lp"]["objective_q"]
    assert np.max(np.abs(np.array(gp["lp"]["x"]) - np.array(npf["lp"]["x"]))) <= 1e-6
    go, no = gold["obstacle_l2_N6"], now["obstacle_l2_N6"]
    assert abs(go["lx"] - no["lx"]) <= 1e-12 * abs(go["lx"]) and np.max(np.abs(np.array(go["dlx"]) - np.array(no["dlx"]))) <= 1e-12
    assert np.array_equal(np.array(go["prox"]), np.array(no["prox"]))
    assert np.max(np.abs(np.array(gold["lp_prox"]["z"]) - np.array(now["lp_prox"]["z"]))) <= 1e-12
    assert abs(gold["lp_prox"]["g"] - now["lp_prox"]["g"]) <= 1e-12


def test_matrix_prox_restatement_properties():
    """rank.jl / nuclearNorm.jl / schattenNormLp.jl restated on numpy's LAPACK svd: closed forms on diagonal input, the
    prox minimises lambda*g(Y) + ||Y - X||^2 / (2 gamma) against random competitors, and the reference's constructor errors."""
    import pytest
    rng = np.random.default_rng(11)
    d = np.array([3.0, 1.0, 0.2])
    Y = np.empty((3, 3))
    assert bo.GNuclearNorm(0.5).prox(Y, np.diag(d), 1.0) == 1.5 and np.allclose(np.diag(Y), [2.5, 0.5, 0.0])
    assert bo.GRank(0.5).prox(Y, np.diag(d), 1.0) == 0.5 and np.allclose(np.diag(Y), [3.0, 0.0, 0.0])   # strict > at the tie
    assert bo.GRank(2.0)(np.diag(d)) == 6.0
    X = rng.standard_normal((6, 6))
    for g in (bo.GNuclearNorm(0.7), bo.GRank(0.7), bo.GSchattenNormLpPower(0.5, 0.8)):
        gamma = 0.4
        Y = np.empty_like(X)
        gy = g.prox(Y, X, gamma)
        # (a singular value the prox set to zero comes back from svd(Y) as ~1e-16, and its square root is 1e-8)
        assert abs(gy - g(Y)) <= 1e-6 * max(1.0, abs(gy))
        best = gy + np.sum((Y - X) ** 2) / (2 * gamma)
        for _ in range(200):
            Z = Y + 0.05 * rng.standard_normal(X.shape)
            assert g(Z) + np.sum((Z - X) ** 2) / (2 * gamma) >= best - 1e-9
    for bad in (lambda: bo.GRank(-1.0), lambda: bo.GNuclearNorm(-1.0), lambda: bo.GSchattenNormLpPower(0.0),
                lambda: bo.GSchattenNormLpPower(1.0), lambda: bo.GSchattenNormLpPower(0.5, -1.0),
                lambda: bo.check_and_reshape_as_matrix(np.zeros(35))):
        with pytest.raises(ValueError):
            bad()


def test_matrix_prox_oracle_reproduces_the_golden_vectors():
    """tests/golden/matrix_prox.json (made by make_golden_matprox.py): the restatement of rank.jl / nuclearNorm.jl /
    schattenNormLp.jl still gives these outputs -- a pin against silent changes; the GPU kernel is held to the same file."""
    import json
    G = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "matrix_prox.json")))
    for ent in G["cases"]:
        X = np.array(ent["X"])
        assert np.allclose(np.linalg.svd(X, compute_uv=False), ent["sigma"], rtol=1e-12, atol=1e-13)
        for key, g in (("rank", bo.GRank(G["lambda"])), ("nuclear", bo.GNuclearNorm(G["lambda"])),
                       ("schatten", bo.GSchattenNormLpPower(G["p"], G["alpha"]))):
            Y = np.empty_like(X)
            val = g.prox(Y, X, ent["gamma"])
            ref = ent["out"][key]
            assert np.max(np.abs(Y - np.array(ref["Y"]))) <= 1e-11 * max(1.0, np.abs(X).max()), (ent["name"], key)
            assert abs(val - ref["value"]) <= 1e-11 * max(1.0, abs(ref["value"]))
