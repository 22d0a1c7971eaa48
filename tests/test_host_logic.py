"""Host-side logic that needs no GPU: oracle-type guards (the reference's error messages), the
subsolver plugin slot, and the row-sharding math of SURVEY.md §8e checked with a world_size-2 gloo
group on CPU against the oracle."""
import os
import socket

import numpy as np
import pytest

import bazinga_b200 as bz
import bazinga_oracle as bo


def test_constructor_guards_match_reference_messages():
    with pytest.raises(ValueError, match="nonnegative"):
        bz.NormL1Box(-1.0, u=np.ones(3))          # normL1Box.jl:18-23
    with pytest.raises(ValueError, match="nonnegative entries"):
        bz.NormL0Box(1.0, u=-np.ones(3))          # normL0Box.jl:19-24
    with pytest.raises(ValueError):
        bz.NormL1Nonneg(-0.5)                     # normL1Nonneg.jl:14-19
    with pytest.raises(bz.BazingaError):
        bz.IndicatorSet(bz.NormL1(1.0))           # not in the catalogue -> error, never a CPU fallback
    with pytest.raises(ValueError, match="smaller than one"):
        bz.NormLpPowerNonneg(1.0)                 # normLpNonneg.jl:17-20
    with pytest.raises(ValueError, match="positive"):
        bz.NormLpPowerBox(0.0, u=np.ones(2))      # normLpBox.jl:17-18
    with pytest.raises(ValueError, match="nonnegative"):
        bz.NormLpPowerBox(0.5, -1.0, u=np.ones(2))


def test_subsolver_slot_call_protocol():
    # alps.jl:64: sub_solver = subsolver(tol = inner_tol, verbose = verbose) -- class, closure or instance
    closure = lambda **kw: bz.PANOCplus(directions=bz.LBFGS(5), maxit=100_000, minimum_gamma=2.2e-16, **kw)  # noqa: E731
    for sub in (bz.PANOCplus, closure, bz.PANOCplus(maxit=7)):
        s = sub(tol=1e-3, verbose=False)
        assert isinstance(s, bz.PANOCplus) and s.tol == 1e-3
    assert closure(tol=1.0).maxit == 100_000 and closure(tol=1.0).directions.memory == 5
    assert bz.NoAcceleration().memory == 0


def test_unknown_oracle_types_raise():
    class MyF:   # an arbitrary user-defined oracle cannot cross the C ABI
        pass
    with pytest.raises(bz.BazingaError):
        bz.Problem(MyF(), bz.Zero(), bz.IdentityFunction(), bz.FreeSet(), nx=2, ny=2, ctx=object())


def test_row_shard_partition():
    for ny in (7, 8, 4096, 32768, 1001):
        for nr in (1, 2, 4, 8):
            if ny < nr:
                continue
            blocks = [bz.row_shard(ny, nr, r) for r in range(nr)]
            assert blocks[0][0] == 0 and sum(b[1] for b in blocks) == ny
            for (b0, l0), (b1, _) in zip(blocks, blocks[1:]):
                assert b0 + l0 == b1 and l0 > 0 and (l0 % 2 == 0 or ny // nr < 2)
    with pytest.raises(ValueError):
        bz.row_shard(4, 8, 0)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _shard_worker(rank, world, port, m, n, ret):
    import torch.distributed as dist
    import torch
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    A, _, b = bo.make_basis_pursuit(m, n)
    rng = np.random.default_rng(3)
    mu, y, x = rng.uniform(0.1, 2.0, m), rng.standard_normal(m), rng.standard_normal(n)
    r0, ml = bz.row_shard(m, world, rank)
    sl = slice(r0, r0 + ml)
    # the per-rank work of the row-sharded AL-gradient: local rows only
    cx = A[sl] @ x - b[sl]
    t = cx + mu[sl] * y[sl]
    w = t / mu[sl]                              # D = ZeroSet
    buf = np.concatenate([A[sl].T @ w, [np.sum(t * t / mu[sl])], [np.sum(mu[sl] * y[sl] * y[sl])]])
    tt = torch.from_numpy(buf)
    dist.all_reduce(tt)                         # the path's ONE exchange step: sum of n + 2 doubles
    nrp = torch.tensor([np.max(np.abs(cx))])
    dist.all_reduce(nrp, op=dist.ReduceOp.MAX)
    ret[rank] = (tt.numpy().copy(), float(nrp))
    dist.barrier()
    dist.destroy_process_group()


def test_row_sharded_gradient_equals_unsharded_gloo_world2():
    import torch.multiprocessing as mp
    m, n, world = 24, 40, 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_shard_worker, args=(world, port, m, n, ret), nprocs=world, join=True)
    A, _, b = bo.make_basis_pursuit(m, n)
    rng = np.random.default_rng(3)
    mu, y, x = rng.uniform(0.1, 2.0, m), rng.standard_normal(m), rng.standard_normal(n)
    al = bo.AugLagFun(bo.FZero(), bo.CDenseAffine(A, b), bo.ZeroSet(), mu.copy(), y.copy(), x)
    dl = np.empty(n)
    lx = al.gradient(dl, x)
    for rank in range(world):
        buf, nrp = ret[rank]
        assert np.max(np.abs(buf[:n] - dl)) <= 1e-12 * np.max(np.abs(dl))
        assert abs(0.5 * buf[n] - 0.5 * buf[n + 1] - lx) <= 1e-12 * abs(lx)
        assert abs(nrp - np.max(np.abs(al.cx))) <= 1e-14
    # identical bits on every rank -> replicated x-space control flow stays in lock-step
    assert np.array_equal(ret[0][0], ret[1][0])


def _slab_worker(rank, world, port, gy, gx, ret):
    """Per-rank work of the slab-decomposed obstacle AL-gradient: local grid rows, the neighbours'
    boundary rows of x2 and of yupd exchanged point to point, scalars summed in rank order."""
    import torch.distributed as dist
    import torch
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    N = gy * gx
    rng = np.random.default_rng(11)
    x, mu, y = rng.standard_normal(3 * N), rng.uniform(0.1, 2.0, N), rng.standard_normal(N)
    gl = gy // world
    r0 = rank * gl
    rows = slice(r0 * gx, (r0 + gl) * gx)
    x1, x2, x3 = (x[b * N:(b + 1) * N][rows].reshape(gl, gx) for b in range(3))

    def halo(v):   # last row of rank-1 / first row of rank+1; zeros beyond the grid (Dirichlet)
        up, dn = np.zeros(gx), np.zeros(gx)
        reqs = []
        if rank > 0:
            reqs.append(dist.isend(torch.from_numpy(v[0].copy()), rank - 1))
        if rank + 1 < world:
            reqs.append(dist.isend(torch.from_numpy(v[-1].copy()), rank + 1))
        if rank > 0:
            t = torch.empty(gx, dtype=torch.float64); dist.recv(t, rank - 1); up = t.numpy()
        if rank + 1 < world:
            t = torch.empty(gx, dtype=torch.float64); dist.recv(t, rank + 1); dn = t.numpy()
        for r in reqs:
            r.wait()
        return up, dn

    def stencil(v):   # 5-point, 4 on the diagonal, same operation order as the oracle / the kernel
        up, dn = halo(v)
        ext = np.vstack([up, v, dn])
        out = 4.0 * v
        out[:, 1:] -= v[:, :-1]
        out[:, :-1] -= v[:, 1:]
        out -= ext[:-2]
        out -= ext[2:]
        return out

    cx = (x1 + stencil(x2)) - x3
    t = cx + (mu * y)[rows].reshape(gl, gx)
    w = t / mu[rows].reshape(gl, gx)                       # D = {0}
    grad = np.concatenate([w.ravel(), ((x2 - 1.0) + stencil(w)).ravel(), (-w).ravel()])
    part = torch.tensor([np.sum(t * t / mu[rows].reshape(gl, gx)), np.sum(x2 * x2), np.sum(x2)])
    allp = [torch.empty(3, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(allp, part)                            # the mailbox all-gather; every rank adds in rank order
    tot = allp[0].clone()
    for q in allp[1:]:
        tot += q
    ret[rank] = (grad, cx.ravel(), tot.numpy().copy())
    dist.barrier()
    dist.destroy_process_group()


def test_slab_decomposed_obstacle_gradient_equals_global_gloo_world2():
    import torch.multiprocessing as mp
    gy, gx, world = 8, 6, 2
    N = gy * gx
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_slab_worker, args=(world, port, gy, gx, ret), nprocs=world, join=True)
    rng = np.random.default_rng(11)
    x, mu, y = rng.standard_normal(3 * N), rng.uniform(0.1, 2.0, N), rng.standard_normal(N)
    al = bo.AugLagFun(bo.FObstacleL1(N), bo.CObstacle(N, (gy, gx)), bo.ZeroSet(), mu.copy(), y.copy(), x)
    dl = np.empty(3 * N)
    lx = al.gradient(dl, x)
    Nl = N // world
    got = np.concatenate([np.concatenate([ret[r][0][b * Nl:(b + 1) * Nl] for r in range(world)]) for b in range(3)])
    assert np.array_equal(got, dl)                                           # elementwise + halo: exact
    assert np.array_equal(np.concatenate([ret[r][1] for r in range(world)]), al.cx)
    tot = ret[0][2]
    assert np.array_equal(ret[0][2], ret[1][2])                              # rank-order sums: identical bits
    fx = 0.5 * tot[1] - tot[2]
    assert abs((0.5 * tot[0] + fx - al.musqy) - lx) <= 1e-13 * abs(lx)


def test_demo_result_files_follow_the_reference_schema(tmp_path):
    """demo/eitheror.jl:256-273,363-383: column names and order of the per-start CSV and of the stats CSV, the
    classification of end points by the two minimisers (tolx = 1e-6), max / median columns.  The solves here come from
    the C oracle (no GPU in this test); the writer is the one the GPU path uses."""
    import csv
    from tiny_oracle_binding import tiny_batch
    v = np.arange(-4.0, 8.0 + 1e-9, 3.0)
    X0 = np.array([(i, j) for j in v for i in v])
    out = tiny_batch(1, X0, min_gamma=bo.EPS, sub_maxit=1000, subsolver_maxit=1000)
    out["kernel_ms"] = 2.5
    base = str(tmp_path / "eitheror_alps_lbfgs")
    rows, stats = bz.write_demo_results(base, "eitheror", X0, out)
    data = list(csv.reader(open(base + ".csv")))
    assert tuple(data[0]) == ("id", "xinit_1", "xinit_2", "xfinal_1", "xfinal_2", "iters", "sub_iters", "runtime", "is_solved")
    assert len(data) == 1 + len(X0) and data[1][0] == "1" and float(data[1][1]) == X0[0, 0]
    st = list(csv.reader(open(base + "_stats.csv")))
    assert tuple(st[0]) == bz.STATS_COLUMNS and len(st) == 2
    n, nsolved, it_max, it_med = stats[0], stats[1], stats[2], stats[3]
    assert n == len(X0) and nsolved == int((out["status"] == 0).sum())
    assert it_max == out["iters"].max() and it_med == float(np.median(out["iters"]))
    assert stats[8] + stats[10] + stats[12] == n and stats[8] > 0          # global (2,-2) / local (4,4) / others
    assert abs(stats[9] + stats[11] + stats[13] - 100.0) <= 1e-9


def test_division_by_fma_is_the_ieee_quotient(tmp_path):
    """bz_batched.cu replaces a / b by two FMAs on a cached reciprocal (fdiv/frecip/fmid).  tests/fdiv_check.c restates
    those three functions in C; every quotient -- everyday magnitudes, adversarial significands, operands either side
    of the range test, the whole exponent range, subnormals, zeros, infinities, NaN -- must carry the bits of the
    hardware division, or the batched kernel could not be bit-identical with the reference arithmetic."""
    import subprocess
    src = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fdiv_check.c")
    exe = str(tmp_path / "fdiv_check")
    flags = ["-O2", "-ffp-contract=off", "-fopenmp"]
    try:
        if " fma " in open("/proc/cpuinfo").read():
            flags.append("-mfma")           # hardware FMA; without it libm's exact software fma() is used (slower)
    except OSError:
        pass
    subprocess.check_call(["gcc", *flags, src, "-o", exe, "-lm"])
    per = "20000000" if "-mfma" in flags else "500000"
    out = subprocess.run([exe, per], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout
    pairs, bad, fast = (int(out.stdout.split()[i]) for i in (1, 3, 5))
    assert bad == 0 and pairs >= int(per)
    assert fast > pairs // 2               # the two-FMA path is the one being exercised


def test_portfolio_host_mirror(tmp_path):
    """demo/portfolio.jl's oracle structs map onto catalogue kinds (SmoothCostP -> dense quadratic, ConstraintP -> the
    two-row dense affine provider, SetP -> its own set), the CSV reader checks sizes like the reference's @asserts, and
    the result tables carry push_out_to_data's columns."""
    from bazinga_b200 import _capi as K
    from bazinga_b200 import portfolio as pf
    Q, rho, mu, ub = bo.make_portfolio(12, seed=1)
    f, c, D = bz.SmoothCostP(Q), bz.ConstraintP(mu), bz.SetP(rho)
    assert f.f_kind == K.F_QUADRATIC and not f.q.any() and f.Q.shape == (12, 12)
    assert c.c_kind == K.C_DENSE_AFFINE and c.shape == (2, 12) and np.array_equal(c.A[0], mu) and np.all(c.A[1] == 1) and not c.b.any()
    assert D.d_kind == K.D_PORTFOLIO and D.lo == rho
    with pytest.raises(ValueError):
        bz.SmoothCostP(np.zeros((3, 4)))
    d = tmp_path / "p1"
    d.mkdir()
    np.savetxt(d / "Q.csv", Q, delimiter=",")
    np.savetxt(d / "rho.csv", [rho], delimiter=",")
    np.savetxt(d / "mu.csv", mu, delimiter=",")
    np.savetxt(d / "ub.csv", ub, delimiter=",")
    Q2, rho2, mu2, ub2 = pf.read_FG_problem_data(str(d))
    assert np.array_equal(Q2, Q) and rho2 == rho and np.array_equal(mu2, mu) and np.array_equal(ub2, ub)
    np.savetxt(d / "ub.csv", ub[:-1], delimiter=",")
    with pytest.raises(ValueError):
        pf.read_FG_problem_data(str(d))
    tables = {k: [(1, 5, 100, 0.01, 1e-7, 0.02, 4), (2, 6, 80, 0.02, 2e-7, 0.03, 6)] for k in pf.KEYS}
    pf.write_portfolio_results(str(tmp_path / "dim12_a100"), tables)
    head = open(tmp_path / "dim12_a100_lp0.csv").readline().strip().split(",")
    assert tuple(head) == pf.COLUMNS == ("id", "iters", "subiters", "runtime", "cviolation", "objectiveq", "nnz")
    s = pf.summary(tables)
    assert s["l0"]["nnz"] == (4.0, 6.0, 5.0) and s["lp"]["subiters"] == (80.0, 100.0, 90.0)
