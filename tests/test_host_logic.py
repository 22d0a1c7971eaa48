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
