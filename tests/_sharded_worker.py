"""Worker of tests/test_gpu_sharded.py::test_row_sharded_two_ranks (one process per GPU, no torch): rank r of `world`
takes its row block of a small basis-pursuit instance, joins the library's NCCL communicator (unique id exchanged through
a file), and writes one AL-gradient, one multiplier update and the first iterates of the row-sharded alps solve."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)


def main():
    rank, world, tmp = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
    import bazinga_b200 as bz
    from bazinga_b200.solver import _options
    d = np.load(os.path.join(tmp, "problem.npz"))
    A, b, mu, y, xp = d["A"], d["b"], d["mu"], d["y"], d["xp"]
    m, n = A.shape
    ctx = bz.Context(rank)
    idf = os.path.join(tmp, "nccl_id.bin")
    if rank == 0:
        with open(idf + ".tmp", "wb") as f:
            f.write(bz.comm_unique_id())
        os.replace(idf + ".tmp", idf)
    t0 = time.time()
    while not os.path.exists(idf):
        if time.time() - t0 > 120:
            raise RuntimeError("no NCCL id from rank 0")
        time.sleep(0.05)
    ctx.comm_init(open(idf, "rb").read(), world, rank)
    r0, ml = bz.row_shard(m, world, rank)
    c = bz.ConstraintBasisPursuit(A[r0:r0 + ml], b[r0:r0 + ml], row_begin=r0, ny_global=m)
    p = bz.Problem(bz.Zero(), bz.NormL1(1.0), c, bz.ZeroSet(), ctx=ctx)
    p.auglag_update(mu[r0:r0 + ml], y[r0:r0 + ml])
    g = p.gradient(xp)
    yo, so, nrp = p.multiplier_update(g["cx"], mu[r0:r0 + ml], y[r0:r0 + ml])
    out = dict(r0=r0, ml=ml, dlx=g["dlx"], lx=g["lx"], cx=g["cx"], yupd=g["yupd"], s=g["s"], y_mult=yo, nrp_local=nrp)
    sub = bz.PANOCplus(directions=bz.LBFGS(5), maxit=100_000, minimum_gamma=np.finfo(float).eps)
    for cap in [int(c) for c in d["caps"]]:
        o = _options(tol=1e-6, tol_prim=None, tol_dual=None, inner_tol=None, maxit=100, theta_penalty=0.8, kappa_penalty=0.5,
                     kappa_tol=0.1, subsolver_maxit=100_000, sub=sub, max_inner_total=cap)
        x, ys, s, mus, res = p.alps(np.zeros(n), np.zeros(m), o)
        out[f"x_{cap}"] = x
        out[f"y_{cap}"] = ys
        out[f"mu_{cap}"] = mus
        out[f"it_{cap}"] = np.array([res.tot_it, res.tot_inner_it, res.n_grad])
        out[f"nrp_{cap}"] = res.norm_res_prim
        out[f"obj_{cap}"] = res.objective
    np.savez(os.path.join(tmp, f"rank{rank}.npz"), **out)
    p.close()
    ctx.close()


if __name__ == "__main__":
    main()
