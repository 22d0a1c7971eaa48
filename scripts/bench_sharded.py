"""Row-sharded basis pursuit over N GPUs: one process per GPU under torchrun,
A generated on the device shard by shard, one NCCL allreduce of the n-vector A_g' w_g (+ scalar) per
AL-gradient.  Prints one JSON line on rank 0.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      scripts/bench_sharded.py --rows 32768 --cols 262144 --inner 40
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
EPS = 2.220446049250313e-16


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=4096)
    ap.add_argument("--cols", type=int, default=16384)
    ap.add_argument("--inner", type=int, default=200)
    ap.add_argument("--check", action="store_true", help="also solve unsharded on rank 0 and compare (small sizes)")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import bazinga_b200 as bz
    import bazinga_oracle as bo
    from bazinga_b200.solver import _options
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = bz.Context(local)
    ids = [bz.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    ctx.comm_init(ids[0], world, rank)
    m, n = args.rows, args.cols
    r0, ml = bz.row_shard(m, world, rank)
    xtrue = bo.bz_sparse_randn(1, n, 0.1)
    # A is generated on the device shard by shard; b = A xtrue comes from the library's own dense pass (c(x) with
    # b = 0), and a sample of its rows is checked against the oracle's generator + a host dot product
    c = bz.ConstraintBasisPursuit(None, np.zeros(ml), gen_seed=0, shape=(ml, n), row_begin=r0, ny_global=m)
    p = bz.Problem(bz.Zero(), bz.NormL1(1.0), c, bz.ZeroSet(), ctx=ctx)
    p.auglag_update(np.ones(ml), np.zeros(ml))
    b = p.gradient(xtrue)["cx"]
    for r in np.linspace(0, ml - 1, 8).astype(int):
        row = bo.bz_randn_like(0, (r0 + int(r)) * n, n)
        assert abs(b[r] - row @ xtrue) <= 1e-10 * (np.linalg.norm(row) * np.linalg.norm(xtrue)), "device b differs from the oracle"
    p.upload_dense(None, b)
    sub = bz.PANOCplus(directions=bz.LBFGS(5), maxit=100_000, minimum_gamma=EPS)
    res = None
    for cap in (8, args.inner):
        o = _options(tol=1e-6, tol_prim=None, tol_dual=None, inner_tol=None, maxit=100, theta_penalty=0.8,
                     kappa_penalty=0.5, kappa_tol=0.1, subsolver_maxit=100_000, sub=sub, max_inner_total=cap)
        ctx.sync(); dist.barrier(); ctx.sync()
        ctx.timer_start()
        x, y, s, mu, res = p.alps(np.zeros(n), np.zeros(m), o)
        ms = ctx.timer_stop()
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    # lock-step check: every rank must hold bit-identical x
    xs = torch.from_numpy(x).cuda()
    xmax, xmin = xs.clone(), xs.clone()
    dist.all_reduce(xmax, op=dist.ReduceOp.MAX)
    dist.all_reduce(xmin, op=dist.ReduceOp.MIN)
    identical = bool(torch.equal(xmax, xmin))
    line = {"config": "row-sharded basis pursuit", "rows": m, "cols": n, "n_gpus": world, "rows_per_gpu": ml,
            "inner_iterations": res.tot_inner_it, "ms": float(t.item()), "iterations_per_s": res.tot_inner_it / (float(t.item()) * 1e-3),
            "n_grad": res.n_grad, "x_bit_identical_across_ranks": identical,
            "kernel_breakdown_us": {nm: (1e3 * res.kind_ms[k] / res.kind_launches[k] if res.kind_launches[k] else None)
                                    for k, nm in enumerate(["al_gradient_local", "lbfgs_step", "fb_step", "decision_step"])},
            "al_gradient_local_GBs": (8 * ml * n + 8 * (2 * n + 5 * ml)) / (1e-6 * (1e3 * res.kind_ms[0] / max(1, res.kind_launches[0]))) / 1e9,
            "iteration_GBs_per_gpu": (res.n_grad * (8 * ml * n + 8 * (2 * n + 5 * ml)) + res.tot_inner_it * 8 * n * 37) / (float(t.item()) * 1e-3) / 1e9,
            "collective": f"ncclAllReduce(sum) of {n + 8} doubles per AL-gradient", "objective": res.objective,
            "norm_res_prim": res.norm_res_prim}
    if args.check and rank == 0:
        A = np.vstack([ctx.randn_fill(0, r * n, min(blk, m - r) * n).reshape(-1, n) for r in range(0, m, blk)])
        ctx1 = bz.Context(local)
        out = bz.alps(bz.Zero(), bz.NormL1(1.0), bz.ConstraintBasisPursuit(A, A @ xtrue), bz.ZeroSet(), np.zeros(n), np.zeros(m),
                      subsolver=sub, subsolver_maxit=100_000, max_inner_total=args.inner, ctx=ctx1, return_result=True)
        line["unsharded_check"] = {"inner_iterations": out[0][3], "max_abs_dx": float(np.max(np.abs(out[0][0] - x))),
                                   "objective": out[1].objective}
    if rank == 0:
        print(json.dumps(line), flush=True)
    p.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
