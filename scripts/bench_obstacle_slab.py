"""Obstacle problem (demo/obstacle.jl, 2-D extension = config C4) slab-decomposed over N GPUs: one
process per GPU under torchrun, the grid split by rows; boundary rows and every reduction scalar of the PANOC+ state
machine are exchanged over peer memory (cudaIpc-mapped slabs, NVLink) from inside the kernels.  Prints one JSON line.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 \
      scripts/bench_obstacle_slab.py --side 4096 --inner 200
  --check        also run the same problem un-decomposed on rank 0 and compare (small sides)
  --same-device  all ranks on cuda:0 (functional test on a one-GPU box; processes are time-sliced)
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


def local_slices(v, N, i0, i1):
    """[x1; x2; x3] (each N) -> this rank's [x1[i0:i1]; x2[i0:i1]; x3[i0:i1]]"""
    return np.concatenate([v[i0:i1], v[N + i0:N + i1], v[2 * N + i0:2 * N + i1]])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--side", type=int, default=256)
    ap.add_argument("--inner", type=int, default=100)
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--same-device", action="store_true")
    ap.add_argument("--cpu-baseline", type=int, default=0, metavar="K",
                    help="rank 0 also times K inner iterations of the CPU restatement (numpy) of the same problem")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import bazinga_b200 as bz
    from bazinga_b200.solver import _options
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    dev = 0 if args.same_device else local
    torch.cuda.set_device(dev)
    dist.init_process_group("gloo" if args.same_device else "nccl")
    S = args.side
    assert S % world == 0, "the grid rows must split evenly over the ranks"
    gy, N = S // world, S * S
    Nl = gy * S
    i0, i1 = rank * Nl, (rank + 1) * Nl
    ctx = bz.Context(dev)
    p = bz.Problem(bz.SmoothCostObstacleL1(Nl), bz.NonsmoothCostObstacleL1(Nl),
                   bz.ConstraintObstacle(Nl, (gy, S), slab=(rank, world) if world > 1 else None), bz.ZeroSet(), ctx=ctx)
    if world > 1:
        handles = [None] * world
        dist.all_gather_object(handles, p.xr_handle())
        p.xr_attach(handles)
    dist.barrier()
    line = {"config": "C4 obstacle, slab-decomposed", "grid": [S, S], "n_gpus": world, "rows_per_gpu": gy, "nx": 3 * N}
    sub = bz.PANOCplus(directions=bz.LBFGS(5), maxit=1_000_000, minimum_gamma=EPS)

    def opts(cap):
        return _options(tol=1e-4, tol_prim=None, tol_dual=None, inner_tol=None, maxit=100, theta_penalty=0.8, kappa_penalty=0.5,
                        kappa_tol=0.1, subsolver_maxit=1_000_000, sub=sub, max_inner_total=cap)

    ref = None
    if args.check:
        # (a) one AL-gradient at a random point: elementwise outputs must equal the un-decomposed kernel's bit for bit
        rng = np.random.default_rng(7)
        xg, mug, yg = rng.standard_normal(3 * N), rng.uniform(0.1, 2.0, N), rng.standard_normal(N)
        p.auglag_update(mug[i0:i1], yg[i0:i1])
        g = p.gradient(local_slices(xg, N, i0, i1))
        if rank == 0:
            ctx1 = bz.Context(dev)
            p1 = bz.Problem(bz.SmoothCostObstacleL1(N), bz.NonsmoothCostObstacleL1(N), bz.ConstraintObstacle(N, (S, S)), bz.ZeroSet(), ctx=ctx1)
            p1.auglag_update(mug, yg)
            ref = p1.gradient(xg)
        parts = [None] * world
        dist.all_gather_object(parts, {k: g[k] for k in ("dlx", "cx", "yupd", "lx", "fx")})
        if rank == 0:
            dl = np.concatenate([np.concatenate([q["dlx"][b * Nl:(b + 1) * Nl] for q in parts]) for b in range(3)])
            line["gradient_check"] = {
                "dlx_bit_identical": bool(np.array_equal(dl, ref["dlx"])),
                "cx_bit_identical": bool(np.array_equal(np.concatenate([q["cx"] for q in parts]), ref["cx"])),
                "yupd_bit_identical": bool(np.array_equal(np.concatenate([q["yupd"] for q in parts]), ref["yupd"])),
                "lx_rel_err": abs(parts[0]["lx"] - ref["lx"]) / abs(ref["lx"]), "fx_rel_err": abs(parts[0]["fx"] - ref["fx"]) / abs(ref["fx"]),
                "lx_identical_on_all_ranks": len({q["lx"] for q in parts}) == 1}
    x0, y0 = 2.0 * np.ones(3 * Nl), np.zeros(Nl)
    for cap in (min(10, args.inner), args.inner):
        ctx.sync(); dist.barrier(); ctx.sync()
        ctx.timer_start()
        x, y, s, mu, res = p.alps(x0, y0, opts(cap))
        ms = ctx.timer_stop()
    t = torch.tensor([ms], dtype=torch.float64)
    if not args.same_device:
        t = t.cuda()
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    bytes_it = (res.n_grad * 64 * N + res.tot_inner_it * 8 * 3 * N * 37) / max(1, res.tot_inner_it)
    line.update({"inner_iterations": res.tot_inner_it, "outer_iterations": res.tot_it, "n_grad": res.n_grad, "ms": ms,
                 "iterations_per_s": res.tot_inner_it / (ms * 1e-3), "objective": res.objective, "norm_res_prim": res.norm_res_prim,
                 "algorithmic_GBs_total": bytes_it * res.tot_inner_it / (ms * 1e-3) / 1e9,
                 "kernel_breakdown_us": {nm: (1e3 * res.kind_ms[k] / res.kind_launches[k] if res.kind_launches[k] else None)
                                         for k, nm in enumerate(["al_gradient", "lbfgs_step", "fb_step", "decision_step"])}})
    if args.check:
        xs = [None] * world
        dist.all_gather_object(xs, (x, res.tot_inner_it, res.objective))
        if rank == 0:
            xg = np.concatenate([np.concatenate([q[0][b * Nl:(b + 1) * Nl] for q in xs]) for b in range(3)])
            x1, _, _, _, r1 = p1.alps(2.0 * np.ones(3 * N), np.zeros(N), opts(args.inner))
            line["solve_check"] = {"inner_iterations_single_gpu": r1.tot_inner_it, "objective_single_gpu": r1.objective,
                                   "max_abs_dx": float(np.max(np.abs(xg - x1))), "max_abs_x": float(np.max(np.abs(x1))),
                                   "same_counts_on_all_ranks": len({q[1] for q in xs}) == 1,
                                   "same_objective_on_all_ranks": len({q[2] for q in xs}) == 1}
    if args.cpu_baseline > 0 and rank == 0:
        # the reference's CPU path cannot run here (no Julia): its numpy restatement on the same grid, a bounded sample
        import time
        import bazinga_oracle as bo
        t0 = time.perf_counter()
        r = bo.solve(bo.problem_obstacle(N, (S, S)), 2.0 * np.ones(3 * N), tol=1e-4, max_inner_total=args.cpu_baseline)
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": r.tot_inner_it / dt, "unit": "iterations/s", "cores": os.cpu_count(), "kind": "port",
                                "sample": f"first {r.tot_inner_it} PANOC+ inner iterations ({r.n_grad} AL-gradients), numpy restatement"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    p.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
