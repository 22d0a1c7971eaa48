#!/usr/bin/env python
"""bench.py -- benchmark of the PANOC+/ALPS hot path (BASELINE.json metric), one JSON line per run.

Headline at every N (configs[1] of BASELINE.json, SURVEY.md §8d "C2"): demo/basispursuit.jl scaled to a
dense A 4096x16384 fp64 (synthetic, counter-based generator shared bit-for-bit with the oracle),
f = Zero, g = NormL1(1), c(x) = A x - b, D = ZeroSet, x0 = y0 = 0, LBFGS(5), minimum_gamma = eps.
A "step" is one `alps` call on that instance, capped at --inner-cap PANOC+ inner iterations in total
(the full solve is thousands of iterations; `--inner-cap 0` runs it to convergence).  The metric is
PANOC+ forward-backward inner iterations per second.

  value   whole-job inner iterations / s with A, b already resident in HBM (device-timed, CUDA events
          on the library's stream, max over ranks)
  e2e     the same steps through the public call `bazinga_b200.alps(f, g, c, D, x0, y0; ...)` with HOST
          buffers: every step allocates the problem, uploads A and b from pinned host memory, solves,
          downloads x, y, s, mu and frees the problem -- everything the reference's `alps` call does
  N > 1   independent instances, one per GPU, no collectives ("scaling": "weak")

`legs` (same JSON line, each with its own clock sample): the multi-GPU workloads of BASELINE.json configs[2..4],
partitioned the way SURVEY.md §8e says --
  c5_row_sharded  basis pursuit 32768x262144 (68.7 GB A generated on the device), STRONG: rows split over the N
                  ranks, one ncclAllReduce of A_g' w_g (+ scalar) per AL-gradient on the LIBRARY's communicator
  c3_batched      65,536 either-or + 65,536 MPVC starting points: strong (65,536 in total) and weak (65,536 per GPU)
  c4_slab         obstacle problem on a 4096x4096 grid, STRONG: grid rows split over the ranks, boundary rows and
                  reduction scalars exchanged over peer memory from inside the kernels
At N > 1 rank 0 also runs each strong workload alone (`single_gpu`), so `strong_efficiency` is measured in this run.

`--impl reference` times the CPU restatement of the reference (oracle/bazinga_oracle.py: numpy +
OpenBLAS, all host threads; the Julia reference itself cannot run here -- no Julia in the image) on the same
configuration (same instance, same cap).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
import traceback

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

M_ROWS, N_COLS = 4096, 16384
METRIC = "panoc_fwd_bwd_inner_iterations_per_s"
UNIT = "iterations/s"
EPS = 2.220446049250313e-16


def algorithmic_bytes_per_grad(m, n):
    """A read once + vector traffic of one fused AL-gradient evaluation."""
    return 8 * m * n + 8 * (2 * n + 5 * m)


def algorithmic_bytes_per_iteration_vectors(n, mem=5):
    """n-vector traffic of one PANOC+ iteration besides its AL-gradients, 8n(4 mem + 17)."""
    return 8 * n * (4 * mem + 17)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index=0):
        self.index, self.samples, self.proc, self.thr = index, [], None, None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return self
        self.thr = threading.Thread(target=self._read, daemon=True)
        self.thr.start()
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append([t.strip() for t in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            try:
                sm.append(float(s[0]))
                smax.append(float(s[1]))
                for nm, v in zip(names, s[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def traffic_from_profile(m, n):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed
    `ncu --set full` capture (profiles/); null for other shapes.  NOT measured in this run."""
    p = os.path.join(ROOT, "profiles", "r2n_dense_traffic.json")
    try:
        d = json.load(open(p))
        return d["traffic_bytes_per_launch"] if d["shape"] == [m, n] else None
    except Exception:
        return None


def dist_env():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1; the CPU legs (reference arm, cpu_baseline) use every host core and say so."""
    cores = os.cpu_count() or 1
    for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[k] = str(cores)
    return cores


# ------------------------------------------------------------------------------------------ reference arm
def run_reference(args):
    rank, _, world = dist_env()
    if rank != 0:
        return 0
    cores = use_all_host_threads()      # before numpy is imported
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import bazinga_oracle as bo
    m, n = args.rows, args.cols
    t0 = time.time()
    A = bo.bz_randn_like(0, 0, m * n).reshape(m, n) if m * n <= 1 << 22 else _gen_host_matrix(bo, m, n)
    xtrue = bo.bz_sparse_randn(1, n, 0.1)
    b = A @ xtrue
    gen_s = time.time() - t0
    prob = bo.problem_basis_pursuit(A, b)
    cap = args.inner_cap
    its, times = [], []
    for step in range(args.warmup + args.steps):
        t = time.perf_counter()
        r = bo.solve(prob, np.zeros(n), max_inner_total=cap if cap > 0 else None)
        dt = time.perf_counter() - t
        if step >= args.warmup:
            its.append(r.tot_inner_it)
            times.append(dt)
    value = sum(its) / sum(times)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": headline_config(m, n, cap, None),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{args.steps} x first {cap} PANOC+ inner iterations of the alps solve; numpy/OpenBLAS "
                                   f"restatement of the reference (Julia unavailable); one instance on all host threads "
                                   f"whatever N; data gen {gen_s:.1f}s excluded"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)
    return 0


def headline_config(m, n, cap, world):
    """identical on both arms (the driver compares the `config` objects)"""
    return {"workload": (f"basis pursuit {m}x{n} fp64 (BASELINE configs[1]), alps capped at {cap} inner iterations per step"
                         if cap else f"basis pursuit {m}x{n} fp64 (BASELINE configs[1]), full alps solve"),
            "rows": m, "cols": n, "inner_cap": cap, "lbfgs_memory": 5,
            "l2_policy": f"A is {8 * m * n / 1e6:.0f} MB > 126 MB L2 and is streamed twice per inner iteration; no explicit flush"}


def _gen_host_matrix(bo, m, n):
    """C form of the oracle's generator (oracle/tiny_oracle.c: oracle_randn_fill), same bits as
    bo.bz_randn_like but fast enough for 67M entries."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from tiny_oracle_binding import oracle_randn_fill
    return oracle_randn_fill(0, 0, m * n).reshape(m, n)


# ------------------------------------------------------------------------------------------ B200 arm
class Env:
    """what a leg needs: ranks, the library context, barriers and max/sum over ranks (device-side reductions)"""

    def __init__(self, args):
        import torch
        self.args = args
        self.rank, self.local_rank, self.world = dist_env()
        self.torch = torch
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            torch.cuda.set_device(self.local_rank)
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
            self.dist = dist
        import bazinga_b200 as bz
        self.bz = bz
        self.ctx = bz.Context(self.local_rank)
        self.peak, self.peak_src = measured_peak_hbm()
        self.comm_ready = False

    def barrier(self):
        self.ctx.sync()
        if self.dist:
            self.dist.barrier()
        self.ctx.sync()

    def _red(self, v, op):
        if not self.dist:
            return v
        t = self.torch.tensor([v], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=op)
        return float(t.item())

    def max(self, v):
        return self._red(v, self.dist.ReduceOp.MAX) if self.dist else v

    def sum(self, v):
        return self._red(v, self.dist.ReduceOp.SUM) if self.dist else v

    def gather(self, obj):
        if not self.dist:
            return [obj]
        out = [None] * self.world
        self.dist.all_gather_object(out, obj)
        return out

    def comm_init(self):
        """the LIBRARY's NCCL communicator (bz_comm_init), id broadcast through torch.distributed"""
        if self.comm_ready or self.world == 1:
            return
        ids = [self.bz.comm_unique_id() if self.rank == 0 else None]
        self.dist.broadcast_object_list(ids, src=0)
        self.ctx.comm_init(ids[0], self.world, self.rank)
        self.comm_ready = True


def bp_options(bz, cap, tol=1e-6, maxit_sub=100_000):
    from bazinga_b200.solver import _options
    sub = bz.PANOCplus(directions=bz.LBFGS(5), maxit=maxit_sub, minimum_gamma=EPS)
    return _options(tol=tol, tol_prim=None, tol_dual=None, inner_tol=None, maxit=100, theta_penalty=0.8,
                    kappa_penalty=0.5, kappa_tol=0.1, subsolver_maxit=maxit_sub, sub=sub, max_inner_total=cap), sub


def kernel_breakdown(res, names=("al_gradient", "lbfgs_step", "fb_step", "decision_step")):
    return {nm: (1e3 * res.kind_ms[k] / res.kind_launches[k] if res.kind_launches[k] else None) for k, nm in enumerate(names)}


def run_headline(env, args):
    import numpy as np
    torch, bz, ctx = env.torch, env.bz, env.ctx
    rank, world = env.rank, env.world
    m, n = args.rows, args.cols
    seed = 100 * rank            # independent instance per rank

    # ---- synthetic instance: A generated on the device (bit-identical to the oracle's generator),
    # copied to pinned host memory for the e2e leg and the CPU baseline
    A_host_t = torch.empty((m, n), dtype=torch.float64, pin_memory=True)
    A_host = A_host_t.numpy()
    rows_blk = max(1, (1 << 24) // n)
    for r in range(0, m, rows_blk):
        k = min(rows_blk, m - r)
        A_host[r:r + k] = ctx.randn_fill(seed, r * n, k * n).reshape(k, n)
    import bazinga_oracle as bo
    xtrue = bo.bz_sparse_randn(seed + 1, n, 0.1)
    b_t = torch.empty(m, dtype=torch.float64, pin_memory=True)
    b_host = b_t.numpy()
    b_host[:] = A_host @ xtrue

    opts, sub = bp_options(bz, args.inner_cap)
    cobj = bz.ConstraintBasisPursuit(A_host, b_host)
    prob = bz.Problem(bz.Zero(), bz.NormL1(1.0), cobj, bz.ZeroSet(), ctx=ctx)
    x0, y0 = np.zeros(n), np.zeros(m)

    # ---- resident leg (value): warm-up, then exactly K timed steps
    ctx.set_profiling(False)
    for _ in range(args.warmup):
        prob.alps(x0, y0, opts)
    ctx.set_profiling(True)
    sampler = ClockSampler(env.local_rank)
    env.barrier()
    if rank == 0:
        sampler.start()
    ctx.timer_start()
    its = grads = launches = dom_launches = 0
    dom_ms = 0.0
    kind_ms, kind_n = [0.0] * 4, [0] * 4
    last = None
    for _ in range(args.steps):
        x, y, s, mu, res = prob.alps(x0, y0, opts)
        its += res.tot_inner_it
        grads += res.n_grad
        launches += res.n_launches
        dom_ms += res.dominant_kernel_ms
        dom_launches += res.dominant_kernel_launches
        for k in range(4):
            kind_ms[k] += res.kind_ms[k]
            kind_n[k] += res.kind_launches[k]
        last = res
    ms = ctx.timer_stop()
    env.barrier()
    clocks = sampler.stop() if rank == 0 else None
    ctx.set_profiling(False)
    ms_max = env.max(ms)
    its_all = env.sum(float(its))
    value = its_all / (ms_max * 1e-3)
    prob.close()

    # ---- e2e leg: the public call with host buffers (allocation + H2D + solve + D2H + free), every step
    h2d = 8 * (m * n + m + n + m)
    d2h = 8 * (n + 3 * m)
    kw = dict(tol=1e-6, subsolver=sub, subsolver_maxit=100_000, max_inner_total=args.inner_cap, ctx=ctx, return_result=True)
    bz.alps(bz.Zero(), bz.NormL1(1.0), cobj, bz.ZeroSet(), x0, y0, **kw)      # warm-up of the allocation path
    env.barrier()
    t0 = time.perf_counter()
    ctx.timer_start()
    e_its = 0
    for _ in range(args.steps):
        out, res = bz.alps(bz.Zero(), bz.NormL1(1.0), cobj, bz.ZeroSet(), x0, y0, **kw)
        e_its += res.tot_inner_it
    e_ms = ctx.timer_stop()
    e_wall = time.perf_counter() - t0
    env.barrier()
    e_ms_max = env.max(max(e_ms, 1e3 * e_wall))
    e_value = env.sum(float(e_its)) / (e_ms_max * 1e-3)

    # ---- roofline of the dominant kernel (fused single-pass AL-gradient, k_dense_algrad)
    peak, peak_src = env.peak, env.peak_src
    bgrad = algorithmic_bytes_per_grad(m, n)
    achieved = (bgrad * dom_launches / (dom_ms * 1e-3)) / 1e9 if dom_ms > 0 else None
    roofline = {"bound": "hbm", "kernel": "k_dense_algrad<true, false> (fused A-pass: r=Ax-b, w, A'w; A read once)",
                "achieved": achieved, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                "frac": (achieved / peak) if achieved else None, "traffic": traffic_from_profile(m, n),
                "traffic_source": "profiles/r2n_dense_traffic.json (ncu --set full capture of two consecutive launches of the same kernel and shape, their mean; not re-measured in this run)",
                "algorithmic_bytes_per_launch": bgrad, "launches_timed": dom_launches,
                "avg_launch_us": (1e3 * dom_ms / dom_launches) if dom_launches else None,
                "events_bracket_every_nth_launch": 31,
                "launches_sampled": "every 31st launch of the kernel: G(x) on all row clusters and G(z) on two clusters fewer beside the speculative L step alternate, so both are sampled alike",
                "kernel_share_of_step": ((dom_ms / dom_launches) * grads / ms) if (ms > 0 and dom_launches) else None}
    # iteration-level roofline: bytes = N_grad * B_grad + N_it * 8n(4 mem + 17)  (SURVEY §8d), this rank's step
    it_bytes = grads * bgrad + its * algorithmic_bytes_per_iteration_vectors(n)
    it_gbs = it_bytes / (ms * 1e-3) / 1e9
    roofline["iteration_level"] = {"achieved": it_gbs, "frac": it_gbs / peak, "unit": "GB/s",
                                   "bytes_model": "n_grad*(8mn+8(2n+5m)) + n_it*8n*37 (SURVEY.md §8d), measured counts"}

    # ---- CPU baseline (oracle = restatement of the reference), bounded sample, rank 0 at N=1 only
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cap = min(args.inner_cap, 600) if args.inner_cap > 0 else 600
        prob_o = bo.problem_basis_pursuit(A_host, b_host)
        t = time.perf_counter()
        r = bo.solve(prob_o, np.zeros(n), max_inner_total=cap)
        dt = time.perf_counter() - t
        cpu = {"value": r.tot_inner_it / dt, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
               "sample": f"first {r.tot_inner_it} PANOC+ inner iterations of the same alps solve ({dt:.1f} s), numpy/OpenBLAS "
                         "restatement of the reference (Julia unavailable on the box)"}

    cfg = headline_config(m, n, args.inner_cap, world)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": cfg,
        "parallelism": {"instances": world, "mode": "independent instances, one per GPU, no collectives (SURVEY.md §8e: C2 = replicas only)"},
        "inner_iterations": its_all, "al_gradients_per_iteration": grads / max(1, its),
        "outer_iterations_last_step": last.tot_it, "status_last_step": bz.STATUS_SYMBOLS[last.status],
        "e2e": {"value": e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "call": "bazinga_b200.alps(f, g, c, D, x0, y0; ...) with host buffers: problem allocation, H2D of A/b/x0/y0 from pinned "
                        "memory, solve, D2H of x/y/s/mu, free -- every step"},
        "gpu_launches": launches,
        "kernel_breakdown_us": {nm: (1e3 * kind_ms[k] / kind_n[k] if kind_n[k] else None) for k, nm in
                                enumerate(["al_gradient", "lbfgs_step", "fb_step", "decision_step"])},
        "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
    }
    return line, {"A_host": A_host, "b_host": b_host, "cpu": cpu}


# ------------------------------------------------------------------------------------------ legs
def leg_c5(env, args, shared):
    """C5: basis pursuit 32768 x 262144 (68.7 GB), row-sharded over the ranks, strong scaling."""
    import numpy as np
    import bazinga_oracle as bo
    bz, ctx, rank, world = env.bz, env.ctx, env.rank, env.world
    m, n = args.c5_rows, args.c5_cols
    env.comm_init()
    xtrue = bo.bz_sparse_randn(1, n, 0.1)

    def one(ctx_, sharded):
        if sharded:
            r0, ml = bz.row_shard(m, world, rank)
            c = bz.ConstraintBasisPursuit(None, np.zeros(ml), gen_seed=0, shape=(ml, n), row_begin=r0, ny_global=m)
        else:
            r0, ml = 0, m
            c = bz.ConstraintBasisPursuit(None, np.zeros(m), gen_seed=0, shape=(m, n))
        p = bz.Problem(bz.Zero(), bz.NormL1(1.0), c, bz.ZeroSet(), ctx=ctx_)
        # b = A xtrue from the library's own dense pass (c(x) with b = 0); sampled rows checked against the oracle's generator
        p.auglag_update(np.ones(ml), np.zeros(ml))
        b = p.gradient(xtrue)["cx"]
        for r in np.linspace(0, ml - 1, 4).astype(int):
            row = bo.bz_randn_like(0, (r0 + int(r)) * n, n)
            assert abs(b[r] - row @ xtrue) <= 1e-10 * (np.linalg.norm(row) * np.linalg.norm(xtrue)), "device b differs from the oracle"
        p.upload_dense(None, b)
        res = ms = x = None
        for cap in (8, args.c5_inner):
            o, _ = bp_options(bz, cap)
            ctx_.sync()
            if sharded:
                env.barrier()
            x, y, s, mu, res = p.alps(np.zeros(n), np.zeros(m), o)
            ms = res.device_ms          # the solve on the device, without the H2D of x0, y0 / D2H of the results
        p.close()
        return res, ms, x, ml

    sampler = ClockSampler(env.local_rank).start() if rank == 0 else None
    res, ms, x, ml = one(ctx, world > 1)
    ms = env.max(ms)
    clocks = sampler.stop() if rank == 0 else None
    bgrad = algorithmic_bytes_per_grad(ml, n)
    per_gpu_bytes = res.n_grad * bgrad + res.tot_inner_it * algorithmic_bytes_per_iteration_vectors(n)
    gbs = per_gpu_bytes / (ms * 1e-3) / 1e9
    g_us = 1e3 * res.kind_ms[0] / max(1, res.kind_launches[0])
    leg = {"workload": f"basis pursuit {m}x{n} fp64 (BASELINE configs[4]), A generated on the device ({8 * m * n / 1e9:.1f} GB), "
                       f"first {res.tot_inner_it} PANOC+ inner iterations of the alps solve",
           "scaling": "strong", "n_gpus": world, "rows_per_gpu": ml,
           "partitioning": "unsharded (one GPU holds all of A)" if world == 1 else
                           "rows of A, b, y, mu, cx, s split over the ranks; x-space state replicated",
           "collective": None if world == 1 else f"ncclAllReduce(sum) of {n + 8} doubles per AL-gradient on the library's communicator (bz_comm_init), "
                                                 "+ 2 scalar allreduces per outer iteration",
           "inner_iterations": res.tot_inner_it, "n_grad": res.n_grad, "ms": ms,
           "iterations_per_s": res.tot_inner_it / (ms * 1e-3),
           "iteration_GBs_per_gpu": gbs, "frac_of_hbm_peak_per_gpu": gbs / env.peak,
           "al_gradient_local_us": g_us, "al_gradient_local_GBs": bgrad / (g_us * 1e-6) / 1e9 if g_us else None,
           "al_gradient_local_frac": (bgrad / (g_us * 1e-6) / 1e9 / env.peak) if g_us else None,
           "kernel_breakdown_us": kernel_breakdown(res, ("al_gradient_local", "lbfgs_step", "fb_step", "decision_step")),
           "roofline_100pct_single_gpu_it_s": env.peak * 1e9 / (2 * algorithmic_bytes_per_grad(m, n) + algorithmic_bytes_per_iteration_vectors(n)),
           "clocks": clocks}
    if world > 1:
        leg["allreduce_us"] = ctx.comm_bench_allreduce(n + 8, 50)
        xs = env.torch.from_numpy(x).cuda()
        xmax, xmin = xs.clone(), xs.clone()
        env.dist.all_reduce(xmax, op=env.dist.ReduceOp.MAX)
        env.dist.all_reduce(xmin, op=env.dist.ReduceOp.MIN)
        leg["x_bit_identical_across_ranks"] = bool(env.torch.equal(xmax, xmin))
        leg["parity_check"] = c5_parity_check(env)
        if not args.no_leg_baselines:
            if rank == 0:
                r1, ms1, _, _ = one(ctx, False)
                leg["single_gpu"] = {"iterations_per_s": r1.tot_inner_it / (ms1 * 1e-3), "ms": ms1, "inner_iterations": r1.tot_inner_it,
                                     "note": "rank 0 alone, the whole 68.7 GB instance on one GPU, same run"}
            env.barrier()
    else:
        cpu = shared.get("cpu")
        if cpu:   # the 68.7 GB instance does not fit host RAM: scale the C2 CPU measurement by bytes, and say so
            scale = (2 * algorithmic_bytes_per_grad(M_ROWS, N_COLS)) / (2 * algorithmic_bytes_per_grad(m, n))
            leg["cpu_baseline"] = {"value": cpu["value"] * scale, "unit": UNIT, "cores": cpu["cores"], "kind": "port",
                                   "sample": "EXTRAPOLATED: the 4096x16384 CPU measurement of this run scaled by matrix bytes "
                                             f"(x{scale:.5f}); a 68.7 GB A does not fit the host's RAM"}
    return leg


def c5_parity_check(env, m=512, n=2048, inner=12):
    """Row-sharded path against the unsharded kernels and the numpy oracle on a small instance: one AL-gradient and the
    first `inner` PANOC+ iterates (before rounding differences of the two summation orders can split the trajectories)."""
    import numpy as np
    import bazinga_oracle as bo
    bz, ctx, rank, world = env.bz, env.ctx, env.rank, env.world
    A = ctx.randn_fill(7, 0, m * n).reshape(m, n)
    xtrue = bo.bz_sparse_randn(8, n, 0.1)
    b = A @ xtrue
    r0, ml = bz.row_shard(m, world, rank)
    rng = np.random.default_rng(3)
    mu, y, xp = rng.uniform(0.1, 2.0, m), rng.standard_normal(m), rng.standard_normal(n)
    c = bz.ConstraintBasisPursuit(A[r0:r0 + ml], b[r0:r0 + ml], row_begin=r0, ny_global=m)
    p = bz.Problem(bz.Zero(), bz.NormL1(1.0), c, bz.ZeroSet(), ctx=ctx)
    p.auglag_update(mu[r0:r0 + ml], y[r0:r0 + ml])
    g = p.gradient(xp)
    o, sub = bp_options(bz, inner)
    xs, ys, ss, mus, res = p.alps(np.zeros(n), np.zeros(m), o)
    p.close()
    parts = env.gather({"cx": g["cx"], "yupd": g["yupd"], "lx": g["lx"], "dlx": g["dlx"], "y": ys})
    out = None
    if rank == 0:
        al = bo.AugLagFun(bo.FZero(), bo.CDenseAffine(A, b), bo.ZeroSet(), mu.copy(), y.copy(), xp)
        dl = np.empty(n)
        lx = al.gradient(dl, xp)
        scale = np.linalg.norm(A, 2) * np.linalg.norm(al.yupd)
        ctx1 = bz.Context(env.local_rank)
        out1, r1 = bz.alps(bz.Zero(), bz.NormL1(1.0), bz.ConstraintBasisPursuit(A, b), bz.ZeroSet(), np.zeros(n), np.zeros(m),
                           subsolver=sub, subsolver_maxit=100_000, max_inner_total=inner, ctx=ctx1, return_result=True)
        ro = bo.solve(bo.problem_basis_pursuit(A, b), np.zeros(n), max_inner_total=inner)
        ctx1.close()
        out = {"shape": [m, n], "inner_iterations": int(res.tot_inner_it),
               "gradient_err_vs_oracle_normwise": float(np.max(np.abs(parts[0]["dlx"] - dl)) / scale),
               "lx_rel_err_vs_oracle": float(abs(parts[0]["lx"] - lx) / abs(lx)),
               "cx_err_vs_oracle": float(np.max(np.abs(np.concatenate([q["cx"] for q in parts]) - al.cx)) / np.max(np.abs(al.cx))),
               "x_max_abs_diff_vs_unsharded_gpu": float(np.max(np.abs(xs - out1[0]))),
               "x_max_abs_diff_vs_oracle": float(np.max(np.abs(xs - ro.x))),
               "y_max_abs_diff_vs_oracle": float(np.max(np.abs(np.concatenate([q["y"] for q in parts]) - ro.y))),
               "x_scale": float(np.max(np.abs(ro.x))),
               "same_inner_count_as_unsharded": bool(res.tot_inner_it == r1.tot_inner_it == ro.tot_inner_it)}
    return out


def _grid(lo, hi, k):
    import numpy as np
    v = np.linspace(lo, hi, k)
    return np.array([(i, j) for j in v for i in v])


def leg_c3(env, args, shared):
    """C3: 65,536 independent either-or + MPVC starting points, batched across GPUs, no collectives."""
    import numpy as np
    bz, ctx, rank, world = env.bz, env.ctx, env.rank, env.world
    import bazinga_oracle as bo
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from tiny_oracle_binding import tiny_batch
    side = args.c3_side
    leg = {"instances": side * side, "partitioning": "instance i -> GPU i mod N; no data-path collective", "problems": {}}
    sampler = ClockSampler(env.local_rank).start() if rank == 0 else None
    for name, kind, lo, hi in (("eitheror", 1, -4.0, 8.0), ("mpvca", 3, -5.0, 20.0)):
        X0 = _grid(lo, hi, side)                                # the demo box, column-major
        bz.alps_batch(name, X0[:1024], ctx=ctx)                 # warm-up (module load)
        entry = {}
        keep = {}
        for mode, mine in (("strong", X0[rank::world]), ("weak", X0)):
            best, out = None, None
            for _ in range(3):
                env.barrier()
                out = bz.alps_batch(name, mine, ctx=ctx)
                t = env.max(out["kernel_ms"])
                best = t if best is None else min(best, t)
            tot = [env.sum(float(len(mine))), env.sum(float(out["sub_iters"].sum())), env.sum(float((out["status"] == 0).sum()))]
            entry[mode] = {"n_gpus": world, "instances_total": int(tot[0]), "solved": int(tot[2]), "inner_iterations_total": int(tot[1]),
                           "kernel_ms_max_over_ranks": best, "solves_per_s": tot[0] / (best * 1e-3),
                           "inner_iterations_per_s": tot[1] / (best * 1e-3)}
            keep[mode] = (mine, out, out["kernel_ms"])
        # strong efficiency measured in this run: the weak launch of rank 0 IS the whole 65,536-instance batch on one GPU
        t1 = keep["weak"][2]
        t1 = env.max(t1 if rank == 0 else 0.0)
        entry["strong_efficiency"] = (t1 / entry["strong"]["kernel_ms_max_over_ranks"]) / world
        entry["single_gpu_ms_all_instances"] = t1
        if rank == 0:      # bit-identity of this rank's strong share against the C oracle (OpenMP, all host cores)
            mine, out, _ = keep["strong"]
            t = time.perf_counter()
            ref = tiny_batch(kind, mine, min_gamma=bo.EPS, sub_maxit=1000, subsolver_maxit=1000)
            dt = time.perf_counter() - t
            entry["bit_identical_vs_c_oracle"] = bool(np.array_equal(out["x"], ref["x"], equal_nan=True) and
                                                      np.array_equal(out["y"], ref["y"], equal_nan=True) and
                                                      np.array_equal(out["sub_iters"], ref["sub_iters"]) and
                                                      np.array_equal(out["iters"], ref["iters"]))
            entry["bit_identical_checked_instances"] = int(len(mine))
            if world == 1:
                entry["cpu_baseline"] = {"value": len(mine) / dt, "unit": "solves/s", "cores": os.cpu_count() or 1, "kind": "port",
                                         "sample": f"all {len(mine)} instances, oracle/tiny_oracle.c with OpenMP ({dt:.2f} s)"}
        if rank == 0:      # the residual bound of a launch: its longest dependent chain = the longest-iterating instance run ALONE
            mine, out, _ = keep["strong"]
            worst = int(np.argmax(out["sub_iters"]))
            bz.alps_batch(name, mine[worst:worst + 1], ctx=ctx)
            alone = min(bz.alps_batch(name, mine[worst:worst + 1], ctx=ctx)["kernel_ms"] for _ in range(2))
            entry["longest_instance_alone"] = {"inner_iterations": int(out["sub_iters"][worst]), "kernel_ms": alone,
                                               "note": "one instance, one launch: the dependent chain no partitioning over GPUs can shorten"}
        leg["problems"][name] = entry
    leg["clocks"] = sampler.stop() if rank == 0 else None
    # secondary roofline: ~250 fp64 flop per inner iteration against the measured fp64 FMA issue peak
    peak_tf = ctx.fp64_fma_peak()
    leg["fp64_fma_peak_tflops_measured"] = peak_tf
    for name, e in leg["problems"].items():
        for mode in ("strong", "weak"):
            fl = 250.0 * e[mode]["inner_iterations_per_s"] / 1e12
            e[mode]["fp64_tflops_model"] = fl
            e[mode]["frac_of_fp64_peak"] = fl / (peak_tf * world)
    leg["bound"] = ("not HBM (9 MB of I/O): dependent-issue latency of the slowest instances (a few warps are live at the end of a "
                    "launch); flop model 250 fp64 flop per inner iteration (SURVEY.md §8d)")
    return leg


def leg_c4(env, args, shared):
    """C4: obstacle problem on a 4096 x 4096 grid (N = 16.8 M, nx = 50.3 M); slab-decomposed over the ranks."""
    import numpy as np
    bz, ctx, rank, world = env.bz, env.ctx, env.rank, env.world
    S = args.c4_side
    if S % world:
        return {"skipped": f"grid side {S} does not split over {world} ranks"}

    def one(ctx_, S_, nranks, cap, tol=1e-4, warm=10):
        gy, N = S_ // nranks, S_ * S_
        Nl = gy * S_
        r = rank if nranks > 1 else 0
        p = bz.Problem(bz.SmoothCostObstacleL1(Nl), bz.NonsmoothCostObstacleL1(Nl),
                       bz.ConstraintObstacle(Nl, (gy, S_), slab=(r, nranks) if nranks > 1 else None), bz.ZeroSet(), ctx=ctx_)
        if nranks > 1:
            p.xr_attach(env.gather(p.xr_handle()))
            env.barrier()
        from bazinga_b200.solver import _options
        sub = bz.PANOCplus(directions=bz.LBFGS(5), maxit=1_000_000, minimum_gamma=EPS)
        x0, y0 = 2.0 * np.ones(3 * Nl), np.zeros(Nl)
        res = ms = x = None
        for c in ((min(warm, cap), cap) if warm else (cap,)):
            o = _options(tol=tol, tol_prim=None, tol_dual=None, inner_tol=None, maxit=100, theta_penalty=0.8, kappa_penalty=0.5,
                         kappa_tol=0.1, subsolver_maxit=1_000_000, sub=sub, max_inner_total=c)
            ctx_.sync()
            if nranks > 1:
                env.barrier()
            ctx_.timer_start()
            x, y, s, mu, res = p.alps(x0, y0, o)
            ms_e2e = ctx_.timer_stop()
            ms = res.device_ms          # the solve on the device; ms_e2e adds the H2D of x0, y0 and the D2H of x, y, s, mu (1.3 GB at 4096^2)
        p.close()
        res.ms_e2e = ms_e2e
        return res, ms, x

    N = S * S
    sampler = ClockSampler(env.local_rank).start() if rank == 0 else None
    res, ms, x = one(ctx, S, world, args.c4_inner)
    ms = env.max(ms)
    clocks = sampler.stop() if rank == 0 else None
    total_bytes = res.n_grad * 64 * N + res.tot_inner_it * 8 * 3 * N * 37          # SURVEY §8(d) C4, measured counts
    gbs = total_bytes / (ms * 1e-3) / 1e9
    leg = {"workload": f"obstacle_l1 on a {S}x{S} grid (N={N}, nx={3 * N}), x0 = 2, first {res.tot_inner_it} PANOC+ inner iterations (BASELINE configs[3])",
           "scaling": "strong", "n_gpus": world, "grid_rows_per_gpu": S // world,
           "partitioning": "one GPU" if world == 1 else "grid rows split over the ranks (all of x, y, mu sharded); boundary rows and every "
                                                        "reduction scalar exchanged over cudaIpc peer memory from inside the kernels",
           "inner_iterations": res.tot_inner_it, "outer_iterations": res.tot_it, "n_grad": res.n_grad,
           "al_gradients_per_iteration": res.n_grad / max(1, res.tot_inner_it), "ms": ms, "ms_with_host_transfers": res.ms_e2e,
           "iterations_per_s": res.tot_inner_it / (ms * 1e-3), "objective": res.objective, "norm_res_prim": res.norm_res_prim,
           "algorithmic_GBs_total": gbs, "frac_of_hbm_peak_per_gpu": gbs / world / env.peak,
           "bytes_model": "n_grad*64N + n_it*8*3N*37 (SURVEY.md §8d), measured counts",
           "roofline_100pct_single_gpu_it_s_at_measured_grads": env.peak * 1e9 / (total_bytes / max(1, res.tot_inner_it)),
           "kernel_breakdown_us": kernel_breakdown(res), "gpu_launches": res.n_launches, "clocks": clocks}
    if world > 1:
        same = env.gather((int(res.tot_inner_it), float(res.objective)))
        leg["same_counts_and_objective_on_all_ranks"] = len(set(same)) == 1
        if not args.no_leg_baselines:
            if rank == 0:
                r1, ms1, _ = one(ctx, S, 1, args.c4_inner)
                leg["single_gpu"] = {"iterations_per_s": r1.tot_inner_it / (ms1 * 1e-3), "ms": ms1, "n_grad": r1.n_grad, "objective": r1.objective,
                                     "note": "rank 0 alone, the un-decomposed grid, same run"}
            env.barrier()
        # convergence of the decomposed solve at a size that converges in seconds: same solution as the un-decomposed one
        cs = args.c4_check_side
        if cs and cs % world == 0:
            # (i) the first 12 iterations, before rounding differences of the reductions can move a decision: the decomposed
            # iterate must be the un-decomposed one to rounding
            rp, _, xp = one(ctx, cs, world, 12, tol=1e-4, warm=0)
            Nl = (cs // world) * cs
            xps = env.gather(xp)
            if rank == 0:
                r1p, _, x1p = one(ctx, cs, 1, 12, tol=1e-4, warm=0)
                xpg = np.concatenate([np.concatenate([q[b * Nl:(b + 1) * Nl] for q in xps]) for b in range(3)])
                leg["prefix_check"] = {"grid": [cs, cs], "inner_iterations": int(rp.tot_inner_it), "n_grad_slab": int(rp.n_grad),
                                       "n_grad_single": int(r1p.n_grad), "x_max_abs_diff": float(np.max(np.abs(xpg - x1p))),
                                       "x_max_abs": float(np.max(np.abs(x1p))),
                                       "objective_rel_diff": abs(rp.objective - r1p.objective) / max(1e-300, abs(r1p.objective))}
            env.barrier()
            # (ii) the whole solve
            rc, _, xc = one(ctx, cs, world, args.c4_check_cap, tol=1e-4, warm=0)
            xs = env.gather(xc)
            chk = None
            if rank == 0:
                r1c, _, x1 = one(ctx, cs, 1, args.c4_check_cap, tol=1e-4, warm=0)
                xg = np.concatenate([np.concatenate([q[b * Nl:(b + 1) * Nl] for q in xs]) for b in range(3)])
                chk = {"grid": [cs, cs], "tol": 1e-4, "status_slab": bz.STATUS_SYMBOLS[rc.status], "status_single": bz.STATUS_SYMBOLS[r1c.status],
                       "inner_iterations_slab": int(rc.tot_inner_it), "inner_iterations_single": int(r1c.tot_inner_it),
                       "objective_slab": rc.objective, "objective_single": r1c.objective,
                       "objective_rel_diff": abs(rc.objective - r1c.objective) / max(1e-300, abs(r1c.objective)),
                       "norm_res_prim_slab": rc.norm_res_prim, "norm_res_prim_single": r1c.norm_res_prim,
                       "x_max_abs_diff": float(np.max(np.abs(xg - x1))), "x_max_abs": float(np.max(np.abs(x1)))}
            env.barrier()
            leg["convergence_check"] = chk
    elif rank == 0 and not args.no_cpu_baseline:
        import bazinga_oracle as bo
        t0 = time.perf_counter()
        r = bo.solve(bo.problem_obstacle(N, (S, S)), 2.0 * np.ones(3 * N), tol=1e-4, max_inner_total=args.c4_cpu_its)
        dt = time.perf_counter() - t0
        leg["cpu_baseline"] = {"value": r.tot_inner_it / dt, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
                               "sample": f"first {r.tot_inner_it} PANOC+ inner iterations ({r.n_grad} AL-gradients, {dt:.1f} s), numpy "
                                         "restatement of the reference on the same grid"}
    return leg


LEGS = {"c5_row_sharded": leg_c5, "c3_batched": leg_c3, "c4_slab": leg_c4}


def run_b200(args):
    env = Env(args)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    line, shared = run_headline(env, args)
    shared.pop("A_host", None)
    shared.pop("b_host", None)
    wanted = [w for w in args.legs.split(",") if w] if args.legs != "none" else []
    legs, printed = {}, threading.Event()

    def emit():
        if env.rank == 0 and not printed.is_set():
            printed.set()
            line["legs"] = legs
            eff = {}
            for nm, lg in legs.items():
                if isinstance(lg, dict) and "single_gpu" in lg and lg.get("iterations_per_s"):
                    eff[nm] = lg["iterations_per_s"] / lg["single_gpu"]["iterations_per_s"] / env.world
                if nm == "c3_batched" and isinstance(lg, dict) and "problems" in lg:
                    for pn, e in lg["problems"].items():
                        if "strong_efficiency" in e:
                            eff[f"c3_batched.{pn}"] = e["strong_efficiency"]
            line["strong_efficiency"] = eff if env.world > 1 else None
            print(json.dumps(line), flush=True)

    # a leg that hangs (a peer that never answers) must not cost the headline: past the deadline rank 0 prints what it has
    def watchdog(deadline_s, name):
        def run():
            # rank 0 prints first; the other ranks leave a little later so that torchrun does not tear rank 0 down mid-print
            if not done.wait(deadline_s if env.rank == 0 else deadline_s + 15):
                legs[name] = {"error": f"leg exceeded its {deadline_s} s deadline"}
                emit()
                os._exit(0)
        done = threading.Event()
        th = threading.Thread(target=run, daemon=True)
        th.start()
        return done

    for name in wanted:
        if name not in LEGS:
            continue
        done = watchdog(args.leg_deadline, name)
        try:
            legs[name] = LEGS[name](env, args, shared)
        except Exception as e:      # noqa: BLE001 -- a failing leg is reported, the headline stands
            legs[name] = {"error": f"{type(e).__name__}: {e}", "traceback": traceback.format_exc(limit=4)}
        done.set()
        try:
            env.barrier()
        except Exception:
            pass
    emit()
    if env.dist:
        env.dist.barrier()
        env.dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--rows", type=int, default=M_ROWS)
    ap.add_argument("--cols", type=int, default=N_COLS)
    ap.add_argument("--inner-cap", type=int, default=1000, help="inner iterations per step, both arms (0 = full solve)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--legs", default="c5_row_sharded,c3_batched,c4_slab", help="comma list, or 'none'")
    ap.add_argument("--no-leg-baselines", action="store_true", help="N > 1: skip rank 0's single-GPU run of the strong legs")
    ap.add_argument("--leg-deadline", type=float, default=420.0)
    ap.add_argument("--c5-rows", type=int, default=32768)
    ap.add_argument("--c5-cols", type=int, default=262144)
    ap.add_argument("--c5-inner", type=int, default=60)
    ap.add_argument("--c3-side", type=int, default=256)
    ap.add_argument("--c4-side", type=int, default=4096)
    ap.add_argument("--c4-inner", type=int, default=200)
    ap.add_argument("--c4-cpu-its", type=int, default=3)
    ap.add_argument("--c4-check-side", type=int, default=64)
    ap.add_argument("--c4-check-cap", type=int, default=200_000)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    use_all_host_threads()      # the CPU checkers (C oracle, numpy) of rank 0 use every host core even under torchrun
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
