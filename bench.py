#!/usr/bin/env python
"""bench.py -- headline benchmark of the PANOC+/ALPS hot path (BASELINE.json metric).

Workload at every N (configs[1] of BASELINE.json, SURVEY.md §8d "C2"): demo/basispursuit.jl scaled to a
dense A 4096x16384 fp64 (synthetic, counter-based generator shared bit-for-bit with the oracle),
f = Zero, g = NormL1(1), c(x) = A x - b, D = ZeroSet, x0 = y0 = 0, LBFGS(5), minimum_gamma = eps.
A "step" is one `alps` call on that instance, capped at --inner-cap PANOC+ inner iterations in total
(the full solve is thousands of iterations; `--inner-cap 0` runs it to convergence).  The metric is
PANOC+ forward-backward inner iterations per second.

  value   whole-job inner iterations / s with A, b already resident in HBM (device-timed, CUDA events
          on the library's stream, max over ranks)
  e2e     the same steps through the public API `bazinga_b200.alps`-level call with HOST buffers:
          every step re-uploads A and b from pinned host memory and downloads x, y, s, mu
  N > 1   independent instances, one per GPU, no collectives ("scaling": "weak")

`--impl reference` times the CPU restatement of the reference (oracle/bazinga_oracle.py: numpy +
OpenBLAS, all host threads; the Julia reference itself cannot run here -- no Julia in the image) on a
bounded sample of the same workload.
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

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

M_ROWS, N_COLS = 4096, 16384
METRIC = "panoc_fwd_bwd_inner_iterations_per_s"
UNIT = "iterations/s"
EPS = 2.220446049250313e-16


def algorithmic_bytes_per_grad(m, n):
    """A read once + vector traffic of one fused AL-gradient evaluation."""
    return 8 * m * n + 8 * (2 * n + 5 * m)


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
            return
        self.thr = threading.Thread(target=self._read, daemon=True)
        self.thr.start()

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
    `ncu --set full` capture (profiles/); null for other shapes."""
    p = os.path.join(ROOT, "profiles", "r1_dense_traffic.json")
    try:
        d = json.load(open(p))
        return d["traffic_bytes_per_launch"] if d["shape"] == [m, n] else None
    except Exception:
        return None


def dist_env():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


# ------------------------------------------------------------------------------------------ reference arm
def run_reference(args):
    rank, _, world = dist_env()
    if rank != 0:
        return 0
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import bazinga_oracle as bo
    cores = os.cpu_count() or 1
    m, n = args.rows, args.cols
    t0 = time.time()
    A = bo.bz_randn_like(0, 0, m * n).reshape(m, n) if m * n <= 1 << 22 else _gen_host_matrix(bo, m, n)
    xtrue = bo.bz_sparse_randn(1, n, 0.1)
    b = A @ xtrue
    gen_s = time.time() - t0
    prob = bo.problem_basis_pursuit(A, b)
    cap = args.ref_inner_cap
    its, times = [], []
    for step in range(args.warmup + args.steps):
        t = time.perf_counter()
        r = bo.solve(prob, np.zeros(n), max_inner_total=cap)
        dt = time.perf_counter() - t
        if step >= args.warmup:
            its.append(r.tot_inner_it)
            times.append(dt)
    value = sum(its) / sum(times)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"basis pursuit {m}x{n} fp64 (BASELINE configs[1]), alps capped at {cap} inner iterations per step",
                   "rows": m, "cols": n, "inner_cap": cap},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{args.steps} x first {cap} PANOC+ inner iterations of the alps solve; numpy/OpenBLAS "
                                   f"restatement of the reference (Julia unavailable); data gen {gen_s:.1f}s excluded"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)
    return 0


def _gen_host_matrix(bo, m, n):
    """C form of the oracle's generator (oracle/tiny_oracle.c: oracle_randn_fill), same bits as
    bo.bz_randn_like but fast enough for 67M entries."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from tiny_oracle_binding import oracle_randn_fill
    return oracle_randn_fill(0, 0, m * n).reshape(m, n)


# ------------------------------------------------------------------------------------------ B200 arm
def run_b200(args):
    rank, local_rank, world = dist_env()
    import numpy as np
    import torch
    import bazinga_b200 as bz

    use_dist = world > 1
    if use_dist:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = bz.Context(local_rank)
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
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import bazinga_oracle as bo
    xtrue = bo.bz_sparse_randn(seed + 1, n, 0.1)
    b_t = torch.empty(m, dtype=torch.float64, pin_memory=True)
    b_host = b_t.numpy()
    b_host[:] = A_host @ xtrue

    sub = bz.PANOCplus(directions=bz.LBFGS(5), maxit=100_000, minimum_gamma=EPS)
    prob = bz.Problem(bz.Zero(), bz.NormL1(1.0), bz.ConstraintBasisPursuit(A_host, b_host), bz.ZeroSet(), ctx=ctx)
    from bazinga_b200.solver import _options
    opts = _options(tol=1e-6, tol_prim=None, tol_dual=None, inner_tol=None, maxit=100, theta_penalty=0.8,
                    kappa_penalty=0.5, kappa_tol=0.1, subsolver_maxit=100_000, sub=sub,
                    max_inner_total=args.inner_cap)
    x0, y0 = np.zeros(n), np.zeros(m)

    def barrier():
        ctx.sync()
        if use_dist:
            dist.barrier()
        ctx.sync()

    def reduce_max(v):
        if not use_dist:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def reduce_sum(v):
        if not use_dist:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---- resident leg (value): warm-up, then exactly K timed steps
    ctx.set_profiling(False)
    for _ in range(args.warmup):
        prob.alps(x0, y0, opts)
    ctx.set_profiling(True)
    sampler = ClockSampler(local_rank)
    barrier()
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
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ctx.set_profiling(False)
    ms_max = reduce_max(ms)
    its_all = reduce_sum(float(its))
    value = its_all / (ms_max * 1e-3)

    # ---- e2e leg: host buffers in, host buffers out, every step
    h2d = 8 * (m * n + m + n + m)
    d2h = 8 * (n + 3 * m)
    barrier()
    t0 = time.perf_counter()
    ctx.timer_start()
    e_its = 0
    for _ in range(args.steps):
        prob.upload_dense(A_host, b_host)
        x, y, s, mu, res = prob.alps(x0, y0, opts)
        e_its += res.tot_inner_it
    e_ms = ctx.timer_stop()
    barrier()
    e_wall = time.perf_counter() - t0
    e_ms_max = reduce_max(max(e_ms, 1e3 * e_wall))
    e_value = reduce_sum(float(e_its)) / (e_ms_max * 1e-3)

    # ---- roofline of the dominant kernel (fused single-pass AL-gradient, k_dense_algrad)
    peak, peak_src = measured_peak_hbm()
    bgrad = algorithmic_bytes_per_grad(m, n)
    achieved = (bgrad * dom_launches / (dom_ms * 1e-3)) / 1e9 if dom_ms > 0 else None
    roofline = {"bound": "hbm", "kernel": "k_dense_algrad<true, false> (fused A-pass: r=Ax-b, w, A'w; A read once)",
                "achieved": achieved, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                "frac": (achieved / peak) if achieved else None, "traffic": traffic_from_profile(m, n),
                "algorithmic_bytes_per_launch": bgrad, "launches_timed": dom_launches,
                "avg_launch_us": (1e3 * dom_ms / dom_launches) if dom_launches else None,
                "events_bracket_every_nth_launch": 8,
                "kernel_share_of_step": ((dom_ms / dom_launches) * grads / ms) if (ms > 0 and dom_launches) else None}

    # ---- CPU baseline (oracle = restatement of the reference), bounded sample, rank 0 at N=1 only
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cap = args.ref_inner_cap
        prob_o = bo.problem_basis_pursuit(A_host, b_host)
        t = time.perf_counter()
        r = bo.solve(prob_o, np.zeros(n), max_inner_total=cap)
        dt = time.perf_counter() - t
        cpu = {"value": r.tot_inner_it / dt, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
               "sample": f"first {r.tot_inner_it} PANOC+ inner iterations of the same alps solve ({dt:.1f} s), numpy/OpenBLAS "
                         "restatement of the reference (Julia unavailable on the box)"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"basis pursuit {m}x{n} fp64 (BASELINE configs[1]), alps capped at {args.inner_cap} inner iterations per step"
                       if args.inner_cap else f"basis pursuit {m}x{n} fp64 (BASELINE configs[1]), full alps solve",
                       "rows": m, "cols": n, "inner_cap": args.inner_cap, "lbfgs_memory": 5,
                       "instances": world, "parallelism": "independent instances, no collectives",
                       "l2_policy": f"A is {8 * m * n / 1e6:.0f} MB > 126 MB L2 and is streamed twice per inner iteration; no explicit flush"},
            "inner_iterations": its_all, "al_gradients_per_iteration": grads / max(1, its),
            "outer_iterations_last_step": last.tot_it, "status_last_step": bz.STATUS_SYMBOLS[last.status],
            "e2e": {"value": e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches,
            "kernel_breakdown_us": {nm: (1e3 * kind_ms[k] / kind_n[k] if kind_n[k] else None) for k, nm in
                                    enumerate(["al_gradient", "lbfgs_step", "fb_step", "decision_step"])},
            "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    prob.close()
    if use_dist:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--rows", type=int, default=M_ROWS)
    ap.add_argument("--cols", type=int, default=N_COLS)
    ap.add_argument("--inner-cap", type=int, default=2000, help="inner iterations per step (0 = full solve)")
    ap.add_argument("--ref-inner-cap", type=int, default=600, help="inner iterations per CPU-baseline step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
