#!/usr/bin/env python
"""Text summary of an `ncu --set full` capture: one block per launch with the numbers the roofline discussion in
DESIGN.md cites (duration, DRAM bytes, throughput percentages, IPC, occupancy, the top warp-stall reasons).

    python scripts/ncu_summary.py gpurun_out/r2i_dense.ncu-rep [--alg-bytes 537296896] > profiles/r2i_dense_summary.txt

Reads the report with `ncu -i <rep> --page raw --csv` (no GPU needed)."""
import argparse
import csv
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_active", "L1/TEX throughput %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__inst_executed.avg.per_cycle_elapsed", "IPC (elapsed)"),
    ("sm__inst_executed.avg.per_cycle_active", "IPC (active)"),
    ("smsp__issue_active.avg.pct", "issue slots busy %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__occupancy_limit_registers", "occupancy limit (regs), CTAs/SM"),
    ("launch__occupancy_limit_shared_mem", "occupancy limit (smem), CTAs/SM"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64 pipe %"),
    ("sm__inst_executed_pipe_fp64.sum", "fp64 warp instructions"),
    ("sm__inst_executed_pipe_tensor_op_dmma.sum", "DMMA warp instructions"),
    ("sm__pipe_tensor_op_dmma_cycles_active.avg.pct_of_peak_sustained_active", "DMMA pipe %"),
    ("smsp__inst_executed_op_local_ld.sum", "local loads (spills)"),
    ("smsp__inst_executed_op_local_st.sum", "local stores (spills)"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
]
STALL_PREFIX = "smsp__average_warps_issue_stalled_"
STALL_SUFFIX = "_per_issue_active.ratio"
# newer ncu names
STALL_PREFIX2 = "smsp__average_warp_latency_issue_stalled_"


def to_float(s):
    try:
        return float(s.replace(",", ""))
    except ValueError:
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--alg-bytes", type=float, default=None, help="algorithmic bytes per launch (DESIGN.md), for the ratio")
    args = ap.parse_args()
    out = subprocess.run(["ncu", "-i", args.report, "--page", "raw", "--csv"], capture_output=True, text=True)
    if out.returncode != 0:
        sys.exit(out.stderr)
    rows = list(csv.reader(out.stdout.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print(f"# {args.report}: {len(rows) - 2} launch(es); ncu --set full --clock-control none; read with --page raw --csv")
    for r in rows[2:]:
        print()
        print(f"kernel: {r[idx['Kernel Name']]}")
        print(f"  grid {r[idx['Grid Size']]} block {r[idx['Block Size']]}")
        tot = 0.0
        dur_us = None
        for key, label in KEYS:
            if key not in idx:
                continue
            val, unit = r[idx[key]], units[idx[key]]
            print(f"  {label:36s} {val} {unit}   [{key}]")
            f = to_float(val)
            if f is None:
                continue
            if key.startswith("dram__bytes"):
                mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
                tot += f * mult
            if key == "gpu__time_duration.sum":
                dur_us = f * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(unit, 1)
        if tot:
            line = f"  dram traffic per launch              {tot:.0f} B"
            if dur_us:
                line += f"  = {tot / dur_us / 1e3:.1f} GB/s over the profiled duration (cold, serialised: not a bench number)"
            print(line)
            if args.alg_bytes:
                print(f"  algorithmic bytes per launch         {args.alg_bytes:.0f} B   traffic/algorithmic = {tot / args.alg_bytes:.3f}")
        stalls = []
        for h, i in idx.items():
            if (h.startswith(STALL_PREFIX) or h.startswith(STALL_PREFIX2)) and h.endswith(STALL_SUFFIX) and "not_issued" not in h:
                f = to_float(r[i])
                if f is not None:
                    name = h.replace(STALL_PREFIX, "").replace(STALL_PREFIX2, "").replace(STALL_SUFFIX, "")
                    stalls.append((f, name))
        stalls.sort(reverse=True)
        if stalls:
            print("  warp stalls (cycles per issued instruction), top 6: " + ", ".join(f"{n} {v:.2f}" for v, n in stalls[:6]))


if __name__ == "__main__":
    main()
