"""Summarise an ncu report: headline metrics (raw page) and the hottest SASS lines with their stall reasons
(source page).  Usage: python scripts/ncu_top.py <file.ncu-rep> [N]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, u, v = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum", "sm__cycles_elapsed.max"]
for n in want:
    if n in h: print(f"{n:70s} {v[h.index(n)]} {u[h.index(n)]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]; data = rows[2:]
ix = {n: i for i, n in enumerate(h)}
stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
tot = sum(int(r[ix["# Samples"]]) for r in data)
agg = sorted(((sum(int(r[ix[s]]) for r in data), s) for s in stalls), reverse=True)[:8]
print("total samples", tot, [(s, c) for c, s in agg])
top = sorted(range(len(data)), key=lambda i: -int(data[i][ix["# Samples"]]))[:N]
for i in sorted(top):
    r = data[i]
    t2 = sorted(((int(r[ix[s]]), s[6:]) for s in stalls), reverse=True)[:2]
    print(f"{i:5d} {r[ix['Source']].strip()[:64]:64s} smp={r[ix['# Samples']]:>6s} exec={r[ix['Instructions Executed']]:>9s} {t2}")
