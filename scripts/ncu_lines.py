"""Aggregate an ncu report's source page by CUDA-C line: executed instructions and stall samples per line.
Usage: python scripts/ncu_lines.py <file.ncu-rep> [N] [source-file]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
h = rows[hi]
iS, iE = h.index("# Samples"), h.index("Instructions Executed")
agg = {}; cur = None; text = {}; fname = ""
for r in rows[hi + 1:]:
    if r and r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if len(r) < len(h) or r[0] == "Line No": continue
    if r[0] != "":
        cur = (fname, int(r[0])); text[cur] = r[1]; agg.setdefault(cur, [0, 0, 0]); continue
    a = agg.setdefault(cur, [0, 0, 0])
    if not r[iS].isdigit(): continue
    a[0] += int(r[iS]); a[1] += int(r[iE]); a[2] += 1
tot_s = sum(a[0] for a in agg.values()); tot_e = sum(a[1] for a in agg.values())
print("total samples", tot_s, "total warp instructions", tot_e)
for ln, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:N]:
    print(f"{ln[0][:14]:14s}:{ln[1]:4d} smp={a[0]:7d} ({100*a[0]/max(tot_s,1):4.1f}%) exec={a[1]:10d} ({100*a[1]/max(tot_e,1):4.1f}%) sass={a[2]:4d} | {text[ln].strip()[:110]}")
