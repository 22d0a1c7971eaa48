"""On-disk result format of the batched demos (demo/eitheror.jl:256-273,363-383; demo/mpvca.jl:262-290): the per-start
CSV `<problem>_<solver>_<subsolver>.csv` and the one-row `_stats.csv` the reference writes next to its plots.
Host-side formatting only -- the solves come from alps_batch / als_batch."""
import csv

import numpy as np

DATA_COLUMNS = ("id", "xinit_1", "xinit_2", "xfinal_1", "xfinal_2", "iters", "sub_iters", "runtime", "is_solved")
STATS_COLUMNS = ("npoints", "nsolved", "iters_max", "iters_median", "subiters_max", "subiters_median", "runtime_max",
                 "runtime_median", "global_nabs", "global_nrel", "local_nabs", "local_nrel", "unkwn_nabs", "unkwn_nrel")
# the minimisers the demos classify the end points by (tolx = 1e-6): eitheror.jl:286-288, mpvca.jl:296-298
MINIMISERS = {"eitheror": ((2.0, -2.0), (4.0, 4.0)), "xor": ((2.0, -2.0), (4.0, 4.0)),
              "eitheror_fullslack": ((2.0, -2.0), (4.0, 4.0)), "xor_fullslack": ((2.0, -2.0), (4.0, 4.0)),
              "mpvca": ((0.0, 0.0), (0.0, 5.0)), "mpvca_slack": ((0.0, 0.0), (0.0, 5.0)), "mpvca_fullslack": ((0.0, 0.0), (0.0, 5.0))}


def demo_rows(X0, out):
    """Rows of the per-start table (1-based id like Julia's loop index).  `runtime` is the launch's kernel time shared
    evenly by the starts: the batched kernel has no per-instance clock (the reference times each sequential solve)."""
    X0 = np.asarray(X0, dtype=np.float64)
    x = np.asarray(out["x"])
    n = x.shape[0]
    rt = float(out.get("kernel_ms", 0.0)) * 1e-3 / max(n, 1)
    return [(i + 1, X0[i, 0], X0[i, 1], x[i, 0], x[i, 1], int(out["iters"][i]), int(out["sub_iters"][i]), rt,
             1 if out["status"][i] == 0 else 0) for i in range(n)]


def demo_stats(problem, rows, tolx=1e-6):
    """The one-row statistics table (eitheror.jl:363-383): counts, max / median, and how many starts end in the global
    minimiser, the local one, or elsewhere."""
    it = np.array([r[5] for r in rows], dtype=float)
    sub = np.array([r[6] for r in rows], dtype=float)
    rt = np.array([r[7] for r in rows], dtype=float)
    xf = np.array([(r[3], r[4]) for r in rows], dtype=float).reshape(-1, 2)
    glb, lcl = (np.array(m) for m in MINIMISERS[problem])
    is_g = np.linalg.norm(xf - glb, axis=1) <= tolx
    is_l = ~is_g & (np.linalg.norm(xf - lcl, axis=1) <= tolx)
    n = len(rows)
    cg, cl = int(is_g.sum()), int(is_l.sum())
    cu = n - cg - cl
    med = lambda a: float(np.median(a)) if n else float("nan")   # noqa: E731
    mx = lambda a: float(np.max(a)) if n else float("nan")       # noqa: E731
    return (n, int(sum(r[8] for r in rows)), mx(it), med(it), mx(sub), med(sub), mx(rt), med(rt),
            cg, 100 * cg / max(n, 1), cl, 100 * cl / max(n, 1), cu, 100 * cu / max(n, 1))


def write_demo_results(filepath, problem, X0, out):
    """Writes `<filepath>.csv` and `<filepath>_stats.csv` with the reference's headers; returns (rows, stats)."""
    rows = demo_rows(X0, out)
    with open(filepath + ".csv", "w", newline="") as fh:
        w = csv.writer(fh)
        w.writerow(DATA_COLUMNS)
        w.writerows(rows)
    stats = demo_stats(problem, rows)
    with open(filepath + "_stats.csv", "w", newline="") as fh:
        w = csv.writer(fh)
        w.writerow(STATS_COLUMNS)
        w.writerow(stats)
    return rows, stats
