import os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bazinga_b200 as bz
v = np.linspace(-4.0, 8.0, 256); G = np.array([(i, j) for j in v for i in v])
out = bz.alps_batch("eitheror", G)
order = np.argsort(-out["sub_iters"])
print("top iters", out["sub_iters"][order[:5]], "outer", out["iters"][order[:5]], "kernel_ms", out["kernel_ms"])
hard = G[order[0]]; easy = G[order[-1]]
for name, X in [("1 hard", hard[None]), ("32 hard", np.tile(hard, (32, 1))), ("64 hard", np.tile(hard, (64, 1))),
                ("1 easy", easy[None]), ("4096 hard", np.tile(hard, (4096, 1))), ("65536 hard", np.tile(hard, (65536, 1))),
                ("65536 easy", np.tile(easy, (65536, 1))), ("sorted grid", G[order]), ("grid", G)]:
    best = 1e9
    for _ in range(3):
        o = bz.alps_batch("eitheror", X); best = min(best, o["kernel_ms"])
    print(f"{name:12s} kernel_ms {best:9.3f}  iters[0] {o['sub_iters'][0]} outer {o['iters'][0]}  total inner {o['sub_iters'].sum()}")
