import os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bazinga_b200 as bz
hard = np.array([[-4.0 + 12.0 * 133 / 255, -4.0 + 12.0 * 0 / 255]])
v = np.linspace(-4.0, 8.0, 256); G = np.array([(i, j) for j in v for i in v])
o = bz.alps_batch("eitheror", G[:4096]); k = int(np.argmax(o["sub_iters"]))
X = np.tile(G[k], (64, 1))
o = bz.alps_batch("eitheror", X)
print("iters", o["sub_iters"][0], "kernel_ms", o["kernel_ms"])
