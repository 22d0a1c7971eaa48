"""Writes tests/golden/oracle_golden.json from the CPU oracle (the Julia reference cannot run in
this image, so these freeze the ORACLE's outputs; the reference-owned golden vector of
test/problems/test_verbose.jl:29 is asserted directly in tests/test_oracle_pins.py)."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import bazinga_oracle as bo  # noqa: E402

A, _, b = bo.make_basis_pursuit(20, 100)
r = bo.solve(bo.problem_basis_pursuit(A, b), np.zeros(100))
rng = np.random.default_rng(42)
mu, y, x = rng.uniform(0.1, 2.0, 20), rng.standard_normal(20), rng.standard_normal(100)
al = bo.AugLagFun(bo.FZero(), bo.CDenseAffine(A, b), bo.ZeroSet(), mu.copy(), y.copy(), x)
dl = np.empty(100)
lx = al.gradient(dl, x)
gold = {
    "bp20x100": {"tot_it": r.tot_it, "tot_inner_it": r.tot_inner_it, "x": r.x.tolist(), "y": r.y.tolist(),
                 "objective": r.objective, "norm_res_prim": r.norm_res_prim},
    "grad": {"mu": mu.tolist(), "y": y.tolist(), "x": x.tolist(), "lx": lx, "dlx": dl.tolist(),
             "cx": al.cx.tolist(), "yupd": al.yupd.tolist()},
}
with open(os.path.join(HERE, "oracle_golden.json"), "w") as fh:
    json.dump(gold, fh)
print("wrote oracle_golden.json")
