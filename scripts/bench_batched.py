"""Batched tiny instances over N GPUs: one process per GPU under torchrun, instance i ->
GPU i mod N, no data-path collective.  Reports strong scaling (65,536 instances in total) and weak scaling (65,536
per GPU) for the either-or and the MPVC demo problems; time = max over ranks of the device-timed kernel.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
      scripts/bench_batched.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def grid(lo, hi, k):
    v = np.linspace(lo, hi, k)
    return np.array([(i, j) for j in v for i in v])


def main():
    import torch
    import torch.distributed as dist
    import bazinga_b200 as bz
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = bz.Context(local)
    for name, lo, hi in (("eitheror", -4.0, 8.0), ("mpvca", -5.0, 20.0)):
        X0 = grid(lo, hi, 256)                                  # 65,536 starts on the demo box, column-major
        bz.alps_batch(name, X0[:1024], ctx=ctx)                 # warm-up (module load)
        for mode, mine in (("strong", X0[rank::world]), ("weak", X0)):
            best = None
            for _ in range(3):
                dist.barrier()
                out = bz.alps_batch(name, mine, ctx=ctx)
                t = torch.tensor([out["kernel_ms"]], dtype=torch.float64, device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                best = float(t.item()) if best is None else min(best, float(t.item()))
            tot = torch.tensor([len(mine), int(out["sub_iters"].sum()), int((out["status"] == 0).sum())], dtype=torch.float64, device="cuda")
            dist.all_reduce(tot)
            if rank == 0:
                print(json.dumps({"config": "C3 batched", "problem": name, "scaling": mode, "n_gpus": world,
                                  "instances_total": int(tot[0].item()), "solved": int(tot[2].item()),
                                  "inner_iterations_total": int(tot[1].item()), "kernel_ms_max_over_ranks": best,
                                  "solves_per_s": tot[0].item() / (best * 1e-3),
                                  "inner_iterations_per_s": tot[1].item() / (best * 1e-3)}), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
