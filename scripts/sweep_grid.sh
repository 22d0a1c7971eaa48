#!/bin/bash
# tuning sweep of the grid-wide dense kernel (bz_grid.cu): warp-role variant, ring slots left in flight, L2 prefetch
# distance.  CFGS="v,i,p v,i,p ..." (comma inside a config, blank between configs)
for cfg in ${CFGS:-0,4,0 1,4,0 2,4,0 3,4,0 0,6,0 2,6,0}; do
  IFS=, read -r v i p <<< "$cfg"
  if [ -n "$DRY" ]; then echo "variant=$v inflight=$i prefetch=$p"; continue; fi
  BZ_GRID_VARIANT=$v BZ_GRID_INFLIGHT=$i BZ_GRID_PREFETCH=$p timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29511 scripts/bench_sharded.py --rows ${ROWS:-8192} --cols ${COLS:-262144} --inner 30 > gpurun_out/sweep_tmp.log 2>&1
  echo "variant=$v inflight=$i prefetch=$p rc=$? $(grep -o 'al_gradient_local.: [0-9.]*' gpurun_out/sweep_tmp.log)"
done
