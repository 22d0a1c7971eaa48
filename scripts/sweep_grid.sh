#!/bin/bash
# timing of the grid-wide dense kernel (bz_grid.cu) on one GPU: CFGS="variant variant ...", ROWS, COLS
for v in ${CFGS:-0 1}; do
  BZ_GRID_VARIANT=$v timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29511 scripts/bench_sharded.py --rows ${ROWS:-8192} --cols ${COLS:-262144} --inner 30 > gpurun_out/sweep_tmp.log 2>&1
  echo "variant=$v rows=${ROWS:-8192} cols=${COLS:-262144} rc=$? $(grep -o 'al_gradient_local.: [0-9.]*' gpurun_out/sweep_tmp.log)"
done
