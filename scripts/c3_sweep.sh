# end-game sweep of the batched launch: phase-1 threads left at which the remaining solves move to persistent warps
for d in 0 -1 200 1200 2400 4800; do for p in 2 4; do
  VARIANT="drain$d p2ctas$p" BZ_BATCH_DRAIN=$d BZ_BATCH_P2CTAS=$p python scripts/c3_quick.py
done; done
