# phase-2 grid size / concurrency sweep of the batched launch (65,536-start demo grids and the tail instances alone)
for c in 1 0; do for p in 1 2 4; do
  VARIANT="conc$c p2ctas$p" BZ_BATCH_CONCURRENT=$c BZ_BATCH_P2CTAS=$p python scripts/c3_quick.py
done; done
