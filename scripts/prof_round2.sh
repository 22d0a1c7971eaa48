# ncu evidence for the kernels of round 2 (run inside gpurun; text summaries are made on the box, the large reports of the
# batched kernels are dropped: gpurun_out/ carries 64 MiB at most)
set -x
O=gpurun_out
S="python scripts/ncu_summary.py"
B="python bench.py --legs none --steps 1 --warmup 1 --inner-cap 60 --no-cpu-baseline"
$B > $O/r2n_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 400 --csv --log-file $O/r2n_launches_bp4096x16384.csv $B > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_dense_algrad -s 30 -c 2 -o $O/r2n_dense_algrad $B > /dev/null 2>&1
$S $O/r2n_dense_algrad.ncu-rep --alg-bytes 537296896 > $O/r2n_dense_algrad_summary.txt
ncu --set full --clock-control none -k "regex:k_fstep|k_dstep" -s 20 -c 2 -o $O/r2n_fd $B > /dev/null 2>&1
$S $O/r2n_fd.ncu-rep > $O/r2n_fd_summary.txt; rm -f $O/r2n_fd.ncu-rep
export BZ_BATCH_CONCURRENT=0
for name in eitheror mpvca; do
  python scripts/c3_full.py $name && ncu --set full --clock-control none --import-source on -k regex:k_batch_alps -c 2 -o $O/r2n_batch_$name python scripts/c3_full.py $name > /dev/null 2>&1
  $S $O/r2n_batch_$name.ncu-rep > $O/r2n_batch_${name}_summary.txt
  python scripts/ncu_lines.py $O/r2n_batch_$name.ncu-rep > $O/r2n_batch_${name}_lines.txt 2>&1
  rm -f $O/r2n_batch_$name.ncu-rep
done
unset BZ_BATCH_CONCURRENT
python scripts/prof_matprox.py 20 4096 | tee $O/r2n_matprox.txt && ncu --set full --clock-control none -k regex:k_matprox -s 1 -c 1 -o $O/r2n_matprox python scripts/prof_matprox.py 20 4096 > /dev/null 2>&1
$S $O/r2n_matprox.ncu-rep > $O/r2n_matprox_summary.txt; rm -f $O/r2n_matprox.ncu-rep
python scripts/prof_bdense.py | tee $O/r2n_bdense.txt && ncu --set full --clock-control none -k regex:k_bd_ -s 4 -c 2 -o $O/r2n_bdense python scripts/prof_bdense.py > /dev/null 2>&1
$S $O/r2n_bdense.ncu-rep > $O/r2n_bdense_summary.txt
ncu -i $O/r2n_bdense.ncu-rep --page raw --csv | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]
for r in rows[2:]:
    print(r[h.index('Kernel Name')][:40])
    for i,k in enumerate(h):
        if 'dmma' in k.lower() or 'tensor' in k.lower(): print('   ',k,r[i],rows[1][i])
" > $O/r2n_bdense_tensor_metrics.txt; rm -f $O/r2n_bdense.ncu-rep
python scripts/prof_c4.py 4096 12 | tee $O/r2n_c4.txt && ncu --set full --clock-control none -k "regex:k_obstacle_fused|k_lstep" -s 6 -c 4 -o $O/r2n_c4 python scripts/prof_c4.py 4096 12 > /dev/null 2>&1
$S $O/r2n_c4.ncu-rep > $O/r2n_c4_summary.txt; rm -f $O/r2n_c4.ncu-rep
du -sh $O; ls -la $O
