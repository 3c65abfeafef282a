#!/bin/bash
# ncu --set full of the seed-sort kernels at the s40 and s32 levels of ONE job (launches 8..15 of the af_* family), exported on the box.
mkdir -p gpurun_out
REP=gpurun_out/prof_sort
timeout 1200 ncu --set full --import-source on --clock-control none \
  -k regex:"lq_af_big_k|lq_af_level_k|lq_af_walk_k|lq_af_walk_small_k" --launch-skip ${NCU_SKIP:-8} --launch-count ${NCU_COUNT:-8} \
  -f -o $REP python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_sort.log 2>&1
tail -3 gpurun_out/ncu_sort.log | cut -c1-200
ncu -i $REP.ncu-rep --page raw --csv > gpurun_out/prof_sort_raw.csv 2> gpurun_out/ncu_export.err
for k in lq_af_walk_k lq_af_big_k lq_af_level_k; do
  ncu -i $REP.ncu-rep --page source --csv -k regex:"^$k" > gpurun_out/prof_sort_src_$k.csv 2>> gpurun_out/ncu_export.err
done
gzip -f gpurun_out/prof_sort_src_*.csv gpurun_out/prof_sort_raw.csv
SZ=$(stat -c %s $REP.ncu-rep); if [ "$SZ" -gt 30000000 ]; then rm -f $REP.ncu-rep; echo "rep too big ($SZ), removed"; fi
ls -la gpurun_out/ | grep prof_sort
