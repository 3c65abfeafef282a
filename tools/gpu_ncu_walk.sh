#!/bin/bash
# ncu --set full (+ source) of the long-walk kernel and the placement kernel at the s40 level of ONE job
mkdir -p gpurun_out
REP=gpurun_out/prof_walk
timeout 1200 ncu --set full --import-source on --clock-control none \
  -k regex:"${NCU_K:-lq_af_walk3_k|lq_af_place_k}" --launch-skip ${NCU_SKIP:-4} --launch-count ${NCU_COUNT:-2} \
  -f -o $REP python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-cli --no-sdust > gpurun_out/ncu_walk.log 2>&1
tail -3 gpurun_out/ncu_walk.log | cut -c1-200
ncu -i $REP.ncu-rep --page raw --csv > gpurun_out/prof_walk_raw.csv 2> gpurun_out/ncu_export.err
for k in lq_af_walk3_k lq_af_place_k; do
  ncu -i $REP.ncu-rep --page source --csv -k regex:"^$k" > gpurun_out/prof_walk_src_$k.csv 2>> gpurun_out/ncu_export.err
done
gzip -f gpurun_out/prof_walk_src_*.csv gpurun_out/prof_walk_raw.csv
SZ=$(stat -c %s $REP.ncu-rep); if [ "$SZ" -gt 30000000 ]; then rm -f $REP.ncu-rep; echo "rep too big ($SZ), removed"; fi
ls -la gpurun_out/ | grep prof_walk
