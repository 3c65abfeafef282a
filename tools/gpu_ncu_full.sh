#!/bin/bash
# One ncu --set full capture of the hot kernels of ONE job (first job of bench.py), exported to CSV on the box
# (the .ncu-rep is only brought back when it fits gpurun's 64 MiB return limit).
mkdir -p gpurun_out
REP=gpurun_out/prof_full
timeout 1200 ncu --set full --clock-control none \
  -k regex:"${NCU_KERNELS:-lq_af_level_k|lq_af_walk_k|lq_fill_filtered_k|lq_filter_count_k|lq_sketch_roll_k|lq_rs_scatter_k|lq_lookup_k|lq_chain_k|lq_gather_k|lq_fill_k}" \
  -c ${NCU_COUNT:-32} -f -o $REP python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log | cut -c1-300
ncu -i $REP.ncu-rep --page raw --csv > gpurun_out/prof_full_raw.csv 2> gpurun_out/ncu_export.err
ncu -i $REP.ncu-rep --page details --csv > gpurun_out/prof_full_details.csv 2>> gpurun_out/ncu_export.err
for k in lq_af_level_k lq_af_walk_k lq_fill_filtered_k lq_sketch_roll_k; do
  ncu -i $REP.ncu-rep --page source --csv -k regex:$k > gpurun_out/prof_src_$k.csv 2>> gpurun_out/ncu_export.err
done
gzip -f gpurun_out/prof_src_*.csv gpurun_out/prof_full_raw.csv
SZ=$(stat -c %s $REP.ncu-rep)
if [ "$SZ" -gt 30000000 ]; then rm -f $REP.ncu-rep; echo "rep too big ($SZ), removed"; fi
ls -la gpurun_out/
du -sh gpurun_out
