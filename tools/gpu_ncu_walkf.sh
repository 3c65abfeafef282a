#!/bin/bash
# ncu --set full (+ source) of the few-region walker on ONE GPU: 200 000 target reads in one part make rid >> 16 take four values
mkdir -p gpurun_out
REP=gpurun_out/prof_walkf
timeout 1200 ncu --set full --import-source on --clock-control none \
  -k regex:"lq_af_walkf_k" --launch-skip 1 --launch-count 1 \
  -f -o $REP python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-cli --no-sdust --reads 200000 --queries 2500 > gpurun_out/ncu_walkf.log 2>&1
tail -3 gpurun_out/ncu_walkf.log | cut -c1-300
ncu -i $REP.ncu-rep --page raw --csv > gpurun_out/prof_walkf_raw.csv 2> gpurun_out/ncu_export.err
ncu -i $REP.ncu-rep --page source --csv > gpurun_out/prof_walkf_src.csv 2>> gpurun_out/ncu_export.err
gzip -f gpurun_out/prof_walkf_src.csv gpurun_out/prof_walkf_raw.csv
rm -f $REP.ncu-rep
ls -la gpurun_out/ | grep prof_walkf
