#!/bin/bash
# 2-GPU visit + the single-GPU case that walks few regions (140 000 targets in one part)
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_zz_c1_consumer.py -m gpu -x -q -k many_targets ) > gpurun_out/pytest_many.log 2>&1
tail -4 gpurun_out/pytest_many.log
bash tools/gpu_multi.sh 2
