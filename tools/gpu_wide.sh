#!/bin/bash
# wide-key (k > 15) parity
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden_cli.py -m gpu -q -k "k19 or k17 or k28 or k16 or sketch_adversarial or fast_k15 or plain_pb" ) > gpurun_out/pytest_wide.log 2>&1
tail -25 gpurun_out/pytest_wide.log
