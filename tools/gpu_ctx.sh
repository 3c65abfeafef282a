#!/bin/bash
mkdir -p gpurun_out; L=gpurun_out/ctx_time.log; : > $L
nvidia-smi -q | grep -i -E "persistence|product name|attached" >> $L
echo "CUDA_VISIBLE_DEVICES=$CUDA_VISIBLE_DEVICES" >> $L
for i in 1 2 3; do ( time tools/probe/ctx_time 0 30 ) >> $L 2>&1; done
echo "--- 2 GB only" >> $L
for i in 1 2; do ( time tools/probe/ctx_time 0 2 ) >> $L 2>&1; done
echo "--- while another process holds a context" >> $L
tools/probe/ctx_time 30 1 >> $L 2>&1 &
sleep 6
for i in 1 2 3; do ( time tools/probe/ctx_time 0 30 ) >> $L 2>&1; done
for i in 1 2; do ( time tools/probe/ctx_time 0 2 ) >> $L 2>&1; done
wait
cat $L
