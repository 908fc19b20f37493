#!/bin/bash
# round 2: ncu --set full of every hot kernel at the batch-100 shapes (one warm launch each), sources imported
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:attn_pair|gemm2_bf16|residual_ln" -s 10 -c 10 \
  -o gpurun_out/prof_r2_kernels -f python tools/ncu_targets.py > gpurun_out/ncu_targets_r2.log 2>&1
tail -3 gpurun_out/ncu_targets_r2.log
