#!/bin/bash
# session 3 of round 2: full GPU suite on the build with NVTX + the config-driven generators, and an ncu capture cut by
# NVTX range (only the kernels of block 1's feed-forward, then only block 0's attn1)
mkdir -p gpurun_out
timeout -s KILL 1200 python -m pytest tests -q -m gpu > gpurun_out/gpu_suite_r2c.log 2>&1; tail -4 gpurun_out/gpu_suite_r2c.log; grep -E "^(FAILED|ERROR)|Error|assert" gpurun_out/gpu_suite_r2c.log | head -20
for r in "b01.ff]" "b00.attn1]"; do
  n=$(echo "$r" | tr -d ']')
  timeout -s KILL 300 ncu --nvtx --nvtx-include "$r" --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/nvtx_${n}.csv python tools/nvtx_demo.py 2>&1 | tail -2
done
wc -l gpurun_out/nvtx_*.csv
