#!/bin/bash
# warm-cache launch list of ONE batch-1 ours_fast generation (config 1), eager, aggregated by kernel
mkdir -p gpurun_out
timeout -s KILL 800 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 4500 -c 2193 --csv \
  --log-file gpurun_out/launches_c1_r2.csv python tools/latency_c1.py 1 > gpurun_out/launches_c1_r2.log 2>&1
echo "rc=$?"
python tools/summarize_launches.py gpurun_out/launches_c1_r2.csv gpurun_out/launches_c1_r2.md | head -30
