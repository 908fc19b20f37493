#!/bin/bash
# refresh the round's ncu evidence with the current kernels (one GPU, ~5 min)
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:attn_pair|gemm2_bf16|residual_ln" -s 10 -c 10 \
  -o gpurun_out/prof_r1_kernels_v2 -f python tools/ncu_targets.py > gpurun_out/ncu_targets_v2.log 2>&1
tail -2 gpurun_out/ncu_targets_v2.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 2500 -c 2450 --csv \
  --log-file gpurun_out/launches_r1_v2.csv python bench.py --steps 1 --warmup 1 --fixed-schedule --no-cpu-baseline \
  > gpurun_out/ncu_bench_v2.log 2>&1
tail -2 gpurun_out/ncu_bench_v2.log | cut -c1-300
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r1_final2_n1.json 2> gpurun_out/bench_r1_final2_n1.err
cut -c1-400 gpurun_out/bench_r1_final2_n1.json
timeout 600 python bench.py --steps 3 --warmup 3 --fixed-schedule --no-cpu-baseline > gpurun_out/bench_r1_final2_n1_oursfast.json 2> /dev/null
cut -c1-400 gpurun_out/bench_r1_final2_n1_oursfast.json
