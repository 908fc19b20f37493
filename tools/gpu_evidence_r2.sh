#!/bin/bash
# round 2 evidence with the final build: kernel / attention times, ncu --set full of the hot kernels, launch list of the
# bench command, config-1 latency, dense forwards of configs 3/4, CPU batch sensitivity, the bench lines
mkdir -p gpurun_out
echo "=== attention times"
{ echo "# python tools/attn_times.py (B200, CUDA events, 3 launches per event pair) - round-2 build"; timeout -s KILL 200 python tools/attn_times.py; \
  echo "# ECADK_ATTN_MODE=pair1 (first-generation 256-query kernel, head-major operands only)"; ECADK_ATTN_MODE=pair1 timeout -s KILL 100 python tools/attn_times.py 2>&1 | grep -v "row-major" | head -2; \
  echo "# phase clocks per (sample, head) item, attn_pair2_kernel, instrumented build (tools/micro/attn_phase_timing.py)"; \
  for m in c2self c2cross; do ECAD_B200_LIB=tools/micro/libecad_b200_timing.so timeout -s KILL 120 python tools/micro/attn_phase_timing.py $m; done; } > gpurun_out/r2_attention_times.txt 2>&1
tail -5 gpurun_out/r2_attention_times.txt
echo "=== kernel times"
timeout -s KILL 300 python tools/kernel_times.py 200 > gpurun_out/r2_kernel_times.txt 2>&1; cp gpurun_out/kernel_times.json gpurun_out/r2_kernel_times_batch100.json; tail -3 gpurun_out/r2_kernel_times.txt
echo "=== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:attn_pair|gemm2_bf16|residual_ln" -s 10 -c 10 \
  -o gpurun_out/prof_r2_kernels -f python tools/ncu_targets.py > gpurun_out/ncu_targets_r2.log 2>&1; tail -2 gpurun_out/ncu_targets_r2.log
echo "=== launch list of the bench command (ours_fast, 1 step)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 2500 -c 2450 --csv \
  --log-file gpurun_out/launches_r2.csv python bench.py --steps 1 --warmup 1 --fixed-schedule --no-cpu-baseline --no-flux --no-population72 \
  > gpurun_out/ncu_bench_r2.log 2>&1; tail -1 gpurun_out/ncu_bench_r2.log | cut -c1-200
echo "=== latency c1"
{ timeout -s KILL 200 python tools/latency_c1.py 1; timeout -s KILL 200 python tools/latency_c1.py 1 graph; timeout -s KILL 200 python tools/latency_c1.py 4 graph; } > gpurun_out/r2_latency_c1.jsonl 2>/dev/null; cat gpurun_out/r2_latency_c1.jsonl
echo "=== dense forwards c2/c3/c4"
{ for c in c2 c3 c4; do timeout -s KILL 300 python tools/forward_bench.py $c 2>/dev/null | tail -1; done; } > gpurun_out/r2_dense_forward.jsonl; cat gpurun_out/r2_dense_forward.jsonl | cut -c1-300
echo "=== CPU batch sensitivity"
timeout -s KILL 600 python tools/cpu_batch_sensitivity.py 1 2 4 > gpurun_out/r2_cpu_batch_sensitivity.txt 2>&1; cat gpurun_out/r2_cpu_batch_sensitivity.txt
echo "=== bench (driver flags)"
timeout -s KILL 1200 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_r2_final_n1.json 2> gpurun_out/bench_r2_final_n1.err; echo "rc=$?"; cut -c1-300 gpurun_out/bench_r2_final_n1.json
timeout -s KILL 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/bench_r2_final_ref.json 2>/dev/null; cut -c1-300 gpurun_out/bench_r2_final_ref.json
