#!/bin/bash
# VAE decode: parity, throughput, and the launch list of one batch-100 decode aggregated by kernel
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_vae.py -x -q 2>&1 | tail -3
timeout -s KILL 200 python tools/vae_bench.py 100 32 2>&1 | tail -1 | tee gpurun_out/vae_bench_b100.json
timeout -s KILL 200 python tools/vae_bench.py 1 32 2>&1 | tail -1 | tee gpurun_out/vae_bench_b1.json
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 1620 -c 534 --csv \
  --log-file gpurun_out/launches_vae.csv python tools/vae_bench.py 100 32 > gpurun_out/launches_vae.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_vae.csv gpurun_out/launches_vae.md | head -24
