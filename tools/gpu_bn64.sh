#!/bin/bash
# 64-wide GEMM tiles for the batch-1 latency configuration: parity of the small shapes, then latency A/B
mkdir -p gpurun_out
timeout -s KILL 400 python -m pytest tests/test_gpu_kernels.py -q -x -k "gemm" 2>&1 | tail -3
for v in 0 1; do
  for mode in "" graph; do
    echo "ECADK_GEMM_BN64=$v $mode"
    ECADK_GEMM_BN64=$v timeout -s KILL 300 python tools/latency_c1.py 1 $mode 2>&1 | tail -1
  done
done
ECADK_GEMM_BN64=1 timeout -s KILL 300 python tools/latency_c1.py 4 graph 2>&1 | tail -1
ECADK_GEMM_BN64=0 timeout -s KILL 300 python tools/latency_c1.py 4 graph 2>&1 | tail -1
