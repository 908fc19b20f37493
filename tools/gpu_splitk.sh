#!/bin/bash
# split-K GEMMs for the batch-1 latency configuration: parity of the small shapes, then latency A/B
mkdir -p gpurun_out
timeout -s KILL 400 python -m pytest tests/test_gpu_kernels.py -q -x -k "gemm" 2>&1 | tail -5
for v in 0 1; do
  echo "ECADK_GEMM_SPLITK=$v graph"
  ECADK_GEMM_SPLITK=$v timeout -s KILL 300 python tools/latency_c1.py 1 graph 2>&1 | tail -1
done
ECADK_GEMM_SPLITK=1 timeout -s KILL 300 python tools/latency_c1.py 1 2>&1 | tail -1
