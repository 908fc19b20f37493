#!/bin/bash
# kernel parity (each file under its own timeout) then per-kernel timings, for both GEMM variants
mkdir -p gpurun_out
for g in 2 1; do
  echo "=== ECADK_GEMM_CTA_GROUP=$g kernel tests"
  ECADK_GEMM_CTA_GROUP=$g timeout -s KILL 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short -p no:cacheprovider -x 2>&1 | tail -15
done
echo "=== ECADK_ATTN_MODE=tile attention tests"
ECADK_ATTN_MODE=tile timeout -s KILL 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short -p no:cacheprovider -k attention 2>&1 | tail -5
echo "=== ECADK_ATTN_MODE=flash attention tests (streaming kernel on the short shapes too)"
ECADK_ATTN_MODE=flash timeout -s KILL 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short -p no:cacheprovider -k attention 2>&1 | tail -8
echo "=== auto mode: model parity"
timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -p no:cacheprovider -x 2>&1 | tail -15
for g in 2; do
  echo "=== kernel times, cta_group=$g"
  ECADK_GEMM_CTA_GROUP=$g timeout -s KILL 300 python tools/kernel_times.py 200 2>&1 | tail -24
done
