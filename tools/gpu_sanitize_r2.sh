#!/bin/bash
# round 2: compute-sanitizer over the kernel-level parity tests (small shapes), logs -> gpurun_out/
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
SEL='test_gemm or test_attention or test_residual_ln or test_patch or test_timestep'
echo "=== memcheck"
timeout -s KILL 900 $CS --tool memcheck --error-exitcode 7 --print-limit 20 \
  python -m pytest tests/test_gpu_kernels.py -m gpu -q -x --tb=line -p no:cacheprovider -k "$SEL and not many_items and not streaming" \
  > gpurun_out/r2_sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -6 gpurun_out/r2_sanitizer_memcheck.log
echo "=== racecheck (shared-memory hazards; glue kernels + attention/GEMM at their smallest shapes)"
timeout -s KILL 900 $CS --tool racecheck --error-exitcode 7 --print-limit 20 \
  python -m pytest tests/test_gpu_kernels.py -m gpu -q -x --tb=line -p no:cacheprovider -k "test_residual_ln or test_attention_all_masked or test_timestep or test_patch" \
  > gpurun_out/r2_sanitizer_racecheck.log 2>&1
echo "racecheck rc=$?"; tail -6 gpurun_out/r2_sanitizer_racecheck.log
