#!/bin/bash
# compute-sanitizer over the kernel-level tests of the late round-2 additions (VAE entry points at small shapes, split-K GEMMs)
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
timeout -s KILL 1200 $CS --tool memcheck --error-exitcode 7 --print-limit 20 \
  python -m pytest tests/test_gpu_vae.py tests/test_gpu_kernels.py -m gpu -q -x --tb=line -p no:cacheprovider \
  -k "(test_conv_nhwc and not 24-64) or (test_conv_up2x and not 20-64) or test_groupnorm or test_upsample or (test_gemm and splitk and not 20480)" \
  > gpurun_out/r2_sanitizer_memcheck_vae_splitk.log 2>&1
echo "memcheck rc=$?"; tail -4 gpurun_out/r2_sanitizer_memcheck_vae_splitk.log
timeout -s KILL 900 $CS --tool racecheck --error-exitcode 7 --print-limit 20 \
  python -m pytest tests/test_gpu_vae.py -m gpu -q -x --tb=line -p no:cacheprovider -k "test_groupnorm or test_upsample" \
  > gpurun_out/r2_sanitizer_racecheck_vae.log 2>&1
echo "racecheck rc=$?"; tail -4 gpurun_out/r2_sanitizer_racecheck_vae.log
