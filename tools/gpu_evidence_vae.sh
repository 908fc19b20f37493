#!/bin/bash
# evidence for the late round-2 additions: ncu --set full of the convolution / GroupNorm / split-K kernels, and
# compute-sanitizer over their kernel-level tests
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
echo "=== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:gemm2_bf16|gemm_splitk|gn_stats|gn_apply|upsample2x" -s 0 -c 34 \
  -o /tmp/prof_r2_vae_splitk -f python tools/ncu_targets_vae.py > gpurun_out/ncu_targets_vae.log 2>&1; tail -2 gpurun_out/ncu_targets_vae.log
python tools/summarize_ncu.py /tmp/prof_r2_vae_splitk.ncu-rep gpurun_out/r2_vae_splitk_ncu_full 2>&1 | tail -30
echo "=== memcheck: VAE entry points (small shapes) + split-K GEMMs"
timeout -s KILL 1200 $CS --tool memcheck --error-exitcode 7 --print-limit 20 \
  python -m pytest tests/test_gpu_vae.py tests/test_gpu_kernels.py -m gpu -q -x --tb=line -p no:cacheprovider \
  -k "(test_conv_nhwc and not 24-64) or (test_conv_up2x and not 20-64) or test_groupnorm or test_upsample or (test_gemm and splitk and not 20480)" \
  > gpurun_out/r2_sanitizer_memcheck_vae_splitk.log 2>&1
echo "memcheck rc=$?"; tail -5 gpurun_out/r2_sanitizer_memcheck_vae_splitk.log
echo "=== racecheck: GroupNorm / softmax / glue (shared-memory reductions)"
timeout -s KILL 900 $CS --tool racecheck --error-exitcode 7 --print-limit 20 \
  python -m pytest tests/test_gpu_vae.py -m gpu -q -x --tb=line -p no:cacheprovider -k "test_groupnorm or test_upsample" \
  > gpurun_out/r2_sanitizer_racecheck_vae.log 2>&1
echo "racecheck rc=$?"; tail -5 gpurun_out/r2_sanitizer_racecheck_vae.log
