#!/bin/bash
# Runs each kernel test in its own process under a timeout so one hung kernel cannot take the others down.
mkdir -p gpurun_out
LOG=gpurun_out/first_light.log
: > $LOG
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv >> $LOG 2>&1
for t in test_residual_ln test_patch_embed_and_final_layer test_timestep_path_and_small_ops test_bad_arguments \
         test_gemm_bias test_gemm_gated_residual_cache test_gemm_headmajor "test_attention" ; do
  echo "=== $t" >> $LOG
  timeout -s KILL 240 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "$t" --tb=short -p no:cacheprovider 2>&1 | tail -40 >> $LOG
  echo "exit=$?" >> $LOG
done
tail -150 $LOG
