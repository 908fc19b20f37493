#!/bin/bash
# session 3 of round 2: compute-sanitizer memcheck over the padded-token-count forwards (ecadk_patch_embed_padded's
# 2-D memset, the biased streaming self-attention, the unpatchify guard) and the NVTX test
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
timeout -s KILL 1200 $CS --tool memcheck --error-exitcode 7 --print-limit 20 \
  python -m pytest tests/test_gpu_parity.py tests/test_nvtx.py -m gpu -q -x --tb=line -p no:cacheprovider \
  -k "not_a_multiple or one_range" > gpurun_out/r2_sanitizer_memcheck_padded.log 2>&1
echo "memcheck rc=$?"; tail -6 gpurun_out/r2_sanitizer_memcheck_padded.log
