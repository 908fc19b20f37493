#!/bin/bash
for lib in ecad_b200/libecad_b200.so tools/micro/variants/lib_*.so; do
  echo "=== $lib"
  ECAD_B200_LIB=$lib timeout -s KILL 200 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_flux_kernels.py -m gpu -q --tb=line -p no:cacheprovider -k "gemm" 2>&1 | tail -4
  ECAD_B200_LIB=$lib timeout -s KILL 200 python tools/micro/epi2_sensitivity.py 2>&1 | tail -14
done
