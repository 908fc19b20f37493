#!/bin/bash
mkdir -p gpurun_out
for v in 0 1; do
  ECADK_CONV_TAP3=$v timeout 600 ncu --set full --clock-control none --import-source on -k "regex:gemm2_bf16" -s 0 -c 6 \
    -o gpurun_out/prof_conv128_tap$v -f python tools/ncu_targets_vae.py conv128 > gpurun_out/ncu_conv128_$v.log 2>&1
  echo "TAP3=$v"; python tools/summarize_ncu.py gpurun_out/prof_conv128_tap$v.ncu-rep gpurun_out/conv128_tap$v 2>&1 | tail -8 | cut -c1-60,95-260
done
