#!/bin/bash
# round 2: second-generation 256-query attention kernel - parity, stand-alone times (new vs ECADK_ATTN_MODE=pair1), phase clocks
mkdir -p gpurun_out
echo "=== attention kernel tests (new kernel)"
timeout -s KILL 240 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short -p no:cacheprovider -k "attention" 2>&1 | tail -15
echo "=== attention kernel tests (pair1)"
ECADK_ATTN_MODE=pair1 timeout -s KILL 240 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short -p no:cacheprovider -k "attention" 2>&1 | tail -5
echo "=== times: new"
timeout -s KILL 200 python tools/attn_times.py 2>&1 | head -3
echo "=== times: pair1"
ECADK_ATTN_MODE=pair1 timeout -s KILL 200 python tools/attn_times.py 2>&1 | head -3
echo "=== phase clocks (new)"
for m in c2self c2cross; do ECAD_B200_LIB=tools/micro/libecad_b200_timing.so timeout -s KILL 120 python tools/micro/attn_phase_timing.py $m 2>&1 | tail -18; done
