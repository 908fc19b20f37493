#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_gpu_kernels.py -q -x -k "residual_ln" 2>&1 | tail -2
echo "default graph"; timeout -s KILL 300 python tools/latency_c1.py 1 graph 2>&1 | tail -1
echo "NO_QKV pair2 graph"; ECAD_B200_NO_QKV=1 timeout -s KILL 300 python tools/latency_c1.py 1 graph 2>&1 | tail -1
echo "NO_QKV tile graph"; ECAD_B200_NO_QKV=1 ECADK_ATTN_MODE=tile timeout -s KILL 300 python tools/latency_c1.py 1 graph 2>&1 | tail -1
bash tools/gpu_c1_launches.sh 2>&1 | head -16
