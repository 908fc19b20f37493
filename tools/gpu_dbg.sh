#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short -p no:cacheprovider -k "attention" 2>&1 | tail -15
timeout -s KILL 200 python tools/attn_times.py 2>&1 | head -12
timeout -s KILL 200 python tools/kernel_times.py 200 2>&1 | head -14
timeout -s KILL 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/suite_r2.log 2>&1
grep -v "^WARNING" gpurun_out/suite_r2.log | tail -30
