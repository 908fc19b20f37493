#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/suite_r2.log 2>&1
grep -v "^WARNING" gpurun_out/suite_r2.log | tail -40
