#!/bin/bash
# round 2, call B: whole GPU suite + smoke + bench (N=1, driver flags)
mkdir -p gpurun_out
echo "=== GPU suite"
timeout -s KILL 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider 2>&1 | tail -30
echo "=== smoke"
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "=== bench"
timeout -s KILL 1200 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_r2_a.json 2> gpurun_out/bench_r2_a.err
echo "rc=$?"; tail -3 gpurun_out/bench_r2_a.err | cut -c1-300; cat gpurun_out/bench_r2_a.json | cut -c1-6000
