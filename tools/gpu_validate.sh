set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r1_e.json 2> gpurun_out/bench_r1_e.err; tail -2 gpurun_out/bench_r1_e.err; cat gpurun_out/bench_r1_e.json
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:attn_flash|gemm2_bf16|residual_ln|qk_norm_rope|strided_unary" -s 9 -c 9 -o gpurun_out/prof_r1_flux -f python tools/ncu_targets_flux.py > gpurun_out/ncu_flux.log 2>&1; tail -3 gpurun_out/ncu_flux.log
