#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 1200 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_r2_b.json 2> gpurun_out/bench_r2_b.err
echo "rc=$?"; tail -2 gpurun_out/bench_r2_b.err | cut -c1-300
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2_b.json'))
r=d['roofline']
print('value',d['value'],'e2e',d['e2e']['value'],'step_frac',r['step_frac_of_sustained'],'gemm frac',r['frac'],'share',r['share_of_step'])
print('attn',r['other_kernels']['attention'],'glue',r['other_kernels']['glue_residual_ln'])
print('ours_fast',d['ours_fast']); print('pop72',d['population72']['images_per_s']); print('flux',d['flux_c5'].get('images_per_s'), d['flux_c5'].get('error'))
print('clocks',d['clocks'])
PY
