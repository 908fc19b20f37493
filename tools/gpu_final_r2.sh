#!/bin/bash
# what the driver runs at round end: GPU tests, smoke, the default bench line, the reference arm (short)
mkdir -p gpurun_out
timeout -s KILL 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout -s KILL 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout -s KILL 1500 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_r2_final.json 2> gpurun_out/bench_r2_final.err
echo "bench rc=$?"; wc -l gpurun_out/bench_r2_final.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2_final.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value'],'clocks',d['clocks'])
print('roofline frac',d['roofline']['frac'],'step_frac',d['roofline']['step_frac_of_sustained'])
print('ours_fast',d['ours_fast']['images_per_s'],'pop72',d['population72']['images_per_s'] if d['population72'] else None)
print('vae',json.dumps(d['vae_decode'])[:400]); print('flux',json.dumps(d['flux_c5'])[:300]); print('cpu',d['cpu_baseline'])
PY
