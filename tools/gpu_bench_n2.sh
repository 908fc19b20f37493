#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 2 --steps 6 --warmup 3 > gpurun_out/bench_r2_n2.json 2> gpurun_out/bench_r2_n2.err
echo "rc=$?"; tail -2 gpurun_out/bench_r2_n2.err | cut -c1-300
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2_n2.json'))
print('value',d['value'],'e2e',d['e2e']['value'],'n',d['n_gpus'])
print('pop72',json.dumps(d['population72'])[:900]); print('cpu',d['cpu_baseline'])
PY
