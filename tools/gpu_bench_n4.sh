#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 \
  bench.py --gpus 4 --steps 6 --warmup 3 > gpurun_out/bench_r2_n4.json 2> gpurun_out/bench_r2_n4.err
echo "rc=$?"; wc -l gpurun_out/bench_r2_n4.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2_n4.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value'],'n',d['n_gpus'])
print('pop72',json.dumps(d['population72'])[:900]); print('cpu',d['cpu_baseline'])
PY
