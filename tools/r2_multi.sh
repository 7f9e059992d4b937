#!/bin/bash
# usage: r2_multi.sh N TAG — strips test + the driver's N-GPU bench launch
N=$1; TAG=$2
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
[ -z "$SKIPTEST" ] && timeout 900 python -m pytest tests/test_gpu_multigpu.py -x -q 2>&1 | tail -5
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_${TAG}_n$N.json 2> gpurun_out/bench_${TAG}_n$N.err
tail -5 gpurun_out/bench_${TAG}_n$N.err | cut -c1-300
python - <<PY
import json
d=json.load(open('gpurun_out/bench_${TAG}_n$N.json'))
print('N', d['n_gpus'], 'ms_per_step', d['ms_per_step'], 'value', d['value'])
for k in ('e2e','e2e_lower','e2e_device_solve'):
    print(k, d[k] and {a:b for a,b in d[k].items() if a not in ('what','with_solve','jacobian_breakdown_ms')})
print('strong', d['strong'])
PY
python - <<PY
import json
d=json.load(open('gpurun_out/bench_${TAG}_n$N.json'))
print('apalm', d.get('apalm'))
PY
