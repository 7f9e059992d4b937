#!/bin/bash
# usage: r2_mp.sh N TAG — the bench line with only the multipatch / strips sub-records (N = 1: python, N > 1: torchrun)
N=$1; TAG=$2
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
FLAGS="--steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-configs --no-apalm --no-solid"
if [ "$N" = "1" ]; then
  timeout 900 python bench.py $FLAGS > gpurun_out/bench_${TAG}_n$N.json 2> gpurun_out/bench_${TAG}_n$N.err
else
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29556 bench.py --gpus $N $FLAGS > gpurun_out/bench_${TAG}_n$N.json 2> gpurun_out/bench_${TAG}_n$N.err
fi
tail -3 gpurun_out/bench_${TAG}_n$N.err | cut -c1-300
python - <<PY
import json
d=json.load(open('gpurun_out/bench_${TAG}_n$N.json'))
print('N', d['n_gpus'], 'ms_per_step', d['ms_per_step'], 'jac', d['jacobian_ms'])
print('multipatch', d.get('multipatch'))
s=d.get('strong')
print('strong', s and {k:v for k,v in s.items() if k not in ('what','nccl_op')})
PY
