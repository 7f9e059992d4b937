#!/bin/bash
# usage: r2_full.sh TAG — all GPU tests, then the default bench line (as the driver runs it)
TAG=$1
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -12
timeout 900 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
tail -3 gpurun_out/bench_$TAG.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_$TAG.json'))
print('ms_per_step', d['ms_per_step'], 'jac_ms', d['jacobian_ms'], 'pts', d['points_residual_ms'], 'frac', d['roofline']['frac'], d['clocks'])
for k in ('e2e','e2e_lower','e2e_device_solve'):
    print(k, d[k] and {a:b for a,b in d[k].items() if a!='what'})
print('cpu', d['cpu_baseline'])
for c in d['configs'] or []:
    print(c)
PY
