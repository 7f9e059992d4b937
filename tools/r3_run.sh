#!/bin/bash
# usage: r3_run.sh TAG [pytest-args...] — selected GPU tests, then the bench's device leg with the per-config records
TAG=$1; shift
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest "$@" -q -m gpu -x 2>&1 | tail -8
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-apalm --no-multipatch --no-solid > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
tail -3 gpurun_out/bench_$TAG.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_$TAG.json'))
print('ms_per_step', d['ms_per_step'], 'jac_ms', d['jacobian_ms'], 'pts', d['points_residual_ms'], 'frac', d['roofline']['frac'], d['clocks'])
for c in d['configs'] or []:
    print({k: c[k] for k in ('workload','step_ms','jacobian_ms','points_residual_ms','max_rel_diff_vs_oracle_12x12') if k in c})
PY
