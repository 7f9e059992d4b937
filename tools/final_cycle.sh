#!/bin/bash
# usage: final_cycle.sh TAG — run on the GPU box: full GPU test suite, smoke, bench (both arms), material variants, ncu launch list
TAG=$1
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 600 python __graft_entry__.py smoke 2>&1 | grep "^smoke" | tail -5
timeout 600 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
python -c "
import json; d=json.load(open('gpurun_out/bench_$TAG.json')); print('ms_per_step', d['ms_per_step'], 'value', d['value'], 'jac_ms', d['jacobian_ms'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'cpu', d['cpu_baseline']['value'], d['clocks'], 'launches', d['gpu_launches'])"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${TAG}_reference.json 2>/dev/null
cut -c1-300 gpurun_out/bench_${TAG}_reference.json
for m in nh mr nh_c mr_c; do
  timeout 300 python bench.py --material $m --no-cpu-baseline --no-e2e > gpurun_out/bench_${TAG}_$m.json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/bench_${TAG}_$m.json')); print('$m', 'ms_per_step', round(d['ms_per_step'],3), 'jac_ms', round(d['jacobian_ms'],3))"
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/b_ncu.log 2>&1
python tools/solid_bench.py 2>/dev/null | tail -1 | cut -c1-600
