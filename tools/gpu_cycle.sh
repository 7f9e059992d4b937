#!/bin/bash
# usage: gpu_cycle.sh TAG  — run on the GPU box: parity tests, bench, ncu launch list + full capture of k_jacobian
TAG=$1
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
python -c "
import json; d=json.load(open('gpurun_out/bench_$TAG.json')); print('ms_per_step', d['ms_per_step'], 'jac_ms', d['jacobian_ms'], 'frac', d['roofline']['frac'], 'e2e_ms', d['e2e']['ms_per_step'], d['clocks'])"
tail -3 gpurun_out/bench_$TAG.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/b_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_jacobian -s 1 -c 1 -o gpurun_out/jac_$TAG timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/b_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_points -s 1 -c 1 -o gpurun_out/pts_$TAG python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/b_ncu3.log 2>&1
