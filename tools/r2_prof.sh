#!/bin/bash
# usage: r2_prof.sh TAG [ENV=VAL...] — bench line + ncu launch list + full capture of the Jacobian kernel
TAG=$1; shift
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-configs --no-apalm --no-multipatch --no-solid > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
python -c "
import json; d=json.load(open('gpurun_out/bench_$TAG.json')); print('$TAG', 'ms_per_step', d['ms_per_step'], 'jac_ms', d['jacobian_ms'], 'frac', d['roofline']['frac'], d['clocks'])" || tail -3 gpurun_out/bench_$TAG.err
env "$@" ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-configs --no-apalm --no-multipatch --no-solid > gpurun_out/b_ncu.log 2>&1
env "$@" ncu --set full --clock-control none --import-source on -k regex:k_jacobian -s 1 -c 1 -o gpurun_out/jac_$TAG timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-configs --no-apalm --no-multipatch --no-solid > gpurun_out/b_ncu2.log 2>&1
ls -la gpurun_out/jac_$TAG.ncu-rep
