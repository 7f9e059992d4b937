#!/bin/bash
# usage: r2_ab.sh TAG name1:ENV=VAL[,ENV=VAL] name2:... — parity once on the default lib, then one bench line per variant
TAG=$1; shift
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -2
for spec in "$@"; do
  name=${spec%%:*}; envs=${spec#*:}; envs=${envs//,/ }
  env $envs timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-configs --no-apalm --no-multipatch --no-solid > gpurun_out/bench_${TAG}_$name.json 2> gpurun_out/bench_${TAG}_$name.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_${TAG}_$name.json')); print('$name', 'ms_per_step', round(d['ms_per_step'],3), 'jac_ms', round(d['jacobian_ms'],3), 'frac', round(d['roofline']['frac'],3), d['clocks']['sm_mhz'])" || tail -3 gpurun_out/bench_${TAG}_$name.err
done
