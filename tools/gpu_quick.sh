#!/bin/bash
# usage: gpu_quick.sh TAG [ENVVAR=VAL ...] — GPU parity tests, then one bench line per extra environment setting (A/B runs)
TAG=$1; shift
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
run() {
  local name=$1; shift
  env "$@" timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_${TAG}_$name.json 2> gpurun_out/bench_${TAG}_$name.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_${TAG}_$name.json')); print('$name', 'ms_per_step', d['ms_per_step'], 'jac_ms', d['jacobian_ms'], 'frac', d['roofline']['frac'], d['clocks'])" || tail -3 gpurun_out/bench_${TAG}_$name.err
}
run default KL_NOP=1
i=0
for kv in "$@"; do i=$((i+1)); run v$i $kv; done
