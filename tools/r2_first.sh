#!/bin/bash
# round 2: first GPU contact of the sliding-window Jacobian kernel: parity, then A/B bench against the shared-memory kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -5
run() {
  local name=$1; shift
  env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_r2a_$name.json 2> gpurun_out/bench_r2a_$name.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_r2a_$name.json')); print('$name', 'ms_per_step', d['ms_per_step'], 'jac_ms', d['jacobian_ms'], 'frac', d['roofline']['frac'], d['clocks'])" || tail -3 gpurun_out/bench_r2a_$name.err
}
run sw KL_NOP=1
run shared KL_JAC_SHARED=1
run seg16 KL_SW_SEG=16
run seg64 KL_SW_SEG=64
run seg144 KL_SW_SEG=144
