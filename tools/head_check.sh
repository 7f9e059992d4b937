#!/bin/bash
# usage: head_check.sh TAG — run on the GPU box: default bench line (both arms) of the current tree
TAG=$1
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 200 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
python -c "
import json; d=json.load(open('gpurun_out/bench_$TAG.json')); print(d['ms_per_step'], d['jacobian_ms'], d['e2e']['ms_per_step'], d['clocks'])"
tail -3 gpurun_out/bench_$TAG.err
timeout 150 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${TAG}_reference.json 2>/dev/null
cut -c1-250 gpurun_out/bench_${TAG}_reference.json
