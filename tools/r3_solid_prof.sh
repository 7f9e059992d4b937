#!/bin/bash
# usage: r3_solid_prof.sh TAG [NEL] — one full ncu capture of the tri-cubic window kernel
TAG=$1; NEL=${2:-67}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k3_jacobian_sw -s 1 -c 1 -o gpurun_out/$TAG -f python tools/solid_bench.py $NEL 2 1 > gpurun_out/prof_$TAG.log 2>&1
tail -2 gpurun_out/prof_$TAG.log
