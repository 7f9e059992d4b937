#!/bin/bash
# usage: r3_solid.sh TAG [NEL] — solid bench, phase ablation (KS_ABLATE bits: 1 scatter, 2 W/acc, 4 Z/U), one full ncu capture of the window kernel
TAG=$1; NEL=${2:-67}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for ab in 0 1 2 4 7; do
  KS_ABLATE=$ab timeout 300 python tools/solid_bench.py $NEL 2 3 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('ablate $ab', 'assembly_ms', round(d['ms_per_assembly'],2), d['kernels_ms'])"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k3_jacobian_sw -s 1 -c 1 -o gpurun_out/$TAG -f python tools/solid_bench.py $NEL 2 1 > gpurun_out/prof_$TAG.log 2>&1
tail -2 gpurun_out/prof_$TAG.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_$TAG.csv python tools/solid_bench.py $NEL 2 1 > /dev/null 2>&1
