#!/bin/bash
# usage: r3_prof.sh TAG WORKLOAD KERNEL_REGEX [SKIP] — timing line, then one full ncu capture of the first matching kernel after SKIP launches
TAG=$1; WL=$2; KR=$3; SKIP=${4:-2}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python tools/prof_shell.py $WL 576 5
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$KR -s $SKIP -c 1 -o gpurun_out/$TAG -f python tools/prof_shell.py $WL 576 1 > gpurun_out/prof_$TAG.log 2>&1
tail -2 gpurun_out/prof_$TAG.log
ls -la gpurun_out/$TAG.ncu-rep
