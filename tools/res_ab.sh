#!/bin/bash
# A/B of residual-path builds: step time of kl_jacobian_device + kl_residual_device (separate calls) per library suffix
cd /root/repo
for v in "$@"; do
  KL_LIB=$PWD/gsstructuralanalysis_b200/libkl_$v.so python bench.py --no-cpu-baseline --no-e2e --separate-calls --steps 10 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', 'ms_per_step', round(d['ms_per_step'],3), 'jacobian_ms', round(d.get('jacobian_ms'),3))"
done
