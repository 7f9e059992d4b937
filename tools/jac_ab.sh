#!/bin/bash
# A/B of Jacobian kernel builds: tools/jac_ab.sh <lib suffix>... ; each is gsstructuralanalysis_b200/libkl_<suffix>.so
cd /root/repo
B=gsstructuralanalysis_b200
for v in "$@"; do
  if [ -n "$JAC_AB_PARITY" ]; then KL_LIB=$PWD/$B/libkl_$v.so python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -2; fi
  for mat in ${JAC_AB_MATS:-svk}; do
    echo "== $v $mat"
    KL_LIB=$PWD/$B/libkl_$v.so python bench.py --no-cpu-baseline --no-e2e --material $mat --steps 10 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms_per_step', round(d['ms_per_step'],3), 'jacobian_ms', round(d.get('jacobian_ms'),3))"
  done
done
