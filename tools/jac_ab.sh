#!/bin/bash
# A/B of the Jacobian kernel variants: parity first (one variant), then timing
cd /root/repo
B=gsstructuralanalysis_b200
echo "== parity of k_jacobian_cd (e1_m3)"
KL_LIB=$PWD/$B/libkl_e1_m3.so KL_JAC=1 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -4
KL_LIB=$PWD/$B/libkl_e2_m2.so KL_JAC=1 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -2
for v in e1_m2 e1_m3 e1_m4 e2_m1 e2_m2; do
  for mat in svk nh; do
    echo "== $v $mat"
    KL_LIB=$PWD/$B/libkl_$v.so KL_JAC=1 python bench.py --no-cpu-baseline --no-e2e --material $mat --steps 10 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms_per_step', d['ms_per_step'], 'jacobian_ms', d.get('jacobian_ms'))"
  done
done
echo "== baseline k_jacobian"
python bench.py --no-cpu-baseline --no-e2e --steps 10 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms_per_step', d['ms_per_step'], 'jacobian_ms', d.get('jacobian_ms'))"
