#!/usr/bin/env python
"""Device CG on the BASELINE 1M-DOF roof: SpMV time / bandwidth and time per iteration (not the headline bench)."""
import json, sys, time
import numpy as np
sys.path.insert(0, ".")
from gsstructuralanalysis_b200 import workloads as W
from gsstructuralanalysis_b200.ops import ShellAssembler

nel = int(sys.argv[1]) if len(sys.argv) > 1 else 576
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 400
prob = W.roof(nel=nel)
asm = ShellAssembler(prob)
n = asm.n_dofs
ok, _ = asm.jacobian(np.zeros(n), fetch=False)
assert ok
f = asm.force()
v = np.random.default_rng(0).standard_normal(n)
for _ in range(3):
    asm.spmv(v)
samples = []
for _ in range(8):
    asm.spmv(v)
    samples.append(round(asm.cg_last_timing()["spmv_ms"], 4))
spmv = min(samples)
x, it, err = asm.cg_solve(f, tol=1e-30, max_iter=iters)
t = asm.cg_last_timing()
bytes_alg = 8 * asm.nnz + 3 * 8 * n      # values + x, y, dot operand; regular columns do not read the row-index array
print(json.dumps({"n_dofs": n, "nnz": asm.nnz, "spmv_ms": spmv, "spmv_GBps": bytes_alg / spmv / 1e6, "spmv_samples_ms": samples, "cg_iters": it,
                  "cg_iter_ms": t["iter_ms"], "cg_total_ms": t["total_ms"], "rel_err": err}))
