#!/usr/bin/env python
"""gsElasticity solid path at ~1M DOF (tri-cubic, 192x192 local blocks): device-resident assembly time of
K and rhs in one pass, per-kernel times and algorithmic rates.  Secondary to bench.py (the KL-shell headline)."""
import json, sys, time
import numpy as np
sys.path.insert(0, ".")
import torch
from gsstructuralanalysis_b200 import solid as S

nel = int(sys.argv[1]) if len(sys.argv) > 1 else 67
law = int(sys.argv[2]) if len(sys.argv) > 2 else S.KS_LAW_NEO_HOOKE_LN
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
v = S.brick(1.0, 1.0, 1.0, degrees=(3, 3, 3), nels=(nel, nel, nel))
pr = S.SolidProblem(v, S.SolidBC().add_condition(S.KS_WEST), law=law, E=5.0, nu=0.3, tractions=[(S.KS_EAST, (0.0, 0.0, 0.01))])
t0 = time.perf_counter()
asm = S.SolidAssembler(pr)
torch.cuda.synchronize()
setup = time.perf_counter() - t0
n = asm.n_dofs
x = torch.from_numpy(1e-3 / nel * np.random.default_rng(20240607).uniform(-1, 1, n)).cuda()
r = torch.empty(n, dtype=torch.float64, device="cuda")
stream = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    asm.assemble_device(x.data_ptr(), r.data_ptr(), True, stream)
assert asm.check(stream) == 0
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record()
for _ in range(steps):
    asm.assemble_device(x.data_ptr(), r.data_ptr(), True, stream)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
# per-kernel times through the host entry (events inside the library)
ok, K, rr = asm.assemble(x.cpu().numpy())
t = asm.last_timing()
fma_per_elem = 8 * 4 * (128 * 81 + 3456 * 4 + 1152 * 12 + 1152 * 8)      # Z, U, W, acc per (column block, slab), tri-cubic
flops = 2.0 * fma_per_elem * asm.n_elements
print(json.dumps({"workload": f"unit cube, tri-cubic, {nel}^3 elements, law {law}", "n_dofs": n, "nnz": asm.nnz, "elements": asm.n_elements,
                  "quad_points": asm.n_qp, "setup_s": setup, "ms_per_assembly": ms, "quad_pts_per_s": asm.n_qp / (ms * 1e-3),
                  "kernels_ms": t, "jacobian_flops": flops, "jacobian_TFLOPs": flops / (t["jacobian_ms"] * 1e-3) / 1e12,
                  "values_GB": 8 * asm.nnz / 1e9, "records_GB": 8 * 100 * asm.n_qp / 1e9}))
