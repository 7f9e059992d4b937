"""compute-sanitizer workload for the kernels of round 2, session 3: k_jacobian_sw<false,true> (follower-pressure tangent in the window;
balloon with its collapsed pole = irregular columns, and a regular cylinder-like patch with pressure), k_points with 4 elements per CTA,
k3_jacobian_sw<true/false> (tri-cubic solid window kernel: TMA slabs behind mbarriers, double-buffered U, segments of 1/4 elements,
constant and shared-memory tables, symmetric and full mode) and the arithmetic k3_mirror; multi-patch assembly on forked streams."""
import os, sys
import numpy as np
sys.path.insert(0, '.')
import torch
from gsstructuralanalysis_b200 import workloads as W, solid as S
from gsstructuralanalysis_b200.ops import ShellAssembler

s = torch.cuda.current_stream().cuda_stream
for mk, n in ((W.balloon, 6), (W.balloon, 11)):
    pr = mk(n)
    a = ShellAssembler(pr)
    L = max(np.ptp(pr.surface.cp[:, k]) for k in range(3))
    x = W.smooth_state(pr, 1e-4 * L)
    ok, K = a.jacobian(x)
    ok2, r = a.residual(x)
    print("balloon", n, ok, ok2, float(np.abs(K.values).max()), float(np.abs(r).max()))
    a.close()
for seg, full, noconst in ((None, False, False), ("4", False, True), ("1", True, False)):
    for k, v in (("KS_SW_SEG", seg), ("KS_FULL", "1" if full else None), ("KS_NO_CONST", "1" if noconst else None)):
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v
    vol = S.brick(2.0, 1.0, 0.5, degrees=(3, 3, 3), nels=(9, 2, 3))
    sp = S.SolidProblem(vol, S.SolidBC().add_condition(S.KS_WEST).add_condition(S.KS_EAST, 1), law=S.KS_LAW_NEO_HOOKE_LN, E=7.0, nu=0.3,
                        tractions=[(S.KS_NORTH, (0.0, -0.3, 0.1))])
    a = S.SolidAssembler(sp)
    xs = 1e-3 * np.random.default_rng(1).standard_normal(a.n_dofs)
    ok, K, r = a.assemble(xs)
    print("solid", seg, full, noconst, ok, float(np.abs(K.values).max()), float(np.abs(r).max()))
    a.close()
# multi-patch on forked streams
single, multi, _ = W.cut(W.tension_sheet(8), [], [0.25, 0.5])
from gsstructuralanalysis_b200.ops import MultiPatchAssembler
m = MultiPatchAssembler(multi)
xm = 1e-6 * np.random.default_rng(2).standard_normal(m.n_dofs)
ok, Km = m.jacobian(xm)
ok2, rm = m.residual(xm)
print("multipatch", ok, ok2, float(np.abs(Km.values).max()))
m.close()
