import copy, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from gsstructuralanalysis_b200 import workloads as W, ops, capi
from gsstructuralanalysis_b200.problem import KL_MAT_SVK
from tests.mp_problems import cut
base = W.tutorial_paraboloid(nel=8, material=KL_MAT_SVK)
base.body_force = (0.0, 0.0, -1e3)
single, multi, cps = cut(base, [0.5], [0.5])
asm = ops.MultiPatchAssembler(multi)
one = ops.ShellAssembler(single)
for name, a in (("multi", asm), ("single", one)):
    for ls in (True, False):
        for tol in (0.0, 1e-12):
            U, info = a.newton_solve(tolU=1e-9, tolF=1e-9, max_it=30, linear_start=ls, cg_tol=tol)
            print(name, "linear_start", ls, "cg_tol", tol, info, capi.lib().kl_last_error().decode()[:200], flush=True)
b = asm.force()
ok, _ = asm.jacobian(np.zeros(asm.n_dofs), fetch=False)
for tol in (1e-10, 1e-14, 0.0):
    try:
        sol, it, err = asm.cg_solve(b, tol=tol)
        print("cg multi tol", tol, it, err)
    except Exception as e:
        print("cg multi tol", tol, "FAILED", e)
