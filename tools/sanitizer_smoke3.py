"""compute-sanitizer workload for the stress / stretch recovery kernels (k_eval_stress, k_residual<P,FULL>, k_side_sum)."""
import numpy as np, sys
sys.path.insert(0, '.')
from gsstructuralanalysis_b200 import workloads as W, capi
from gsstructuralanalysis_b200.ops import ShellAssembler
from gsstructuralanalysis_b200.problem import KL_MAT_SVK, KL_MAT_NH, KL_MAT_MR
for degree, mat, comp in ((3, KL_MAT_NH, True), (2, KL_MAT_SVK, False), (4, KL_MAT_MR, False)):
    a = ShellAssembler(W.tutorial_paraboloid(3, degree, mat, comp))
    x = 1e-3 * np.random.default_rng(3).uniform(-1, 1, a.n_dofs)
    uv = np.random.default_rng(4).uniform(0, 1, (37, 2))
    uv[0] = (0.0, 0.0); uv[1] = (1.0, 1.0)
    tot = 0.0
    for name, t in capi.STRESS_TYPES.items():
        tot += float(np.abs(a.eval_stress(x, t, uv, 0.001)).sum())
    f = [a.boundaryForce(x, s) for s in range(4)]
    print("stress", degree, mat, comp, tot, np.abs(np.array(f)).max())
    a.close()
