import numpy as np, sys
sys.path.insert(0, '.')
from gsstructuralanalysis_b200 import workloads as W
from gsstructuralanalysis_b200.ops import ShellAssembler
from gsstructuralanalysis_b200.problem import KL_MAT_MR
for pr in (W.tutorial_paraboloid(5, 3, KL_MAT_MR, True), W.balloon(6), W.tension_sheet(5), W.tutorial_paraboloid(4, 2), W.tutorial_paraboloid(3, 4)):
    a = ShellAssembler(pr)
    x = W.displacement_state(a.n_dofs, 1e-4)
    ok, K = a.jacobian(x); ok2, r = a.residual(x); m = a.mass(1.0); l = a.mass(1.0, lumped=True)
    print(pr.surface.name, pr.surface.p, a.n_dofs, ok, ok2, float(np.abs(K.values).max()))
    a.close()
