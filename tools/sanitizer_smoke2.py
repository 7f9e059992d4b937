"""compute-sanitizer workload for the widened rows: device CG / Newton and the solid kernels (small meshes)."""
import numpy as np, sys
sys.path.insert(0, '.')
from gsstructuralanalysis_b200 import workloads as W, solid as S
from gsstructuralanalysis_b200.ops import ShellAssembler
from gsstructuralanalysis_b200.problem import KL_MAT_SVK, KL_MAT_NH
a = ShellAssembler(W.tutorial_paraboloid(4, 3, KL_MAT_SVK))
ok, K = a.jacobian(np.zeros(a.n_dofs))
x, it, err = a.cg_solve(a.force(), tol=1e-10, max_iter=5000)
y = a.spmv(x)
print("cg", it, err, float(np.abs(y - a.force()).max()))
a.close()
pr = W.tutorial_paraboloid(3, 3, KL_MAT_NH)
pr.point_loads = [((0.5, 0.5), (0.0, 0.0, -2e3))]
a = ShellAssembler(pr)
U, info = a.newton_solve(tolU=1e-8, tolF=1e-8, cg_tol=1e-12, cg_max_iter=20000)
print("newton", info["status"], info["iterations"], info["cg_iterations"])
a.close()
for degrees, nels, law in (((3, 3, 3), (2, 2, 2), S.KS_LAW_NEO_HOOKE_LN), ((2, 2, 2), (3, 2, 1), S.KS_LAW_SVK), ((3, 2, 1), (2, 2, 3), S.KS_LAW_NEO_HOOKE_QUAD),
                           ((1, 1, 1), (3, 3, 2), S.KS_LAW_HOOKE), ((3, 3, 2), (2, 2, 1), S.KS_LAW_SVK)):
    sp = S.SolidProblem(S.brick(1.0, 0.5, 0.5, degrees=degrees, nels=nels), S.SolidBC().add_condition(S.KS_WEST).add_corner_value(7, 1),
                        law=law, E=3.0, nu=0.3, tractions=[(S.KS_EAST, (0.0, 0.0, 0.01))], body_force=(0.0, 0.0, -0.1))
    s = S.SolidAssembler(sp)
    xs = 1e-3 * np.random.default_rng(1).standard_normal(s.n_dofs)
    ok, Ks, rs = s.assemble(xs)
    print("solid", degrees, nels, s.n_dofs, ok, float(np.abs(Ks.values).max()))
    s.close()
