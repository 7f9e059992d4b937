"""GPU parity of the stress / stretch recovery through the C ABI: kl_eval_stress, kl_principal_stretches,
kl_boundary_force against the oracle (1e-12 of the largest entry: FP64 re-association only), plus the reference's own
uniaxial-tension test of these calls solved and post-processed on the device
(unittests/gsStaticSolver_test.cpp:313-324,355-385)."""
import numpy as np
import pytest

from gsstructuralanalysis_b200 import geometry as G
from gsstructuralanalysis_b200 import capi
from gsstructuralanalysis_b200.problem import (ShellProblem, BoundaryConditions, KL_MAT_SVK, KL_MAT_NH, KL_MAT_MR,
                                                KL_BC_DIRICHLET, KL_BC_CLAMPED, WEST, EAST, SOUTH, NORTH, SW, SE, NW, NE)
from tests import kat_problems as kp

pytestmark = pytest.mark.gpu


def _problems():
    out = []
    bc = BoundaryConditions()
    for c in (SW, SE, NW, NE):
        bc.add_corner_value(c)
    sph = G.eighth_sphere(1.0).degree_elevate(1).uniform_refine(2)
    free = BoundaryConditions()          # no fixed corner: the coincident control points of the pole must move together
    out.append(("sphere_svk", ShellProblem(sph, free, material=KL_MAT_SVK, E=1.0, nu=0.3, thickness=0.05), 5e-2))
    out.append(("sphere_nh", ShellProblem(sph, free, material=KL_MAT_NH, E=1.0, nu=0.5, thickness=0.05), 5e-2))
    out.append(("sphere_mr_comp_z2", ShellProblem(sph, free, material=KL_MAT_MR, compressible=True, metric_z2=True, E=1.0, nu=0.3,
                                                 thickness=0.05, mr_ratio=7.0), 5e-2))
    par = G.paraboloid(0.2).degree_elevate(2).uniform_refine(2)           # degree 4
    bc2 = BoundaryConditions()
    bc2.add_condition(WEST, KL_BC_DIRICHLET, 0).add_condition(WEST, KL_BC_DIRICHLET, 1).add_condition(WEST, KL_BC_DIRICHLET, 2)
    bc2.add_condition(WEST, KL_BC_CLAMPED, 2)
    out.append(("paraboloid_p4_nh_comp", ShellProblem(par, bc2, material=KL_MAT_NH, compressible=True, E=2.0, nu=0.4, thickness=0.02), 2e-2))
    pl = G.plate(2.0, 1.0).degree_elevate(1).uniform_refine(3)           # degree 2, membrane only
    out.append(("plate_p2_membrane", ShellProblem(pl, bc, material=KL_MAT_NH, E=1.0, nu=0.5, thickness=0.01, bending=False), 2e-2))
    return out


def _state(pr, n_dofs, amp, rng):
    """Smooth displacement field sampled at the control points (coincident control points of a collapsed edge move together,
    so the pole of the eighth sphere stays a pole) plus a little noise."""
    cp = pr.surface.cp
    ncp = len(cp)
    u = amp * (cp * np.array([1.0, 0.6, 1.4]) + 0.5 * np.sin(2.0 * np.roll(cp, 1, axis=1)))
    x = np.zeros(n_dofs)
    for c in range(3):
        g = pr.dof_map[c * ncp:(c + 1) * ncp]
        x[g[g < n_dofs]] = u[g < n_dofs, c]
    if pr.surface.w is None:
        x += 0.05 * amp * rng.uniform(-1, 1, n_dofs)
    return x


@pytest.mark.parametrize("name,pr,amp", _problems(), ids=[p[0] for p in _problems()])
def test_eval_stress_matches_oracle(name, pr, amp):
    from gsstructuralanalysis_b200.ops import ShellAssembler
    from oracle.binding import Oracle
    if pr.dof_map is None:
        pr.number_dofs(capi.lib().kl_build_dofmap)
    asm, orc = ShellAssembler(pr, device=0), Oracle(pr)
    rng = np.random.default_rng(11)
    x = _state(pr, asm.n_dofs, amp, rng)
    uv = np.concatenate([rng.uniform(0.03, 0.92, (61, 2)), [[0.0, 0.0], [0.5, 0.0], [0.25, 0.5], [0.0, 0.9]]])   # incl. knots / boundary
    for tname, t in capi.STRESS_TYPES.items():
        for z in ((0.0, 0.3 * pr.thickness) if "stretch" in tname else (0.0,)):
            g, o = asm.eval_stress(x, t, uv, z), orc.eval_stress(x, t, uv, z)
            assert g.shape == o.shape == (len(uv), capi.lib().kl_stress_dim(t))
            if tname == "principal_stretch_dir":
                g3, o3 = g.reshape(-1, 3, 3), o.reshape(-1, 3, 3)
                sgn = np.sign(np.einsum("kic,kic->ki", g3, o3))          # eigenvector sign is arbitrary
                assert np.abs(g3 - sgn[..., None] * o3).max() < 1e-9, (name, tname, z)
            elif tname == "tension_field":
                assert np.array_equal(g, o), (name, tname)
            else:
                scale = max(np.abs(o).max(), 1e-300)
                assert np.abs(g - o).max() <= 1e-11 * scale, (name, tname, z, np.abs(g - o).max() / scale)
    assert np.abs(asm.computePrincipalStretches(uv, x, 0.0) - orc.computePrincipalStretches(uv, x, 0.0)).max() < 1e-12
    for side in (WEST, EAST, SOUTH, NORTH):
        fg, fo = asm.boundaryForce(x, side), orc.boundaryForce(x, side)
        ref = np.abs(orc.residual(x) - orc.force()).max()
        assert np.abs(fg - fo).max() <= 1e-11 * max(np.abs(fo).max(), ref), (name, side, fg, fo)
    # the recovery calls do not disturb the assembly state of the context
    ok, r = asm.residual(x)
    assert ok and np.abs(r - orc.residual(x)).max() <= 1e-12 * max(np.abs(orc.residual(x)).max(), 1e-300)


def test_eval_stress_argument_errors():
    from gsstructuralanalysis_b200.ops import ShellAssembler
    pr = _problems()[0][1]
    asm = ShellAssembler(pr, device=0)
    x = np.zeros(asm.n_dofs)
    with pytest.raises(capi.KLError):
        asm.eval_stress(x, 99, [[0.5, 0.5]])
    with pytest.raises(capi.KLError):
        asm.eval_stress(x, "membrane", [[1.5, 0.5]])            # outside the parametric domain
    with pytest.raises(capi.KLError):
        asm.boundaryForce(x, 7)
    assert asm.eval_stress(x, "membrane", np.zeros((0, 2))).shape == (0, 3)
    # undeformed state: unit stretches, zero stress
    uv = [[0.3, 0.4], [0.8, 0.1]]
    assert np.abs(asm.computePrincipalStretches(uv, x) - 1.0).max() < 1e-13
    assert np.abs(asm.eval_stress(x, "membrane", uv)).max() < 1e-13


@pytest.mark.parametrize("material,compressible", [(KL_MAT_NH, False), (KL_MAT_MR, False), (KL_MAT_NH, True), (KL_MAT_MR, True)])
def test_gpu_uat_numerical_like_the_reference(material, compressible):
    """UAT_numerical on the device path: Newton with the GPU closures, then computePrincipalStretches(pt = (1,0)) and
    boundaryForce(east) exactly as unittests/gsStaticSolver_test.cpp:313-324 uses them."""
    from gsstructuralanalysis_b200.ops import ShellAssembler
    pr, _ = kp.uat_problem(material, compressible, capi.lib().kl_build_dofmap)
    asm, x = kp.newton(lambda p: ShellAssembler(p, device=0), pr, load_steps=np.linspace(0.25, 1.0, 4), scale_fixed=1.0)
    lambdas = asm.computePrincipalStretches([[1.0, 0.0]], x, 0.0)[0]
    side_force = asm.boundaryForce(x, EAST).sum()
    S = -side_force / (pr.thickness * lambdas[0] * lambdas[2])
    L = lambdas[0]
    J = kp.UAT_J[(material, compressible)]
    assert abs(L - np.sqrt(J / 2.0)) / np.sqrt(J / 2.0) < 1e-7          # the reference's tolerance (:415)
    San = kp.uat_analytical_cauchy_stress(material, compressible)
    assert abs(S - San) / San < 1e-6, (S, San)
    sig = asm.eval_stress(x, "membrane", [[0.5, 0.5]])[0]
    assert abs(sig[0] - San) / San < 1e-6 and abs(sig[1]) / San < 1e-6 and abs(sig[2]) / San < 1e-6


def test_gpu_homogeneous_states_known_answers():
    """Hand-computed answers on the device (same cases as tests/test_oracle_stress.py): rigid motion, rotated biaxial stretch of
    an incompressible neo-Hookean sheet (sigma_a = mu (lambda_a^2 - lambda_3^2)), taut / wrinkled / slack states."""
    from gsstructuralanalysis_b200.ops import ShellAssembler
    from tests.test_oracle_stress import _plate_problem, _affine_state
    pr = _plate_problem()
    asm = ShellAssembler(pr, device=0)
    uv = np.array([[0.2, 0.3], [0.77, 0.5]])
    a = 0.7
    Rz = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1.0]])
    Rx = np.array([[1.0, 0, 0], [0, np.cos(0.4), -np.sin(0.4)], [0, np.sin(0.4), np.cos(0.4)]])
    x = _affine_state(pr, Rx @ Rz, (0.3, -0.2, 1.0))
    assert np.abs(asm.computePrincipalStretches(uv, x) - 1.0).max() < 1e-13
    for t in ("membrane", "flexural", "membrane_strain", "flexural_strain", "membrane_force", "flexural_moment"):
        assert np.abs(asm.eval_stress(x, t, uv)).max() < 1e-12, t
    assert np.abs(asm.boundaryForce(x, EAST)).max() < 1e-13
    x = _affine_state(pr, Rx @ Rz @ np.diag([1.2, 1.1, 1.0]))
    assert np.abs(asm.computePrincipalStretches(uv, x) - np.array([1.1, 1.2, 1.0 / 1.32])).max() < 1e-13
    mu = pr.E / 3.0
    sig = asm.eval_stress(x, "membrane", uv)
    assert np.abs(sig - np.array([mu * (1.44 - 1 / 1.32 ** 2), mu * (1.21 - 1 / 1.32 ** 2), 0.0])).max() < 1e-12 * mu
    d = asm.eval_stress(x, "principal_stretch_dir", uv).reshape(-1, 3, 3)
    R = Rx @ Rz
    for k in range(len(uv)):
        assert abs(abs(d[k, 0] @ R[:, 1]) - 1) < 1e-12 and abs(abs(d[k, 1] @ R[:, 0]) - 1) < 1e-12
        assert np.abs(d[k, 2] - R[:, 2]).max() < 1e-12
    assert np.all(asm.eval_stress(x, "tension_field", uv) == 1.0)
    assert np.all(asm.eval_stress(_affine_state(pr, np.diag([1.2, 0.7, 1.0])), "tension_field", uv) == 0.0)
    assert np.all(asm.eval_stress(_affine_state(pr, np.diag([0.9, 0.8, 1.0])), "tension_field", uv) == -1.0)
    x = _affine_state(pr, np.diag([1.3, 1.0, 1.0]))
    fe = asm.boundaryForce(x, EAST)
    sig, area = mu * (1.69 - 1 / 1.69), pr.thickness / 1.3
    assert abs(-fe[0] - sig * area) < 1e-11 * sig * area


def test_paraview_sized_sampling():
    """A 1000 x 1000 sampling grid (what gsWriteParaview(field, name, 1000) asks for, benchmarks/benchmark_Balloon.cpp:388) is one
    kernel launch: spans and basis functions are evaluated on the device.  Checked against the oracle at a random subset."""
    import time
    from gsstructuralanalysis_b200.ops import ShellAssembler
    from oracle.binding import Oracle
    name, pr, amp = _problems()[1]
    if pr.dof_map is None:
        pr.number_dofs(capi.lib().kl_build_dofmap)
    asm, orc = ShellAssembler(pr, device=0), Oracle(pr)
    rng = np.random.default_rng(2)
    x = _state(pr, asm.n_dofs, amp, rng)
    g = np.linspace(0.0, 0.97, 1000)
    uv = np.stack(np.meshgrid(g, g, indexing="ij"), -1).reshape(-1, 2)
    t0 = time.perf_counter()
    sig = asm.eval_stress(x, "membrane", uv)
    lam = asm.eval_stress(x, "principal_stretch", uv)
    dt = time.perf_counter() - t0
    assert sig.shape == (1000000, 3) and np.isfinite(sig).all() and np.isfinite(lam).all()
    pick = rng.choice(len(uv), 64, replace=False)
    so, lo = orc.eval_stress(x, "membrane", uv[pick]), orc.eval_stress(x, "principal_stretch", uv[pick])
    assert np.abs(sig[pick] - so).max() <= 1e-11 * np.abs(so).max()
    assert np.abs(lam[pick] - lo).max() <= 1e-12
    print(f"2 x 1e6 evaluation points in {dt * 1e3:.1f} ms (host call, incl. PCIe)")
