"""Stress / stretch recovery of the oracle (definitions in include/kl_shell.h) pinned by
   (1) the reference's own test of these calls: computePrincipalStretches + boundaryForce on the uniaxial-tension sheet
       (unittests/gsStaticSolver_test.cpp:313-324, analytical values :355-385),
   (2) an independent numpy evaluation (scipy-free basis matrices, dense eigen-solvers) on a curved NURBS shell,
   (3) rigid-body motions and homogeneous in-plane deformations with hand-computed answers."""
import numpy as np
import pytest

from gsstructuralanalysis_b200 import geometry as G
from gsstructuralanalysis_b200.problem import (ShellProblem, BoundaryConditions, KL_MAT_SVK, KL_MAT_NH, KL_MAT_MR, EAST,
                                                SW, SE, NW, NE)
from oracle.binding import Oracle, OracleOps, lib as olib
from tests import kat_problems as kp
from tests import stress_model as sm


@pytest.mark.parametrize("material,compressible", [(KL_MAT_NH, False), (KL_MAT_MR, False), (KL_MAT_NH, True), (KL_MAT_MR, True)])
def test_uat_numerical_like_the_reference(material, compressible):
    """UAT_numerical of the reference, line by line: lambdas = computePrincipalStretches(pt, mp_def, 0);
    sideForce = boundaryForce(mp_def, east).sum(); S = -sideForce / (thickness lambdas(0) lambdas(2)); L = lambdas(0)."""
    pr, _ = kp.uat_problem(material, compressible, olib().klo_build_dofmap)
    asm, x = kp.newton(lambda p: OracleOps(p), pr, load_steps=np.linspace(0.25, 1.0, 4), scale_fixed=1.0)
    lambdas = asm.o.computePrincipalStretches([[1.0, 0.0]], x, 0.0)[0]
    side_force = asm.o.boundaryForce(x, EAST).sum()
    S = -side_force / (pr.thickness * lambdas[0] * lambdas[2])
    J = kp.UAT_J[(material, compressible)]
    # UAT_analytical: L = sqrt(J / lambda) (tolerance of the reference: 1e-7), San
    assert abs(lambdas[0] - np.sqrt(J / 2.0)) / np.sqrt(J / 2.0) < 1e-7
    assert abs(lambdas[1] - 2.0) < 1e-9                      # the imposed stretch
    assert abs(lambdas[2] - np.sqrt(J / 2.0)) / np.sqrt(J / 2.0) < 1e-7   # isotropy: thickness stretch = lateral stretch
    San = kp.uat_analytical_cauchy_stress(material, compressible)
    assert abs(S - San) / San < 1e-6, (S, San)
    # the recovered Cauchy membrane stress field is homogeneous and equals the same known answer
    uv = np.array([[0.1, 0.2], [0.5, 0.5], [0.93, 0.71]])
    sig = asm.o.eval_stress(x, "membrane", uv)
    assert np.abs(sig[:, 0] - San).max() / San < 1e-6
    assert np.abs(sig[:, 1:]).max() / San < 1e-6            # lateral and shear stress vanish
    assert np.abs(asm.o.eval_stress(x, "von_mises_membrane", uv)[:, 0] - San).max() / San < 1e-6
    ps = asm.o.eval_stress(x, "principal_stress_membrane", uv)
    assert np.abs(ps[:, 1] - San).max() / San < 1e-6 and np.abs(ps[:, 0]).max() / San < 1e-6


def _curved_problem(mat, comp, z2=False, bending=True):
    s = G.eighth_sphere(1.0).degree_elevate(1).uniform_refine(1)
    bc = BoundaryConditions()
    for c in (SW, SE, NW, NE):
        bc.add_corner_value(c)
    return ShellProblem(s, bc, material=mat, compressible=comp, metric_z2=z2, E=1.0, nu=0.3, thickness=0.05, bending=bending)


CASES = [("svk", KL_MAT_SVK, False, False), ("nh_inc", KL_MAT_NH, False, False), ("mr_comp", KL_MAT_MR, True, False),
         ("nh_comp_z2", KL_MAT_NH, True, True)]


@pytest.mark.parametrize("name,mat,comp,z2", CASES)
def test_against_independent_numpy_model(name, mat, comp, z2):
    """Every kinematic output against tests/stress_model.py (3-D tensors from numpy basis matrices + numpy.linalg.eigh);
    the stress outputs through sigma = F S F^T / J with S taken from MEMBRANE_FORCE / FLEXURAL_MOMENT."""
    pr = _curved_problem(mat, comp, z2)
    o = Oracle(pr)
    rng = np.random.default_rng(5)
    x = 1e-2 * rng.uniform(-1, 1, o.n_dofs)
    uv = rng.uniform(0.05, 0.9, (7, 2))       # away from the collapsed pole of the eighth sphere
    for z in (0.0, 0.02):
        ref = sm.kinematics(pr, x, uv, z)
        lam = o.eval_stress(x, "principal_stretch", uv, z)
        assert np.abs(lam[:, :2] - ref["stretch"]).max() < 1e-11
        dirs = o.eval_stress(x, "principal_stretch_dir", uv, z).reshape(-1, 3, 3)
        for k in range(len(uv)):
            for i in range(2):
                assert abs(abs(dirs[k, i] @ ref["dirs"][k, i]) - 1.0) < 1e-9       # same line, sign free
            assert np.abs(dirs[k, 2] - ref["normal"][k]).max() < 1e-12
    ref = sm.kinematics(pr, x, uv, 0.0)
    assert np.abs(o.eval_stress(x, "displacement", uv) - ref["disp"]).max() < 1e-13
    Em = o.eval_stress(x, "membrane_strain", uv)
    Ef = o.eval_stress(x, "flexural_strain", uv)
    assert np.abs(Em - ref["Em"]).max() < 1e-12 and np.abs(Ef - ref["Ef"]).max() < 1e-11
    assert np.abs(o.eval_stress(x, "principal_membrane_strain", uv) - ref["Em_p"]).max() < 1e-12
    assert np.abs(o.eval_stress(x, "principal_flexural_strain", uv) - ref["Ef_p"]).max() < 1e-11
    # thickness stretch: 1/J0 unless the law is compressible (then sqrt(C33) with S33(C33) = 0, checked by the UAT test)
    lam = o.eval_stress(x, "principal_stretch", uv, 0.0)
    if not comp:
        assert np.abs(lam[:, 2] * lam[:, 0] * lam[:, 1] - 1.0).max() < 1e-12
    # Cauchy stresses from the resultants
    N = o.eval_stress(x, "membrane_force", uv)
    M = o.eval_stress(x, "flexural_moment", uv)
    sig_m, sig_f = sm.cauchy_from_resultants(pr, x, uv, N, M, lam[:, 2])
    got_m, got_f = o.eval_stress(x, "membrane", uv), o.eval_stress(x, "flexural", uv)
    assert np.abs(got_m - sig_m).max() <= 1e-12 * np.abs(sig_m).max()
    assert np.abs(got_f - sig_f).max() <= 1e-12 * max(np.abs(sig_f).max(), 1e-300)
    pm = o.eval_stress(x, "principal_stress_membrane", uv)
    for k in range(len(uv)):
        w = np.linalg.eigvalsh(np.array([[got_m[k, 0], got_m[k, 2]], [got_m[k, 2], got_m[k, 1]]]))
        assert np.abs(pm[k] - w).max() <= 1e-12 * np.abs(got_m).max()


def test_membrane_force_is_what_the_assembly_integrates():
    """MEMBRANE_FORCE / FLEXURAL_MOMENT are the N, M of gsMaterialMatrixIntegrate: for the linear law N = t C : E, M = t^3/12 C : K
    with C^abcd = lam_ps A^ab A^cd + mu (A^ac A^bd + A^ad A^bc)."""
    pr = _curved_problem(KL_MAT_SVK, False)
    o = Oracle(pr)
    rng = np.random.default_rng(9)
    x = 2e-2 * rng.uniform(-1, 1, o.n_dofs)
    uv = rng.uniform(0.05, 0.95, (5, 2))
    N, M = o.eval_stress(x, "membrane_force", uv), o.eval_stress(x, "flexural_moment", uv)
    Nr, Mr = sm.svk_resultants(pr, x, uv)
    assert np.abs(N - Nr).max() <= 1e-12 * np.abs(Nr).max()
    assert np.abs(M - Mr).max() <= 1e-12 * np.abs(Mr).max()


def _plate_problem(mat=KL_MAT_NH, comp=False):
    s = G.plate(1.0, 1.0).degree_elevate(2).uniform_refine(1)
    bc = BoundaryConditions()
    pr = ShellProblem(s, bc, material=mat, compressible=comp, E=3.0, nu=0.5 if not comp else 0.3, thickness=0.01)
    pr.number_dofs(olib().klo_build_dofmap)
    return pr


def _affine_state(pr, Fm, shift=(0.0, 0.0, 0.0)):
    """DoF vector of the affine map X -> Fm X + shift (the spline space reproduces it exactly)."""
    cp = pr.surface.cp
    u = cp @ np.asarray(Fm).T + np.asarray(shift) - cp
    ncp = len(cp)
    x = np.zeros(pr.n_free)
    for c in range(3):
        g = pr.dof_map[c * ncp:(c + 1) * ncp]
        x[g[g < pr.n_free]] = u[g < pr.n_free, c]
    return x


def test_rigid_body_motion_and_homogeneous_states():
    pr = _plate_problem()
    o = Oracle(pr)
    uv = np.array([[0.2, 0.3], [0.77, 0.5]])
    a = 0.7
    Rz = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1.0]])
    Rx = np.array([[1.0, 0, 0], [0, np.cos(0.4), -np.sin(0.4)], [0, np.sin(0.4), np.cos(0.4)]])
    x = _affine_state(pr, Rx @ Rz, (0.3, -0.2, 1.0))
    assert np.abs(o.eval_stress(x, "principal_stretch", uv) - 1.0).max() < 1e-13
    for t in ("membrane", "flexural", "membrane_strain", "flexural_strain", "membrane_force", "flexural_moment",
              "von_mises_membrane"):
        assert np.abs(o.eval_stress(x, t, uv)).max() < 1e-12, t
    assert np.abs(o.boundaryForce(x, EAST)).max() < 1e-13
    # biaxial stretch 1.2 x 1.1, rotated: stretches sorted, thickness stretch 1/(1.2*1.1), taut
    x = _affine_state(pr, Rx @ Rz @ np.diag([1.2, 1.1, 1.0]))
    lam = o.eval_stress(x, "principal_stretch", uv)
    assert np.abs(lam - np.array([1.1, 1.2, 1.0 / 1.32])).max() < 1e-13
    d = o.eval_stress(x, "principal_stretch_dir", uv).reshape(-1, 3, 3)
    R = Rx @ Rz
    for k in range(len(uv)):
        assert abs(abs(d[k, 0] @ R[:, 1]) - 1) < 1e-12 and abs(abs(d[k, 1] @ R[:, 0]) - 1) < 1e-12
        assert np.abs(d[k, 2] - R[:, 2]).max() < 1e-12
    assert np.all(o.eval_stress(x, "tension_field", uv) == 1.0)
    E = o.eval_stress(x, "membrane_strain", uv)
    assert np.abs(E - np.array([0.5 * (1.44 - 1), 0.5 * (1.21 - 1), 0.0])).max() < 1e-13
    # incompressible neo-Hooke, plane stress: sigma_a = mu (lambda_a^2 - lambda_3^2) in the deformed frame e1 = R e_x
    mu = pr.E / 3.0
    sig = o.eval_stress(x, "membrane", uv)
    assert np.abs(sig - np.array([mu * (1.44 - 1 / 1.32 ** 2), mu * (1.21 - 1 / 1.32 ** 2), 0.0])).max() < 1e-12 * mu
    # stretched in x, compressed in y: wrinkled (0); compressed in both: slack (-1)
    assert np.all(o.eval_stress(_affine_state(pr, np.diag([1.2, 0.7, 1.0])), "tension_field", uv) == 0.0)
    assert np.all(o.eval_stress(_affine_state(pr, np.diag([0.9, 0.8, 1.0])), "tension_field", uv) == -1.0)


def test_boundary_force_balances():
    """Internal forces are self-equilibrated: the four side sums minus the double-counted corners add up to minus the
    sum over interior control points; with all sides summed over a free sheet under a homogeneous stretch the total
    vanishes, and the east side of a sheet stretched in x carries the Cauchy stress times the deformed area."""
    pr = _plate_problem()
    o = Oracle(pr)
    x = _affine_state(pr, np.diag([1.3, 1.0, 1.0]))
    fe, fw = o.boundaryForce(x, EAST), o.boundaryForce(x, 0)
    assert np.abs(fe + fw).max() < 1e-12 * np.abs(fe).max()
    mu = pr.E / 3.0
    sig = mu * (1.69 - 1 / 1.69)                     # lateral stretch held at 1 => lambda3 = 1/1.3
    area = 1.0 * pr.thickness / 1.3
    assert abs(-fe[0] - sig * area) < 1e-11 * sig * area
