"""CPU checks that pin the solid oracle (oracle/ks_oracle.c, gsElasticity path) without gsElasticity:
   consistency of K with the residual, of the residual with an independently coded energy, and beam theory."""
import numpy as np
import pytest
import scipy.sparse.linalg as spla

from gsstructuralanalysis_b200 import solid as S
from oracle.binding_solid import SolidOracle

LAWS = [S.KS_LAW_HOOKE, S.KS_LAW_SVK, S.KS_LAW_NEO_HOOKE_LN, S.KS_LAW_NEO_HOOKE_QUAD]


def tutorial_problem(law, nels=(2, 2, 1), degrees=(3, 3, 2)):
    """tutorials/nonlinear_solid_static.cpp:70-97: thick paraboloid, front corners pinned, traction on a face."""
    v = S.paraboloid_volume(nels=nels, degrees=degrees)
    bc = S.SolidBC()
    for k in range(4):
        bc.add_corner_value(k)
    return S.SolidProblem(v, bc, law=law, E=1e3, nu=0.3, tractions=[(S.KS_BACK, (0.0, 0.0, -1.0))], body_force=(0.0, 0.1, -0.5))


@pytest.mark.parametrize("law", LAWS)
def test_tangent_is_derivative_of_residual_and_symmetric(law):
    o = SolidOracle(tutorial_problem(law))
    rng = np.random.default_rng(0)
    x = 2e-3 * rng.standard_normal(o.n_dofs)
    K = o.jacobian(x)
    assert abs(K - K.T).max() <= 1e-12 * abs(K).max()
    d = rng.standard_normal(o.n_dofs)
    eps = 1e-6
    fd = -(o.residual(x + eps * d) - o.residual(x - eps * d)) / (2 * eps)
    assert np.abs(K @ d - fd).max() <= 1e-7 * np.abs(fd).max()


@pytest.mark.parametrize("law", LAWS)
def test_internal_force_is_gradient_of_stored_energy(law):
    o = SolidOracle(tutorial_problem(law))
    rng = np.random.default_rng(1)
    x = 2e-3 * rng.standard_normal(o.n_dofs)
    fint = o.force() - o.residual(x)
    d = rng.standard_normal(o.n_dofs)
    eps = 1e-6
    g = (o.energy(x + eps * d) - o.energy(x - eps * d)) / (2 * eps)
    assert abs(g - fint @ d) <= 1e-6 * abs(g)


def test_energy_against_independent_numpy_model():
    """Stored energy of a homogeneous stretch computed by hand: x -> (1+a) x on the unit cube, F = (1+a) I."""
    v = S.brick(1.0, 1.0, 1.0, degrees=(2, 2, 2), nels=(2, 1, 1))
    pr = S.SolidProblem(v, S.SolidBC(), law=S.KS_LAW_NEO_HOOKE_LN, E=10.0, nu=0.25)
    o = SolidOracle(pr)
    a = 0.1
    x = np.zeros(o.n_dofs)
    ncp = v.cp.shape[0]
    for c in range(3):
        x[pr.dof_map[c * ncp:(c + 1) * ncp]] = a * v.cp[:, c]
    lam, mu = 10.0 * 0.25 / (1.25 * 0.5), 10.0 / 2.5
    J = (1 + a) ** 3
    psi = 0.5 * mu * (3 * (1 + a) ** 2 - 3) - mu * np.log(J) + 0.5 * lam * np.log(J) ** 2
    assert abs(o.energy(x) - psi) <= 1e-12 * psi
    # SvK on the same state
    pr2 = S.SolidProblem(v, S.SolidBC(), law=S.KS_LAW_SVK, E=10.0, nu=0.25)
    E = 0.5 * ((1 + a) ** 2 - 1)
    psi2 = 0.5 * lam * (3 * E) ** 2 + mu * 3 * E * E
    assert abs(SolidOracle(pr2).energy(x) - psi2) <= 1e-12 * psi2


def test_rigid_body_motion_and_patch_test():
    v = S.brick(2.0, 1.0, 0.5, degrees=(2, 2, 1), nels=(3, 2, 1))
    pr = S.SolidProblem(v, S.SolidBC(), law=S.KS_LAW_NEO_HOOKE_QUAD, E=5.0, nu=0.3)
    o = SolidOracle(pr)
    ncp = v.cp.shape[0]
    th = 0.3
    R = np.array([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1.0]])
    disp = v.cp @ R.T + np.array([0.1, -0.2, 0.3]) - v.cp
    x = np.zeros(o.n_dofs)
    for c in range(3):
        x[pr.dof_map[c * ncp:(c + 1) * ncp]] = disp[:, c]
    K = o.jacobian(np.zeros(o.n_dofs))
    assert np.abs(o.residual(x)).max() <= 1e-12 * abs(K).max()
    for c in range(3):
        t = np.zeros(o.n_dofs)
        t[pr.dof_map[c * ncp:(c + 1) * ncp]] = 1.0
        assert np.abs(K @ t).max() <= 1e-12 * abs(K).max()


def test_pattern_is_the_structural_stencil():
    v = S.brick(1, 1, 1, degrees=(2, 3, 1), nels=(4, 3, 2))
    pr = S.SolidProblem(v, S.SolidBC().add_condition(S.KS_WEST), law=S.KS_LAW_SVK)
    o = SolidOracle(pr)
    assert o.outer[0] == 0 and o.outer[-1] == o.nnz
    import scipy.sparse as sp
    # sorted columns, structurally symmetric, and no assembled value falls outside the stored pattern
    for col in range(0, o.n_dofs, 7):
        rows = o.inner[o.outer[col]:o.outer[col + 1]]
        assert np.all(np.diff(rows) > 0)
    P = sp.csc_matrix((np.ones(o.nnz), o.inner, o.outer), shape=(o.n_dofs, o.n_dofs))
    assert abs(P - P.T).sum() == 0
    # an interior control point couples with (2p+1) functions per direction, clipped at the boundary
    n1, n2, n3 = v.n
    J = 3 + n1 * (2 + n2 * 1)
    col = pr.dof_map[2 * n1 * n2 * n3 + J]
    expected = sum(int(pr.dof_map[c * n1 * n2 * n3 + i1 + n1 * (i2 + n2 * i3)] < o.n_dofs)
                   for c in range(3) for i1 in range(1, 6) for i2 in range(0, 6) for i3 in range(0, 3))
    assert o.outer[col + 1] - o.outer[col] == expected


def test_cantilever_tip_deflection_matches_beam_theory():
    """benchmarks/benchmark_Elasticity_Beam_APALM.cpp:196-236 (testCase 2): clamped beam L=1, B=H=0.01, E=1, nu=0 under a
    vertical end traction; linear solution against Euler-Bernoulli  w = P L^3 / (3 E I)."""
    L, B, H, E = 1.0, 0.01, 0.01, 1.0
    v = S.brick(L, B, H, degrees=(3, 2, 2), nels=(8, 1, 1))
    Pload = 1e-9
    pr = S.SolidProblem(v, S.SolidBC().add_condition(S.KS_WEST), law=S.KS_LAW_HOOKE, E=E, nu=0.0,
                        tractions=[(S.KS_EAST, (0.0, 0.0, Pload / (B * H)))])
    o = SolidOracle(pr)
    K = o.jacobian(np.zeros(o.n_dofs))
    f = o.force()
    assert abs(f.sum() - Pload) <= 1e-12 * Pload
    u = spla.spsolve(K.tocsc(), f)
    n1, n2, n3 = v.n
    ncp = n1 * n2 * n3
    tip = pr.dof_map[2 * ncp + (n1 - 1)]          # z-displacement of the corner control point at x = L
    w_beam = Pload * L ** 3 / (3 * E * B * H ** 3 / 12)
    assert abs(u[tip] - w_beam) <= 0.01 * w_beam


def test_inverted_state_is_reported():
    pr = tutorial_problem(S.KS_LAW_NEO_HOOKE_LN)
    o = SolidOracle(pr)
    with pytest.raises(RuntimeError):
        o.residual(np.full(o.n_dofs, 1.0) * np.random.default_rng(3).standard_normal(o.n_dofs) * 10.0)


def test_mass_matrix_sums_to_density_times_volume():
    """gsMassAssembler (tutorials/nonlinear_solid_dynamic.cpp:98-109): partition of unity => sum of one component block of M
    = density x volume; M is symmetric positive definite on the pattern of K."""
    import scipy.sparse as sp
    v = S.brick(2.0, 1.0, 0.5, degrees=(2, 3, 1), nels=(3, 2, 2))
    pr = S.SolidProblem(v, S.SolidBC(), law=S.KS_LAW_SVK)
    o = SolidOracle(pr)
    M = sp.csc_matrix((o.mass(7.0), o.inner, o.outer), shape=(o.n_dofs, o.n_dofs))
    assert abs(M.sum() - 3 * 7.0 * 2.0 * 1.0 * 0.5) <= 1e-12 * M.sum()
    assert abs(M - M.T).max() <= 1e-14 * abs(M).max()
    x = np.random.default_rng(0).standard_normal(o.n_dofs)
    assert x @ (M @ x) > 0
