"""GPU parity tests of the gsElasticity solid path through the C ABI (include/ks_solid.h) against the
CPU oracle (oracle/ks_oracle.c): pattern bit-exact, K / rhs / F within 1e-12 of the largest entry."""
import numpy as np
import pytest

from gsstructuralanalysis_b200 import solid as S

pytestmark = pytest.mark.gpu
RTOL = 1e-12
LAWS = [S.KS_LAW_HOOKE, S.KS_LAW_SVK, S.KS_LAW_NEO_HOOKE_LN, S.KS_LAW_NEO_HOOKE_QUAD]


@pytest.fixture(scope="module")
def gpu():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("CUDA device required for -m gpu tests (no CPU fallback exists)")
    from gsstructuralanalysis_b200 import build as kbuild
    kbuild.build()
    return S.SolidAssembler


def tutorial_problem(law, nels=(3, 3, 1), degrees=(3, 3, 2)):
    v = S.paraboloid_volume(nels=nels, degrees=degrees)
    bc = S.SolidBC()
    for k in range(4):
        bc.add_corner_value(k)
    return S.SolidProblem(v, bc, law=law, E=1e3, nu=0.3, tractions=[(S.KS_BACK, (0.0, 0.0, -1.0))], body_force=(0.0, 0.1, -0.5))


def _compare(Asm, prob, scale, tag):
    from oracle.binding_solid import SolidOracle
    asm, orc = Asm(prob), SolidOracle(prob)
    assert asm.n_dofs == orc.n_dofs and asm.nnz == orc.nnz, tag
    outer, inner = asm.pattern()
    assert np.array_equal(outer, orc.outer) and np.array_equal(inner, orc.inner), tag     # bit-exact pattern
    f, fo = asm.force(), orc.force()
    assert np.abs(f - fo).max() <= RTOL * max(np.abs(fo).max(), 1e-300), tag
    rng = np.random.default_rng(11)
    for x in (np.zeros(asm.n_dofs), scale * rng.standard_normal(asm.n_dofs)):
        ok, K, r = asm.assemble(x)
        assert ok, (tag, getattr(asm, "last_error", ""))
        Ko, ro = orc.assemble(x)
        assert np.abs(K.values - Ko).max() <= RTOL * np.abs(Ko).max(), (tag, np.abs(K.values - Ko).max() / np.abs(Ko).max())
        assert np.abs(r - ro).max() <= RTOL * max(np.abs(ro).max(), np.abs(fo).max()), tag
        ok, K2 = asm.jacobian(x)
        ok2, r2 = asm.residual(x)
        assert ok and ok2 and np.abs(K2.values - Ko).max() <= RTOL * np.abs(Ko).max() and np.abs(r2 - ro).max() <= RTOL * max(np.abs(ro).max(), np.abs(fo).max())
        ok, ra = asm.al_residual(x, 0.37)
        assert ok and np.abs(ra - (fo - 0.37 * fo - ro)).max() <= RTOL * max(np.abs(ro).max(), np.abs(fo).max()), tag
    Mg, Mo = asm.mass(3.5), orc.mass(3.5)
    assert np.abs(Mg.values - Mo).max() <= RTOL * np.abs(Mo).max(), tag
    return asm


@pytest.mark.parametrize("law", LAWS)
def test_tutorial_paraboloid_volume(gpu, law):
    _compare(gpu, tutorial_problem(law), 2e-3, f"tutorial-law{law}")


@pytest.mark.parametrize("degrees,nels", [((1, 1, 1), (3, 2, 2)), ((2, 2, 2), (4, 2, 1)), ((3, 3, 3), (5, 5, 5)), ((3, 2, 1), (2, 3, 4)),
                                          ((3, 3, 3), (1, 1, 1))])
def test_degrees_and_meshes(gpu, degrees, nels):
    v = S.brick(2.0, 1.0, 0.5, degrees=degrees, nels=nels)
    bc = S.SolidBC().add_condition(S.KS_WEST).add_condition(S.KS_FRONT, 2)
    pr = S.SolidProblem(v, bc, law=S.KS_LAW_NEO_HOOKE_LN, E=7.0, nu=0.3, tractions=[(S.KS_EAST, (0.1, 0.0, 0.2)), (S.KS_NORTH, (0.0, -0.3, 0.0))])
    _compare(gpu, pr, 5e-3, f"brick-{degrees}-{nels}")


@pytest.mark.parametrize("nels,seg,full", [((11, 2, 3), None, False), ((11, 2, 3), "4", False), ((9, 3, 2), "8", True), ((2, 4, 2), "1", False),
                                           ((90, 1, 1), None, False)])       # more than 85 elements per direction: tables in shared memory
def test_tricubic_sliding_window(gpu, monkeypatch, nels, seg, full):
    """k3_jacobian_sw (the tri-cubic production kernel): long element rows, segments of 1 / 4 / 8 elements (partial windows at both
    ends of a segment), I <= J + mirror and the full assembly, Dirichlet faces on both ends of the walking direction."""
    if seg:
        monkeypatch.setenv("KS_SW_SEG", seg)
    if full:
        monkeypatch.setenv("KS_FULL", "1")
    v = S.brick(3.0, 1.0, 0.7, degrees=(3, 3, 3), nels=nels)
    bc = S.SolidBC().add_condition(S.KS_WEST).add_condition(S.KS_EAST, 1).add_condition(S.KS_FRONT, 2)
    pr = S.SolidProblem(v, bc, law=S.KS_LAW_NEO_HOOKE_QUAD, E=7.0, nu=0.3, tractions=[(S.KS_NORTH, (0.0, -0.3, 0.1))], body_force=(0.0, 0.0, -0.2))
    _compare(gpu, pr, 5e-3, f"sw-{nels}-{seg}-{full}")


def test_tricubic_window_with_nonuniform_and_repeated_knots(gpu):
    """Walking direction with a non-uniform knot vector and interior knots of multiplicity 2 and 3 (C1 / C0 lines): the window shifts by
    2 or 3 functions at such an element boundary."""
    from gsstructuralanalysis_b200 import geometry as G
    degrees = (3, 3, 3)
    U = (np.array([0, 0, 0, 0, 0.15, 0.4, 0.4, 0.55, 0.8, 0.8, 0.8, 0.9, 1, 1, 1, 1.0]), G.open_uniform_knots(3, 2), G.open_uniform_knots(3, 2))
    gr = [G.greville(p, u) for p, u in zip(degrees, U)]
    B = [G.basis_matrix(p, u, g) for p, u, g in zip(degrees, U, gr)]
    g3, g2, g1 = np.meshgrid(gr[2], gr[1], gr[0], indexing="ij")
    X = np.stack((2.0 * g1 + 0.1 * g2 * g3, 1.0 * g2 + 0.05 * g1 * g1, 0.5 * g3 + 0.05 * g1 * g2), axis=-1)
    for axis, Bm in ((0, B[2]), (1, B[1]), (2, B[0])):
        Xm = np.moveaxis(X, axis, 0)
        X = np.moveaxis(np.linalg.solve(Bm, Xm.reshape(Bm.shape[0], -1)).reshape(Xm.shape), 0, axis)
    v = S.Volume(degrees, tuple(U), np.ascontiguousarray(X.reshape(-1, 3)))
    bc = S.SolidBC().add_condition(S.KS_WEST).add_condition(S.KS_FRONT, 2)
    pr = S.SolidProblem(v, bc, law=S.KS_LAW_NEO_HOOKE_LN, E=7.0, nu=0.3, tractions=[(S.KS_EAST, (0.1, 0.0, 0.2))])
    _compare(gpu, pr, 5e-3, "sw-repeated-knots")


def test_beam_with_prescribed_displacement(gpu):
    """benchmark_Elasticity_Beam_APALM.cpp:226-236 style beam; non-zero fixedDofs on the clamped face."""
    v = S.brick(1.0, 0.01, 0.01, degrees=(3, 2, 2), nels=(8, 1, 1))
    bc = S.SolidBC().add_condition(S.KS_WEST).add_condition(S.KS_EAST, 0)
    pr = S.SolidProblem(v, bc, law=S.KS_LAW_SVK, E=1.0, nu=0.0, tractions=[(S.KS_EAST, (0.0, 0.0, 1e-5))])
    pr.number_dofs(__import__("gsstructuralanalysis_b200.capi", fromlist=["lib"]).lib().ks_build_dofmap)
    pr.fixed_values = 1e-3 * np.random.default_rng(2).standard_normal(pr.n_fixed)
    _compare(gpu, pr, 1e-4, "beam-fixed")


def test_inverted_state_returns_false(gpu):
    asm = gpu(tutorial_problem(S.KS_LAW_NEO_HOOKE_LN))
    x = 10.0 * np.random.default_rng(3).standard_normal(asm.n_dofs)
    ok, _ = asm.residual(x)
    assert not ok and "inverted" in asm.last_error
    ok, _ = asm.residual(np.zeros(asm.n_dofs))      # the flag is cleared: the next call works
    assert ok


def test_full_size_properties(gpu):
    """Large mesh (48^3 elements, tri-cubic, 0.4M DOFs): size-independent properties instead of an oracle run."""
    import scipy.sparse as sp
    v = S.brick(1.0, 1.0, 1.0, degrees=(3, 3, 3), nels=(48, 48, 48))
    pr = S.SolidProblem(v, S.SolidBC(), law=S.KS_LAW_NEO_HOOKE_QUAD, E=5.0, nu=0.3)
    asm = gpu(pr)
    n = asm.n_dofs
    assert n == 3 * 51 ** 3
    ncp = 51 ** 3
    th = 0.2
    R = np.array([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1.0]])
    disp = v.cp @ R.T + np.array([0.1, -0.2, 0.3]) - v.cp
    x = np.zeros(n)
    for c in range(3):
        x[pr.dof_map[c * ncp:(c + 1) * ncp]] = disp[:, c]
    ok, K, r = asm.assemble(np.zeros(n))
    assert ok
    Ks = sp.csc_matrix((K.values.copy(), K.inner, K.outer), shape=(n, n))
    scale = np.abs(K.values).max()
    rng = np.random.default_rng(5)
    u, w = rng.standard_normal(n), rng.standard_normal(n)
    a, b = u @ (Ks @ w), w @ (Ks @ u)
    assert abs(a - b) <= 1e-10 * max(abs(a), abs(b))
    for c in range(3):
        t = np.zeros(n)
        t[pr.dof_map[c * ncp:(c + 1) * ncp]] = 1.0
        assert np.abs(Ks @ t).max() <= 1e-10 * scale
    ok, rr = asm.residual(x)                           # rigid motion: no internal force
    assert ok and np.abs(rr).max() <= 1e-10 * scale


def test_too_large_mesh_is_refused(gpu):
    """nnz must fit index_t = int32: a 92^3 tri-cubic mesh has 2.4 G entries; ks_create fails before the big allocations."""
    from gsstructuralanalysis_b200.capi import KLError
    v = S.Volume((3, 3, 3), tuple(__import__("gsstructuralanalysis_b200.geometry", fromlist=["x"]).open_uniform_knots(3, 92) for _ in range(3)),
                 np.zeros((95 ** 3, 3)))
    g = np.linspace(0.0, 1.0, 95)
    v.cp[:] = np.stack(np.meshgrid(g, g, g, indexing="ij")[::-1], axis=-1).reshape(-1, 3)
    with pytest.raises(KLError) as ei:
        gpu(S.SolidProblem(v, S.SolidBC(), law=S.KS_LAW_SVK))
    assert ei.value.rc == -1 and "int32" in str(ei.value)
