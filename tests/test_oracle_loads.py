"""Known answers for the load vector Force = assemble().rhs() of the oracle (and, in the -m gpu twin, of the CUDA path):
Neumann edge tractions, follower pressure on the undeformed surface, lifting of non-zero Dirichlet values.
Reference call sites: benchmarks/benchmark_Frustrum_APALM.cpp:236-242,267; benchmark_Cylinder.cpp:118-129;
benchmark_Balloon.cpp:258-263,285."""
import numpy as np
import pytest
import scipy.sparse as sp

from gsstructuralanalysis_b200 import geometry as G, workloads as W
from gsstructuralanalysis_b200.problem import (ShellProblem, BoundaryConditions, KL_MAT_NH, KL_MAT_SVK, KL_BC_DIRICHLET,
                                               WEST, EAST, SOUTH, NORTH)
from oracle.binding import Oracle


def _free_problem(surface, **kw):
    return ShellProblem(surface, BoundaryConditions(), **kw)


def test_neumann_total_force_is_traction_times_edge_length():
    # partition of unity: the nodal forces of a side sum to traction * length of the undeformed edge
    s = W._uniform(G.frustrum(), 3, 6)
    t = np.array([0.3, -0.2, -1.0])
    for side, length in ((NORTH, 0.5 * np.pi * 1.0), (SOUTH, 0.5 * np.pi * 2.0), (WEST, np.sqrt(2.0)), (EAST, np.sqrt(2.0))):
        pr = _free_problem(s, E=1.0, nu=0.3, thickness=0.1, neumann=[(side, tuple(t))])
        o = Oracle(pr)
        F = o.force().reshape(3, -1)
        assert np.allclose(F.sum(axis=1), t * length, rtol=2e-6, atol=0)       # NURBS arc by 4-point Gauss per element
        # only the control points of that side carry load
        n1, n2 = s.n
        idx = np.arange(n1 * n2).reshape(n2, n1)
        edge = {NORTH: idx[-1, :], SOUTH: idx[0, :], WEST: idx[:, 0], EAST: idx[:, -1]}[side]
        mask = np.ones(n1 * n2, bool)
        mask[edge] = False
        assert np.abs(F[:, mask]).max() == 0.0
        o.close()


def test_neumann_against_independent_quadrature():
    """F_i = int N_i t |X'| on a polynomial edge, against scipy's BSpline basis and a dense Gauss rule."""
    from scipy.interpolate import BSpline
    s = W._uniform(G.paraboloid(), 3, 5)
    pr = _free_problem(s, E=1.0, nu=0.3, thickness=0.1, neumann=[(EAST, (0.0, 2.0, 0.0))])
    o = Oracle(pr)
    n1, n2 = s.n
    F = o.force().reshape(3, n2, n1)[1, :, -1]
    U = s.U[1]
    cp_edge = s.cp.reshape(n2, n1, 3)[:, -1, :]
    xg, wg = np.polynomial.legendre.leggauss(12)
    ref = np.zeros(n2)
    for a, b in zip(np.unique(U)[:-1], np.unique(U)[1:]):
        v = 0.5 * (a + b) + 0.5 * (b - a) * xg
        spl = BSpline(U, np.eye(n2), 3)
        B, dB = spl(v).T, spl.derivative()(v).T
        tan = dB.T @ cp_edge
        ref += B @ (0.5 * (b - a) * wg * np.linalg.norm(tan, axis=1)) * 2.0
    assert np.allclose(F, ref, rtol=1e-9, atol=1e-14)
    o.close()


def test_pressure_force_on_unconstrained_octant_is_p_times_projected_area():
    """Force includes the follower pressure on the UNDEFORMED surface: sum_i F_i^c = p * int n_c dA = p pi R^2 / 4."""
    R, p = 10.0, 3.0
    s = W._uniform(G.eighth_sphere(R), 3, 8)
    pr = _free_problem(s, material=KL_MAT_NH, E=3.0, nu=0.5, thickness=0.1, pressure=p)
    o = Oracle(pr)
    F = o.force().reshape(3, -1).sum(axis=1)
    assert np.allclose(F, p * np.pi * R * R / 4.0, rtol=1e-6)
    # arc-length closure: Force - lam*Force - rhs(x) (benchmark_Balloon.cpp:285)
    x = W.dilation_state(pr, 1e-3)
    lam = 0.37
    assert np.allclose(o.al_residual(x, lam), (1.0 - lam) * o.force() - o.residual(x), rtol=0, atol=1e-12 * np.abs(o.force()).max())
    o.close()


def test_dirichlet_lifting_equals_minus_Kfd_g():
    """Force of a problem with prescribed non-zero displacements = F_dead - K_L[free, eliminated] g, with K_L taken from the
    same problem WITHOUT boundary conditions (every DoF free) at the undeformed configuration."""
    s = W._uniform(G.paraboloid(), 3, 4)
    kw = dict(material=KL_MAT_SVK, E=1e3, nu=0.3, thickness=0.05, point_loads=[((0.5, 0.5), (0.0, 0.0, -2.0))])
    full = Oracle(_free_problem(s, **kw))
    Kf = sp.csc_matrix((full.jacobian_values(np.zeros(full.n_dofs)), full.inner, full.outer), shape=(full.n_dofs,) * 2).toarray()
    bc = BoundaryConditions().add_condition(WEST, KL_BC_DIRICHLET).add_condition(EAST, KL_BC_DIRICHLET, 0)
    pr = ShellProblem(s, bc, **kw)
    pr.number_dofs(__import__("oracle.binding", fromlist=["lib"]).lib().klo_build_dofmap)
    rng = np.random.default_rng(3)
    pr.fixed_values = 1e-3 * rng.uniform(-1, 1, pr.n_fixed)
    o = Oracle(pr)
    n1, n2 = s.n
    ncp = n1 * n2
    dm = np.asarray(pr.dof_map)                 # [3*ncp] -> constrained numbering; the free problem numbers DoF k as k
    free = dm < pr.n_free
    gvec = np.zeros(3 * ncp)
    gvec[~free] = pr.fixed_values[dm[~free] - pr.n_free]
    lift = Kf @ gvec
    expect = np.zeros(pr.n_free)
    expect[dm[free]] = full.force()[free] - lift[free]
    assert np.allclose(o.force(), expect, rtol=0, atol=1e-11 * np.abs(expect).max())
    full.close(); o.close()
