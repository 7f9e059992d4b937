"""Named workloads = BASELINE.json configs / SURVEY §8(d) synthetic inputs S1..S5, as ShellProblem builders.
Only problem *definitions* live here (geometry, BCs, material, loads); no arithmetic of the hot path."""
from __future__ import annotations

import copy as _copy

import numpy as np

from . import geometry as G
from .problem import (ShellProblem, MultiPatchProblem, BoundaryConditions, KL_MAT_SVK, KL_MAT_NH, KL_MAT_MR, KL_BC_DIRICHLET,
                      KL_BC_CLAMPED, KL_BC_COLLAPSED, WEST, EAST, SOUTH, NORTH, SW, SE, NW, NE)


def _uniform(surface, degree, nel):
    s = surface.degree_elevate(degree - surface.p[0]) if surface.p[0] == surface.p[1] else surface
    if s.p != (degree, degree):
        s = surface.respace((degree, degree), (G.elevate_knots(surface.p[0], surface.U[0], degree - surface.p[0]),
                                              G.elevate_knots(surface.p[1], surface.U[1], degree - surface.p[1])))
    return s.refine_to(nel)


def tutorial_paraboloid(nel=4, degree=3, material=KL_MAT_NH, compressible=False):
    """configs[0]: tutorials/nonlinear_shell_static.cpp — corner-pinned paraboloid, E=1e9, nu=0.45, t=1e-2,
    point load (0,0,-1e4) at (0.5,0.5), Material=1 (NH), Implementation=1 (:71-80,92-105)."""
    s = _uniform(G.paraboloid(), degree, nel)
    bc = BoundaryConditions()
    for c in (SW, SE, NW, NE):
        bc.add_corner_value(c)
    return ShellProblem(s, bc, material=material, compressible=compressible, E=1e9, nu=0.45, thickness=1e-2,
                        point_loads=[((0.5, 0.5), (0.0, 0.0, -1e4))])


def roof(nel=8, degree=3, thickness=6.35):
    """configs[1]: benchmarks/benchmark_Roof.cpp — shallow Scordelis-Lo roof, E=3102.75, nu=0.3, SvK,
    north/south edges fixed in all components, point load at (0.5,0.5) (:128-144,201-232)."""
    s = _uniform(G.scordelis_lo_roof_shallow(), degree, nel)
    bc = BoundaryConditions()
    bc.add_condition(NORTH, KL_BC_DIRICHLET).add_condition(SOUTH, KL_BC_DIRICHLET)
    return ShellProblem(s, bc, material=KL_MAT_SVK, E=3102.75, nu=0.3, thickness=thickness,
                        point_loads=[((0.5, 0.5), (0.0, 0.0, -1e1))])


def balloon(nel=8, degree=3, material=KL_MAT_NH, pressure=1e3):
    """configs[2]: benchmarks/benchmark_Balloon.cpp — NURBS eighth sphere, incompressible NH, mu=4.225e5,
    t=0.1, follower pressure, symmetry BCs (:50-51,93-110,149,258)."""
    s = _uniform(G.eighth_sphere(10.0), degree, nel)
    mu = 4.225e5
    bc = BoundaryConditions()
    # symmetry planes: x=0 on the east edge (u=1), y=0 on the west edge (u=0), z=0 on the south edge (v=0)
    bc.add_condition(WEST, KL_BC_DIRICHLET, 1).add_condition(WEST, KL_BC_CLAMPED, 0).add_condition(WEST, KL_BC_CLAMPED, 2)
    bc.add_condition(EAST, KL_BC_DIRICHLET, 0).add_condition(EAST, KL_BC_CLAMPED, 1).add_condition(EAST, KL_BC_CLAMPED, 2)
    bc.add_condition(SOUTH, KL_BC_DIRICHLET, 2).add_condition(SOUTH, KL_BC_CLAMPED, 0).add_condition(SOUTH, KL_BC_CLAMPED, 1)
    bc.add_condition(NORTH, KL_BC_COLLAPSED, 2).add_condition(NORTH, KL_BC_DIRICHLET, 0).add_condition(NORTH, KL_BC_DIRICHLET, 1)
    return ShellProblem(s, bc, material=material, compressible=False, E=2 * mu * 1.5, nu=0.5, thickness=0.1, pressure=pressure)


def tension_sheet(nel=8, degree=3):
    """configs[3]: benchmarks/benchmark_TensionWrinkling.cpp — Rectangle(0.14,0.07) with clamping knots, MR
    C10=6.21485502e4, C01=15.8114570e4, t=0.14e-3 (:160-205,229-234)."""
    s = G.rectangle_with_clamping(0.14, 0.07, degree, nel, nel, 1e-2)
    C10, C01 = 6.21485502e4, 15.8114570e4
    mu = 2 * (C10 + C01)
    nu = 0.5
    bc = BoundaryConditions()
    bc.add_condition(WEST, KL_BC_DIRICHLET)
    bc.add_condition(EAST, KL_BC_COLLAPSED, 0).add_condition(EAST, KL_BC_DIRICHLET, 1).add_condition(EAST, KL_BC_DIRICHLET, 2)
    bc.add_condition(WEST, KL_BC_CLAMPED, 2).add_condition(EAST, KL_BC_CLAMPED, 2)
    bc.add_condition(WEST, KL_BC_DIRICHLET)
    return ShellProblem(s, bc, material=KL_MAT_MR, compressible=False, E=2 * mu * (1 + nu), nu=nu, thickness=0.14e-3,
                        mr_ratio=C10 / C01, point_loads=[((1.0, 0.5), (1.0, 0.0, 0.0))])


def frustrum(nel=8, degree=3):
    """configs[4]: benchmarks/benchmark_Frustrum_APALM.cpp testCase 0 — quarter frustrum R1=2, R2=1, h=1, MR mu=4.225,
    ratio 7, t=0.1 (:195-207); north: Neumann traction (0,0,-1), x/y dirichlet, z collapsed; south fixed; east x=0 symmetry
    (dirichlet x, clamped y,z); west y=0 symmetry (clamped x, dirichlet y, clamped z) (:216-259)."""
    s = _uniform(G.frustrum(), degree, nel)
    mu = 4.225
    bc = BoundaryConditions()
    bc.add_condition(NORTH, KL_BC_DIRICHLET, 0).add_condition(NORTH, KL_BC_DIRICHLET, 1).add_condition(NORTH, KL_BC_COLLAPSED, 2)
    bc.add_condition(SOUTH, KL_BC_DIRICHLET)
    bc.add_condition(EAST, KL_BC_DIRICHLET, 0).add_condition(EAST, KL_BC_CLAMPED, 1).add_condition(EAST, KL_BC_CLAMPED, 2)
    bc.add_condition(WEST, KL_BC_CLAMPED, 0).add_condition(WEST, KL_BC_DIRICHLET, 1).add_condition(WEST, KL_BC_CLAMPED, 2)
    return ShellProblem(s, bc, material=KL_MAT_MR, compressible=False, E=2 * mu * 1.5, nu=0.5, thickness=0.1, mr_ratio=7.0,
                        neumann=[(NORTH, (0.0, 0.0, -1.0))])


def cylinder(nel=8, degree=3, material=KL_MAT_NH):
    """configs[2]: benchmarks/benchmark_Cylinder.cpp with -M 1 — half cylinder (filedata/surface/half_cylinder.xml), E=168e9,
    nu=0.4, t=2e-3 (:97-100), incompressible Neo-Hookean; north: Neumann traction (0,0,-1), y dirichlet, z clamped; east:
    x dirichlet, z clamped; south: fixed (+ clamped z) (:118-139)."""
    s = _uniform(G.half_cylinder(), degree, nel)
    bc = BoundaryConditions()
    bc.add_condition(NORTH, KL_BC_DIRICHLET, 1).add_condition(NORTH, KL_BC_CLAMPED, 2)
    bc.add_condition(EAST, KL_BC_DIRICHLET, 0).add_condition(EAST, KL_BC_CLAMPED, 2)
    bc.add_condition(SOUTH, KL_BC_DIRICHLET)      # the additional clamped z on an eliminated side changes nothing
    return ShellProblem(s, bc, material=material, compressible=False, E=168e9, nu=0.4, thickness=2e-3, mr_ratio=4.0,
                        neumann=[(NORTH, (0.0, 0.0, -1.0))])


def plate_1m(nel=576, degree=3, material=KL_MAT_SVK, compressible=False):
    """SURVEY §8 canonical 1M-DOF case: single patch, degree 3, 576x576 elements, corner-pinned shallow
    paraboloid with the tutorial's material constants."""
    return tutorial_paraboloid(nel, degree, material, compressible)


def displacement_state(n_dofs, scale, seed=20240607):
    """x = scale * U(-1,1) per DoF, seed 20240607.  numpy's PCG64 stands in for mt19937_64:
    the same generator feeds the oracle and the GPU path, so parity does not depend on it."""
    rng = np.random.default_rng(seed)
    return scale * rng.uniform(-1.0, 1.0, n_dofs)


def dilation_state(prob, eps, noise=0.0, seed=20240607):
    """Smooth synthetic state u = eps * X (a homogeneous dilation of the control net) on the free DoFs plus optional noise:
    valid on degenerate parametrisations (the pole of the balloon) and on thick, finely meshed shells where seeded noise
    of a fixed fraction of the element size would flip the through-thickness metric.  Coupled DoFs take the value of the
    last control point mapped to them."""
    n1, n2 = prob.surface.n
    ncp = n1 * n2
    x = np.zeros(prob.n_free)
    dm = np.asarray(prob.dof_map).reshape(3, ncp)
    for c in range(3):
        free = dm[c] < prob.n_free
        x[dm[c][free]] = eps * prob.surface.cp[free, c]
    if noise:
        x += noise * np.random.default_rng(seed).uniform(-1.0, 1.0, prob.n_free)
    return x


def smooth_state(prob, amp, noise=0.0, seed=20240607):
    """Smooth synthetic state u_c = amp_c sin(pi xi) sin(pi eta) sampled at the Greville abscissae of the control points
    (+ optional seeded noise): it vanishes on every side, so it is compatible with any homogeneous Dirichlet / symmetry
    condition, with the degenerate pole of the balloon and with thick, finely meshed shells, where seeded noise of a fixed
    fraction of the element size would flip the through-thickness metric.  amp: scalar or 3 components."""
    s = prob.surface
    g1, g2 = G.greville(s.p[0], s.U[0]), G.greville(s.p[1], s.U[1])
    phi = np.outer(np.sin(np.pi * g2), np.sin(np.pi * g1)).reshape(-1)          # control point i = i1 + n1 * i2
    a = np.broadcast_to(np.asarray(amp, dtype=np.float64), (3,)) * np.array([1.0, -0.7, 0.5])
    ncp = phi.size
    x = np.zeros(prob.n_free)
    dm = np.asarray(prob.dof_map).reshape(3, ncp)
    for c in range(3):
        free = dm[c] < prob.n_free
        x[dm[c][free]] = a[c] * phi[free]
    if noise:
        x += noise * np.random.default_rng(seed).uniform(-1.0, 1.0, prob.n_free)
    return x


# ---- multi-patch workloads: a single-patch problem cut along parameter lines into conforming patches (what gsMultiPatch::
#      computeTopology glues again), next to the uncut patch with C0 lines at the same places — both span the same function
#      space, so their matrices / residuals agree up to the DoF permutation of dof_permutation()
def cut(prob: ShellProblem, cuts1, cuts2):
    """-> (single: ShellProblem on the surface with C0 lines, multi: MultiPatchProblem, cps: per patch the control-point index of the
    single patch each patch control point coincides with)"""
    s = prob.surface
    p = s.p
    c0 = s
    for u in cuts1:
        c0 = c0.insert_knot(0, u, p[0])
    for v in cuts2:
        c0 = c0.insert_knot(1, v, p[1])
    single = _copy.copy(prob)
    single.surface = c0
    single.dof_map = None
    patches, interfaces = G.split_grid(s, cuts1, cuts2)
    m1, m2 = len(cuts1) + 1, len(cuts2) + 1
    plist, cps = [], []
    N1 = c0.n[0]
    start1 = np.concatenate([[0], np.cumsum([patches[i].n[0] - 1 for i in range(m1)])])
    start2 = np.concatenate([[0], np.cumsum([patches[m1 * j].n[1] - 1 for j in range(m2)])])
    for j in range(m2):
        for i in range(m1):
            q = i + m1 * j
            pp = _copy.copy(prob)
            pp.surface = patches[q]
            pp.dof_map = None
            bc = BoundaryConditions()
            if i == 0:
                bc.side[WEST] = prob.bc.side[WEST]
            if i == m1 - 1:
                bc.side[EAST] = prob.bc.side[EAST]
            if j == 0:
                bc.side[SOUTH] = prob.bc.side[SOUTH]
            if j == m2 - 1:
                bc.side[NORTH] = prob.bc.side[NORTH]
            if i == 0 and j == 0:
                bc.corner[0] = prob.bc.corner[0]
            if i == m1 - 1 and j == 0:
                bc.corner[1] = prob.bc.corner[1]
            if i == 0 and j == m2 - 1:
                bc.corner[2] = prob.bc.corner[2]
            if i == m1 - 1 and j == m2 - 1:
                bc.corner[3] = prob.bc.corner[3]
            pp.bc = bc
            # Neumann sides stay with the patches that own that side; point loads go to the patch that contains them
            pp.neumann = [(sd, t) for sd, t in prob.neumann
                          if (sd == WEST and i == 0) or (sd == EAST and i == m1 - 1) or (sd == SOUTH and j == 0) or (sd == NORTH and j == m2 - 1)]
            U1, U2 = patches[q].U
            pls = []
            for (u, v), f in prob.point_loads:
                in1 = (U1[0] <= u < U1[-1]) or (i == m1 - 1 and u == U1[-1])
                in2 = (U2[0] <= v < U2[-1]) or (j == m2 - 1 and v == U2[-1])
                if in1 and in2:
                    pls.append(((u, v), f))
            pp.point_loads = pls
            plist.append(pp)
            n1, n2 = patches[q].n
            i1 = start1[i] + np.arange(n1)
            i2 = start2[j] + np.arange(n2)
            cps.append((i1[None, :] + N1 * i2[:, None]).reshape(-1))
    return single, MultiPatchProblem(plist, interfaces), cps


def dof_permutation(single: ShellProblem, multi: MultiPatchProblem, cps):
    """perm[g_multi] = g_single for every free DoF (asserts that the two numberings describe the same space)"""
    ncp_s = single.surface.n[0] * single.surface.n[1]
    perm = np.full(multi.n_free, -1, dtype=np.int64)
    for pp, cp in zip(multi.patches, cps):
        ncp = len(cp)
        for c in range(3):
            gm = pp.dof_map[c * ncp:(c + 1) * ncp]
            gs = single.dof_map[c * ncp_s + cp]
            free = gm < multi.n_free
            assert np.array_equal(free, gs < single.n_free)
            prev = perm[gm[free]]
            assert np.all((prev < 0) | (prev == gs[free]))
            perm[gm[free]] = gs[free]
    assert multi.n_free == single.n_free and np.array_equal(np.sort(perm), np.arange(single.n_free))
    return perm
