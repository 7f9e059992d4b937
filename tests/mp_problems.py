"""Multi-patch test problems: a single-patch problem cut along parameter lines into conforming patches (what
gsMultiPatch::computeTopology would glue again), next to the uncut patch with C0 lines at the same places — both span the same
function space, so their matrices / residuals agree up to the DoF permutation returned here."""
from __future__ import annotations

import copy
import numpy as np

from gsstructuralanalysis_b200.geometry import split_grid
from gsstructuralanalysis_b200.problem import ShellProblem, MultiPatchProblem, BoundaryConditions, WEST, EAST, SOUTH, NORTH


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
    single = copy.copy(prob)
    single.surface = c0
    single.dof_map = None
    patches, interfaces = split_grid(s, cuts1, cuts2)
    m1, m2 = len(cuts1) + 1, len(cuts2) + 1
    plist, cps = [], []
    N1 = c0.n[0]
    start1 = np.concatenate([[0], np.cumsum([patches[i].n[0] - 1 for i in range(m1)])])
    start2 = np.concatenate([[0], np.cumsum([patches[m1 * j].n[1] - 1 for j in range(m2)])])
    for j in range(m2):
        for i in range(m1):
            q = i + m1 * j
            pp = copy.copy(prob)
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
