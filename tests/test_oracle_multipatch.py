"""CPU tests of the multi-patch path: the product's DoF numbering (C, kl_mp_build_dofmap) against the oracle's Python
restatement, and the multi-patch oracle against the single-patch oracle on the same function space (a patch cut in two and
glued C0 == the uncut patch with a knot of multiplicity p on the cut)."""
import ctypes as C
import numpy as np
import pytest

from gsstructuralanalysis_b200 import capi, workloads
from gsstructuralanalysis_b200.problem import (MultiPatchProblem, BoundaryConditions, KL_BC_DIRICHLET, KL_BC_CLAMPED, KL_BC_COLLAPSED,
                                               WEST, EAST, SOUTH, NORTH, KL_MAT_NH, KL_MAT_SVK)
from oracle import binding
from oracle.multipatch import build_dofmap_mp, MultiPatchOracle
from tests.mp_problems import cut, dof_permutation


def _maps(mprob):
    return np.concatenate([p.dof_map for p in mprob.patches]), mprob.n_free, mprob.n_fixed


@pytest.mark.parametrize("case", ["2x1", "2x2_bcs", "reversed", "ring"])
def test_dofmap_product_equals_restatement(case):
    import copy
    base = workloads.tutorial_paraboloid(nel=4)
    if case == "2x1":
        _, mp, _ = cut(base, [0.5], [])
    elif case == "2x2_bcs":
        base.bc = BoundaryConditions()
        base.bc.add_condition(WEST, KL_BC_DIRICHLET).add_condition(EAST, KL_BC_CLAMPED, 2).add_condition(NORTH, KL_BC_COLLAPSED, 1)
        base.bc.add_corner_value(1, 0)
        _, mp, _ = cut(base, [0.5], [0.25])
    elif case == "reversed":
        _, mp, _ = cut(base, [0.5], [])
        mp.interfaces = [(0, EAST, 1, WEST, 1)]
    else:   # a closed ring of two patches: east-west and west-east glued (benchmarks/benchmark_Wrinkling.cpp:485 addInterface)
        _, mp, _ = cut(base, [0.5], [])
        mp.interfaces = [(0, EAST, 1, WEST, 0), (1, EAST, 0, WEST, 0)]
    a = copy.deepcopy(mp).number_dofs(capi.lib().kl_mp_build_dofmap)
    b = copy.deepcopy(mp).number_dofs(build_dofmap_mp)
    ma, fa, xa = _maps(a)
    mb, fb, xb = _maps(b)
    assert (fa, xa) == (fb, xb)
    assert np.array_equal(ma, mb)


def test_single_patch_numbering_unchanged():
    """kl_build_dofmap is the one-patch case of the same routine: identical to the oracle's C numbering"""
    pr = workloads.frustrum(nel=4)
    a = np.array(pr.number_dofs(capi.lib().kl_build_dofmap).dof_map)
    b = np.array(pr.number_dofs(binding.lib().klo_build_dofmap).dof_map)
    assert np.array_equal(a, b)


@pytest.mark.parametrize("name,cuts", [("paraboloid_nh", ([0.5], [])), ("roof", ([], [0.5])), ("balloon", ([0.5], [0.5])),
                                        ("frustrum", ([0.25, 0.75], []))])
def test_cut_patch_equals_uncut_patch_with_c0_lines(name, cuts):
    base = {"paraboloid_nh": lambda: workloads.tutorial_paraboloid(nel=4, material=KL_MAT_NH),
            "roof": lambda: workloads.roof(nel=4), "balloon": lambda: workloads.balloon(nel=4),
            "frustrum": lambda: workloads.frustrum(nel=4)}[name]()
    single, multi, cps = cut(base, *cuts)
    single.number_dofs(binding.lib().klo_build_dofmap)
    multi.number_dofs(build_dofmap_mp)
    perm = dof_permutation(single, multi, cps)
    o1 = binding.Oracle(single)
    om = MultiPatchOracle(multi)
    assert om.n_dofs == o1.n_dofs
    rng = np.random.default_rng(5)
    L = np.abs(base.surface.cp).max()
    xs = 1e-3 * L * rng.uniform(-1, 1, o1.n_dofs)
    xm = xs[perm]
    K1 = o1.jacobian(xs).tocsr()
    Km = om.jacobian(xm)
    K1p = K1[perm][:, perm].tocsc()
    K1p.sort_indices()
    assert om.nnz == K1p.nnz and np.array_equal(Km.indices, K1p.indices) and np.array_equal(Km.indptr, K1p.indptr)
    scale = np.abs(K1.data).max()
    assert np.abs(Km.data - K1p.data).max() <= 1e-12 * scale
    r1, rm = o1.residual(xs), om.residual(xm)
    assert np.abs(rm - r1[perm]).max() <= 1e-12 * max(np.abs(r1).max(), 1e-300)
    f1, fm = o1.force(), om.force()
    assert np.abs(fm - f1[perm]).max() <= 1e-12 * max(np.abs(f1).max(), 1e-300)
    a1, am = o1.al_residual(xs, 0.3), om.al_residual(xm, 0.3)
    assert np.abs(am - a1[perm]).max() <= 1e-12 * max(np.abs(a1).max(), 1e-300)
