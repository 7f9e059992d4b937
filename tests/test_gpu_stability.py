"""kl_stability (device banded LDL^T) against the oracle's dense LDL^T in the same ordering and against numpy's eigenvalues:
pivots 1e-9, number of negative pivots and sign of the indicator exact (gsALMBase::_computeStability, "Determinant" method)."""
import numpy as np
import pytest

from gsstructuralanalysis_b200 import workloads as W
from gsstructuralanalysis_b200.problem import KL_MAT_SVK

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("CUDA device required for -m gpu tests (no CPU fallback exists)")
    from gsstructuralanalysis_b200 import build as kbuild
    kbuild.build()
    from gsstructuralanalysis_b200.ops import ShellAssembler
    return ShellAssembler


@pytest.mark.parametrize("name", ["roof", "paraboloid_wide", "tension", "frustrum"])
def test_pivots_and_inertia(gpu, name):
    from oracle.stability import stability, node_major_permutation
    pr = {"roof": lambda: W.roof(7), "tension": lambda: W.tension_sheet(6), "frustrum": lambda: W.frustrum(6),
          "paraboloid_wide": lambda: W.tutorial_paraboloid(6, material=KL_MAT_SVK)}[name]()
    if name == "paraboloid_wide":      # n1 > n2: the ordering runs along the second direction
        pr.surface = pr.surface.insert_knot(0, 0.37, 1).insert_knot(0, 0.81, 1)
        pr.dof_map = None
    asm = gpu(pr)
    n = asm.n_dofs
    x = W.displacement_state(n, 1e-4 * np.abs(pr.surface.cp).max())
    ok, K = asm.jacobian(x)
    assert ok
    Kd = K.to_scipy().toarray()
    Kd = 0.5 * (Kd + Kd.T)
    perm = node_major_permutation(pr)
    ind, neg, D = asm.stability(return_D=True)
    indo, nego, Do = stability(Kd, perm)
    ev = np.linalg.eigvalsh(Kd)
    assert neg == nego == int((ev < 0).sum())
    assert np.abs(D - Do).max() <= 1e-9 * np.abs(Do).max(), np.abs(D - Do).max() / np.abs(Do).max()
    assert abs(ind - indo) <= 1e-9 * abs(indo)
    # an indefinite matrix on the same pattern: K - sigma I with sigma between two eigenvalues
    for k in (1, 5):
        sigma = 0.5 * (ev[k - 1] + ev[k])
        Ks = K.to_scipy().tolil()
        Ks.setdiag(Ks.diagonal() - sigma)
        Ks = Ks.tocsc(); Ks.sort_indices()
        assert Ks.nnz == asm.nnz
        asm.set_values(Ks.data)
        ind, neg = asm.stability()
        indo, nego, _ = stability(Kd - sigma * np.eye(n), perm)
        assert neg == nego == k and ind < 0 and abs(ind - indo) <= 1e-6 * abs(indo)
    asm.close()


def test_refused_where_it_does_not_apply(gpu):
    asm = gpu(W.balloon(4))           # follower pressure: unsymmetric tangent
    ok, _ = asm.jacobian(np.zeros(asm.n_dofs), fetch=False)
    with pytest.raises(Exception):
        asm.stability()
    asm.close()
