"""Device CG on an indefinite tangent: Eigen's ConjugateGradient has no p.Ap sign test, so kl_cg_solve must not have one either (the
arc-length solvers solve with an indefinite K past a limit point, gsALMBase.hpp:249-258); an unassembled matrix must still fail at once."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_cg_runs_on_an_indefinite_tangent_and_fails_fast_on_an_empty_matrix():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("CUDA device required for -m gpu tests (no CPU fallback exists)")
    from gsstructuralanalysis_b200 import workloads as W
    from gsstructuralanalysis_b200.capi import KLError
    from gsstructuralanalysis_b200.ops import ShellAssembler
    from gsstructuralanalysis_b200.problem import KL_MAT_NH
    from oracle.binding import Oracle, cg_solve
    prob = W.tutorial_paraboloid(8, 3, KL_MAT_NH, False)
    asm, orc = ShellAssembler(prob, device=0), Oracle(prob)
    x = W.displacement_state(asm.n_dofs, 2e-3)           # compressive membrane stresses in a thin plate: K(x) has negative eigenvalues
    ok, K = asm.jacobian(x)
    assert ok
    outer, inner = asm.pattern()
    import scipy.sparse as sp
    A = sp.csc_matrix((K.values, inner, outer), shape=(asm.n_dofs, asm.n_dofs))
    assert np.linalg.eigvalsh(A.toarray()).min() < 0.0
    f = asm.force()
    xg, itg, errg = asm.cg_solve(f, tol=1e-10, max_iter=100000)
    xo, ito, erro = cg_solve(asm.n_dofs, outer, inner, K.values, f, tol=1e-10, max_iter=100000)
    # both converge (as Eigen does); on an indefinite matrix the iteration counts may differ with the summation order, the solutions agree
    assert errg < 1e-10 and erro < 1e-10
    res = np.abs(A @ xg - f).max() / np.abs(f).max()
    assert res < 1e-8, res
    # an all-zero matrix: p.Ap = 0 -> alpha = inf -> non-finite, reported at once instead of running max_iter iterations
    asm.set_values(np.zeros(asm.nnz))
    with pytest.raises(KLError):
        asm.cg_solve(f, tol=1e-10, max_iter=100000)
    asm.close()
