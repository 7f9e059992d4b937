"""Oracle of the stability indicator: the pivots of LDL^T carry the inertia of the matrix (Sylvester), whatever the ordering."""
import numpy as np

from oracle.stability import ldlt_pivots, stability, node_major_permutation


def test_pivot_signs_equal_eigenvalue_signs():
    rng = np.random.default_rng(0)
    for n, shift in ((12, 0.0), (30, 1.5), (30, -0.7), (45, 4.0)):
        A = rng.standard_normal((n, n))
        A = A + A.T - shift * np.eye(n)
        ev = np.linalg.eigvalsh(A)
        D = ldlt_pivots(A)
        assert (D < 0).sum() == (ev < 0).sum()
        L = np.eye(n)        # reconstruct to make sure the pivots belong to a factorisation of A
        B = A.copy()
        for k in range(n):
            L[k + 1:, k] = B[k + 1:, k] / B[k, k]
            B[k + 1:, k + 1:] -= np.outer(L[k + 1:, k], L[k + 1:, k]) * B[k, k]
        assert np.abs(L @ np.diag(D) @ L.T - A).max() <= 1e-9 * np.abs(A).max() * max(1.0, np.abs(L).max() ** 2)
        p = rng.permutation(n)
        ind, neg, Dp = stability(A, p)
        assert neg == (ev < 0).sum() and (ind < 0) == (ev.min() < 0)


def test_roof_tangent_inertia():
    from gsstructuralanalysis_b200 import workloads as W
    from oracle.binding import Oracle, lib as olib
    pr = W.roof(5)
    pr.number_dofs(olib().klo_build_dofmap)
    orc = Oracle(pr)
    K = orc.jacobian(np.zeros(orc.n_dofs)).toarray()
    perm = node_major_permutation(pr)
    ind, neg, D = stability(K, perm)
    assert neg == 0 and ind > 0
    ev = np.linalg.eigvalsh(K)
    sigma = 0.5 * (ev[3] + ev[4])
    ind, neg, D = stability(K - sigma * np.eye(len(K)), perm)
    assert neg == 4 and ind < 0
