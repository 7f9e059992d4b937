"""CPU checks of the oracle's restatement of the reference's linear solve and Newton loop:
gsSparseSolver<>::CGDiagonal = Eigen::ConjugateGradient + DiagonalPreconditioner
(src/gsStaticSolvers/gsStaticNewton.hpp:23) and gsStaticNewton::_solveNonlinear (:141-196)."""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from gsstructuralanalysis_b200 import workloads as W
from gsstructuralanalysis_b200.problem import KL_MAT_SVK, KL_MAT_NH
from oracle.binding import Oracle, OracleOps, cg_solve, newton_solve


def _csc(A):
    A = sp.csc_matrix(A)
    A.sort_indices()
    return A.shape[0], A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data.astype(np.float64)


def test_cg_matches_direct_solve_on_shell_stiffness():
    prob = W.tutorial_paraboloid(nel=4, material=KL_MAT_SVK)
    o = Oracle(prob)
    K = o.jacobian(np.zeros(o.n_dofs))
    f = o.force()
    x, it, err = cg_solve(o.n_dofs, o.outer, o.inner, K.data, f, tol=1e-13, max_iter=20000)
    xd = spla.spsolve(K.tocsc(), f)
    assert err < 1e-13 and 0 < it < 20000
    assert np.abs(x - xd).max() <= 1e-8 * np.abs(xd).max()


def test_cg_iteration_count_convention():
    """Eigen leaves the loop before counting the converging iteration: a diagonal system converges inside the
    first loop body, so iterations() == 0; a hit of maxIterations reports maxIterations."""
    d = np.array([2.0, 4.0, 5.0, 10.0])
    n, o, i, v = _csc(sp.diags(d))
    b = np.array([1.0, 2.0, 3.0, 4.0])
    x, it, err = cg_solve(n, o, i, v, b)
    assert it == 0 and np.allclose(x, b / d, rtol=1e-15) and err <= 2.3e-16
    rng = np.random.default_rng(0)
    M = rng.standard_normal((30, 30))
    n, o, i, v = _csc(M @ M.T + 30 * np.eye(30))
    b = rng.standard_normal(30)
    x, it, err = cg_solve(n, o, i, v, b, tol=1e-30, max_iter=5)
    assert it == 5 and err > 1e-30


def test_cg_zero_rhs_and_zero_diagonal():
    n, o, i, v = _csc(sp.diags([1.0, 2.0, 3.0]))
    x, it, err = cg_solve(n, o, i, v, np.zeros(3))
    assert it == 0 and err == 0.0 and not x.any()
    # a structurally present but vanishing diagonal entry is preconditioned with 1 (Eigen::DiagonalPreconditioner)
    A = sp.csc_matrix(np.array([[0.0, 1.0], [1.0, 0.0]]))
    x, it, err = cg_solve(2, np.array([0, 1, 2], np.int32), np.array([1, 0], np.int32), np.array([1.0, 1.0]), np.array([1.0, 1.0]))
    assert np.allclose(x, [1.0, 1.0])


def test_cg_defaults_are_eigens():
    """tol <= 0 -> machine epsilon, max_iter <= 0 -> 2 n."""
    rng = np.random.default_rng(1)
    M = rng.standard_normal((12, 12))
    A = M @ M.T + np.diag(np.logspace(0, 6, 12))
    n, o, i, v = _csc(A)
    b = rng.standard_normal(12)
    x, it, err = cg_solve(n, o, i, v, b)
    assert it <= 24
    assert np.linalg.norm(A @ x - b) <= 1e-9 * np.linalg.norm(b)


def test_newton_converges_like_a_direct_newton():
    """The restated gsStaticNewton loop (linear start, |dU|/|DU| < tolU and |R|/|R0| < tolF) against a plain
    Newton iteration with a direct solver on the same oracle closures."""
    prob = W.tutorial_paraboloid(nel=4, material=KL_MAT_NH)
    prob.point_loads = [((0.5, 0.5), (0.0, 0.0, -2e3))]
    ops = OracleOps(prob)
    U, info = newton_solve(ops, tolU=1e-8, tolF=1e-8, max_it=30, cg_tol=1e-13, cg_max_iter=50000)
    assert info["status"] == 0 and 1 <= info["iterations"] < 30
    ok, r = ops.residual(U)
    assert ok and np.linalg.norm(r) <= 1e-8 * info["residual_ini"]
    x = np.zeros(ops.n_dofs)
    for _ in range(40):
        ok, r = ops.residual(x)
        if np.linalg.norm(r) <= 1e-10 * info["residual_ini"]:
            break
        ok, K = ops.jacobian(x)
        x = x + spla.spsolve(K.tocsc(), r)
    assert np.abs(U - x).max() <= 1e-6 * np.abs(x).max()


def test_newton_reports_not_converged():
    prob = W.tutorial_paraboloid(nel=3, material=KL_MAT_NH)
    ops = OracleOps(prob)
    U, info = newton_solve(ops, tolU=1e-14, tolF=1e-14, max_it=2, cg_tol=1e-12, cg_max_iter=20000)
    assert info["status"] == 1 and info["iterations"] == 2
