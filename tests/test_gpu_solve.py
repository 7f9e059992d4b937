"""GPU tests of the device-resident linear solve and Newton loop, through the C ABI, against the
oracle's restatement of Eigen's ConjugateGradient + DiagonalPreconditioner (= gsSparseSolver<>::CGDiagonal,
src/gsStaticSolvers/gsStaticNewton.hpp:23) and of gsStaticNewton::_solveNonlinear (:141-196)."""
import numpy as np
import pytest
import scipy.sparse as sp

from gsstructuralanalysis_b200 import workloads as W
from gsstructuralanalysis_b200.problem import KL_MAT_SVK, KL_MAT_NH, KL_MAT_MR

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


def _scipy(K, n):
    return sp.csc_matrix((K.values.copy(), K.inner, K.outer), shape=(n, n))


@pytest.mark.parametrize("builder", [lambda: W.tutorial_paraboloid(nel=6, material=KL_MAT_NH),
                                     lambda: W.roof(nel=9),
                                     lambda: W.tension_sheet(nel=5)])
def test_spmv_matches_scipy(gpu, builder):
    prob = builder()
    asm = gpu(prob)
    n = asm.n_dofs
    x = W.displacement_state(n, 1e-4)
    ok, K = asm.jacobian(x)
    assert ok
    Ks = _scipy(K, n)
    v = np.random.default_rng(3).standard_normal(n)
    y = asm.spmv(v)
    ref = Ks @ v
    assert np.abs(y - ref).max() <= 1e-13 * (abs(Ks) @ np.abs(v)).max()


@pytest.mark.parametrize("tol", [1e-6, 1e-12])
def test_cg_matches_oracle_cg(gpu, tol):
    from oracle.binding import cg_solve
    prob = W.tutorial_paraboloid(nel=6, material=KL_MAT_SVK)
    asm = gpu(prob)
    n = asm.n_dofs
    ok, K = asm.jacobian(np.zeros(n))
    assert ok
    f = asm.force()
    x, it, err = asm.cg_solve(f, tol=tol, max_iter=100000)
    xo, ito, erro = cg_solve(n, K.outer, K.inner, K.values, f, tol=tol, max_iter=100000)
    assert err < tol and erro < tol
    # same algorithm, different summation order: the count may differ by rounding only
    assert abs(it - ito) <= max(3, ito // 50), (it, ito)
    Ks = _scipy(K, n)
    assert np.linalg.norm(Ks @ x - f) <= 2 * tol * np.linalg.norm(f) + 1e-12 * np.linalg.norm(f)
    assert np.abs(x - xo).max() <= 1e3 * tol * np.abs(xo).max()


def test_cg_iteration_conventions(gpu):
    """Zero right-hand side -> x = 0, 0 iterations; max_iter is honoured and reported (Eigen's convention)."""
    prob = W.roof(nel=6)
    asm = gpu(prob)
    n = asm.n_dofs
    ok, K = asm.jacobian(np.zeros(n))
    assert ok
    x, it, err = asm.cg_solve(np.zeros(n))
    assert it == 0 and err == 0.0 and not x.any()
    b = np.random.default_rng(0).standard_normal(n)
    x, it, err = asm.cg_solve(b, tol=1e-30, max_iter=7)
    from oracle.binding import cg_solve
    xo, ito, erro = cg_solve(n, K.outer, K.inner, K.values, b, tol=1e-30, max_iter=7)
    assert it == 7 == ito
    assert np.abs(x - xo).max() <= 1e-10 * np.abs(xo).max()
    assert abs(err - erro) <= 1e-9 * erro


def test_cg_defaults_and_reproducibility(gpu):
    """tol <= 0 / max_iter <= 0 are Eigen's defaults (eps, 2n); two solves are bit-identical (fixed-order reductions)."""
    prob = W.tutorial_paraboloid(nel=3, material=KL_MAT_SVK)
    asm = gpu(prob)
    n = asm.n_dofs
    ok, K = asm.jacobian(np.zeros(n))
    f = asm.force()
    x1, it1, err1 = asm.cg_solve(f)
    x2, it2, err2 = asm.cg_solve(f)
    assert it1 <= 2 * n and it1 == it2 and np.array_equal(x1, x2)
    Ks = _scipy(K, n)
    assert np.linalg.norm(Ks @ x1 - f) <= 1e-8 * np.linalg.norm(f)


def test_cg_rejects_unsymmetric_follower_pressure_tangent(gpu):
    from gsstructuralanalysis_b200.capi import KLError
    prob = W.balloon(nel=4)
    asm = gpu(prob)
    ok, K = asm.jacobian(np.zeros(asm.n_dofs))
    assert ok
    with pytest.raises(KLError):
        asm.cg_solve(np.ones(asm.n_dofs), tol=1e-8)


def test_cg_on_mass_matrix(gpu):
    """The solver works on whatever matrix the last assembly left on the device (here M, SPD)."""
    prob = W.tutorial_paraboloid(nel=5, material=KL_MAT_SVK)
    asm = gpu(prob)
    n = asm.n_dofs
    M = asm.mass(7.0)
    Ms = _scipy(M, n)
    b = np.random.default_rng(2).standard_normal(n)
    x, it, err = asm.cg_solve(b, tol=1e-12, max_iter=10000)
    assert np.linalg.norm(Ms @ x - b) <= 1e-11 * np.linalg.norm(b)


@pytest.mark.parametrize("material", [KL_MAT_SVK, KL_MAT_NH, KL_MAT_MR])
def test_newton_matches_oracle_newton(gpu, material):
    from oracle.binding import OracleOps, newton_solve
    prob = W.tutorial_paraboloid(nel=4, material=material)
    prob.point_loads = [((0.5, 0.5), (0.0, 0.0, -2e3))]
    asm = gpu(prob)
    kw = dict(tolU=1e-8, tolF=1e-8, max_it=30, cg_tol=1e-13, cg_max_iter=50000)
    U, info = asm.newton_solve(**kw)
    Uo, infoo = newton_solve(OracleOps(prob), **kw)
    assert info["status"] == 0 == infoo["status"]
    assert info["iterations"] == infoo["iterations"]
    assert np.abs(U - Uo).max() <= 1e-8 * np.abs(Uo).max()
    assert abs(info["residual_ini"] - infoo["residual_ini"]) <= 1e-9 * infoo["residual_ini"]
    ok, r = asm.residual(U)
    assert ok and np.linalg.norm(r) <= 1e-8 * info["residual_ini"]


def test_newton_without_linear_start_and_not_converged(gpu):
    from oracle.binding import OracleOps, newton_solve
    prob = W.tutorial_paraboloid(nel=3, material=KL_MAT_NH)
    asm = gpu(prob)
    kw = dict(tolU=1e-14, tolF=1e-14, max_it=2, linear_start=False, cg_tol=1e-12, cg_max_iter=20000)
    U, info = asm.newton_solve(**kw)
    Uo, infoo = newton_solve(OracleOps(prob), **kw)
    assert info["status"] == 1 == infoo["status"] and info["iterations"] == 2 == infoo["iterations"]
    assert np.abs(U - Uo).max() <= 1e-7 * np.abs(Uo).max()


def test_newton_reports_assembly_error(gpu):
    """A non-finite start makes the closure fail -> status AssemblyError (gsStaticNewton.hpp:126-127)."""
    prob = W.tutorial_paraboloid(nel=3, material=KL_MAT_NH)
    asm = gpu(prob)
    U0 = np.full(asm.n_dofs, np.nan)
    U, info = asm.newton_solve(U=U0, linear_start=False, cg_tol=1e-10, cg_max_iter=100)
    assert info["status"] == 2
