"""GPU parity of the multi-patch assembler (include/kl_shell.h: kl_mp_*), through the C ABI:
  * against the multi-patch oracle (per-patch single-patch oracle + scipy sum): pattern bit-exact, K / R / AL residual / Force 1e-12;
  * against the SINGLE-PATCH GPU assembly of the same function space (the uncut patch with C0 lines): the identity that pins the
    gluing without any reference;
  * the matrix-level consumers on the matrix context: CG, Newton, lower-triangular copy-out, lazy fetch, mass;
  * patch -> GPU partition: the partial sums of the active patches add up to the whole."""
import copy
import numpy as np
import pytest

from gsstructuralanalysis_b200 import workloads as W
from gsstructuralanalysis_b200.problem import KL_MAT_NH, KL_MAT_SVK
from tests.mp_problems import cut, dof_permutation

pytestmark = pytest.mark.gpu
RTOL = 1e-12


@pytest.fixture(scope="module")
def gpu():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("CUDA device required for -m gpu tests (no CPU fallback exists)")
    from gsstructuralanalysis_b200 import build as kbuild
    kbuild.build()
    from gsstructuralanalysis_b200 import ops
    return ops


CASES = {"paraboloid_nh": (lambda: W.tutorial_paraboloid(nel=6, material=KL_MAT_NH), ([0.5], [])),
         "roof": (lambda: W.roof(nel=6), ([], [0.5])),
         "balloon": (lambda: W.balloon(nel=6), ([0.5], [0.5])),
         "frustrum": (lambda: W.frustrum(nel=8), ([0.25, 0.75], [])),
         "tension_2x3": (lambda: W.tension_sheet(nel=6), ([0.5], [1.0 / 3, 2.0 / 3]))}


@pytest.mark.parametrize("name", list(CASES))
def test_multipatch_matches_oracle_and_uncut_patch(gpu, name):
    from oracle.multipatch import MultiPatchOracle
    from gsstructuralanalysis_b200 import capi
    make, cuts = CASES[name]
    base = make()
    single, multi, cps = cut(base, *cuts)
    asm = gpu.MultiPatchAssembler(multi)
    orc = MultiPatchOracle(copy.deepcopy(multi))
    assert asm.n_dofs == orc.n_dofs and asm.nnz == orc.nnz
    outer, inner = asm.pattern()
    assert np.array_equal(outer, orc.outer) and np.array_equal(inner, orc.inner)      # bit-exact union pattern
    single.number_dofs(capi.lib().kl_build_dofmap)
    perm = dof_permutation(single, multi, cps)
    one = gpu.ShellAssembler(single)
    f, fo = asm.force(), orc.force()
    assert np.abs(f - fo).max() <= RTOL * max(np.abs(fo).max(), 1e-300)
    L = np.abs(base.surface.cp).max()
    rng = np.random.default_rng(11)
    for amp in (0.0, 1e-3):
        xm = amp * L * rng.uniform(-1, 1, asm.n_dofs)
        ok, K = asm.jacobian(xm)
        assert ok, asm.last_error
        Ko = orc.jacobian_values(xm)
        sc = np.abs(Ko).max()
        assert np.abs(K.values - Ko).max() <= RTOL * sc, name
        ok, r = asm.residual(xm)
        ro = orc.residual(xm)
        rs = max(np.abs(ro).max(), np.abs(fo).max(), 1e-4 * sc * L, 1e-300)
        assert np.abs(r - ro).max() <= RTOL * rs
        ok, ra = asm.al_residual(xm, 0.37)
        assert np.abs(ra - orc.al_residual(xm, 0.37)).max() <= RTOL * rs
        # the same space assembled as ONE patch on the GPU
        xs = np.zeros(one.n_dofs)
        xs[perm] = xm
        ok, K1 = one.jacobian(xs)
        K1p = K1.to_scipy().tocsr()[perm][:, perm].tocsc()
        K1p.sort_indices()
        assert np.array_equal(K1p.indptr, outer) and np.array_equal(K1p.indices, inner)
        assert np.abs(K1p.data - K.values).max() <= RTOL * sc
        ok, r1 = one.residual(xs)
        assert np.abs(r1[perm] - r).max() <= RTOL * rs
    asm.close(); one.close(); orc.close()


def test_matrix_level_consumers_on_the_matrix_context(gpu):
    from oracle.multipatch import MultiPatchOracle
    base = W.tutorial_paraboloid(nel=8, material=KL_MAT_SVK)
    base.point_loads = [((0.3, 0.6), (0.0, 0.0, -2e3))]
    base.body_force = (0.0, 0.0, -5.0)
    single, multi, cps = cut(base, [0.5], [0.5])
    asm = gpu.MultiPatchAssembler(multi)
    orc = MultiPatchOracle(copy.deepcopy(multi))
    n = asm.n_dofs
    x = 1e-4 * np.random.default_rng(3).uniform(-1, 1, n)
    ok, K = asm.jacobian(x)
    Ks = K.to_scipy()
    # lazy fetch == direct copy-out; lower-triangular view
    ok, _ = asm.jacobian(x, fetch=False)
    assert ok and np.abs(asm.fetch_values() - K.values).max() <= 1e-13 * np.abs(K.values).max()   # two atomically assembled matrices
    ok, Kl = asm.jacobian_lower(x)
    import scipy.sparse as sp
    low = sp.tril(Ks).tocsc(); low.sort_indices()
    assert np.array_equal(Kl.inner, low.indices) and np.abs(Kl.values - low.data).max() <= 1e-13 * np.abs(low.data).max()
    # CG and SpMV on the device matrix
    b = asm.force()
    ok, _ = asm.jacobian(x, fetch=False)
    y = asm.spmv(b)
    assert np.abs(y - Ks @ b).max() <= 1e-12 * np.abs(Ks @ b).max()
    sol, it, err = asm.cg_solve(b, tol=1e-12)
    import scipy.sparse.linalg as spl
    ref = spl.spsolve(Ks.tocsc(), b)
    assert np.linalg.norm(Ks @ sol - b) <= 1e-3 * np.linalg.norm(b)           # true residual; the recurrence residual met Eigen's stop test (ill-conditioned shell matrix)
    assert np.abs(sol - ref).max() <= 1e-5 * np.abs(ref).max()               # ill-conditioned shell matrix: cond * tol
    # Newton on the multi-patch == Newton with the multi-patch oracle closures
    kw = dict(tolU=1e-8, tolF=1e-8, max_it=30, cg_tol=1e-13, cg_max_iter=50000)
    U, info = asm.newton_solve(**kw)
    assert info["status"] == 0, info
    ok, r = asm.residual(U)
    assert np.linalg.norm(r) <= 1e-8 * info["residual_ini"]
    assert np.linalg.norm(orc.residual(U)) <= 1e-7 * info["residual_ini"]
    # ... and == the single-patch Newton on the same function space
    from gsstructuralanalysis_b200 import capi
    single.number_dofs(capi.lib().kl_build_dofmap)
    perm = dof_permutation(single, multi, cps)
    one = gpu.ShellAssembler(single)
    U1, info1 = one.newton_solve(**kw)
    assert info1["status"] == 0 and info1["iterations"] == info["iterations"]
    assert np.abs(U1[perm] - U).max() <= 1e-7 * np.abs(U).max()
    one.close()
    # mass matrix and lumped mass
    M = asm.mass(7.0)
    vo, lo = orc.mass(7.0)
    assert np.abs(M.values - vo).max() <= RTOL * np.abs(vo).max()
    assert np.abs(asm.mass(7.0, lumped=True) - lo).max() <= RTOL * np.abs(lo).max()
    # geometry-level calls go to a patch view
    uv = np.array([[0.25, 0.25], [0.4, 0.1]])
    s = asm.patch(0).eval_stress(U, "displacement", uv)
    from oracle.binding import Oracle
    so = orc.parts[0].eval_stress(U, "displacement", uv)
    assert np.abs(s - so).max() <= 1e-12 * max(np.abs(so).max(), 1e-300)
    with pytest.raises(Exception):
        asm.eval_stress(U, "displacement", uv)
    asm.close(); orc.close()


def test_patch_partition_partials_add_up(gpu):
    """patch -> GPU partition on one device: two assemblers with complementary active sets"""
    base = W.roof(nel=8)
    _, multi, _ = cut(base, [0.5], [0.5])
    whole = gpu.MultiPatchAssembler(copy.deepcopy(multi))
    a = gpu.MultiPatchAssembler(copy.deepcopy(multi))
    b = gpu.MultiPatchAssembler(copy.deepcopy(multi))
    a.set_active([1, 0, 0, 1])
    b.set_active([0, 1, 1, 0])
    x = 1e-2 * np.random.default_rng(2).uniform(-1, 1, whole.n_dofs)
    ok, K = whole.jacobian(x)
    ok, Ka = a.jacobian(x)
    va = Ka.values.copy()
    ok, Kb = b.jacobian(x)
    sc = np.abs(K.values).max()
    assert np.abs(va + Kb.values - K.values).max() <= RTOL * sc
    # only interface columns receive contributions from both sides
    both = (va != 0) & (Kb.values != 0)
    cols = np.repeat(np.arange(whole.n_dofs), np.diff(K.outer))
    assert set(np.unique(cols[both])) <= set(whole.interface_dofs().tolist())
    ok, r = whole.residual(x)
    ok, ra = a.residual(x)
    ra = ra.copy()
    ok, rb = b.residual(x)
    assert np.abs(ra + rb - r).max() <= RTOL * max(np.abs(r).max(), 1e-300)
    assert np.abs(a.force() + b.force() - whole.force()).max() <= RTOL * np.abs(whole.force()).max()
    for m in (whole, a, b):
        m.close()
