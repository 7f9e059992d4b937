"""Copy-out variants of the Jacobian closure: pageable / page-locked destination, lazy fetch, lower-triangular view,
and the same-state fusion behind the separate Residual / Jacobian calls (all through the C ABI)."""
import numpy as np
import pytest

from gsstructuralanalysis_b200 import workloads as W
from gsstructuralanalysis_b200.problem import KL_MAT_NH, KL_MAT_SVK

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def asm_orc():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("CUDA device required")
    from gsstructuralanalysis_b200 import build as kbuild
    kbuild.build()
    from gsstructuralanalysis_b200.ops import ShellAssembler
    from oracle.binding import Oracle
    pr = W.roof(40, 3)          # 40 element rows: the pipelined strip copy-out is active
    return ShellAssembler(pr), Oracle(pr)


def test_pageable_pinned_lazy_agree(asm_orc):
    import torch
    asm, orc = asm_orc
    x = W.displacement_state(asm.n_dofs, 0.3)
    Ko = orc.jacobian_values(x)
    scale = np.abs(Ko).max()
    ok, K = asm.jacobian(x)                       # pageable numpy destination: staged copy-out
    assert ok and np.abs(K.values - Ko).max() <= 1e-12 * scale
    pinned = torch.empty(asm.nnz, dtype=torch.float64).pin_memory().numpy()
    asm._values = pinned                          # page-locked destination: direct DMA
    ok, K2 = asm.jacobian(x)
    assert ok and np.abs(K2.values - Ko).max() <= 1e-12 * scale
    asm._values = None
    ok, _ = asm.jacobian(x, fetch=False)          # values stay on the device ...
    assert ok
    v = asm.fetch_values()                        # ... until somebody asks for them
    assert np.abs(v - Ko).max() <= 1e-12 * scale
    own = np.zeros(asm.nnz)
    asm.pin_values(own)                           # the owner of a long-lived array pins it explicitly
    asm._values = own
    ok, K3 = asm.jacobian(x)
    asm.unpin_values(own)
    asm._values = None
    assert ok and np.abs(K3.values - Ko).max() <= 1e-12 * scale


def test_lower_triangular_view(asm_orc):
    import scipy.sparse as sp
    asm, orc = asm_orc
    x = W.displacement_state(asm.n_dofs, 0.3)
    Ko = sp.csc_matrix((orc.jacobian_values(x), orc.inner, orc.outer), shape=(asm.n_dofs, asm.n_dofs))
    Lo = sp.tril(Ko, format="csc")
    Lo.sort_indices()
    ok, KL = asm.jacobian_lower(x)
    assert ok
    assert np.array_equal(KL.outer, Lo.indptr) and np.array_equal(KL.inner, Lo.indices)      # bit-exact lower pattern
    assert np.abs(KL.values - Lo.data).max() <= 1e-12 * np.abs(Lo.data).max()
    assert 2 * asm.nnz_lower - asm.n_dofs == asm.nnz


def test_same_state_fusion_and_misses(asm_orc):
    """Residual(x) leaves the per-point records of x behind; Jacobian(x) reuses them, Jacobian(y) must not."""
    asm, orc = asm_orc
    x = W.displacement_state(asm.n_dofs, 0.3)
    y = W.displacement_state(asm.n_dofs, 0.3, seed=7)
    Kx, Ky = orc.jacobian_values(x), orc.jacobian_values(y)
    rx, ry = orc.residual(x), orc.residual(y)
    sK, sR = np.abs(Kx).max(), max(np.abs(rx).max(), np.abs(orc.force()).max())
    for (a, Ka, ra), (b, Kb) in (((x, Kx, rx), (x, Kx)), ((x, Kx, rx), (y, Ky)), ((y, Ky, ry), (y, Ky)), ((y, Ky, ry), (x, Kx))):
        ok, r = asm.residual(a)
        assert ok and np.abs(r - ra).max() <= 1e-12 * sR
        ok, K = asm.jacobian(b)
        assert ok and np.abs(K.values - Kb).max() <= 1e-12 * sK
    # residual-only callers (explicit dynamics, DR): repeated residuals stay correct while speculation switches off
    for a, ra in ((x, rx), (y, ry), (x, rx)):
        ok, r = asm.residual(a)
        assert ok and np.abs(r - ra).max() <= 1e-12 * sR
    ok, K = asm.jacobian(x)
    assert ok and np.abs(K.values - Kx).max() <= 1e-12 * sK
    ok, K = asm.jacobian(np.zeros(asm.n_dofs))
    assert ok and np.abs(K.values - orc.jacobian_values(np.zeros(asm.n_dofs))).max() <= 1e-12 * sK
