"""Device-resident Crisfield arc-length step (kl_alm_step) against the oracle's restatement of gsALMCrisfield (oracle/alm.py):
same path, same iteration counts; the GPU solves with CGDiagonal (as benchmarks/benchmark_Frustrum_APALM.cpp:435), the oracle
with a sparse direct solver."""
import numpy as np
import pytest

from gsstructuralanalysis_b200 import workloads as W

pytestmark = pytest.mark.gpu


def test_alm_steps_follow_the_oracle_path():
    import torch
    assert torch.cuda.is_available()
    from gsstructuralanalysis_b200.ops import ShellAssembler
    from oracle.alm import crisfield_step
    from oracle.binding import Oracle
    pr = W.frustrum(6)
    asm, orc = ShellAssembler(pr), Oracle(pr)
    n = asm.n_dofs
    Ug, Lg, DUg, DLg = np.zeros(n), 0.0, None, 0.0
    Uo, Lo, DUo, DLo = np.zeros(n), 0.0, None, 0.0
    arc = 5e-2
    for k in range(3):
        sg, Ug, Lg, DUg, DLg, ig = asm.alm_step(Ug, Lg, DUg, DLg, arc_length=arc, phi=0.0, cg_tol=1e-14)
        so, Uo, Lo, DUo, DLo, io = crisfield_step(orc, Uo, Lo, DUo, DLo, arc_length=arc, phi=0.0)
        assert sg == 0 and so == 0
        assert ig["iterations"] == io["iterations"]
        assert abs(Lg - Lo) <= 1e-8 * abs(Lo)
        assert np.abs(Ug - Uo).max() <= 1e-7 * np.abs(Uo).max()
        assert abs(np.linalg.norm(DUg) - arc) <= 1e-5 * arc
        print(f"step {k}: L={Lg:.8f} iterations={ig['iterations']} cg_iterations={ig['cg_iterations']} "
              f"assembly {ig['ms_assembly']:.2f} ms solve {ig['ms_solve']:.2f} ms")
    # a step that cannot converge leaves the state alone and reports NotConverged (the caller bisects, gsAPALM.hpp:975-983)
    s, U2, L2, _, _, inf = asm.alm_step(Ug, Lg, DUg, DLg, arc_length=arc, phi=0.0, max_it=2, cg_tol=1e-14)
    assert s == 1 and np.array_equal(U2, Ug) and L2 == Lg
    asm.close(); orc.close()
