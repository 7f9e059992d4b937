"""Committed golden fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py).
CPU: the oracle reproduces them bit-for-bit in the pattern and to 1e-13 in the values.
GPU (-m gpu): the CUDA path, through the C ABI, matches them to 1e-12 without the oracle in the loop."""
import os

import numpy as np
import pytest

from tests.golden.make_golden import CASES, build

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_matches_golden(name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    pr, o, x = build(name)
    assert np.array_equal(pr.dof_map, g["dof_map"]) and pr.n_free == int(g["n_free"])
    assert np.array_equal(o.outer, g["outer"]) and np.array_equal(o.inner, g["inner"])
    assert np.array_equal(x, g["x"])
    K, R = o.jacobian_values(x), o.residual(x)
    assert np.abs(K - g["K"]).max() <= 1e-13 * np.abs(g["K"]).max()
    assert np.abs(R - g["R"]).max() <= 1e-13 * max(np.abs(g["R"]).max(), np.abs(g["F"]).max())


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CASES))
def test_gpu_matches_golden(name):
    import torch
    assert torch.cuda.is_available()
    from gsstructuralanalysis_b200 import capi
    from gsstructuralanalysis_b200.ops import ShellAssembler
    g = np.load(os.path.join(GOLD, name + ".npz"))
    mk, _ = CASES[name]
    pr = mk()
    asm = ShellAssembler(pr)          # numbers DoFs with the product's kl_build_dofmap
    assert np.array_equal(pr.dof_map, g["dof_map"])
    outer, inner = asm.pattern()
    assert np.array_equal(outer, g["outer"]) and np.array_equal(inner, g["inner"])
    ok, K = asm.jacobian(g["x"])
    assert ok
    ok, R = asm.residual(g["x"])
    assert ok
    assert np.abs(K.values - g["K"]).max() <= 1e-12 * np.abs(g["K"]).max()
    assert np.abs(R - g["R"]).max() <= 1e-12 * max(np.abs(g["R"]).max(), np.abs(g["F"]).max())
    assert np.abs(asm.force() - g["F"]).max() <= 1e-12 * max(np.abs(g["F"]).max(), 1e-300)
