#!/usr/bin/env python
"""Generates tests/golden/*.npz: oracle outputs (pattern, K values, residual, force) for small seeded cases.
The real reference (gismo + gsKLShell) cannot be imported or built in this container, so these fixtures freeze the
ORACLE (which is itself pinned by tests/test_oracle_kat.py and tests/test_oracle_energy.py); they guard against
drift of the oracle and give the GPU tests a comparison that does not need the oracle at run time.
Run from the repo root:  python tests/golden/make_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from gsstructuralanalysis_b200 import workloads as W  # noqa: E402
from gsstructuralanalysis_b200.problem import KL_MAT_SVK, KL_MAT_NH, KL_MAT_MR  # noqa: E402
from oracle.binding import Oracle  # noqa: E402

CASES = {
    "paraboloid_svk_n3": (lambda: W.tutorial_paraboloid(3, 3, KL_MAT_SVK, False), 2e-3),
    "paraboloid_nh_n3": (lambda: W.tutorial_paraboloid(3, 3, KL_MAT_NH, False), 2e-3),
    "paraboloid_mrc_n3": (lambda: W.tutorial_paraboloid(3, 3, KL_MAT_MR, True), 2e-3),
    "roof_svk_n4": (lambda: W.roof(4), 0.5),
    "balloon_nh_n3": (lambda: W.balloon(3), 2e-2),
    "frustrum_mr_n3": (lambda: W.frustrum(3), 2e-3),
    "tension_mr_n3": (lambda: W.tension_sheet(3), 1e-5),
}


def build(name):
    mk, scale = CASES[name]
    pr = mk()
    o = Oracle(pr)
    x = W.displacement_state(o.n_dofs, scale)
    return pr, o, x


if __name__ == "__main__":
    out = os.path.dirname(os.path.abspath(__file__))
    for name in CASES:
        pr, o, x = build(name)
        np.savez_compressed(os.path.join(out, name + ".npz"), outer=o.outer, inner=o.inner, x=x,
                            K=o.jacobian_values(x), R=o.residual(x), F=o.force(), dof_map=pr.dof_map,
                            n_free=pr.n_free)
        print(name, o.n_dofs, o.nnz)
