"""Arc-length step of the oracle (oracle/alm.py, restating gsALMCrisfield): invariants that need no reference run."""
import numpy as np

from gsstructuralanalysis_b200 import workloads as W
from oracle.alm import crisfield_step
from oracle.binding import Oracle


def test_crisfield_steps_on_the_frustrum():
    pr = W.frustrum(4)
    o = Oracle(pr)
    n = o.n_dofs
    U, L, DU, DL = np.zeros(n), 0.0, None, 0.0
    arc = 5e-2
    path = []
    for k in range(4):
        st, U, L, DU, DL, info = crisfield_step(o, U, L, DU, DL, arc_length=arc, phi=0.0)
        assert st == 0 and 1 <= info["iterations"] < 20
        # the converged point is in equilibrium at its load factor: Force - L Force - rhs(U) = 0
        F = o.force()
        assert np.linalg.norm(o.al_residual(U, L)) <= 1e-3 * abs(L) * np.linalg.norm(F)
        # cylindrical constraint (phi = 0): |DeltaU| = arc length
        assert abs(np.linalg.norm(DU) - arc) <= 1e-6 * arc * 10
        path.append(L)
    # the load factor follows the path monotonically at the start (load -1 on the top edge, stable branch)
    assert all(b > a for a, b in zip(path[:-1], path[1:])) or all(b < a for a, b in zip(path[:-1], path[1:]))
    o.close()
