"""CPU restatement (TEST INFRASTRUCTURE) of one arc-length step of the reference:
gsALMBase<T>::_step (src/gsALMSolvers/gsALMBase.hpp:354-416) with gsALMCrisfield<T>
(src/gsALMSolvers/gsALMCrisfield.hpp: predictor :111-164, iteration :75-99, computeLambdas :328-363, computeLambdasSimple :206-226,
computeLambdasModified :244-301, computeLambdasEta :229-242, computeLambdasComplex :302-325, computeLambdaMU :386-401,
computeLambdaDOT :404-425, iterationFinish :192-203), AngleMethod = step, no quasi-Newton.
The linear solves use a sparse direct solver (the reference's default SimplicialLDLT); Jacobian / ALResidual / Force come from
the assembly oracle.  PARITY UNPINNED like the rest of oracle/ (no G+Smo here); checked by its own invariants in
tests/test_oracle_alm.py (arc-length constraint, equilibrium of the converged point, path continuity)."""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla


def crisfield_step(orc, U, L, DUold=None, DLold=0.0, arc_length=1e-2, tolU=1e-6, tolF=1e-3, max_it=100, phi=-1.0, relaxation=1.0):
    """returns (status, U, L, DeltaU, DeltaL, info); status 0 = Success, 1 = NotConverged, 2 = AssemblyError"""
    n = orc.n_dofs
    F = orc.force()
    FF = float(F @ F)
    U = np.array(U, dtype=np.float64)
    DUold = np.zeros(n) if DUold is None else np.array(DUold, dtype=np.float64)
    phi_user = phi >= 0.0
    phi = phi if phi_user else 0.0

    def factor(x):
        K = sp.csc_matrix((orc.jacobian_values(x), orc.inner, orc.outer), shape=(n, n))
        return spla.splu(K)

    info = {"iterations": 0, "residueF": 0.0, "residueU": 0.0}
    try:
        DU = np.zeros(n); dUbar = np.zeros(n); DL = 0.0; dL = 0.0
        # predictor
        lu = factor(U)
        dUt = lu.solve(F)
        if DUold @ DUold == 0 and DLold * DLold == 0:
            dL = arc_length / np.sqrt(2.0 * (dUt @ dUt))
            if not phi_user:
                phi = np.sqrt((dUt @ dUt) / FF)
        else:
            if not phi_user:
                phi = np.sqrt((U @ U) / (L * L * FF))
            A0 = phi * phi * FF
            direction = np.sign(DUold @ dUt + A0 * DLold)
            den = np.sqrt(dUt @ dUt + A0)
            dL = direction * (arc_length if den == 0 else arc_length / den)
        dU = dUbar + dL * dUt
        DU = DU + dU; DL += dL
        A0 = phi * phi * FF
        R = orc.al_residual(U + DU, L + DL)
        basisF, basisU = np.linalg.norm((L + DL) * F), np.linalg.norm(DU)
        info["residueF"] = np.linalg.norm(R) / basisF
        for it in range(1, max_it):
            lu = factor(U + DU)
            dUt = lu.solve(F)
            dUbar = lu.solve(-R)
            eta, lamold = 1.0, dL
            a0 = dUt @ dUt + A0; b0 = 2 * (dUt @ DU + DL * A0); b1 = 2 * (dUbar @ dUt)
            c0 = DU @ DU + DL * DL * A0 - arc_length ** 2; c1 = 2 * (DU @ dUbar); c2 = dUbar @ dUbar
            al1, al2, al3 = a0, b0 + eta * b1, c0 + eta * c1 + eta * eta * c2
            disc = al2 * al2 - 4 * al1 * al3
            complex_root = False
            if disc >= 0:
                dLs = [(-al2 + np.sqrt(disc)) / (2 * al1), (-al2 - np.sqrt(disc)) / (2 * al1)]
            else:
                m1, m2, m3 = b1 * b1 - 4 * a0 * c2, 2 * b0 * b1 - 4 * a0 * c1, b0 * b0 - 4 * a0 * c0
                disc = m2 * m2 - 4 * m1 * m3
                if disc >= 0:
                    e = [(-m2 + np.sqrt(disc)) / (2 * m1), (-m2 - np.sqrt(disc)) / (2 * m1)]
                    eta1, eta2 = min(e), max(e)
                    xi = 0.05 * abs(eta2 - eta1)
                    if eta2 < 1.0: eta = eta2 - xi
                    elif eta2 > 1.0 and -m2 / m1 < 1.0: eta = eta2 + xi
                    elif eta1 < 1.0 and -m2 / m1 > 1.0: eta = eta1 - xi
                    elif eta1 > 1.0: eta = eta1 + xi
                if disc >= 0 and eta > 0.05:
                    al2 = b0 + eta * b1
                    dLs = [-al2 / (2 * al1)] * 2
                else:
                    eta, complex_root = 1.0, True
            if not complex_root:
                t = DUold @ dUt + phi * phi * DLold
                D1, D2 = dLs[0] * t, dLs[1] * t
                dL = dLs[1] if D1 < D2 else dLs[0]
                dU = eta * dUbar + dUt * dL
            else:
                DUcr = DU + dUbar
                K = sp.csc_matrix((orc.jacobian_values(U + DU), orc.inner, orc.outer), shape=(n, n))
                Fint = K @ (U + DU)
                DLcr = (Fint @ F) / FF - L
                mu = arc_length / np.sqrt(DUcr @ DUcr + A0 * DLcr ** 2)
                dL = mu * DLcr - DL
                dU = mu * DUcr - DU
            if lamold * dL < 0 and abs(dL) <= abs(lamold) and relaxation != 1.0:
                dU = relaxation * (dL * dUt + eta * dUbar)
                dL = relaxation * dL
            DU = DU + dU; DL += dL
            R = orc.al_residual(U + DU, L + DL)
            info["residueF"] = np.linalg.norm(R) / basisF
            info["residueU"] = np.linalg.norm(dU) / basisU
            info["iterations"] = it
            if info["residueF"] < tolF and info["residueU"] < tolU:
                info["phi"] = phi
                return 0, U + DU, L + DL, DU, DL, info
        return 1, U, L, DUold, DLold, info
    except RuntimeError:
        return 2, U, L, DUold, DLold, info
