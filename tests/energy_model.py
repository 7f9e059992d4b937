"""Independent numpy restatement of the DISCRETE STRAIN ENERGY of the KL shell (no variations, no
tangents): W(u) = sum_qp w |A1 x A2| int_z psi(C(z)) dz.   F_int must equal dW/du — this pins the
oracle's stress resultants and first variations without re-using any hand-derived formula
(SURVEY §8c item 4).  Uses scipy's BSpline for the basis, so it also cross-checks the basis code."""
import numpy as np
from scipy.interpolate import BSpline
from scipy.optimize import brentq

from gsstructuralanalysis_b200.problem import KL_MAT_SVK, KL_MAT_NH, KL_MAT_MR


def _basis(p, U, u, der):
    n = len(U) - p - 1
    spl = BSpline(U, np.eye(n), p, extrapolate=False)
    if der:
        spl = spl.derivative(der)
    return np.nan_to_num(spl(u))


def _psi_incomp(prob, Gc, gc):
    mu = prob.E / (2 * (1 + prob.nu))
    c1, c2 = mu, 0.0
    if prob.material == KL_MAT_MR:
        c2 = mu / (prob.mr_ratio + 1)
        c1 = prob.mr_ratio * c2
    Gi = np.linalg.inv(Gc)
    J0sq = np.linalg.det(gc) / np.linalg.det(Gc)
    c33 = 1.0 / J0sq
    trs = np.sum(gc * Gi)
    I1 = trs + c33
    I2 = c33 * trs + J0sq
    return 0.5 * c1 * (I1 - 3) + 0.5 * c2 * (I2 - 3)


def _psi_comp(prob, Gc, gc, c33):
    mu = prob.E / (2 * (1 + prob.nu))
    K = 2 * mu * (1 + prob.nu) / (3 - 6 * prob.nu)
    c1, c2 = mu, 0.0
    if prob.material == KL_MAT_MR:
        c2 = mu / (prob.mr_ratio + 1)
        c1 = prob.mr_ratio * c2
    G3 = np.eye(3); G3[:2, :2] = Gc
    C3 = np.eye(3); C3[:2, :2] = gc; C3[2, 2] = c33
    Cm = np.linalg.solve(G3, C3)            # mixed components G^-1 C
    I1 = np.trace(Cm)
    I2 = 0.5 * (I1 ** 2 - np.trace(Cm @ Cm))
    J = np.sqrt(np.linalg.det(Cm))
    return 0.5 * c1 * (J ** (-2 / 3) * I1 - 3) + 0.5 * c2 * (J ** (-4 / 3) * I2 - 3) + 0.25 * K * (J * J - 1 - 2 * np.log(J))


def energy(prob, x):
    s = prob.surface
    p1, p2 = s.p
    U1, U2 = s.U
    n1, n2 = s.n
    ncp = n1 * n2
    disp = np.zeros((ncp, 3))
    for c in range(3):
        g = prob.dof_map[c * ncp:(c + 1) * ncp]
        free = g < prob.n_free
        disp[free, c] = x[g[free]]
        disp[~free, c] = prob.fixed_values[g[~free] - prob.n_free]
    w = np.ones(ncp) if s.w is None else s.w
    Hh = np.concatenate([s.cp * w[:, None], w[:, None]], 1).reshape(n2, n1, 4)
    Dd = disp.reshape(n2, n1, 3)
    xg1, wg1 = np.polynomial.legendre.leggauss(prob.quA * p1 + prob.quB)
    xg2, wg2 = np.polynomial.legendre.leggauss(prob.quA * p2 + prob.quB)
    b1 = np.unique(U1); b2 = np.unique(U2)
    u = np.concatenate([0.5 * (a + b) + 0.5 * (b - a) * xg1 for a, b in zip(b1[:-1], b1[1:])])
    wu = np.concatenate([0.5 * (b - a) * wg1 for a, b in zip(b1[:-1], b1[1:])])
    v = np.concatenate([0.5 * (a + b) + 0.5 * (b - a) * xg2 for a, b in zip(b2[:-1], b2[1:])])
    wv = np.concatenate([0.5 * (b - a) * wg2 for a, b in zip(b2[:-1], b2[1:])])
    B1 = [_basis(p1, U1, u, d) for d in range(3)]
    B2 = [_basis(p2, U2, v, d) for d in range(3)]

    def ev(F, d1, d2):
        return np.einsum("ka,lb,bad->lkd", B1[d1], B2[d2], F)

    # rational undeformed geometry via quotient rule on the homogeneous surface
    H = {(a, b): ev(Hh, a, b) for (a, b) in [(0, 0), (1, 0), (0, 1), (2, 0), (0, 2), (1, 1)]}
    W0 = H[0, 0][..., 3:]
    X = H[0, 0][..., :3] / W0
    X1 = (H[1, 0][..., :3] - H[1, 0][..., 3:] * X) / W0
    X2 = (H[0, 1][..., :3] - H[0, 1][..., 3:] * X) / W0
    X11 = (H[2, 0][..., :3] - H[2, 0][..., 3:] * X - 2 * H[1, 0][..., 3:] * X1) / W0
    X22 = (H[0, 2][..., :3] - H[0, 2][..., 3:] * X - 2 * H[0, 1][..., 3:] * X2) / W0
    X12 = (H[1, 1][..., :3] - H[1, 1][..., 3:] * X - H[1, 0][..., 3:] * X2 - H[0, 1][..., 3:] * X1) / W0
    x1 = X1 + ev(Dd, 1, 0); x2 = X2 + ev(Dd, 0, 1)
    x11 = X11 + ev(Dd, 2, 0); x22 = X22 + ev(Dd, 0, 2); x12 = X12 + ev(Dd, 1, 1)

    def metric(a1, a2, h11, h22, h12):
        nn = np.cross(a1, a2)
        J = np.linalg.norm(nn, axis=-1, keepdims=True)
        nn = nn / J
        acov = np.stack([np.stack([np.sum(a1 * a1, -1), np.sum(a1 * a2, -1)], -1),
                         np.stack([np.sum(a1 * a2, -1), np.sum(a2 * a2, -1)], -1)], -2)
        bcov = np.stack([np.stack([np.sum(h11 * nn, -1), np.sum(h12 * nn, -1)], -1),
                         np.stack([np.sum(h12 * nn, -1), np.sum(h22 * nn, -1)], -1)], -2)
        return acov, bcov, J[..., 0]

    Ac, Bc, JA = metric(X1, X2, X11, X22, X12)
    ac, bc, _ = metric(x1, x2, x11, x22, x12)
    if not prob.bending:
        Bc = Bc * 0; bc = bc * 0
    t = prob.thickness
    Wtot = 0.0
    zg, wz = np.polynomial.legendre.leggauss(prob.num_gauss_thickness)
    for l in range(len(v)):
        for k in range(len(u)):
            wq = wu[k] * wv[l] * JA[l, k]
            A, B, a, b = Ac[l, k], Bc[l, k], ac[l, k], bc[l, k]
            if prob.material == KL_MAT_SVK:
                Ai = np.linalg.inv(A)
                mu = prob.E / (2 * (1 + prob.nu))
                lam = prob.E * prob.nu / ((1 + prob.nu) * (1 - 2 * prob.nu))
                lps = 2 * lam * mu / (lam + 2 * mu)
                Cm = lps * np.einsum("ab,cd->abcd", Ai, Ai) + mu * (np.einsum("ac,bd->abcd", Ai, Ai) + np.einsum("ad,bc->abcd", Ai, Ai))
                eps = 0.5 * (a - A); kap = B - b
                Wtot += wq * 0.5 * (t * np.einsum("ab,abcd,cd", eps, Cm, eps) + t ** 3 / 12 * np.einsum("ab,abcd,cd", kap, Cm, kap))
                continue
            for zz, ww in zip(zg, wz):
                z = 0.5 * t * zz
                Gz = A - 2 * z * B; gz = a - 2 * z * b
                if prob.metric_z2:
                    Gz = Gz + z * z * B @ np.linalg.inv(A) @ B
                    gz = gz + z * z * b @ np.linalg.inv(a) @ b
                if not prob.compressible:
                    psi = _psi_incomp(prob, Gz, gz)
                else:
                    h = 1e-6
                    f = lambda c: (_psi_comp(prob, Gz, gz, c + h) - _psi_comp(prob, Gz, gz, c - h)) / (2 * h)
                    c33 = brentq(f, 0.2, 5.0, xtol=1e-14, rtol=1e-14)
                    psi = _psi_comp(prob, Gz, gz, c33)
                Wtot += wq * 0.5 * t * ww * psi
    return Wtot
