"""Independent numpy evaluation of the kinematic post-processing quantities of the KL shell (include/kl_shell.h, SURVEY 8f
rank 4): scipy BSpline basis, full 3-D tensors, numpy.linalg.eigh.  Test infrastructure only.  Shares no code with
oracle/kl_oracle.c or the CUDA kernels; the through-thickness metric convention (g_ab - 2 z b_ab [+ z^2 b_ac a^cd b_db]) is
the assembly's own."""
import numpy as np

from tests.energy_model import _basis


def _disp_net(prob, x):
    n1, n2 = prob.surface.n
    ncp = n1 * n2
    disp = np.zeros((ncp, 3))
    fv = prob.fixed_values if prob.fixed_values is not None else np.zeros(max(prob.n_fixed, 1))
    for c in range(3):
        g = prob.dof_map[c * ncp:(c + 1) * ncp]
        free = g < prob.n_free
        disp[free, c] = x[g[free]]
        disp[~free, c] = fv[g[~free] - prob.n_free]
    return disp


def _frames(prob, x, uv):
    """Per point: undeformed / deformed tangents and second derivatives, displacement."""
    s = prob.surface
    (p1, p2), (U1, U2), (n1, n2) = s.p, s.U, s.n
    ncp = n1 * n2
    w = np.ones(ncp) if s.w is None else s.w
    Hh = np.concatenate([s.cp * w[:, None], w[:, None]], 1).reshape(n2, n1, 4)
    Dd = _disp_net(prob, x).reshape(n2, n1, 3)
    out = []
    for (u, v) in uv:
        B1 = [_basis(p1, U1, np.array([u]), d)[0] for d in range(3)]
        B2 = [_basis(p2, U2, np.array([v]), d)[0] for d in range(3)]

        def ev(F, d1, d2):
            return np.einsum("a,b,bad->d", B1[d1], B2[d2], F)

        H = {k: ev(Hh, *k) for k in [(0, 0), (1, 0), (0, 1), (2, 0), (0, 2), (1, 1)]}
        W0 = H[0, 0][3]
        X = H[0, 0][:3] / W0
        X1 = (H[1, 0][:3] - H[1, 0][3] * X) / W0
        X2 = (H[0, 1][:3] - H[0, 1][3] * X) / W0
        X11 = (H[2, 0][:3] - H[2, 0][3] * X - 2 * H[1, 0][3] * X1) / W0
        X22 = (H[0, 2][:3] - H[0, 2][3] * X - 2 * H[0, 1][3] * X2) / W0
        X12 = (H[1, 1][:3] - H[1, 1][3] * X - H[1, 0][3] * X2 - H[0, 1][3] * X1) / W0
        out.append(dict(A=(X1, X2), H=(X11, X22, X12), a=(X1 + ev(Dd, 1, 0), X2 + ev(Dd, 0, 1)),
                        h=(X11 + ev(Dd, 2, 0), X22 + ev(Dd, 0, 2), X12 + ev(Dd, 1, 1)), u=ev(Dd, 0, 0)))
    return out


def _surface(t1, t2, hh, bending=True):
    n = np.cross(t1, t2)
    n = n / np.linalg.norm(n)
    cov = np.array([[t1 @ t1, t1 @ t2], [t1 @ t2, t2 @ t2]])
    cur = np.array([[hh[0] @ n, hh[2] @ n], [hh[2] @ n, hh[1] @ n]]) * (1.0 if bending else 0.0)
    con = np.linalg.inv(cov)
    up = [con[0, 0] * t1 + con[0, 1] * t2, con[1, 0] * t1 + con[1, 1] * t2]
    return n, cov, cur, con, up


def kinematics(prob, x, uv, z):
    res = dict(stretch=[], dirs=[], normal=[], disp=[], Em=[], Ef=[], Em_p=[], Ef_p=[])
    for f in _frames(prob, x, uv):
        N, Ac, Bc, Ai, Au = _surface(*f["A"], f["H"], prob.bending)
        n, ac, bc, ai, au = _surface(*f["a"], f["h"], prob.bending)
        # metrics at height z (assembly convention) and base vectors g_a(z) = a_a - z b_a^c a_c
        Gz = Ac - 2 * z * Bc + (z * z * Bc @ Ai @ Bc if prob.metric_z2 else 0.0)
        gz = ac - 2 * z * bc + (z * z * bc @ ai @ bc if prob.metric_z2 else 0.0)
        # generalised symmetric eigenproblem g v = lam^2 G v through G^-1/2
        wG, VG = np.linalg.eigh(Gz)
        Gmh = VG @ np.diag(wG ** -0.5) @ VG.T
        lam2, W = np.linalg.eigh(Gmh @ gz @ Gmh)
        V = Gmh @ W                                  # columns: contravariant components of the material directions
        bmix = bc @ ai
        gvec = [f["a"][0] - z * (bmix[0, 0] * f["a"][0] + bmix[0, 1] * f["a"][1]),
                f["a"][1] - z * (bmix[1, 0] * f["a"][0] + bmix[1, 1] * f["a"][1])]
        dirs = []
        for i in range(2):
            d = V[0, i] * gvec[0] + V[1, i] * gvec[1]
            dirs.append(d / np.linalg.norm(d))
        res["stretch"].append(np.sqrt(lam2))
        res["dirs"].append(dirs)
        res["normal"].append(n)
        res["disp"].append(f["u"])
        # 3-D strain tensors on the undeformed frame E1 = A_1/|A_1|, E2 = N x E1
        E1 = f["A"][0] / np.linalg.norm(f["A"][0])
        E2 = np.cross(N, E1)
        Q = np.array([[Au[0] @ E1, Au[1] @ E1], [Au[0] @ E2, Au[1] @ E2]])   # Q[i, alpha] = A^alpha . E_i
        Em = Q @ (0.5 * (ac - Ac)) @ Q.T
        Ef = Q @ (Bc - bc) @ Q.T
        res["Em"].append([Em[0, 0], Em[1, 1], Em[0, 1]])
        res["Ef"].append([Ef[0, 0], Ef[1, 1], Ef[0, 1]])
        res["Em_p"].append(np.linalg.eigvalsh(Em))
        res["Ef_p"].append(np.linalg.eigvalsh(Ef))
    return {k: np.array(v) for k, v in res.items()}


def cauchy_from_resultants(prob, x, uv, Nres, Mres, lam3):
    """sigma = F S F^T / det F with S = N^ab / t A_a (x) A_b (membrane) and 6 M^ab / t^2 (outer-fibre bending), projected on
    the deformed frame e1 = a_1/|a_1|, e2 = n x e1."""
    t = prob.thickness
    sm, sf = [], []
    for f, Nv, Mv, l3 in zip(_frames(prob, x, uv), Nres, Mres, lam3):
        N, Ac, Bc, Ai, Au = _surface(*f["A"], f["H"], prob.bending)
        n, ac, bc, ai, au = _surface(*f["a"], f["h"], prob.bending)
        F = np.outer(f["a"][0], Au[0]) + np.outer(f["a"][1], Au[1]) + l3 * np.outer(n, N)
        J = np.linalg.det(F)
        e1 = f["a"][0] / np.linalg.norm(f["a"][0])
        e2 = np.cross(n, e1)
        row = []
        for R, scale in ((Nv, 1.0 / t), (Mv, 6.0 / (t * t))):
            S = np.zeros((3, 3))
            comp = np.array([[R[0], R[2]], [R[2], R[1]]])
            for a in range(2):
                for b in range(2):
                    S += scale * comp[a, b] * np.outer(f["A"][a], f["A"][b])
            sig = F @ S @ F.T / J
            row.append([e1 @ sig @ e1, e2 @ sig @ e2, e1 @ sig @ e2])
        sm.append(row[0])
        sf.append(row[1])
    return np.array(sm), np.array(sf)


def svk_resultants(prob, x, uv):
    """gsMaterialMatrixLinear: N^ab = t C^abcd E_cd, M^ab = t^3/12 C^abcd K_cd, C^abcd = lam_ps A^ab A^cd + mu (A^ac A^bd + A^ad A^bc)."""
    t, E, nu = prob.thickness, prob.E, prob.nu
    mu = E / (2 * (1 + nu))
    lam = E * nu / ((1 + nu) * (1 - 2 * nu))
    lps = 2 * lam * mu / (lam + 2 * mu)
    Ns, Ms = [], []
    for f in _frames(prob, x, uv):
        N, Ac, Bc, Ai, Au = _surface(*f["A"], f["H"], prob.bending)
        n, ac, bc, ai, au = _surface(*f["a"], f["h"], prob.bending)
        C = lps * np.einsum("ab,cd->abcd", Ai, Ai) + mu * (np.einsum("ac,bd->abcd", Ai, Ai) + np.einsum("ad,bc->abcd", Ai, Ai))
        Nt = t * np.einsum("abcd,cd->ab", C, 0.5 * (ac - Ac))
        Mt = t ** 3 / 12 * np.einsum("abcd,cd->ab", C, Bc - bc)
        Ns.append([Nt[0, 0], Nt[1, 1], Nt[0, 1]])
        Ms.append([Mt[0, 0], Mt[1, 1], Mt[0, 1]])
    return np.array(Ns), np.array(Ms)
