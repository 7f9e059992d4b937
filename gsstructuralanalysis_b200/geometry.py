"""Workload definitions: tensor-product B-spline / NURBS surfaces, refinement, benchmark geometries.

These are the *inputs* of the hot path (what the reference's drivers build before they create
the assembler), not part of it.  numpy only.

Reference anchors
  degreeElevate + uniformRefine order   tutorials/nonlinear_shell_static.cpp:53-57
  Scordelis-Lo shallow roof             filedata/surface/scordelis_lo_roof_shallow.xml:29-52
  roof material / BCs / load            benchmarks/benchmark_Roof.cpp:128-144,201-207,223-232
  FrustrumDomain                        benchmarks/benchmark_Frustrum_APALM.cpp:751-810
  Rectangle + addClamping               benchmarks/benchmark_TensionWrinkling.cpp:579-668
"""
from __future__ import annotations

from dataclasses import dataclass, field
import numpy as np


# --------------------------------------------------------------------------------------
# 1-D B-spline machinery (host-side set-up only)
# --------------------------------------------------------------------------------------
def open_uniform_knots(p: int, nel: int) -> np.ndarray:
    inner = np.linspace(0.0, 1.0, nel + 1)[1:-1]
    return np.concatenate([np.zeros(p + 1), inner, np.ones(p + 1)])


def basis_matrix(p: int, U: np.ndarray, u: np.ndarray) -> np.ndarray:
    """Dense collocation matrix B[k, i] = N_{i,p}(u_k) (Cox-de Boor, right end included)."""
    U = np.asarray(U, dtype=np.float64)
    u = np.asarray(u, dtype=np.float64)
    n = len(U) - p - 1
    m = len(U) - 1
    B = np.zeros((len(u), m))
    for i in range(m):
        if U[i + 1] > U[i]:
            B[:, i] = (u >= U[i]) & (u < U[i + 1])
    # close the right end: the last non-empty span owns u == U[-1]
    last = max(i for i in range(m) if U[i + 1] > U[i])
    B[u >= U[-1], :] = 0.0
    B[u >= U[-1], last] = 1.0
    for q in range(1, p + 1):
        Bn = np.zeros((len(u), m - q))
        for i in range(m - q):
            d1 = U[i + q] - U[i]
            d2 = U[i + q + 1] - U[i + 1]
            t = 0.0
            if d1 > 0:
                t = t + (u - U[i]) / d1 * B[:, i]
            if d2 > 0:
                t = t + (U[i + q + 1] - u) / d2 * B[:, i + 1]
            Bn[:, i] = t
        B = Bn
    return B[:, :n]


def greville(p: int, U: np.ndarray) -> np.ndarray:
    n = len(U) - p - 1
    return np.array([U[i + 1:i + p + 1].sum() / p for i in range(n)])


def elevate_knots(p: int, U: np.ndarray, times: int = 1) -> np.ndarray:
    vals, mult = np.unique(U, return_counts=True)
    return np.repeat(vals, mult + times)


def refine_knots(U: np.ndarray, times: int = 1) -> np.ndarray:
    U = np.asarray(U, dtype=np.float64)
    for _ in range(times):
        vals = np.unique(U)
        mids = 0.5 * (vals[:-1] + vals[1:])
        U = np.sort(np.concatenate([U, mids]))
    return U


# --------------------------------------------------------------------------------------
@dataclass
class Surface:
    """Tensor-product B-spline (weights None) or NURBS surface in R^3.
    cp[i1 + n1*i2] — first parametric direction fastest (G+Smo tensor index)."""
    p: tuple
    U: tuple
    cp: np.ndarray            # [n1*n2, 3]
    w: np.ndarray | None = None
    name: str = "surface"

    @property
    def n(self):
        return (len(self.U[0]) - self.p[0] - 1, len(self.U[1]) - self.p[1] - 1)

    def homogeneous(self):
        w = np.ones(len(self.cp)) if self.w is None else self.w
        return np.concatenate([self.cp * w[:, None], w[:, None]], axis=1)

    def evaluate(self, u, v):
        """Points on the tensor grid u x v -> [len(v), len(u), 3]."""
        B1 = basis_matrix(self.p[0], self.U[0], u)
        B2 = basis_matrix(self.p[1], self.U[1], v)
        n1, n2 = self.n
        H = self.homogeneous().reshape(n2, n1, 4)
        P = np.einsum("ka,lb,bad->lkd", B1, B2, H)
        return P[..., :3] / P[..., 3:4]

    def respace(self, p_new, U_new, name=None):
        """Exact re-expression in a finer/higher-degree spline space that contains this one
        (= degreeElevate / uniformRefine / knot insertion), by interpolation of the homogeneous
        coordinates at the Greville abscissae of the new space."""
        g1, g2 = greville(p_new[0], U_new[0]), greville(p_new[1], U_new[1])
        B1o = basis_matrix(self.p[0], self.U[0], g1)
        B2o = basis_matrix(self.p[1], self.U[1], g2)
        n1, n2 = self.n
        H = self.homogeneous().reshape(n2, n1, 4)
        F = np.einsum("ka,lb,bad->lkd", B1o, B2o, H)          # [m2, m1, 4]
        B1n = basis_matrix(p_new[0], U_new[0], g1)
        B2n = basis_matrix(p_new[1], U_new[1], g2)
        T = np.linalg.solve(B1n, F.transpose(1, 0, 2).reshape(len(g1), -1)).reshape(len(g1), len(g2), 4)
        C = np.linalg.solve(B2n, T.transpose(1, 0, 2).reshape(len(g2), -1)).reshape(len(g2), len(g1), 4)
        C = C.reshape(-1, 4)
        if self.w is None:
            return Surface(tuple(p_new), (np.asarray(U_new[0]), np.asarray(U_new[1])), C[:, :3].copy(), None,
                           name or self.name)
        w = C[:, 3].copy()
        return Surface(tuple(p_new), (np.asarray(U_new[0]), np.asarray(U_new[1])), C[:, :3] / w[:, None], w,
                       name or self.name)

    def degree_elevate(self, times=1):
        if times == 0:
            return self
        p = (self.p[0] + times, self.p[1] + times)
        U = (elevate_knots(self.p[0], self.U[0], times), elevate_knots(self.p[1], self.U[1], times))
        return self.respace(p, U)

    def uniform_refine(self, times=1):
        if times == 0:
            return self
        return self.respace(self.p, (refine_knots(self.U[0], times), refine_knots(self.U[1], times)))

    def refine_to(self, nel1, nel2=None):
        """Uniform n x n elements (non-dyadic sizes, e.g. 576 for the 1M-DOF case); the source must
        have a single element per direction."""
        nel2 = nel1 if nel2 is None else nel2
        return self.respace(self.p, (open_uniform_knots(self.p[0], nel1), open_uniform_knots(self.p[1], nel2)))


    def insert_knot(self, direction, u, multiplicity):
        """Raise the multiplicity of the knot u in one direction to `multiplicity` by Boehm's knot insertion on the homogeneous
        control net (exact, the surface does not change): multiplicity = p gives a C0 line, p + 1 separates the two sides."""
        p = self.p[direction]
        U = np.asarray(self.U[direction], dtype=np.float64).copy()
        n1, n2 = self.n
        H = self.homogeneous().reshape(n2, n1, 4)
        H = H.transpose(1, 0, 2).copy() if direction == 0 else H.copy()     # insertion axis first
        have = int(np.sum(np.abs(U - u) < 1e-14))
        for _ in range(max(0, multiplicity - have)):
            k = int(np.searchsorted(U, u, side="right")) - 1               # U[k] <= u < U[k+1]
            n = H.shape[0]
            Q = np.empty((n + 1,) + H.shape[1:])
            for i in range(n + 1):
                if i <= k - p:
                    Q[i] = H[i]
                elif i >= k + 1:
                    Q[i] = H[i - 1]
                else:
                    a = (u - U[i]) / (U[i + p] - U[i])
                    Q[i] = a * H[i] + (1.0 - a) * H[i - 1]
            H = Q
            U = np.insert(U, k + 1, float(u))
        H = H.transpose(1, 0, 2) if direction == 0 else H
        C = np.ascontiguousarray(H).reshape(-1, 4)
        UU = (U, np.asarray(self.U[1])) if direction == 0 else (np.asarray(self.U[0]), U)
        if self.w is None:
            return Surface(self.p, UU, C[:, :3].copy(), None, self.name)
        w = C[:, 3].copy()
        return Surface(self.p, UU, C[:, :3] / w[:, None], w, self.name)

    def split(self, direction, u):
        """The two patches left / right (below / above) of the parameter line u: same parametrisation, each with an open knot
        vector; their facing sides carry identical control points (a conforming C0 interface)."""
        s = self.insert_knot(direction, u, self.p[direction] + 1)
        p = s.p[direction]
        U = s.U[direction]
        k = int(np.searchsorted(U, u, side="left"))          # first of the p + 1 copies of u
        n1, n2 = s.n
        cp = s.cp.reshape(n2, n1, 3)
        w = None if s.w is None else s.w.reshape(n2, n1)
        Ua, Ub = U[:k + p + 1], U[k:]
        na = len(Ua) - p - 1
        parts = []
        for lo, hi, Ud in ((0, na, Ua), (na, (n1 if direction == 0 else n2), Ub)):
            if direction == 0:
                c, ww, UU = cp[:, lo:hi], (None if w is None else w[:, lo:hi]), (Ud, s.U[1])
            else:
                c, ww, UU = cp[lo:hi, :], (None if w is None else w[lo:hi, :]), (s.U[0], Ud)
            parts.append(Surface(s.p, (np.array(UU[0]), np.array(UU[1])), np.ascontiguousarray(c).reshape(-1, 3).copy(),
                                 None if ww is None else np.ascontiguousarray(ww).reshape(-1).copy(), self.name))
        return parts[0], parts[1]


def split_grid(surface, cuts1, cuts2):
    """Cut a surface into a grid of conforming patches along the parameter lines cuts1 (first direction) x cuts2 (second):
    returns (patches, interfaces) with patches ordered first direction fastest and interfaces as
    (patch0, side0, patch1, side1, reversed) tuples in G+Smo side numbering (west 0, east 1, south 2, north 3)."""
    cols = [surface]
    for u in sorted(cuts1):
        a, b = cols[-1].split(0, u)
        cols[-1:] = [a, b]
    rows = []
    for c in cols:
        col = [c]
        for v in sorted(cuts2):
            a, b = col[-1].split(1, v)
            col[-1:] = [a, b]
        rows.append(col)
    m1, m2 = len(cols), len(cuts2) + 1
    patches = [rows[i][j] for j in range(m2) for i in range(m1)]
    interfaces = []
    for j in range(m2):
        for i in range(m1):
            q = i + m1 * j
            if i + 1 < m1:
                interfaces.append((q, 1, q + 1, 0, 0))          # east of q meets west of its right neighbour
            if j + 1 < m2:
                interfaces.append((q, 3, q + m1, 2, 0))         # north of q meets south of the patch above
    return patches, interfaces


# --------------------------------------------------------------------------------------
# benchmark geometries
# --------------------------------------------------------------------------------------
def _grid(xs, ys, zfun):
    pts = []
    for j, y in enumerate(ys):
        for i, x in enumerate(xs):
            pts.append((x, y, zfun(i, j)))
    return np.array(pts, dtype=np.float64)


def plate(L=1.0, W=1.0):
    U = np.array([0, 0, 1, 1.0])
    return Surface((1, 1), (U, U.copy()), _grid([0, L], [0, W], lambda i, j: 0.0), None, "plate")


def paraboloid(c=0.25):
    """Stand-in for upstream surfaces/paraboloid.xml (not in the reference tree):
    biquadratic, z = 4c u(1-u) * 4 v(1-v) / 4."""
    U = np.array([0, 0, 0, 1, 1, 1.0])
    return Surface((2, 2), (U, U.copy()),
                   _grid([0, 0.5, 1], [0, 0.5, 1], lambda i, j: (4 * c if (i == 1 and j == 1) else 0.0)),
                   None, "paraboloid")


def scordelis_lo_roof_shallow():
    """filedata/surface/scordelis_lo_roof_shallow.xml:29-52 (degree 2x2, 9 control points)."""
    U = np.array([0, 0, 0, 1, 1, 1.0])
    cp = np.array([[0, 0, 0], [254, 0, 0], [508, 0, 0],
                   [0, -253.577, 25.443], [254, -253.577, 25.443], [508, -253.577, 25.443],
                   [0, -507.154, 0], [254, -507.154, 0], [508, -507.154, 0]], dtype=np.float64)
    return Surface((2, 2), (U, U.copy()), cp, None, "scordelis_lo_roof_shallow")


def scordelis_lo_roof_classic(R=25.0, L=50.0, phi_deg=40.0):
    """Textbook Scordelis-Lo roof as an exact NURBS (degree 1 x 2; cf. the commented block at
    filedata/surface/scordelis_lo_roof_shallow.xml:1-27).  u: length, v: arc from -phi to +phi."""
    phi = np.deg2rad(phi_deg)
    U1 = np.array([0, 0, 1, 1.0])
    U2 = np.array([0, 0, 0, 1, 1, 1.0])
    wm = np.cos(phi)
    ys = [-R * np.sin(phi), 0.0, R * np.sin(phi)]
    zs = [R * np.cos(phi), R / np.cos(phi), R * np.cos(phi)]
    cp, w = [], []
    for j in range(3):
        for x in (0.0, L):
            cp.append((x, ys[j], zs[j]))
            w.append(wm if j == 1 else 1.0)
    return Surface((1, 2), (U1, U2), np.array(cp), np.array(w), "scordelis_lo_roof_classic")


def eighth_sphere(R=10.0):
    """Exact NURBS octant of a sphere (degree 2x2; same construction as
    filedata/surface/eighth_sphere.xml used by benchmarks/benchmark_Balloon.cpp:110): a quarter
    circle in the x-z plane revolved by 90 degrees about z.  The pole is a degenerate edge."""
    s = 1.0 / np.sqrt(2.0)
    U = np.array([0, 0, 0, 1, 1, 1.0])
    prof = [(R, 0.0, 1.0), (R, R, s), (0.0, R, 1.0)]          # (radius, z, weight): equator -> pole
    rev = [((1.0, 0.0), 1.0), ((1.0, 1.0), s), ((0.0, 1.0), 1.0)]
    cp, w = [], []
    for (rad, z, wp) in prof:               # second direction: meridian
        for ((cx, cy), wr) in rev:          # first direction: revolution
            cp.append((rad * cx, rad * cy, z))
            w.append(wp * wr)
    return Surface((2, 2), (U, U.copy()), np.array(cp), np.array(w), "eighth_sphere")


def frustrum(R1=2.0, R2=1.0, h=1.0):
    """Quarter conical frustrum as an exact NURBS, degree 2 (angle) x 1 (height)
    (benchmarks/benchmark_Frustrum_APALM.cpp:751-810: quarter circle with weight 0.70711)."""
    s = 1.0 / np.sqrt(2.0)
    U1 = np.array([0, 0, 0, 1, 1, 1.0])
    U2 = np.array([0, 0, 1, 1.0])
    cp, w = [], []
    for (rad, z) in ((R1, 0.0), (R2, h)):
        for ((cx, cy), wr) in (((1.0, 0.0), 1.0), ((1.0, 1.0), s), ((0.0, 1.0), 1.0)):
            cp.append((rad * cx, rad * cy, z))
            w.append(wr)
    return Surface((2, 1), (U1, U2), np.array(cp), np.array(w), "frustrum")


def half_cylinder():
    """filedata/surface/half_cylinder.xml of the reference (benchmarks/benchmark_Cylinder.cpp:70): NURBS, degree 2x2, knots
    [0 0 0 1 1 1] x [0 0 0 0.5 1 1 1], 3 x 4 control points (first direction = axis, fastest), the active (uncommented)
    geometry of the file."""
    U1 = np.array([0, 0, 0, 1, 1, 1.0])
    U2 = np.array([0, 0, 0, 0.5, 1, 1, 1.0])
    w = np.array([1, 1, 1, 0.5, 0.5, 0.5, 0.5, 0.5, 0.5, 1, 1, 1.0])
    cp = np.array([[0.00, 0.00, -0.09], [0.075, 0.00, -0.09], [0.15, 0.00, -0.09],
                   [0.00, 0.09, -0.09], [0.075, 0.09, -0.09], [0.15, 0.09, -0.09],
                   [0.00, 0.09, 0.09], [0.075, 0.09, 0.09], [0.15, 0.09, 0.09],
                   [0.00, 0.00, 0.09], [0.075, 0.00, 0.09], [0.15, 0.00, 0.09]])
    return Surface((2, 2), (U1, U2), cp, w, "half_cylinder")


def rectangle_with_clamping(L=0.14, B=0.07, p=3, nel1=8, nel2=8, clamp=1e-2):
    """Rectangle(L,B) with extra knots at `clamp` from the west/east edges
    (benchmarks/benchmark_TensionWrinkling.cpp:160-205,579-617) -> non-uniform knot vector."""
    base = plate(L, B).degree_elevate(p - 1)
    U1 = open_uniform_knots(p, nel1)
    U1 = np.sort(np.concatenate([U1, [clamp, 1.0 - clamp]]))
    U2 = open_uniform_knots(p, nel2)
    return base.respace((p, p), (U1, U2), name="tension_sheet")
