"""Problem description shared by the C ABI (include/kl_shell.h: kl_problem, kl_bc).

Mirrors what a reference driver hands to gsThinShellAssembler<3,real_t,true>(ori,bases,bc,force,
materialMatrix) + setPointLoads/setPressure (tutorials/nonlinear_shell_static.cpp:62-114).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
import numpy as np

from .geometry import Surface

KL_MAT_SVK, KL_MAT_NH, KL_MAT_MR = 0, 1, 3
KL_BC_FREE, KL_BC_DIRICHLET, KL_BC_CLAMPED, KL_BC_COLLAPSED = 0, 1, 2, 3
WEST, EAST, SOUTH, NORTH = 0, 1, 2, 3
SW, SE, NW, NE = 0, 1, 2, 3

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int32)


class kl_bc(C.Structure):
    _fields_ = [("side", (C.c_int32 * 3) * 4), ("corner", (C.c_int32 * 3) * 4)]


class kl_interface(C.Structure):
    _fields_ = [("patch", C.c_int32 * 2), ("side", C.c_int32 * 2), ("reversed", C.c_int32)]


class kl_problem(C.Structure):
    _fields_ = [
        ("degree", C.c_int32 * 2),
        ("n_knots", C.c_int32 * 2),
        ("knots", c_double_p * 2),
        ("cp", c_double_p),
        ("weights", c_double_p),
        ("dof_map", c_int_p),
        ("n_free", C.c_int32),
        ("n_fixed", C.c_int32),
        ("fixed_values", c_double_p),
        ("material", C.c_int32),
        ("compressible", C.c_int32),
        ("num_gauss_thickness", C.c_int32),
        ("bending", C.c_int32),
        ("E", C.c_double),
        ("nu", C.c_double),
        ("thickness", C.c_double),
        ("mr_ratio", C.c_double),
        ("metric_z2", C.c_int32),
        ("quA", C.c_int32),
        ("quB", C.c_int32),
        ("body_force", C.c_double * 3),
        ("pressure", C.c_double),
        ("n_point_loads", C.c_int32),
        ("point_load_uv", c_double_p),
        ("point_load_val", c_double_p),
        ("n_neumann", C.c_int32),
        ("neumann_side", c_int_p),
        ("neumann_val", c_double_p),
    ]


@dataclass
class BoundaryConditions:
    side: np.ndarray = field(default_factory=lambda: np.zeros((4, 3), dtype=np.int32))
    corner: np.ndarray = field(default_factory=lambda: np.zeros((4, 3), dtype=np.int32))

    def add_condition(self, side, kind, comp=-1):
        """BCs.addCondition(boundary::side, condition_type::kind, 0, 0, false, comp)"""
        for c in (range(3) if comp < 0 else [comp]):
            self.side[side, c] = kind
        return self

    def add_corner_value(self, corner, comp=-1):
        """bc.addCornerValue(corner, 0.0, patch, comp)  (tutorials/nonlinear_shell_static.cpp:71-74)"""
        for c in (range(3) if comp < 0 else [comp]):
            self.corner[corner, c] = 1
        return self

    def to_c(self):
        b = kl_bc()
        for s in range(4):
            for c in range(3):
                b.side[s][c] = int(self.side[s, c])
                b.corner[s][c] = int(self.corner[s, c])
        return b


@dataclass
class ShellProblem:
    surface: Surface
    bc: BoundaryConditions = field(default_factory=BoundaryConditions)
    material: int = KL_MAT_SVK
    compressible: bool = False
    num_gauss_thickness: int = 4
    bending: bool = True
    E: float = 1.0
    nu: float = 0.0
    thickness: float = 1.0
    mr_ratio: float = 7.0
    metric_z2: bool = False
    quA: int = 1
    quB: int = 1
    body_force: tuple = (0.0, 0.0, 0.0)
    pressure: float = 0.0
    point_loads: list = field(default_factory=list)   # [((u,v),(fx,fy,fz)), ...]
    neumann: list = field(default_factory=list)       # [(side, (tx,ty,tz)), ...]  BCs.addCondition(side, condition_type::neumann, &neuData)
    # filled by number_dofs()
    dof_map: np.ndarray | None = None
    n_free: int = 0
    n_fixed: int = 0
    fixed_values: np.ndarray | None = None

    def number_dofs(self, build_dofmap_fn):
        """build_dofmap_fn: the C symbol kl_build_dofmap (product) or klo_build_dofmap (oracle)."""
        n1, n2 = self.surface.n
        m = np.zeros(3 * n1 * n2, dtype=np.int32)
        nf, nx = C.c_int32(0), C.c_int32(0)
        bc = self.bc.to_c()
        rc = build_dofmap_fn(n1, n2, C.byref(bc), m.ctypes.data_as(c_int_p), C.byref(nf), C.byref(nx))
        if rc != 0:
            raise RuntimeError(f"build_dofmap failed rc={rc}")
        self.dof_map, self.n_free, self.n_fixed = m, nf.value, nx.value
        if self.fixed_values is None or len(self.fixed_values) != self.n_fixed:
            self.fixed_values = np.zeros(self.n_fixed)
        return self

    def to_c(self):
        """Returns (kl_problem, keepalive list)."""
        s = self.surface
        assert self.dof_map is not None, "call number_dofs first"
        keep = []

        def dp(a):
            a = np.ascontiguousarray(a, dtype=np.float64)
            keep.append(a)
            return a.ctypes.data_as(c_double_p)

        P = kl_problem()
        P.degree[0], P.degree[1] = s.p
        P.n_knots[0], P.n_knots[1] = len(s.U[0]), len(s.U[1])
        P.knots[0], P.knots[1] = dp(s.U[0]), dp(s.U[1])
        P.cp = dp(s.cp.reshape(-1))
        P.weights = dp(s.w) if s.w is not None else None
        m = np.ascontiguousarray(self.dof_map, dtype=np.int32)
        keep.append(m)
        P.dof_map = m.ctypes.data_as(c_int_p)
        P.n_free, P.n_fixed = self.n_free, self.n_fixed
        P.fixed_values = dp(self.fixed_values) if self.n_fixed > 0 else None
        P.material, P.compressible = int(self.material), int(self.compressible)
        P.num_gauss_thickness, P.bending = int(self.num_gauss_thickness), int(self.bending)
        P.E, P.nu, P.thickness, P.mr_ratio = self.E, self.nu, self.thickness, self.mr_ratio
        P.metric_z2, P.quA, P.quB = int(self.metric_z2), self.quA, self.quB
        for k in range(3):
            P.body_force[k] = float(self.body_force[k])
        P.pressure = float(self.pressure)
        P.n_point_loads = len(self.point_loads)
        if self.point_loads:
            P.point_load_uv = dp(np.array([pl[0] for pl in self.point_loads]).reshape(-1))
            P.point_load_val = dp(np.array([pl[1] for pl in self.point_loads]).reshape(-1))
        P.n_neumann = len(self.neumann)
        if self.neumann:
            sides = np.ascontiguousarray([nm[0] for nm in self.neumann], dtype=np.int32)
            keep.append(sides)
            P.neumann_side = sides.ctypes.data_as(c_int_p)
            P.neumann_val = dp(np.array([nm[1] for nm in self.neumann]).reshape(-1))
        return P, keep

    def save(self, path):
        """Binary dump read by examples/newton_shell.cpp (little endian):
        'KLP1', 16 int32 header, 8 doubles, then the arrays in the order of kl_problem; Neumann sides follow at the end
        (int32 count, sides, traction vectors) and are optional for readers of the older layout."""
        import struct
        s = self.surface
        assert self.dof_map is not None
        n1, n2 = s.n
        ncp = n1 * n2
        with open(path, "wb") as f:
            f.write(b"KLP1")
            f.write(struct.pack("<16i", s.p[0], s.p[1], len(s.U[0]), len(s.U[1]), ncp, int(s.w is not None), self.n_free,
                                self.n_fixed, int(self.material), int(self.compressible), int(self.num_gauss_thickness),
                                int(self.bending), int(self.metric_z2), self.quA, self.quB, len(self.point_loads)))
            f.write(struct.pack("<8d", self.E, self.nu, self.thickness, self.mr_ratio, *[float(v) for v in self.body_force],
                                float(self.pressure)))
            for a in (s.U[0], s.U[1], s.cp.reshape(-1)):
                f.write(np.ascontiguousarray(a, dtype="<f8").tobytes())
            if s.w is not None:
                f.write(np.ascontiguousarray(s.w, dtype="<f8").tobytes())
            f.write(np.ascontiguousarray(self.dof_map, dtype="<i4").tobytes())
            f.write(np.ascontiguousarray(self.fixed_values if self.n_fixed else np.zeros(0), dtype="<f8").tobytes())
            if self.point_loads:
                f.write(np.array([pl[0] for pl in self.point_loads], dtype="<f8").tobytes())
                f.write(np.array([pl[1] for pl in self.point_loads], dtype="<f8").tobytes())
            f.write(struct.pack("<i", len(self.neumann)))
            if self.neumann:
                f.write(np.array([nm[0] for nm in self.neumann], dtype="<i4").tobytes())
                f.write(np.array([nm[1] for nm in self.neumann], dtype="<f8").tobytes())


@dataclass
class MultiPatchProblem:
    """Several conforming patches glued C0 along whole sides: a gsMultiPatch with computeTopology() / addInterface()
    (benchmarks/benchmark_Wrinkling.cpp:446-522) + one ShellProblem (material, loads, boundary conditions) per patch.
    interfaces: [(patch0, side0, patch1, side1, reversed), ...]."""
    patches: list
    interfaces: list = field(default_factory=list)
    n_free: int = 0
    n_fixed: int = 0

    def number_dofs(self, build_dofmap_mp_fn):
        """build_dofmap_mp_fn: the C symbol kl_mp_build_dofmap (product) or a callable with the same signature (oracle)."""
        npatch = len(self.patches)
        n1 = np.ascontiguousarray([p.surface.n[0] for p in self.patches], dtype=np.int32)
        n2 = np.ascontiguousarray([p.surface.n[1] for p in self.patches], dtype=np.int32)
        bcs = (kl_bc * npatch)(*[p.bc.to_c() for p in self.patches])
        ifs = (kl_interface * max(len(self.interfaces), 1))()
        for k, (pa, sa, pb, sb, rev) in enumerate(self.interfaces):
            ifs[k].patch[0], ifs[k].patch[1], ifs[k].side[0], ifs[k].side[1], ifs[k].reversed = pa, pb, sa, sb, int(rev)
        total = int(3 * np.sum(n1.astype(np.int64) * n2))
        m = np.zeros(total, dtype=np.int32)
        nf, nx = C.c_int32(0), C.c_int32(0)
        rc = build_dofmap_mp_fn(npatch, n1.ctypes.data_as(c_int_p), n2.ctypes.data_as(c_int_p), bcs, len(self.interfaces), ifs,
                                m.ctypes.data_as(c_int_p), C.byref(nf), C.byref(nx))
        if rc != 0:
            raise RuntimeError(f"build_dofmap_mp failed rc={rc}")
        self.set_dof_maps(m, nf.value, nx.value)
        return self

    def set_dof_maps(self, m, n_free, n_fixed, fixed_values=None):
        self.n_free, self.n_fixed = int(n_free), int(n_fixed)
        fv = np.zeros(self.n_fixed) if fixed_values is None else np.ascontiguousarray(fixed_values, dtype=np.float64)
        off = 0
        for p in self.patches:
            ncp = p.surface.n[0] * p.surface.n[1]
            p.dof_map = np.ascontiguousarray(m[off:off + 3 * ncp], dtype=np.int32)
            p.n_free, p.n_fixed, p.fixed_values = self.n_free, self.n_fixed, fv
            off += 3 * ncp
        return self

    def to_c(self):
        arr = (kl_problem * len(self.patches))()
        keep = []
        for k, p in enumerate(self.patches):
            P, kp = p.to_c()
            arr[k] = P
            keep.append((P, kp))
        return arr, keep
