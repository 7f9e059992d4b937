"""Host side of the gsElasticity solid path: trivariate B-spline volumes, the ks_problem structure
of include/ks_solid.h and SolidAssembler, the mirror of the solid closures of
tutorials/nonlinear_solid_static.cpp:101-114.  All arithmetic of the path happens in libkl_shell.so on the GPU."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import capi
from . import geometry as G
from .ops import SparseView
from .problem import c_double_p, c_int_p

KS_LAW_HOOKE, KS_LAW_SVK, KS_LAW_NEO_HOOKE_LN, KS_LAW_NEO_HOOKE_QUAD = 0, 1, 2, 3
KS_WEST, KS_EAST, KS_SOUTH, KS_NORTH, KS_FRONT, KS_BACK = range(6)


class ks_bc(C.Structure):
    _fields_ = [("side", (C.c_int32 * 3) * 6), ("corner", (C.c_int32 * 3) * 8)]


class ks_problem(C.Structure):
    _fields_ = [("degree", C.c_int32 * 3), ("n_knots", C.c_int32 * 3), ("knots", c_double_p * 3), ("cp", c_double_p),
                ("weights", c_double_p), ("dof_map", c_int_p), ("n_free", C.c_int32), ("n_fixed", C.c_int32),
                ("fixed_values", c_double_p), ("material_law", C.c_int32), ("E", C.c_double), ("nu", C.c_double),
                ("body_force", C.c_double * 3), ("n_tractions", C.c_int32), ("traction_side", c_int_p),
                ("traction_val", c_double_p)]


@dataclass
class Volume:
    """Tensor-product B-spline volume; control points numbered i1 + n1*(i2 + n2*i3)."""
    p: tuple
    U: tuple
    cp: np.ndarray          # [ncp, 3]

    @property
    def n(self):
        return tuple(len(self.U[d]) - self.p[d] - 1 for d in range(3))

    @staticmethod
    def from_function(fun, degrees, nels):
        """Interpolate x(u,v,w) at the Greville points (exact when x is a polynomial of the given degrees)."""
        U = [G.open_uniform_knots(p, n) for p, n in zip(degrees, nels)]
        gr = [G.greville(p, u) for p, u in zip(degrees, U)]
        B = [G.basis_matrix(p, u, g) for p, u, g in zip(degrees, U, gr)]
        g3, g2, g1 = np.meshgrid(gr[2], gr[1], gr[0], indexing="ij")
        X = np.stack(fun(g1, g2, g3), axis=-1)                 # [n3, n2, n1, 3]
        for axis, Bm in ((0, B[2]), (1, B[1]), (2, B[0])):
            Xm = np.moveaxis(X, axis, 0)
            X = np.moveaxis(np.linalg.solve(Bm, Xm.reshape(Bm.shape[0], -1)).reshape(Xm.shape), 0, axis)
        return Volume(tuple(degrees), tuple(U), np.ascontiguousarray(X.reshape(-1, 3)))


def brick(L=1.0, B=0.01, H=0.01, degrees=(2, 2, 2), nels=(8, 1, 1)):
    """BrickDomain(n, m, o, p, q, r, L, B, H) of benchmarks/benchmark_Elasticity_Beam_APALM.cpp:204."""
    return Volume.from_function(lambda u, v, w: (L * u, B * v, H * w), degrees, nels)


def paraboloid_volume(c=0.25, t=0.05, degrees=(3, 3, 2), nels=(4, 4, 1)):
    """A thick paraboloid like paraboloid_volume.xml of tutorials/nonlinear_solid_static.cpp:49 (that file is upstream
    filedata and not in the reference tree): mid-surface z = c(1-(2u-1)^2)(1-(2v-1)^2) on the unit square, thickness t."""
    return Volume.from_function(lambda u, v, w: (u, v, c * (1 - (2 * u - 1) ** 2) * (1 - (2 * v - 1) ** 2) + t * (w - 0.5)),
                                degrees, nels)


@dataclass
class SolidBC:
    side: np.ndarray = field(default_factory=lambda: np.zeros((6, 3), dtype=np.int32))
    corner: np.ndarray = field(default_factory=lambda: np.zeros((8, 3), dtype=np.int32))

    def add_condition(self, side, component=None):
        """bc.addCondition(side, condition_type::dirichlet, nullptr, component); component None = all three."""
        for c in (range(3) if component is None else [component]):
            self.side[side, c] = 1
        return self

    def add_corner_value(self, corner, component=None):
        for c in (range(3) if component is None else [component]):
            self.corner[corner, c] = 1
        return self

    def to_c(self):
        b = ks_bc()
        for s in range(6):
            for c in range(3):
                b.side[s][c] = int(self.side[s, c])
        for k in range(8):
            for c in range(3):
                b.corner[k][c] = int(self.corner[k, c])
        return b


@dataclass
class SolidProblem:
    volume: Volume
    bc: SolidBC
    law: int = KS_LAW_SVK
    E: float = 1e9
    nu: float = 0.45
    body_force: tuple = (0.0, 0.0, 0.0)
    tractions: list = field(default_factory=list)      # [(side, (tx,ty,tz)), ...]  dead Neumann loads
    fixed_values: np.ndarray | None = None
    dof_map: np.ndarray | None = None
    n_free: int = 0
    n_fixed: int = 0

    def number_dofs(self, build_fn):
        n1, n2, n3 = self.volume.n
        m = np.zeros(3 * n1 * n2 * n3, dtype=np.int32)
        nf, nx = C.c_int32(), C.c_int32()
        b = self.bc.to_c()
        rc = build_fn(n1, n2, n3, C.byref(b), m.ctypes.data_as(c_int_p), C.byref(nf), C.byref(nx))
        assert rc == 0
        self.dof_map, self.n_free, self.n_fixed = m, nf.value, nx.value
        return self

    def to_c(self):
        v = self.volume
        keep = []

        def dp(a):
            a = np.ascontiguousarray(a, dtype=np.float64)
            keep.append(a)
            return a.ctypes.data_as(c_double_p)

        def ip(a):
            a = np.ascontiguousarray(a, dtype=np.int32)
            keep.append(a)
            return a.ctypes.data_as(c_int_p)

        P = ks_problem()
        for d in range(3):
            P.degree[d] = v.p[d]
            P.n_knots[d] = len(v.U[d])
            P.knots[d] = dp(v.U[d])
        P.cp = dp(v.cp.reshape(-1))
        P.weights = None
        P.dof_map = ip(self.dof_map)
        P.n_free, P.n_fixed = self.n_free, self.n_fixed
        P.fixed_values = dp(self.fixed_values) if self.fixed_values is not None else None
        P.material_law = self.law
        P.E, P.nu = self.E, self.nu
        for k in range(3):
            P.body_force[k] = self.body_force[k]
        P.n_tractions = len(self.tractions)
        if self.tractions:
            P.traction_side = ip([t[0] for t in self.tractions])
            P.traction_val = dp(np.array([t[1] for t in self.tractions]).reshape(-1))
        return P, keep


def _bind(L):
    if getattr(L, "_ks_bound", False):
        return L
    vp = C.c_void_p
    L.ks_build_dofmap.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.POINTER(ks_bc), c_int_p, c_int_p, c_int_p]
    L.ks_create.argtypes = [C.POINTER(ks_problem), C.c_int, C.POINTER(vp)]
    L.ks_destroy.argtypes = [vp]
    L.ks_destroy.restype = None
    L.ks_sizes.argtypes = [vp, c_int_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    L.ks_pattern_host.argtypes = [vp, c_int_p, c_int_p]
    L.ks_assemble.argtypes = [vp, c_double_p, c_double_p, c_double_p]
    L.ks_jacobian.argtypes = [vp, c_double_p, c_double_p]
    L.ks_residual.argtypes = [vp, c_double_p, c_double_p]
    L.ks_al_residual.argtypes = [vp, c_double_p, C.c_double, c_double_p]
    L.ks_force.argtypes = [vp, c_double_p]
    L.ks_mass.argtypes = [vp, C.c_double, c_double_p]
    L.ks_assemble_device.argtypes = [vp, vp, C.c_int, vp, vp]
    L.ks_values_device.argtypes = [vp]
    L.ks_values_device.restype = vp
    L.ks_check.argtypes = [vp, vp]
    L.ks_last_timing.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float)]
    L.ks_kernel_launches.argtypes = [vp]
    L._ks_bound = True
    return L


# every symbol include/ks_solid.h declares
SYMBOLS = ["ks_build_dofmap", "ks_create", "ks_destroy", "ks_sizes", "ks_pattern_host", "ks_assemble", "ks_jacobian",
           "ks_residual", "ks_al_residual", "ks_force", "ks_mass", "ks_assemble_device", "ks_values_device", "ks_check",
           "ks_last_timing", "ks_kernel_launches"]


def _dp(a):
    return a.ctypes.data_as(c_double_p)


class SolidAssembler:
    """gsElasticityAssembler<real_t> behind the closures of tutorials/nonlinear_solid_static.cpp:101-114."""

    def __init__(self, prob: SolidProblem, device: int = -1):
        self.L = _bind(capi.lib())
        if prob.dof_map is None:
            prob.number_dofs(self.L.ks_build_dofmap)
        self.prob = prob
        P, self._keep = prob.to_c()
        h = C.c_void_p()
        capi.check(self.L.ks_create(C.byref(P), device, C.byref(h)))
        self.h = h
        nd, nnz, ne, nq = C.c_int32(), C.c_int64(), C.c_int64(), C.c_int64()
        capi.check(self.L.ks_sizes(self.h, C.byref(nd), C.byref(nnz), C.byref(ne), C.byref(nq)))
        self.n_dofs, self.nnz, self.n_elements, self.n_qp = nd.value, nnz.value, ne.value, nq.value
        self._pattern = None
        self._values = None

    def pattern(self):
        if self._pattern is None:
            outer = np.zeros(self.n_dofs + 1, dtype=np.int32)
            inner = np.zeros(max(self.nnz, 1), dtype=np.int32)
            capi.check(self.L.ks_pattern_host(self.h, outer.ctypes.data_as(c_int_p), inner.ctypes.data_as(c_int_p)))
            self._pattern = (outer, inner[:self.nnz])
        return self._pattern

    def _vals(self):
        if self._values is None:
            self._values = np.zeros(max(self.nnz, 1))
        return self._values

    def _fail(self):
        self.last_error = self.L.kl_last_error().decode()
        return False, None

    def assemble(self, x):
        """assembler.assemble(x, fixedDofs): (ok, K, rhs) in one pass."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        v, r = self._vals(), np.zeros(self.n_dofs)
        if self.L.ks_assemble(self.h, _dp(x), _dp(v), _dp(r)) != 0:
            self.last_error = self.L.kl_last_error().decode()
            return False, None, None
        outer, inner = self.pattern()
        return True, SparseView(self.n_dofs, outer, inner, v[:self.nnz]), r

    def jacobian(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        v = self._vals()
        if self.L.ks_jacobian(self.h, _dp(x), _dp(v)) != 0:
            return self._fail()
        outer, inner = self.pattern()
        return True, SparseView(self.n_dofs, outer, inner, v[:self.nnz])

    def residual(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        r = np.zeros(self.n_dofs)
        if self.L.ks_residual(self.h, _dp(x), _dp(r)) != 0:
            return self._fail()
        return True, r

    def al_residual(self, x, lam):
        x = np.ascontiguousarray(x, dtype=np.float64)
        r = np.zeros(self.n_dofs)
        if self.L.ks_al_residual(self.h, _dp(x), float(lam), _dp(r)) != 0:
            return self._fail()
        return True, r

    def force(self):
        f = np.zeros(self.n_dofs)
        capi.check(self.L.ks_force(self.h, _dp(f)))
        return f

    def mass(self, density):
        """gsMassAssembler with option Density: consistent mass matrix on the pattern of K."""
        v = np.zeros(max(self.nnz, 1))
        capi.check(self.L.ks_mass(self.h, float(density), _dp(v)))
        outer, inner = self.pattern()
        return SparseView(self.n_dofs, outer, inner, v[:self.nnz])

    def assemble_device(self, x_dev_ptr, r_dev_ptr=0, want_matrix=True, stream=0):
        capi.check(self.L.ks_assemble_device(self.h, C.c_void_p(x_dev_ptr), 1 if want_matrix else 0,
                                             C.c_void_p(r_dev_ptr) if r_dev_ptr else None, C.c_void_p(stream)))

    def check(self, stream=0):
        return self.L.ks_check(self.h, C.c_void_p(stream))

    def last_timing(self):
        a, b, c = C.c_float(), C.c_float(), C.c_float()
        self.L.ks_last_timing(self.h, C.byref(a), C.byref(b), C.byref(c))
        return {"points_ms": a.value, "jacobian_ms": b.value, "residual_ms": c.value}

    def kernel_launches(self):
        return self.L.ks_kernel_launches(self.h)

    def close(self):
        if getattr(self, "h", None):
            self.L.ks_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
