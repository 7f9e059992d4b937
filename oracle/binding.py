"""ctypes binding of oracle/libkl_oracle.so (TEST INFRASTRUCTURE ONLY — never imported by the
product package gsstructuralanalysis_b200)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import numpy as np
import scipy.sparse as sp

from gsstructuralanalysis_b200.problem import kl_problem, kl_bc, c_double_p, c_int_p, ShellProblem

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "libkl_oracle.so")
    src = os.path.join(_HERE, "kl_oracle.c")
    hdr = os.path.join(_HERE, "..", "include", "kl_shell.h")
    if force or not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libkl_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.klo_create.restype = C.c_void_p
        L.klo_create.argtypes = [C.POINTER(kl_problem)]
        L.klo_destroy.argtypes = [C.c_void_p]
        L.klo_set_threads.argtypes = [C.c_void_p, C.c_int]
        L.klo_set_strip.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.klo_get_threads.argtypes = [C.c_void_p]
        L.klo_sizes.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_long), C.POINTER(C.c_long), C.POINTER(C.c_long)]
        L.klo_pattern.argtypes = [C.c_void_p, c_int_p, c_int_p]
        L.klo_jacobian.argtypes = [C.c_void_p, c_double_p, c_double_p]
        L.klo_residual.argtypes = [C.c_void_p, c_double_p, c_double_p]
        L.klo_al_residual.argtypes = [C.c_void_p, c_double_p, C.c_double, c_double_p]
        L.klo_force.argtypes = [C.c_void_p, c_double_p]
        L.klo_mass.argtypes = [C.c_void_p, C.c_double, c_double_p, c_double_p]
        L.klo_jacobian_residual.argtypes = [C.c_void_p, c_double_p, c_double_p, c_double_p]
        L.klo_build_dofmap.argtypes = [C.c_int, C.c_int, C.POINTER(kl_bc), c_int_p, c_int_p, c_int_p]
        L.klo_material.argtypes = [C.POINTER(kl_problem)] + [c_double_p] * 9
        L.klo_basis_ders.argtypes = [C.c_int, C.c_int, c_double_p, C.c_double, C.POINTER(C.c_int), c_double_p]
        L.klo_gauss.argtypes = [C.c_int, c_double_p, c_double_p]
        _LIB = L
    return _LIB


def _dp(a):
    return a.ctypes.data_as(c_double_p)


class Oracle:
    def __init__(self, prob: ShellProblem, threads: int | None = None):
        self.L = lib()
        if prob.dof_map is None:
            prob.number_dofs(self.L.klo_build_dofmap)
        self.prob = prob
        P, self._keep = prob.to_c()
        self.h = self.L.klo_create(C.byref(P))
        if threads is not None:
            self.L.klo_set_threads(self.h, threads)
        nd, nnz, ne, nq = C.c_int(), C.c_long(), C.c_long(), C.c_long()
        self.L.klo_sizes(self.h, C.byref(nd), C.byref(nnz), C.byref(ne), C.byref(nq))
        self.n_dofs, self.nnz, self.n_elements, self.n_qp = nd.value, nnz.value, ne.value, nq.value
        self.outer = np.zeros(self.n_dofs + 1, dtype=np.int32)
        self.inner = np.zeros(max(self.nnz, 1), dtype=np.int32)
        self.L.klo_pattern(self.h, self.outer.ctypes.data_as(c_int_p), self.inner.ctypes.data_as(c_int_p))
        self.inner = self.inner[:self.nnz]

    def set_strip(self, e2_begin, e2_end):
        self.L.klo_set_strip(self.h, e2_begin, e2_end)

    @property
    def threads(self):
        return self.L.klo_get_threads(self.h)

    def close(self):
        if self.h:
            self.L.klo_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def jacobian_values(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        v = np.zeros(self.nnz)
        rc = self.L.klo_jacobian(self.h, _dp(x), _dp(v))
        if rc:
            raise RuntimeError(f"oracle jacobian rc={rc}")
        return v

    def jacobian(self, x):
        """scipy CSC matrix (the layout of gsSparseMatrix)."""
        return sp.csc_matrix((self.jacobian_values(x), self.inner, self.outer), shape=(self.n_dofs, self.n_dofs))

    def residual(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        r = np.zeros(self.n_dofs)
        rc = self.L.klo_residual(self.h, _dp(x), _dp(r))
        if rc:
            raise RuntimeError(f"oracle residual rc={rc}")
        return r

    def al_residual(self, x, lam):
        x = np.ascontiguousarray(x, dtype=np.float64)
        r = np.zeros(self.n_dofs)
        rc = self.L.klo_al_residual(self.h, _dp(x), float(lam), _dp(r))
        if rc:
            raise RuntimeError(f"oracle al_residual rc={rc}")
        return r

    def force(self):
        f = np.zeros(self.n_dofs)
        self.L.klo_force(self.h, _dp(f))
        return f

    def mass(self, density):
        v, l = np.zeros(self.nnz), np.zeros(self.n_dofs)
        self.L.klo_mass(self.h, float(density), _dp(v), _dp(l))
        return v, l

    def jacobian_residual(self, x, values=None, r=None):
        x = np.ascontiguousarray(x, dtype=np.float64)
        values = np.zeros(self.nnz) if values is None else values
        r = np.zeros(self.n_dofs) if r is None else r
        rc = self.L.klo_jacobian_residual(self.h, _dp(x), _dp(values), _dp(r))
        if rc:
            raise RuntimeError(f"oracle rc={rc}")
        return values, r


def material(prob: ShellProblem, Ac, Bc, ac, bc):
    """A,B,D (3x3), N,M (3) at one surface point from covariant metrics [11,22,12]."""
    L = lib()
    if prob.dof_map is None:
        prob.number_dofs(L.klo_build_dofmap)
    P, keep = prob.to_c()
    arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in (Ac, Bc, ac, bc)]
    A, B, D, N, M = np.zeros((3, 3)), np.zeros((3, 3)), np.zeros((3, 3)), np.zeros(3), np.zeros(3)
    rc = L.klo_material(C.byref(P), *[_dp(a) for a in arrs], _dp(A), _dp(B), _dp(D), _dp(N), _dp(M))
    return rc, A, B, D, N, M


class OracleOps:
    """The oracle in the closure shapes of gsStructuralAnalysisOps (returns (ok, value)), so that the same test
    drivers run on the oracle and on the GPU path."""

    def __init__(self, prob, threads=None):
        self.o = Oracle(prob, threads)
        self.n_dofs, self.nnz = self.o.n_dofs, self.o.nnz

    def jacobian(self, x):
        try:
            return True, self.o.jacobian(x)
        except RuntimeError:
            return False, None

    def residual(self, x):
        try:
            return True, self.o.residual(x)
        except RuntimeError:
            return False, None

    def force(self):
        return self.o.force()

    def close(self):
        self.o.close()
