"""ctypes binding of the CPU oracle for the gsElasticity solid path (oracle/ks_oracle.c).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np
import scipy.sparse as sp

from gsstructuralanalysis_b200.problem import c_double_p, c_int_p
from gsstructuralanalysis_b200.solid import SolidProblem, ks_problem, ks_bc

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        subprocess.check_call(["make", "-s", "-C", _HERE, "libks_oracle.so"])
        L = C.CDLL(os.path.join(_HERE, "libks_oracle.so"))
        vp = C.c_void_p
        L.kso_create.restype = vp
        L.kso_create.argtypes = [C.POINTER(ks_problem)]
        L.kso_destroy.argtypes = [vp]
        L.kso_set_threads.argtypes = [vp, C.c_int]
        L.kso_get_threads.argtypes = [vp]
        L.kso_sizes.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_long), C.POINTER(C.c_long), C.POINTER(C.c_long)]
        L.kso_pattern.argtypes = [vp, c_int_p, c_int_p]
        L.kso_force.argtypes = [vp, c_double_p]
        L.kso_assemble.argtypes = [vp, c_double_p, c_double_p, c_double_p, c_double_p]
        L.kso_mass.argtypes = [vp, C.c_double, c_double_p]
        L.kso_build_dofmap.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(ks_bc), c_int_p, c_int_p, c_int_p]
        _LIB = L
    return _LIB


def _dp(a):
    return a.ctypes.data_as(c_double_p) if a is not None else None


class SolidOracle:
    def __init__(self, prob: SolidProblem, threads: int | None = None):
        self.L = lib()
        if prob.dof_map is None:
            prob.number_dofs(self.L.kso_build_dofmap)
        P, self._keep = prob.to_c()
        self.h = self.L.kso_create(C.byref(P))
        if not self.h:
            raise RuntimeError("kso_create failed")
        if threads:
            self.L.kso_set_threads(self.h, threads)
        nd, nnz, ne, nq = C.c_int(), C.c_long(), C.c_long(), C.c_long()
        self.L.kso_sizes(self.h, C.byref(nd), C.byref(nnz), C.byref(ne), C.byref(nq))
        self.n_dofs, self.nnz, self.n_elements, self.n_qp = nd.value, nnz.value, ne.value, nq.value
        self.outer = np.zeros(self.n_dofs + 1, dtype=np.int32)
        self.inner = np.zeros(max(self.nnz, 1), dtype=np.int32)
        self.L.kso_pattern(self.h, self.outer.ctypes.data_as(c_int_p), self.inner.ctypes.data_as(c_int_p))
        self.inner = self.inner[:self.nnz]

    @property
    def threads(self):
        return self.L.kso_get_threads(self.h)

    def assemble(self, x, matrix=True, residual=True, energy=False):
        x = np.ascontiguousarray(x, dtype=np.float64)
        v = np.zeros(max(self.nnz, 1)) if matrix else None
        r = np.zeros(self.n_dofs) if residual else None
        e = np.zeros(1) if energy else None
        rc = self.L.kso_assemble(self.h, _dp(x), _dp(v), _dp(r), _dp(e))
        if rc:
            raise RuntimeError(f"solid oracle rc={rc}")
        out = []
        if matrix:
            out.append(v[:self.nnz])
        if residual:
            out.append(r)
        if energy:
            out.append(float(e[0]))
        return out[0] if len(out) == 1 else tuple(out)

    def jacobian(self, x):
        return sp.csc_matrix((self.assemble(x, True, False), self.inner, self.outer), shape=(self.n_dofs, self.n_dofs))

    def residual(self, x):
        return self.assemble(x, False, True)

    def energy(self, x):
        return self.assemble(x, False, False, True)

    def mass(self, density):
        v = np.zeros(max(self.nnz, 1))
        self.L.kso_mass(self.h, float(density), _dp(v))
        return v[:self.nnz]

    def force(self):
        f = np.zeros(self.n_dofs)
        self.L.kso_force(self.h, _dp(f))
        return f

    def close(self):
        if getattr(self, "h", None):
            self.L.kso_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
