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
        L.klo_dead_force.argtypes = [C.c_void_p, c_double_p]
        L.klo_mass.argtypes = [C.c_void_p, C.c_double, c_double_p, c_double_p]
        L.klo_jacobian_residual.argtypes = [C.c_void_p, c_double_p, c_double_p, c_double_p]
        L.klo_build_dofmap.argtypes = [C.c_int, C.c_int, C.POINTER(kl_bc), c_int_p, c_int_p, c_int_p]
        L.klo_material.argtypes = [C.POINTER(kl_problem)] + [c_double_p] * 9
        L.klo_basis_ders.argtypes = [C.c_int, C.c_int, c_double_p, C.c_double, C.POINTER(C.c_int), c_double_p]
        L.klo_gauss.argtypes = [C.c_int, c_double_p, c_double_p]
        L.klo_cg_solve.argtypes = [C.c_int, c_int_p, c_int_p, c_double_p, c_double_p, c_double_p, C.c_double, C.c_int,
                                   C.POINTER(C.c_int), c_double_p]
        L.klo_stress_dim.argtypes = [C.c_int]
        L.klo_eval_stress.argtypes = [C.c_void_p, c_double_p, C.c_int, C.c_int, c_double_p, C.c_double, c_double_p]
        L.klo_boundary_force.argtypes = [C.c_void_p, c_double_p, C.c_int, c_double_p]
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

    def internal_force(self, x):
        """(F_int - P)(x) of the elements assembled by this oracle (the whole mesh, or its strip): dead loads - rhs(x)"""
        f = np.zeros(self.n_dofs)
        self.L.klo_dead_force(self.h, _dp(f))
        return f - self.residual(x)

    def diagonal(self, values):
        """diagonal of a matrix stored on the oracle's pattern (0 where the pattern has no diagonal entry)"""
        cols = np.repeat(np.arange(self.n_dofs), np.diff(self.outer))
        d = np.zeros(self.n_dofs)
        m = self.inner == cols
        d[cols[m]] = values[m]
        return d

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

    def eval_stress(self, x, stress_type, uv, z=0.0):
        from gsstructuralanalysis_b200.capi import STRESS_TYPES
        t = STRESS_TYPES[stress_type] if isinstance(stress_type, str) else int(stress_type)
        dim = self.L.klo_stress_dim(t)
        x = np.ascontiguousarray(x, dtype=np.float64)
        uv = np.ascontiguousarray(uv, dtype=np.float64).reshape(-1, 2)
        out = np.zeros((uv.shape[0], max(dim, 1)))
        rc = self.L.klo_eval_stress(self.h, _dp(x), t, uv.shape[0], _dp(uv), float(z), _dp(out))
        if rc:
            raise RuntimeError(f"oracle eval_stress rc={rc}")
        return out

    def computePrincipalStretches(self, uv, x, z=0.0):
        return self.eval_stress(x, "principal_stretch", uv, z)

    def boundaryForce(self, x, side):
        x = np.ascontiguousarray(x, dtype=np.float64)
        out = np.zeros(3)
        rc = self.L.klo_boundary_force(self.h, _dp(x), int(side), _dp(out))
        if rc:
            raise RuntimeError(f"oracle boundary_force rc={rc}")
        return out

    def jacobian_residual(self, x, values=None, r=None):
        x = np.ascontiguousarray(x, dtype=np.float64)
        values = np.zeros(self.nnz) if values is None else values
        r = np.zeros(self.n_dofs) if r is None else r
        rc = self.L.klo_jacobian_residual(self.h, _dp(x), _dp(values), _dp(r))
        if rc:
            raise RuntimeError(f"oracle rc={rc}")
        return values, r


def cg_solve(n, outer, inner, values, b, tol=0.0, max_iter=0):
    """Eigen's ConjugateGradient + DiagonalPreconditioner (= gsSparseSolver<>::CGDiagonal) restated on the CPU:
    (x, iterations, error)."""
    L = lib()
    outer = np.ascontiguousarray(outer, dtype=np.int32)
    inner = np.ascontiguousarray(inner, dtype=np.int32)
    values = np.ascontiguousarray(values, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    x = np.zeros(n)
    it, err = C.c_int(), C.c_double()
    L.klo_cg_solve(n, outer.ctypes.data_as(c_int_p), inner.ctypes.data_as(c_int_p), _dp(values), _dp(b), _dp(x), float(tol),
                   int(max_iter), C.byref(it), C.byref(err))
    return x, it.value, err.value


def newton_solve(ops, U=None, tolU=1e-6, tolF=1e-6, max_it=25, relaxation=1.0, linear_start=True, cg_tol=0.0, cg_max_iter=0):
    """gsStaticNewton<T>::_solveNonlinear (src/gsStaticSolvers/gsStaticNewton.hpp:141-196, _start :282-325) with the
    CGDiagonal default, on any object with jacobian(x) -> (ok, SparseView-like) / residual(x) / force()."""
    n = ops.n_dofs
    U = np.zeros(n) if U is None else np.array(U, dtype=np.float64)
    info = dict(status=1, iterations=0, cg_iterations=0, residual=0.0, residual_ini=0.0, dU_norm=0.0, DU_norm=0.0)

    def solve(K, rhs):
        o, i, v = (K.indptr, K.indices, K.data) if hasattr(K, "indptr") else (K.outer, K.inner, K.values)
        x, it, _ = cg_solve(n, o, i, v, rhs, cg_tol, cg_max_iter)
        info["cg_iterations"] += it
        return x

    def res(x):
        ok, r = ops.residual(x)
        if not ok:
            raise FloatingPointError
        return r

    def jac(x):
        ok, K = ops.jacobian(x)
        if not ok:
            raise FloatingPointError
        return K

    try:
        if linear_start:
            DU = solve(jac(np.zeros(n)), ops.force())
            U = np.zeros(n)
            R = res(U + DU)
            residual = np.linalg.norm(R) or 1.0
            residual_ini = np.linalg.norm(res(U)) or 1.0
        else:
            DU = np.zeros(n)
            R = res(U)
            residual = np.linalg.norm(R) or 1.0
            residual_ini = residual
        info["residual_ini"] = residual_ini
        k = 0
        while k != max_it:
            dU = solve(jac(U + DU), R)
            DU = DU + relaxation * dU
            R = res(U + DU)
            residual = np.linalg.norm(R)
            info.update(residual=residual, dU_norm=relaxation * np.linalg.norm(dU), DU_norm=np.linalg.norm(DU))
            if info["dU_norm"] / info["DU_norm"] < tolU and residual / residual_ini < tolF:
                info["status"] = 0
                break
            k += 1
        info["iterations"] = k
        U = U + DU
    except FloatingPointError:
        info["status"] = 2
    return U, info


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
