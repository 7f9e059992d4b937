"""Host-side mirror of the reference's operator interface for the hot path.

`ShellAssembler` plays the role of gsThinShellAssembler<3,real_t,true> behind the closures of
tutorials/nonlinear_shell_static.cpp:120-136, and `operators()` returns callables with the shapes of
gsStructuralAnalysisOps<T>::{Jacobian_t, Residual_t, ALResidual_t}
(src/gsStructuralAnalysisTools/gsStructuralAnalysisTypes.h:70-88): they return (ok, result) where the
reference returns bool and fills an out-parameter.  All arithmetic happens in libkl_shell.so on the GPU.
"""
from __future__ import annotations

import ctypes as C
import numpy as np

from . import capi
from .problem import ShellProblem, MultiPatchProblem, c_double_p, c_int_p


def _dp(a):
    return a.ctypes.data_as(c_double_p)


class SparseView:
    """Compressed column storage (outer/inner/values), layout-compatible with
    Eigen::SparseMatrix<double, ColMajor, int> = gsSparseMatrix<real_t>."""

    def __init__(self, n, outer, inner, values):
        self.n, self.outer, self.inner, self.values = n, outer, inner, values

    def to_scipy(self):
        import scipy.sparse as sp
        return sp.csc_matrix((self.values, self.inner, self.outer), shape=(self.n, self.n))


class ShellAssembler:
    def __init__(self, prob: ShellProblem, device: int = -1, _handle=None):
        self.L = capi.lib()
        self._owns = _handle is None
        if _handle is not None:          # non-owning view (the matrix context or one patch of a MultiPatchAssembler)
            self.prob, self.h = prob, C.c_void_p(_handle)
            self._init_sizes()
            return
        if prob.dof_map is None:
            prob.number_dofs(self.L.kl_build_dofmap)
        self.prob = prob
        P, self._keep = prob.to_c()
        h = C.c_void_p()
        capi.check(self.L.kl_create(C.byref(P), device, C.byref(h)))
        self.h = h
        self._init_sizes()

    def _init_sizes(self):
        nd, nnz, ne, nq = C.c_int32(), C.c_int64(), C.c_int64(), C.c_int64()
        capi.check(self.L.kl_sizes(self.h, C.byref(nd), C.byref(nnz), C.byref(ne), C.byref(nq)))
        self.n_dofs, self.nnz, self.n_elements, self.n_qp = nd.value, nnz.value, ne.value, nq.value
        self._pattern = None
        self._values = None

    # -- numDofs(), pattern -------------------------------------------------------------------
    def numDofs(self):
        return self.n_dofs

    def pattern(self):
        if self._pattern is None:
            outer = np.zeros(self.n_dofs + 1, dtype=np.int32)
            inner = np.zeros(max(self.nnz, 1), dtype=np.int32)
            capi.check(self.L.kl_pattern_host(self.h, outer.ctypes.data_as(c_int_p), inner.ctypes.data_as(c_int_p)))
            self._pattern = (outer, inner[:self.nnz])
        return self._pattern

    def values_buffer(self):
        """Host value array that the matrix view aliases (pageable unless the owner pins it with pin_values)."""
        if self._values is None:
            self._values = np.zeros(max(self.nnz, 1))
        return self._values

    # -- the closure bodies -------------------------------------------------------------------
    def jacobian(self, x, fetch=True):
        """constructSolution(x,def); assembleMatrix(def); m = matrix()."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        assert x.shape == (self.n_dofs,)
        v = self.values_buffer() if fetch else None
        rc = self.L.kl_jacobian(self.h, _dp(x), _dp(v) if fetch else None)
        if rc != 0:
            self.last_error = self.L.kl_last_error().decode()
            return False, None
        if not fetch:
            return True, None
        outer, inner = self.pattern()
        return True, SparseView(self.n_dofs, outer, inner, v[:self.nnz])

    def pattern_lower(self):
        """(outer, inner) of the lower-triangular view (row >= col) handed to one-triangle consumers (SimplicialLDLT)."""
        if getattr(self, "_pattern_lower", None) is None:
            nl = C.c_int64()
            capi.check(self.L.kl_pattern_lower_host(self.h, None, None, C.byref(nl)))
            outer = np.zeros(self.n_dofs + 1, dtype=np.int32)
            inner = np.zeros(max(nl.value, 1), dtype=np.int32)
            capi.check(self.L.kl_pattern_lower_host(self.h, outer.ctypes.data_as(c_int_p), inner.ctypes.data_as(c_int_p), C.byref(nl)))
            self.nnz_lower = nl.value
            self._pattern_lower = (outer, inner[:nl.value])
        return self._pattern_lower

    def jacobian_lower(self, x, out=None):
        """K(x), lower triangle only (half the PCIe bytes): SparseView on the lower pattern."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        outer, inner = self.pattern_lower()
        v = np.zeros(max(self.nnz_lower, 1)) if out is None else out
        rc = self.L.kl_jacobian_lower(self.h, _dp(x), _dp(v))
        if rc != 0:
            self.last_error = self.L.kl_last_error().decode()
            return False, None
        return True, SparseView(self.n_dofs, outer, inner, v[:self.nnz_lower])

    def fetch_values(self, out=None):
        """lazy fetch of the matrix the last jacobian(fetch=False) / jacobian_device call left on the GPU"""
        v = np.zeros(max(self.nnz, 1)) if out is None else out
        capi.check(self.L.kl_fetch_values(self.h, _dp(v)))
        return v[:self.nnz]

    def set_values(self, values):
        v = np.ascontiguousarray(values, dtype=np.float64)
        assert v.shape[0] >= self.nnz
        capi.check(self.L.kl_set_values(self.h, _dp(v)))

    def pin_values(self, arr):
        capi.check(self.L.kl_pin_values(self.h, _dp(arr), arr.shape[0]))

    def unpin_values(self, arr):
        capi.check(self.L.kl_unpin_values(self.h, _dp(arr)))

    def residual(self, x, out=None):
        """constructSolution; assembleVector(def); v = rhs()   (= F_ext - F_int).  `out`: the caller's result vector (the
        reference fills a gsVector the solver owns); page-locked memory is written by the D2H copy directly."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        r = np.zeros(self.n_dofs) if out is None else out
        assert r.shape == (self.n_dofs,) and r.dtype == np.float64 and r.flags.c_contiguous
        rc = self.L.kl_residual(self.h, _dp(x), _dp(r))
        if rc != 0:
            self.last_error = self.L.kl_last_error().decode()
            return False, None
        return True, r

    def al_residual(self, x, lam):
        """Force - lam*Force - rhs()  (benchmarks/benchmark_Roof.cpp:335-344)."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        r = np.zeros(self.n_dofs)
        rc = self.L.kl_al_residual(self.h, _dp(x), float(lam), _dp(r))
        if rc != 0:
            self.last_error = self.L.kl_last_error().decode()
            return False, None
        return True, r

    def force(self):
        f = np.zeros(self.n_dofs)
        capi.check(self.L.kl_force(self.h, _dp(f)))
        return f

    def mass(self, density, lumped=False):
        """assembleMass(lumped): consistent mass matrix on the stiffness pattern, or the lumped mass vector."""
        if lumped:
            m = np.zeros(self.n_dofs)
            capi.check(self.L.kl_mass(self.h, float(density), None, _dp(m)))
            return m
        v = np.zeros(max(self.nnz, 1))
        capi.check(self.L.kl_mass(self.h, float(density), _dp(v), None))
        outer, inner = self.pattern()
        return SparseView(self.n_dofs, outer, inner, v[:self.nnz])

    def operators(self):
        """(Jacobian_t, Residual_t, ALResidual_t) — cheap-to-copy handles like the reference's lambdas."""
        return (lambda x: self.jacobian(x)), (lambda x: self.residual(x)), (lambda x, lam: self.al_residual(x, lam))

    # -- device-resident leg ------------------------------------------------------------------
    def jacobian_device(self, x_dev_ptr, stream=0):
        capi.check(self.L.kl_jacobian_device(self.h, C.c_void_p(x_dev_ptr), C.c_void_p(stream)))

    def residual_device(self, x_dev_ptr, r_dev_ptr, lam_fext=1.0, sign_fint=-1.0, stream=0):
        capi.check(self.L.kl_residual_device(self.h, C.c_void_p(x_dev_ptr), lam_fext, sign_fint, C.c_void_p(r_dev_ptr),
                                             C.c_void_p(stream)))

    def assemble_device(self, x_dev_ptr, r_dev_ptr, lam_fext=1.0, sign_fint=-1.0, stream=0):
        """K(x) and the residual at the same state in one pass (the point kernel also integrates the internal force)."""
        capi.check(self.L.kl_assemble_device(self.h, C.c_void_p(x_dev_ptr), lam_fext, sign_fint, C.c_void_p(r_dev_ptr),
                                             C.c_void_p(stream)))

    def strip_begin_device(self, x_dev_ptr, r_dev_ptr, lam_fext, sign_fint, tail_rows, stream=0):
        capi.check(self.L.kl_strip_begin_device(self.h, C.c_void_p(x_dev_ptr), lam_fext, sign_fint, C.c_void_p(r_dev_ptr), int(tail_rows),
                                                C.c_void_p(stream)))

    def jacobian_rows_device(self, e2_begin, e2_end, stream=0):
        capi.check(self.L.kl_jacobian_rows_device(self.h, int(e2_begin), int(e2_end), C.c_void_p(stream)))

    def check(self, stream=0):
        return self.L.kl_check(self.h, C.c_void_p(stream))

    def values_device_ptr(self):
        return self.L.kl_values_device(self.h)

    # -- device-resident linear solve / Newton loop ---------------------------
    def cg_solve(self, b, tol=0.0, max_iter=0):
        """gsSparseSolver<>::CGDiagonal on the matrix of the last jacobian()/mass() call: (x, iterations, error).
        tol <= 0 / max_iter <= 0 select Eigen's defaults (machine epsilon, 2 n)."""
        b = np.ascontiguousarray(b, dtype=np.float64)
        assert b.shape == (self.n_dofs,)
        x = np.zeros(self.n_dofs)
        it, err = C.c_int32(), C.c_double()
        capi.check(self.L.kl_cg_solve(self.h, _dp(b), _dp(x), float(tol), int(max_iter), C.byref(it), C.byref(err)))
        return x, it.value, err.value

    def spmv(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.zeros(self.n_dofs)
        capi.check(self.L.kl_spmv(self.h, _dp(x), _dp(y)))
        return y

    def cg_last_timing(self):
        a, b, c = C.c_float(), C.c_float(), C.c_float()
        self.L.kl_cg_last_timing(self.h, C.byref(a), C.byref(b), C.byref(c))
        return {"total_ms": a.value, "iter_ms": b.value, "spmv_ms": c.value}

    def newton_solve(self, U=None, tolU=1e-6, tolF=1e-6, max_it=25, relaxation=1.0, linear_start=True, cg_tol=0.0,
                     cg_max_iter=0):
        """gsStaticNewton::solveNonlinear with the CGDiagonal default, device resident: (U, info dict)."""
        U = np.zeros(self.n_dofs) if U is None else np.array(U, dtype=np.float64)
        opt = capi.kl_newton_options(tolU, tolF, relaxation, max_it, 1 if linear_start else 0, cg_tol, cg_max_iter)
        info = capi.kl_newton_info()
        capi.check(self.L.kl_newton_solve(self.h, _dp(U), C.byref(opt), C.byref(info)))
        return U, {k: getattr(info, k) for k, _ in capi.kl_newton_info._fields_}

    def alm_step(self, U, L, DUold=None, DLold=0.0, arc_length=1e-2, tolU=1e-6, tolF=1e-3, max_it=100, phi=-1.0, relaxation=1.0,
                 cg_tol=0.0, cg_max_iter=0):
        """gsALMCrisfield::step(), device resident: returns (status, U, L, DeltaU, DeltaL, info); the inputs are not modified."""
        U = np.array(U, dtype=np.float64)
        DU = np.zeros(self.n_dofs) if DUold is None else np.array(DUold, dtype=np.float64)
        Lc, DLc = C.c_double(L), C.c_double(DLold)
        opt = capi.kl_alm_options(tolU, tolF, max_it, phi, relaxation, cg_tol, cg_max_iter)
        info = capi.kl_alm_info()
        capi.check(self.L.kl_alm_step(self.h, _dp(U), C.byref(Lc), _dp(DU), C.byref(DLc), float(arc_length), C.byref(opt), C.byref(info)))
        return info.status, U, Lc.value, DU, DLc.value, {k: getattr(info, k) for k, _ in capi.kl_alm_info._fields_}

    def stability(self, return_D=False):
        """gsALMBase::_computeStability, "Determinant" method, on the matrix of the last jacobian() call (device LDL^T):
        (indicator = min D, negatives = number of negative pivots[, D per DoF])."""
        ind, neg = C.c_double(), C.c_int32()
        D = np.zeros(self.n_dofs) if return_D else None
        capi.check(self.L.kl_stability(self.h, C.byref(ind), C.byref(neg), _dp(D) if return_D else None))
        return (ind.value, neg.value, D) if return_D else (ind.value, neg.value)

    # -- stress / stretch recovery --------------------------------------------
    def eval_stress(self, x, stress_type, uv, z=0.0):
        """constructStress(mp_def, field, stress_type::X) evaluated at the parametric points uv [n,2]: array [n, dim]
        (benchmarks/benchmark_Balloon.cpp:381-408).  stress_type: a name of capi.STRESS_TYPES or its number."""
        t = capi.STRESS_TYPES[stress_type] if isinstance(stress_type, str) else int(stress_type)
        dim = self.L.kl_stress_dim(t)
        x = np.ascontiguousarray(x, dtype=np.float64)
        uv = np.ascontiguousarray(uv, dtype=np.float64).reshape(-1, 2)
        out = np.zeros((uv.shape[0], max(dim, 1)))
        capi.check(self.L.kl_eval_stress(self.h, _dp(x), t, uv.shape[0], _dp(uv), float(z), _dp(out)))
        return out

    def computePrincipalStretches(self, uv, x, z=0.0):
        """assembler->computePrincipalStretches(pts, mp_def, z) (unittests/gsStaticSolver_test.cpp:317): [n,3],
        lambda(0) <= lambda(1) in-plane, lambda(2) the thickness stretch."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        uv = np.ascontiguousarray(uv, dtype=np.float64).reshape(-1, 2)
        out = np.zeros((uv.shape[0], 3))
        capi.check(self.L.kl_principal_stretches(self.h, _dp(x), uv.shape[0], _dp(uv), float(z), _dp(out)))
        return out

    def boundaryForce(self, x, side):
        """assembler->boundaryForce(mp_def, patchSide(0, side)) (unittests/gsStaticSolver_test.cpp:321): (Fx, Fy, Fz)."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        out = np.zeros(3)
        capi.check(self.L.kl_boundary_force(self.h, _dp(x), int(side), _dp(out)))
        return out

    def set_strip(self, e2_begin, e2_end):
        capi.check(self.L.kl_set_strip(self.h, e2_begin, e2_end))

    def last_timing(self):
        a, b, c = C.c_float(), C.c_float(), C.c_float()
        self.L.kl_last_timing(self.h, C.byref(a), C.byref(b), C.byref(c))
        return {"kernel_ms": a.value, "h2d_ms": b.value, "d2h_ms": c.value}

    def kernel_launches(self):
        return self.L.kl_kernel_launches(self.h)

    def close(self):
        if getattr(self, "h", None):
            if self._owns:
                self.L.kl_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class MultiPatchAssembler(ShellAssembler):
    """gsThinShellAssembler on a gsMultiPatch with C0-matched interfaces (include/kl_shell.h: kl_mp_*): one DoF numbering, one
    matrix.  Every matrix / vector level method of ShellAssembler works on the whole multi-patch; patch(q) is the view of ONE
    patch for the geometry-level calls (eval_stress, computePrincipalStretches, boundaryForce) with the global solution vector."""

    def __init__(self, mprob: MultiPatchProblem, device: int = -1):
        L = capi.lib()
        if any(p.dof_map is None for p in mprob.patches):
            mprob.number_dofs(L.kl_mp_build_dofmap)
        self.mprob = mprob
        arr, self._keep = mprob.to_c()
        mp = C.c_void_p()
        capi.check(L.kl_mp_create(len(mprob.patches), arr, device, C.byref(mp)))
        self.mp = mp
        super().__init__(None, _handle=L.kl_mp_context(mp))

    def patch(self, q):
        return ShellAssembler(self.mprob.patches[q], _handle=self.L.kl_mp_patch(self.mp, q))

    def set_active(self, active=None):
        """patch -> GPU partition: assemble only the patches with active[q] != 0 (None = all)"""
        if active is None:
            capi.check(self.L.kl_mp_set_active(self.mp, None))
        else:
            a = np.ascontiguousarray(active, dtype=np.int32)
            capi.check(self.L.kl_mp_set_active(self.mp, a.ctypes.data_as(c_int_p)))

    def interface_dofs(self):
        n = C.c_int32()
        capi.check(self.L.kl_mp_interface_dofs(self.mp, C.byref(n), None))
        d = np.zeros(max(n.value, 1), dtype=np.int32)
        capi.check(self.L.kl_mp_interface_dofs(self.mp, C.byref(n), d.ctypes.data_as(c_int_p)))
        return d[:n.value]

    def close(self):
        if getattr(self, "mp", None):
            self.L.kl_mp_destroy(self.mp)
            self.mp = None
            self.h = None
