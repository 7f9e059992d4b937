/** @file gsStructuralAnalysisOps_b200.h

    Header-only C++ adapter: turns a kl_ctx (include/kl_shell.h, libkl_shell.so) into the std::function operators
    every gsStructuralAnalysis solver consumes.

    Mirrors, name for name, the reference interface
        gismo::gsStatus                               src/gsStructuralAnalysisTools/gsStructuralAnalysisTypes.h:22-30
        gismo::gsStructuralAnalysisOps<T>::*_t        src/gsStructuralAnalysisTools/gsStructuralAnalysisTypes.h:58-93
    and replaces the closure bodies of the drivers
        Jacobian / Residual                           tutorials/nonlinear_shell_static.cpp:120-136
        ALResidual                                    benchmarks/benchmark_Roof.cpp:335-344

    With G+Smo on the include path (__has_include(<gismo.h>)) the real gsVector / gsSparseMatrix are used and the
    reference's solvers (gsStaticNewton, gsALMCrisfield, gsAPALM, ...) take these callables unchanged.  Without it a
    minimal stand-in with the same storage layout (Eigen compressed column-major: outerIndexPtr / innerIndexPtr /
    valuePtr) is provided so that the adapter, its tests and the examples build on their own.
*/
#pragma once

#include <algorithm>
#include <cmath>
#include <cstring>
#include <functional>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "kl_shell.h"
#include "ks_solid.h"

#if defined(__has_include)
#if __has_include(<gismo.h>)
#include <gismo.h>
#define KL_HAVE_GISMO 1
#endif
#endif

#ifndef KL_HAVE_GISMO
namespace gismo {

typedef int index_t;
typedef double real_t;

/// Dense column vector (stand-in for gsVector<T> = Eigen::Matrix<T,Dynamic,1>)
template <class T = real_t>
class gsVector {
public:
    gsVector() {}
    explicit gsVector(index_t n) : m_v(n, T(0)) {}
    void resize(index_t n) { m_v.resize(n); }
    void setZero() { std::fill(m_v.begin(), m_v.end(), T(0)); }
    void setZero(index_t n) { m_v.assign(n, T(0)); }
    index_t size() const { return (index_t)m_v.size(); }
    index_t rows() const { return size(); }
    T* data() { return m_v.data(); }
    const T* data() const { return m_v.data(); }
    T& operator[](index_t i) { return m_v[i]; }
    const T& operator[](index_t i) const { return m_v[i]; }
    T& at(index_t i) { return m_v[i]; }
    T norm() const { T s = 0; for (T v : m_v) s += v * v; return std::sqrt(s); }
    gsVector& operator+=(const gsVector& o) { for (index_t i = 0; i < size(); ++i) m_v[i] += o[i]; return *this; }
private:
    std::vector<T> m_v;
};

/// Compressed column-major sparse matrix (stand-in for gsSparseMatrix<T> = Eigen::SparseMatrix<T,ColMajor,index_t>)
template <class T = real_t>
class gsSparseMatrix {
public:
    gsSparseMatrix() : m_rows(0), m_cols(0) {}
    index_t rows() const { return m_rows; }
    index_t cols() const { return m_cols; }
    index_t nonZeros() const { return (index_t)m_val.size(); }
    bool isCompressed() const { return true; }
    const index_t* outerIndexPtr() const { return m_outer.data(); }
    const index_t* innerIndexPtr() const { return m_inner.data(); }
    index_t* outerIndexPtr() { return m_outer.data(); }
    index_t* innerIndexPtr() { return m_inner.data(); }
    const T* valuePtr() const { return m_val.data(); }
    T* valuePtr() { return m_val.data(); }
    /// adopt a compressed pattern (what `m = assembler.matrix()` leaves behind)
    void setPattern(index_t n, const index_t* outer, const index_t* inner) {
        m_rows = m_cols = n;
        m_outer.assign(outer, outer + n + 1);
        m_inner.assign(inner, inner + outer[n]);
        m_val.assign((size_t)outer[n], T(0));
    }
    /// y = A x
    void apply(const gsVector<T>& x, gsVector<T>& y) const {
        y.setZero(m_rows);
        for (index_t j = 0; j < m_cols; ++j) {
            const T xj = x[j];
            for (index_t k = m_outer[j]; k < m_outer[j + 1]; ++k) y[m_inner[k]] += m_val[k] * xj;
        }
    }
    T diagonal(index_t j) const {
        for (index_t k = m_outer[j]; k < m_outer[j + 1]; ++k) if (m_inner[k] == j) return m_val[k];
        return T(0);
    }
private:
    index_t m_rows, m_cols;
    std::vector<index_t> m_outer, m_inner;
    std::vector<T> m_val;
};

/// src/gsStructuralAnalysisTools/gsStructuralAnalysisTypes.h:22-30
enum struct gsStatus { Success, NotConverged, AssemblyError, SolverError, NotStarted, OtherError };

/// src/gsStructuralAnalysisTools/gsStructuralAnalysisTypes.h:58-93
template <class T>
struct gsStructuralAnalysisOps {
    typedef std::function<bool(gsVector<T> const&, T&)> Energy_t;
    typedef std::function<bool(gsVector<T>&)> Force_t;
    typedef std::function<bool(const T, gsVector<T>&)> TForce_t;
    typedef std::function<bool(gsVector<T> const&, gsVector<T>&)> Residual_t;
    typedef std::function<bool(gsVector<T> const&, const T, gsVector<T>&)> ALResidual_t;
    typedef std::function<bool(gsVector<T> const&, const T, gsVector<T>&)> TResidual_t;
    typedef std::function<bool(gsSparseMatrix<T>&)> Mass_t;
    typedef std::function<bool(const T, gsSparseMatrix<T>&)> TMass_t;
    typedef std::function<bool(gsVector<T> const&, gsSparseMatrix<T>&)> Damping_t;
    typedef std::function<bool(gsVector<T> const&, const T, gsSparseMatrix<T>&)> TDamping_t;
    typedef std::function<bool(gsSparseMatrix<T>&)> Stiffness_t;
    typedef std::function<bool(gsVector<T> const&, gsSparseMatrix<T>&)> Jacobian_t;
    typedef std::function<bool(gsVector<T> const&, const T, gsSparseMatrix<T>&)> TJacobian_t;
    typedef std::function<bool(gsVector<T> const&, gsVector<T> const&, gsSparseMatrix<T>&)> dJacobian_t;
};

}  // namespace gismo
#endif  // !KL_HAVE_GISMO

namespace gismo {

/** Owns one device context and hands out cheap-to-copy operator handles (the reference stores its closures by
    value inside the solvers: src/gsStaticSolvers/gsStaticNewton.h:232-235), so the handles share the context through
    a shared_ptr.  One object per GPU / solver thread; calls on one object are serialised by the caller, as in the
    reference (closures are not re-entrant, benchmarks/benchmark_Frustrum_APALM.cpp:403-412). */
class gsThinShellAssemblerB200 {
public:
    typedef double T;
    typedef gsStructuralAnalysisOps<T> Ops;

    gsThinShellAssemblerB200(const kl_problem& prob, int device = -1) {
        kl_ctx* c = nullptr;
        const int rc = kl_create(&prob, device, &c);
        if (rc != KL_OK) throw std::runtime_error(std::string("kl_create: ") + kl_last_error());
        m_s = std::make_shared<Shared>();
        m_s->ctx = c;
        int64_t nnz = 0, ne = 0, nq = 0;
        kl_sizes(c, &m_s->ndofs, &nnz, &ne, &nq);
        m_s->nnz = nnz;
        m_s->outer.resize((size_t)m_s->ndofs + 1);
        m_s->inner.resize((size_t)nnz);
        if (kl_pattern_host(c, m_s->outer.data(), m_s->inner.data()) != KL_OK)
            throw std::runtime_error(std::string("kl_pattern_host: ") + kl_last_error());
    }

    index_t numDofs() const { return m_s->ndofs; }
    int64_t nonZeros() const { return m_s->nnz; }
    kl_ctx* context() const { return m_s->ctx; }

    /// K(x): constructSolution(x,def); assembleMatrix(def); m = matrix()
    Ops::Jacobian_t jacobian() const {
        auto s = m_s;
        return [s](gsVector<T> const& x, gsSparseMatrix<T>& m) {
            adoptPattern(m, s->ndofs, s->nnz, s->outer.data(), s->inner.data());
            return kl_jacobian(s->ctx, x.data(), m.valuePtr()) == KL_OK;
        };
    }
    /// dJacobian_t ignores dx, like the wrappers in gsStaticNewton.h:79-82 / gsALMBase.h:57-60
    Ops::dJacobian_t djacobian() const {
        Ops::Jacobian_t J = jacobian();
        return [J](gsVector<T> const& x, gsVector<T> const&, gsSparseMatrix<T>& m) { return J(x, m); };
    }
    /// R(x) = F_ext - F_int: constructSolution; assembleVector(def); v = rhs()
    Ops::Residual_t residual() const {
        auto s = m_s;
        return [s](gsVector<T> const& x, gsVector<T>& r) {
            r.resize(s->ndofs);
            return kl_residual(s->ctx, x.data(), r.data()) == KL_OK;
        };
    }
    /// F_int - lam F_ext = Force - lam*Force - rhs()
    Ops::ALResidual_t alResidual() const {
        auto s = m_s;
        return [s](gsVector<T> const& x, const T lam, gsVector<T>& r) {
            r.resize(s->ndofs);
            return kl_al_residual(s->ctx, x.data(), lam, r.data()) == KL_OK;
        };
    }
    /** Time-dependent signatures of the dynamic solvers (gsDynamicBase.h:294,330; TJacobian_t / TResidual_t / TForce_t,
        gsStructuralAnalysisTypes.h:66-74,90).  The shell drivers of the reference ignore the time argument
        (examples/example_DynamicShellNL.cpp:217-236: "to do: add time dependency of forcing"); so do these. */
    Ops::TJacobian_t tJacobian() const {
        Ops::Jacobian_t J = jacobian();
        return [J](gsVector<T> const& x, const T /*time*/, gsSparseMatrix<T>& m) { return J(x, m); };
    }
    Ops::TResidual_t tResidual() const {
        Ops::Residual_t R = residual();
        return [R](gsVector<T> const& x, const T /*time*/, gsVector<T>& r) { return R(x, r); };
    }
    Ops::TForce_t tForce() const {
        Ops::Force_t F = force();
        return [F](const T /*time*/, gsVector<T>& f) { return F(f); };
    }
    /// C = 0 on the stiffness pattern: `C = gsSparseMatrix<>(numDofs, numDofs)` of examples/example_DynamicShellNL.cpp:251,255
    Ops::Damping_t damping() const {
        auto s = m_s;
        return [s](gsVector<T> const&, gsSparseMatrix<T>& m) {
            adoptPattern(m, s->ndofs, s->nnz, s->outer.data(), s->inner.data());
            std::fill(m.valuePtr(), m.valuePtr() + s->nnz, T(0));
            return true;
        };
    }
    /// M: assembler.assembleMass(); m = matrix()  (Mass_t, gsStructuralAnalysisTypes.h:77)
    Ops::Mass_t mass(T density) const {
        auto s = m_s;
        return [s, density](gsSparseMatrix<T>& m) {
            adoptPattern(m, s->ndofs, s->nnz, s->outer.data(), s->inner.data());
            return kl_mass(s->ctx, density, m.valuePtr(), nullptr) == KL_OK;
        };
    }
    /// lumped mass vector: assembleMass(true); rhs()  (used by gsStaticDR, unittests/gsStaticSolver_test.cpp:252-253)
    Ops::Force_t lumpedMass(T density) const {
        auto s = m_s;
        return [s, density](gsVector<T>& v) {
            v.resize(s->ndofs);
            return kl_mass(s->ctx, density, nullptr, v.data()) == KL_OK;
        };
    }
    /// linear stiffness (Stiffness_t): assemble(); K = matrix()  (tutorials/nonlinear_shell_dynamic.cpp:116-118) = K(0)
    Ops::Stiffness_t stiffness() const {
        auto s = m_s;
        return [s](gsSparseMatrix<T>& m) {
            adoptPattern(m, s->ndofs, s->nnz, s->outer.data(), s->inner.data());
            return kl_jacobian(s->ctx, nullptr, m.valuePtr()) == KL_OK;
        };
    }
    /// F (what assemble(); rhs() gives at u = 0)
    Ops::Force_t force() const {
        auto s = m_s;
        return [s](gsVector<T>& f) {
            f.resize(s->ndofs);
            return kl_force(s->ctx, f.data()) == KL_OK;
        };
    }

    /** Device-resident replacement of `m_solver->compute(jacMat); m_solver->solve(F)` for the CGDiagonal default
        (gsStaticBase.h:164, gsStaticNewton.hpp:230-240): solves with the matrix the last jacobian()/mass() call left on
        the GPU; only the two vectors cross PCIe.  tol <= 0 / maxIter <= 0 = Eigen's defaults.  Returns false on error. */
    bool cgSolve(gsVector<T> const& rhs, gsVector<T>& x, T tol = 0, index_t maxIter = 0, index_t* iterations = nullptr,
                 T* error = nullptr) const {
        x.resize(m_s->ndofs);
        int32_t it = 0;
        double err = 0;
        const int rc = kl_cg_solve(m_s->ctx, rhs.data(), x.data(), tol, maxIter, &it, &err);
        if (iterations) *iterations = it;
        if (error) *error = err;
        return rc == KL_OK;
    }
    /** gsStaticNewton::solveNonlinear (gsStaticNewton.hpp:101-196) run entirely on the device (Jacobian, CGDiagonal solve,
        residual, norms); U is m_U on entry and the solution on exit.  The returned status is the solver's gsStatus. */
    gsStatus newtonSolve(gsVector<T>& U, const kl_newton_options& options, kl_newton_info* info = nullptr) const {
        kl_newton_info local;
        if (!info) info = &local;
        if ((index_t)U.size() != (index_t)m_s->ndofs) { U.resize(m_s->ndofs); U.setZero(); }
        if (kl_newton_solve(m_s->ctx, U.data(), &options, info) != KL_OK) return gsStatus::OtherError;
        switch (info->status) {
            case 0: return gsStatus::Success;
            case 1: return gsStatus::NotConverged;
            case 2: return gsStatus::AssemblyError;
            case 3: return gsStatus::SolverError;
        }
        return gsStatus::OtherError;
    }
    /** assembler->computePrincipalStretches(pts, mp_def, z) (unittests/gsStaticSolver_test.cpp:317): `uv` holds the
        parametric points column by column (u0,v0,u1,v1,...), the result 3 stretches per point — lambda(0) <= lambda(1)
        in-plane, lambda(2) the thickness stretch.  `solVector` is the DoF vector that constructSolution would take. */
    bool computePrincipalStretches(const std::vector<T>& uv, gsVector<T> const& solVector, T z, std::vector<T>& lambdas) const {
        const int32_t n = (int32_t)(uv.size() / 2);
        lambdas.assign((size_t)3 * n, T(0));
        return kl_principal_stretches(m_s->ctx, solVector.data(), n, uv.data(), z, lambdas.data()) == KL_OK;
    }
    /** assembler->boundaryForce(mp_def, patchSide(0, side)) (unittests/gsStaticSolver_test.cpp:321); side = KL_WEST.. */
    bool boundaryForce(gsVector<T> const& solVector, int side, T force[3]) const {
        return kl_boundary_force(m_s->ctx, solVector.data(), side, force) == KL_OK;
    }
    /** assembler->constructStress(mp_def, field, stress_type::X) evaluated at parametric points
        (benchmarks/benchmark_Balloon.cpp:381-408): `type` = KL_STRESS_*, result kl_stress_dim(type) values per point. */
    bool evalStress(gsVector<T> const& solVector, int type, const std::vector<T>& uv, std::vector<T>& result, T z = 0) const {
        const int32_t n = (int32_t)(uv.size() / 2);
        result.assign((size_t)kl_stress_dim(type) * n, T(0));
        return kl_eval_stress(m_s->ctx, solVector.data(), type, n, uv.data(), z, result.data()) == KL_OK;
    }
    /// gsStaticBase::defaultOptions (gsStaticBase.h:66-75) + gsStaticNewton::defaultOptions (gsStaticNewton.hpp:20-25)
    static kl_newton_options defaultNewtonOptions() {
        kl_newton_options o;
        o.tolU = 1e-6; o.tolF = 1e-6; o.relaxation = 1.0; o.max_it = 25; o.linear_start = 1; o.cg_tol = 0.0; o.cg_max_iter = 0;
        return o;
    }

private:
    static void adoptPattern(gsSparseMatrix<T>& m, index_t n, int64_t nnz, const int32_t* outer, const int32_t* inner) {
        if (m.rows() == n && m.cols() == n && (int64_t)m.nonZeros() == nnz && m.isCompressed()) return;   // values only
#ifdef KL_HAVE_GISMO
        // one-time deep copy of the symbolic pattern into the solver's matrix; afterwards only valuePtr() is refreshed
        typedef gsEigen::Map<const gsEigen::SparseMatrix<T, gsEigen::ColMajor, index_t> > MapT;
        std::vector<T> zeros((size_t)nnz, T(0));
        m = MapT(n, n, (index_t)nnz, outer, inner, zeros.data());
        m.makeCompressed();
#else
        m.setPattern(n, outer, inner);
#endif
    }

    struct Shared {
        kl_ctx* ctx = nullptr;
        int32_t ndofs = 0;
        int64_t nnz = 0;
        std::vector<int32_t> outer, inner;
        ~Shared() { if (ctx) kl_destroy(ctx); }
    };
    std::shared_ptr<Shared> m_s;   // shared by every handle handed out, so handles may outlive this object
};

/** gsElasticityAssembler<real_t> behind the solid closures (tutorials/nonlinear_solid_static.cpp:101-114,
    benchmarks/benchmark_Elasticity_Beam_APALM.cpp:307-325) on libkl_shell.so (include/ks_solid.h).  The reference's two
    closures each run the full assemble(x, fixedDofs); here each computes only its half, and assemble() gives both in one pass. */
class gsElasticityAssemblerB200 {
public:
    typedef double T;
    typedef gsStructuralAnalysisOps<T> Ops;

    gsElasticityAssemblerB200(const ks_problem& prob, int device = -1) {
        ks_ctx* c = nullptr;
        if (ks_create(&prob, device, &c) != KL_OK) throw std::runtime_error(std::string("ks_create: ") + kl_last_error());
        m_s = std::make_shared<Shared>();
        m_s->ctx = c;
        int64_t nnz = 0, ne = 0, nq = 0;
        ks_sizes(c, &m_s->ndofs, &nnz, &ne, &nq);
        m_s->nnz = nnz;
        m_s->outer.resize((size_t)m_s->ndofs + 1);
        m_s->inner.resize((size_t)nnz);
        if (ks_pattern_host(c, m_s->outer.data(), m_s->inner.data()) != KL_OK)
            throw std::runtime_error(std::string("ks_pattern_host: ") + kl_last_error());
    }
    index_t numDofs() const { return m_s->ndofs; }
    int64_t nonZeros() const { return m_s->nnz; }
    ks_ctx* context() const { return m_s->ctx; }

    /// assembler.assemble(x, fixedDofs); m = assembler.matrix();
    Ops::Jacobian_t jacobian() const {
        auto s = m_s;
        return [s](gsVector<T> const& x, gsSparseMatrix<T>& m) {
            adopt(m, *s);
            return ks_jacobian(s->ctx, x.data(), m.valuePtr()) == KL_OK;
        };
    }
    /// assembler.assemble(x, fixedDofs); v = assembler.rhs();
    Ops::Residual_t residual() const {
        auto s = m_s;
        return [s](gsVector<T> const& x, gsVector<T>& r) {
            r.resize(s->ndofs);
            return ks_residual(s->ctx, x.data(), r.data()) == KL_OK;
        };
    }
    /// Force - lam*Force - assembler.rhs()
    Ops::ALResidual_t alResidual() const {
        auto s = m_s;
        return [s](gsVector<T> const& x, const T lam, gsVector<T>& r) {
            r.resize(s->ndofs);
            return ks_al_residual(s->ctx, x.data(), lam, r.data()) == KL_OK;
        };
    }
    /// assembler.assemble(); F = assembler.rhs()
    Ops::Force_t force() const {
        auto s = m_s;
        return [s](gsVector<T>& f) {
            f.resize(s->ndofs);
            return ks_force(s->ctx, f.data()) == KL_OK;
        };
    }
    /// gsMassAssembler with option "Density": assemble(); M = matrix()  (tutorials/nonlinear_solid_dynamic.cpp:98-109,137)
    Ops::Mass_t mass(T density) const {
        auto s = m_s;
        return [s, density](gsSparseMatrix<T>& m) {
            adopt(m, *s);
            return ks_mass(s->ctx, density, m.valuePtr()) == KL_OK;
        };
    }
    /// one pass for both outputs of assemble(x, fixedDofs)
    bool assemble(gsVector<T> const& x, gsSparseMatrix<T>& m, gsVector<T>& rhs) const {
        adopt(m, *m_s);
        rhs.resize(m_s->ndofs);
        return ks_assemble(m_s->ctx, x.data(), m.valuePtr(), rhs.data()) == KL_OK;
    }

private:
    struct Shared {
        ks_ctx* ctx = nullptr;
        int32_t ndofs = 0;
        int64_t nnz = 0;
        std::vector<int32_t> outer, inner;
        ~Shared() { if (ctx) ks_destroy(ctx); }
    };
    static void adopt(gsSparseMatrix<T>& m, const Shared& s) {
        if (m.rows() == s.ndofs && m.cols() == s.ndofs && (int64_t)m.nonZeros() == s.nnz && m.isCompressed()) return;
#ifdef KL_HAVE_GISMO
        typedef gsEigen::Map<const gsEigen::SparseMatrix<T, gsEigen::ColMajor, index_t> > MapT;
        std::vector<T> zeros((size_t)s.nnz, T(0));
        m = MapT(s.ndofs, s.ndofs, (index_t)s.nnz, s.outer.data(), s.inner.data(), zeros.data());
        m.makeCompressed();
#else
        m.setPattern(s.ndofs, s.outer.data(), s.inner.data());
#endif
    }
    std::shared_ptr<Shared> m_s;
};

}  // namespace gismo
