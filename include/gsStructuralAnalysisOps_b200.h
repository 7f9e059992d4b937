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

/// Stand-in for the abstract gsSparseSolver<T> (upstream gismo/src/gsMatrix/gsSparseSolver.h, not in /root/reference
/// [UPSTREAM-RECALLED]): what the reference's solvers hold as `typename gsSparseSolver<T>::uPtr m_solver` and drive with
/// compute(jacMat) / info() / solve(F) (src/gsStaticSolvers/gsStaticBase.h:85,164-166,260, gsStaticNewton.hpp:214-240,
/// src/gsALMSolvers/gsALMBase.hpp:73,186-205).  info() follows Eigen::ComputationInfo: 0 Success, 1 NumericalIssue,
/// 2 NoConvergence, 3 InvalidInput.
template <class T = real_t>
class gsSparseSolver {
public:
    typedef std::unique_ptr<gsSparseSolver> uPtr;
    virtual ~gsSparseSolver() {}
    virtual gsSparseSolver& compute(const gsSparseMatrix<T>& matrix) = 0;
    virtual gsVector<T> solve(const gsVector<T>& rhs) const = 0;
    virtual int info() const = 0;
    virtual bool succeed() const = 0;
    virtual std::string detail() const = 0;
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

/** How the Jacobian_t closure hands K to the solver.
      Full   : every value into m (what `m = assembler.matrix()` does; 8 nnz bytes over PCIe per call)
      Lower  : m is the lower-triangular view (row >= col) — all that SimplicialLDLT / selfadjointView<Lower> read
               (benchmarks/benchmark_Roof.cpp:359-360); half the bytes
      Device : K stays on the GPU; m receives the pattern and a tag only.  The consumer is gsSparseSolverB200 (below), which
               recognises the tag in compute(m) and solves on the device; any other consumer calls fetch(m) first. */
enum class gsB200CopyOut { Full, Lower, Device };
template <class T = real_t> class gsSparseSolverB200;

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
        init();
    }
    /** Multi-patch: one kl_problem per patch, all numbered by ONE mapper (kl_mp_build_dofmap) — the reference's
        gsThinShellAssembler(mp, dbasis, bc, force, materialMatrix) on a gsMultiPatch with computeTopology()
        (benchmarks/benchmark_Wrinkling.cpp:446-522).  Every operator below then acts on the whole multi-patch. */
    gsThinShellAssemblerB200(const std::vector<kl_problem>& patches, int device = -1) {
        kl_mp* mp = nullptr;
        const int rc = kl_mp_create((int32_t)patches.size(), patches.data(), device, &mp);
        if (rc != KL_OK) throw std::runtime_error(std::string("kl_mp_create: ") + kl_last_error());
        m_s = std::make_shared<Shared>();
        m_s->mp = mp;
        m_s->ctx = kl_mp_context(mp);
        init();
    }

private:
    void init() {
        kl_ctx* c = m_s->ctx;
        int64_t nnz = 0, ne = 0, nq = 0;
        kl_sizes(c, &m_s->ndofs, &nnz, &ne, &nq);
        m_s->nnz = nnz;
        m_s->outer.resize((size_t)m_s->ndofs + 1);
        m_s->inner.resize((size_t)nnz);
        if (kl_pattern_host(c, m_s->outer.data(), m_s->inner.data()) != KL_OK)
            throw std::runtime_error(std::string("kl_pattern_host: ") + kl_last_error());
    }

public:
    /// geometry-level calls of a multi-patch (computePrincipalStretches, boundaryForce, evalStress) address one patch
    void selectPatch(int32_t q) {
        if (!m_s->mp || !kl_mp_patch(m_s->mp, q)) throw std::runtime_error("selectPatch: not a multi-patch assembler / no such patch");
        m_s->geo = kl_mp_patch(m_s->mp, q);
    }
    void setCopyOut(gsB200CopyOut mode) {
        if (mode == gsB200CopyOut::Lower && m_s->outer_lower.empty()) {
            int64_t nl = 0;
            if (kl_pattern_lower_host(m_s->ctx, nullptr, nullptr, &nl) != KL_OK) throw std::runtime_error(kl_last_error());
            m_s->outer_lower.resize((size_t)m_s->ndofs + 1);
            m_s->inner_lower.resize((size_t)nl);
            m_s->nnz_lower = nl;
            if (kl_pattern_lower_host(m_s->ctx, m_s->outer_lower.data(), m_s->inner_lower.data(), &nl) != KL_OK)
                throw std::runtime_error(kl_last_error());
        }
        m_s->mode = mode;
    }
    gsB200CopyOut copyOut() const { return m_s->mode; }
    /// Device mode: bring the values of the matrix behind the placeholder m to the host after all (lazy fetch)
    bool fetch(gsSparseMatrix<T>& m) const {
        adoptPattern(m, m_s->ndofs, m_s->nnz, m_s->outer.data(), m_s->inner.data());
        return kl_fetch_values(m_s->ctx, m.valuePtr()) == KL_OK;
    }

    index_t numDofs() const { return m_s->ndofs; }
    int64_t nonZeros() const { return m_s->nnz; }
    kl_ctx* context() const { return m_s->ctx; }

    /// K(x): constructSolution(x,def); assembleMatrix(def); m = matrix()
    Ops::Jacobian_t jacobian() const {
        auto s = m_s;
        return [s](gsVector<T> const& x, gsSparseMatrix<T>& m) {
            if (s->mode == gsB200CopyOut::Lower) {
                adoptPattern(m, s->ndofs, s->nnz_lower, s->outer_lower.data(), s->inner_lower.data());
                return kl_jacobian_lower(s->ctx, x.data(), m.valuePtr()) == KL_OK;
            }
            adoptPattern(m, s->ndofs, s->nnz, s->outer.data(), s->inner.data());
            if (s->mode == gsB200CopyOut::Device) {
                if (kl_jacobian(s->ctx, x.data(), nullptr) != KL_OK) return false;
                if (s->nnz > 0) m.valuePtr()[0] = s->newTag();      // placeholder: pattern + tag, the values live on the GPU
                return true;
            }
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
        return kl_principal_stretches(m_s->geoCtx(), solVector.data(), n, uv.data(), z, lambdas.data()) == KL_OK;
    }
    /** assembler->boundaryForce(mp_def, patchSide(0, side)) (unittests/gsStaticSolver_test.cpp:321); side = KL_WEST.. */
    bool boundaryForce(gsVector<T> const& solVector, int side, T force[3]) const {
        return kl_boundary_force(m_s->geoCtx(), solVector.data(), side, force) == KL_OK;
    }
    /** assembler->constructStress(mp_def, field, stress_type::X) evaluated at parametric points
        (benchmarks/benchmark_Balloon.cpp:381-408): `type` = KL_STRESS_*, result kl_stress_dim(type) values per point. */
    bool evalStress(gsVector<T> const& solVector, int type, const std::vector<T>& uv, std::vector<T>& result, T z = 0) const {
        const int32_t n = (int32_t)(uv.size() / 2);
        result.assign((size_t)kl_stress_dim(type) * n, T(0));
        return kl_eval_stress(m_s->geoCtx(), solVector.data(), type, n, uv.data(), z, result.data()) == KL_OK;
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

    template <class U> friend class gsSparseSolverB200;
    struct Shared {
        kl_ctx* ctx = nullptr;
        kl_mp* mp = nullptr;             // multi-patch: owns ctx (= kl_mp_context)
        kl_ctx* geo = nullptr;           // patch addressed by the geometry-level calls (multi-patch)
        int32_t ndofs = 0;
        int64_t nnz = 0, nnz_lower = 0;
        std::vector<int32_t> outer, inner, outer_lower, inner_lower;
        gsB200CopyOut mode = gsB200CopyOut::Full;
        uint64_t generation = 0;         // Device mode: which Jacobian call the device matrix belongs to
        kl_ctx* geoCtx() const { return geo ? geo : ctx; }
        /// the tag written into values[0] of a Device-mode placeholder: a quiet NaN whose payload is the generation
        T newTag() { ++generation; return tagOf(generation); }
        static T tagOf(uint64_t g) {
            const uint64_t bits = 0x7ff8000000000000ULL | (0x0000b20000000000ULL) | (g & 0xffffffffffULL);
            T v; std::memcpy(&v, &bits, sizeof(v)); return v;
        }
        bool isCurrentPlaceholder(const gsSparseMatrix<T>& m) const {
            if (mode != gsB200CopyOut::Device || generation == 0 || (int64_t)m.nonZeros() != nnz || nnz == 0) return false;
            const T want = tagOf(generation);
            return std::memcmp(m.valuePtr(), &want, sizeof(T)) == 0;
        }
        ~Shared() { if (mp) kl_mp_destroy(mp); else if (ctx) kl_destroy(ctx); }
    };
    std::shared_ptr<Shared> m_s;   // shared by every handle handed out, so handles may outlive this object
};

/** gsSparseSolver-shaped front of the device-resident CGDiagonal (kl_cg_solve), for the solvers' `m_solver`
    (src/gsStaticSolvers/gsStaticBase.h:85,164-166,260; src/gsALMSolvers/gsALMBase.hpp:73,186-205).
      compute(m): m is the Device-mode placeholder of the assembler's current matrix -> nothing moves;
                  m is any other matrix on the assembler's full pattern (Full mode, or a matrix the solver has modified,
                  e.g. K - shift*M) -> its values are uploaded once (8 nnz bytes H2D);
                  anything else -> info() = InvalidInput.
      solve(b)  : Jacobi-preconditioned CG with Eigen's conventions on the GPU; only b and x cross PCIe.
    The matrix must be symmetric positive definite, as for gsSparseSolver<>::CGDiagonal. */
template <class T>
class gsSparseSolverB200 : public gsSparseSolver<T> {
public:
    explicit gsSparseSolverB200(const gsThinShellAssemblerB200& assembler) : m_s(assembler.m_s) {}
    void setTolerance(T tol) { m_tol = tol; }              ///< Eigen default: machine epsilon
    void setMaxIterations(index_t it) { m_maxit = it; }   ///< Eigen default: 2 n
    gsSparseSolverB200& compute(const gsSparseMatrix<T>& m) override {
        m_info = 0;
        if (m_s->isCurrentPlaceholder(m)) return *this;
        if ((index_t)m.rows() != (index_t)m_s->ndofs || (int64_t)m.nonZeros() != m_s->nnz) { m_info = 3; return *this; }
        if (kl_set_values(m_s->ctx, m.valuePtr()) != KL_OK) m_info = 1;
        return *this;
    }
    gsVector<T> solve(const gsVector<T>& rhs) const override {
        gsVector<T> x;
        x.resize(m_s->ndofs);
        int32_t it = 0;
        double err = 0;
        const int rc = kl_cg_solve(m_s->ctx, rhs.data(), x.data(), m_tol, m_maxit, &it, &err);
        m_iterations = it; m_error = err;
        const double tol = m_tol > 0 ? m_tol : 2.220446049250313e-16;
        m_info = rc != KL_OK ? 1 : (err <= tol ? 0 : 2);
        return x;
    }
    int info() const override { return m_info; }
    bool succeed() const override { return m_info == 0; }
    std::string detail() const override { return "CGDiagonal on the B200 (libkl_shell kl_cg_solve)"; }
    index_t iterations() const { return m_iterations; }
    T error() const { return m_error; }

private:
    std::shared_ptr<gsThinShellAssemblerB200::Shared> m_s;
    T m_tol = 0;
    index_t m_maxit = 0;
    mutable int m_info = 0;
    mutable index_t m_iterations = 0;
    mutable T m_error = 0;
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
