/*
 * kl_shell.h — C ABI of the B200-native Kirchhoff–Love shell assembly path.
 *
 * This is the drop-in boundary for the ONE hot path of gismo/gsStructuralAnalysis:
 * the bodies of the Jacobian_t / Residual_t / ALResidual_t closures
 * (reference: src/gsStructuralAnalysisTools/gsStructuralAnalysisTypes.h:58-93,
 *  closure bodies tutorials/nonlinear_shell_static.cpp:120-136,
 *  benchmarks/benchmark_Roof.cpp:324-344).  Every entry point below names the
 * reference call it replaces.  Plain pointers and sizes only; no C++ / torch types.
 *
 * Conventions
 *   - real_t = double, index_t = int32_t (G+Smo defaults).
 *   - Control points are numbered tensor-style, first parametric direction fastest:
 *       i = i1 + n1*i2,  n_d = n_knots[d] - degree[d] - 1.
 *   - DoF numbering follows gsDofMapper (component-major; free first, coupled after the
 *     plain free DoFs of their component, eliminated after ALL free DoFs).
 *   - The sparse matrix is handed out as compressed arrays (outer/inner/values) that are
 *     layout-compatible with Eigen::SparseMatrix<double,ColMajor,int> in compressed mode,
 *     i.e. gsSparseMatrix<real_t>.  The pattern is structurally symmetric, so the same
 *     arrays are a CSR of the transpose; values are written column-compressed (entry
 *     (row i, col j) lives in outer[j]..outer[j+1]).
 *   - All functions return 0 on success, a negative KL_E_* code on failure.  Nothing
 *     throws across this boundary.  A non-zero return from kl_jacobian / kl_residual
 *     maps to "closure returns false" => gsStatus::AssemblyError in the solvers
 *     (reference: src/gsStaticSolvers/gsStaticNewton.hpp:196-212).
 */
#ifndef KL_SHELL_H
#define KL_SHELL_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- error codes ---------------------------------------------------------------- */
enum {
    KL_OK            =  0,
    KL_E_ARG         = -1,  /* bad argument / inconsistent sizes                      */
    KL_E_CUDA        = -2,  /* CUDA runtime error (see kl_last_error)                  */
    KL_E_NONFINITE   = -3,  /* a quadrature point produced a non-finite value          */
    KL_E_JACOBIAN    = -4,  /* |a1 x a2| <= 0 at a quadrature point (inverted element) */
    KL_E_C33         = -5,  /* plane-stress Newton on C33 did not converge             */
    KL_E_NOGPU       = -6   /* no CUDA device: there is NO CPU fallback                */
};

/* ---- material selection (reference option strings:
 *      tutorials/nonlinear_shell_static.cpp:103-104, unittests/gsStaticSolver_test.cpp:215-217) */
enum {
    KL_MAT_SVK = 0,   /* Material=0  St.Venant–Kirchhoff / gsMaterialMatrixLinear       */
    KL_MAT_NH  = 1,   /* Material=1  Neo-Hookean                                          */
    KL_MAT_MR  = 3    /* Material=3  Mooney–Rivlin (Ratio = c1/c2)                        */
};

/* boundary-condition kinds per side and component (gsBoundaryConditions condition_type) */
enum { KL_BC_FREE = 0, KL_BC_DIRICHLET = 1, KL_BC_CLAMPED = 2, KL_BC_COLLAPSED = 3 };
/* side order = G+Smo boxSide: west(u=0), east(u=1), south(v=0), north(v=1);
 * corner order = southwest, southeast, northwest, northeast                              */
enum { KL_WEST = 0, KL_EAST = 1, KL_SOUTH = 2, KL_NORTH = 3 };

typedef struct kl_bc {
    int32_t side[4][3];    /* KL_BC_* for [side][component]                              */
    int32_t corner[4][3];  /* 1 = addCornerValue(corner, 0.0, patch, component)          */
} kl_bc;

/* ---- problem description: what the reference passes to
 *      gsThinShellAssembler<3,real_t,true>(ori,bases,bc,force,materialMatrix)
 *      (tutorials/nonlinear_shell_static.cpp:113) + setPointLoads/setPressure --------- */
typedef struct kl_problem {
    int32_t degree[2];          /* p1, p2 of the solution basis (= geometry basis)       */
    int32_t n_knots[2];
    const double* knots[2];     /* open knot vectors, possibly non-uniform                */
    const double* cp;           /* [n_cp*3] undeformed control net, xyz interleaved       */
    const double* weights;      /* [n_cp] NURBS weights of the GEOMETRY or NULL.  The     */
                                /* displacement basis stays polynomial (gsMultiBasis(mp,true),*/
                                /* benchmarks/benchmark_Balloon.cpp:122)                   */
    const int32_t* dof_map;     /* [3*n_cp], dof_map[c*n_cp+i] = global index; an index    */
                                /* >= n_free is eliminated; (index-n_free) addresses       */
                                /* fixed_values (gsDofMapper::global_to_bindex)            */
    int32_t n_free;
    int32_t n_fixed;
    const double* fixed_values; /* [n_fixed] Dirichlet displacement values or NULL (=0)    */

    int32_t material;           /* KL_MAT_*                                               */
    int32_t compressible;       /* option "Compressibility"                               */
    int32_t num_gauss_thickness;/* option "NumGauss", default 4                           */
    int32_t bending;            /* third template argument of gsThinShellAssembler        */
    double  E, nu, thickness;   /* gsConstantFunction parameters                          */
    double  mr_ratio;           /* Mooney–Rivlin c1/c2                                    */
    int32_t metric_z2;          /* 1: add z^2 n,a.n,b to the through-thickness metric     */
    int32_t quA, quB;           /* Gauss nodes per direction = quA*p+quB (solver_options.xml:11-12) */

    double  body_force[3];      /* constant surface force (gsFunctionExpr force)          */
    double  pressure;           /* follower pressure, setPressure (benchmark_Balloon.cpp:258) */
    int32_t n_point_loads;
    const double* point_load_uv;   /* [n_point_loads*2] parametric positions              */
    const double* point_load_val;  /* [n_point_loads*3] load vectors                      */
    /* BCs.addCondition(side, condition_type::neumann, &neuData) with a constant traction vector per unit UNDEFORMED edge
     * length (gsConstantFunction neuData; benchmarks/benchmark_Frustrum_APALM.cpp:236-242,267, benchmark_Cylinder.cpp:118-129):
     * F[i,c] += int_side N_i t_c |dX/dxi| dxi, a dead load                                                          */
    int32_t n_neumann;
    const int32_t* neumann_side;   /* [n_neumann] KL_WEST .. KL_NORTH                     */
    const double* neumann_val;     /* [n_neumann*3] traction vectors                      */
} kl_problem;

typedef struct kl_ctx kl_ctx;

/* Number the DoFs of one patch the way gsFeSpace::setupMapper + gsDofMapper::finalize do
 * (SURVEY Appendix A.6).  Replaces: the mapper set-up inside the gsThinShellAssembler ctor.
 * dof_map must hold 3*n1*n2 entries. */
int kl_build_dofmap(int32_t n1, int32_t n2, const kl_bc* bc,
                    int32_t* dof_map, int32_t* n_free, int32_t* n_fixed);

/* Create a device context: uploads geometry, builds quadrature/basis tables, the symbolic
 * pattern and the scatter tables ON THE GPU.  device < 0 selects the current device.
 * Replaces: gsThinShellAssembler ctor + gsExprAssembler::initSystem. */
int kl_create(const kl_problem* prob, int device, kl_ctx** out);
void kl_destroy(kl_ctx* ctx);

/* Sizes: numDofs() (tutorials/nonlinear_shell_static.cpp:153), nnz, elements, quadrature points */
int kl_sizes(const kl_ctx* ctx, int32_t* n_dofs, int64_t* nnz, int64_t* n_elements, int64_t* n_qp);

/* Copy the symbolic pattern to host arrays (outer[n_dofs+1], inner[nnz]). */
int kl_pattern_host(const kl_ctx* ctx, int32_t* outer, int32_t* inner);
/* Device-resident view (valid for the life of the context). */
int kl_pattern_device(const kl_ctx* ctx, const int32_t** outer_dev, const int32_t** inner_dev);

/* K(x): replaces constructSolution(x,def); assembleMatrix(def); m = matrix()
 * (tutorials/nonlinear_shell_static.cpp:120-127).  x_host has n_dofs entries.
 * values_host (nnz doubles) may be NULL: then the values stay on the device only. */
int kl_jacobian(kl_ctx* ctx, const double* x_host, double* values_host);

/* Copy-out rules of kl_jacobian / kl_jacobian_lower / kl_fetch_values: page-locked caller memory (cudaMallocHost, or
 * kl_pin_values) is written by the DMA engine directly; pageable memory is served through a page-locked staging buffer the
 * context owns and copied out by host threads.  The library never page-locks memory it does not own: a solver that builds
 * a fresh gsSparseMatrix per call (gsStaticNewton::_computeJacobian, src/gsStaticSolvers/gsStaticNewton.hpp:205-212) may free
 * it at any time.  The OWNER of a long-lived value array may pin it once with kl_pin_values and must unpin it before freeing. */
int kl_pin_values(kl_ctx* ctx, double* values_host, int64_t count);
int kl_unpin_values(kl_ctx* ctx, double* values_host);
/* values of the matrix left on the device by the last kl_jacobian(.., NULL) / kl_jacobian_device / kl_mass call -> host (lazy
 * fetch for consumers that need the host matrix after all), and the reverse (a host matrix for kl_cg_solve). */
int kl_fetch_values(kl_ctx* ctx, double* values_host);
int kl_set_values(kl_ctx* ctx, const double* values_host);

/* Lower-triangular view (row >= col) of K for consumers that read one triangle only (SimplicialLDLT,
 * benchmarks/benchmark_Roof.cpp:359-360; Eigen's selfadjointView<Lower>): the same compressed column layout without the
 * upper entries.  kl_pattern_lower_host: outer_lower[n_dofs+1], inner_lower[nnz_lower] (either may be NULL to query
 * nnz_lower first); kl_jacobian_lower writes nnz_lower values, half the bytes of kl_jacobian.  Refused with follower pressure. */
int kl_pattern_lower_host(kl_ctx* ctx, int32_t* outer_lower, int32_t* inner_lower, int64_t* nnz_lower);
int kl_jacobian_lower(kl_ctx* ctx, const double* x_host, double* values_lower_host);

/* R(x) = F_ext - F_int(x): replaces constructSolution; assembleVector(def); v = rhs()
 * (tutorials/nonlinear_shell_static.cpp:129-136). */
int kl_residual(kl_ctx* ctx, const double* x_host, double* r_host);

/* Arc-length form  Force - lam*Force - rhs(x)  (benchmarks/benchmark_Roof.cpp:335-344); equal to F_int(x) - lam*F_ext for
 * dead loads and homogeneous Dirichlet values. */
int kl_al_residual(kl_ctx* ctx, const double* x_host, double lam, double* r_host);

/* Mass matrix M_ij^{cd} = delta_cd * density * thickness * int N_i N_j dA on the SAME pattern as K (values of the
 * off-diagonal component blocks are zero) and/or the lumped mass vector (row sums over all basis functions):
 * replaces assembler.assembleMass(); M = matrix()  /  assembleMass(true); rhs()
 * (unittests/gsStaticSolver_test.cpp:249-253; Mass_t in gsStructuralAnalysisTypes.h:77).  Either output may be NULL. */
int kl_mass(kl_ctx* ctx, double density, double* values_host, double* lumped_host);

/* Force = assembler->assemble(); assembler->rhs(): the right-hand side of the LINEAR system at the undeformed geometry
 * (tutorials/nonlinear_shell_static.cpp:141-143, benchmarks/benchmark_Balloon.cpp:262-263): body force, point loads and
 * Neumann tractions, the follower pressure evaluated on the undeformed surface (p N_i n meas), and the lifting of non-zero
 * Dirichlet values, -K_L(free, eliminated) g.  kl_al_residual uses the same vector:
 *     Force - lam*Force - rhs(x),   rhs(x) = kl_residual(x)                       (benchmarks/benchmark_Balloon.cpp:285) */
int kl_force(kl_ctx* ctx, double* f_host);

/* Device-resident variants: x_dev / out pointers are device memory; work is enqueued on
 * `stream` (a cudaStream_t passed as void*) and NOT synchronised — the error flag is
 * fetched by kl_check().  These are what a device-resident solver (or bench.py's
 * HBM-resident leg) calls. */
int kl_jacobian_device(kl_ctx* ctx, const double* x_dev, void* stream);
int kl_residual_device(kl_ctx* ctx, const double* x_dev, double lam_fext, double sign_fint,
                       double* r_dev, void* stream);
int kl_al_residual_device(kl_ctx* ctx, const double* x_dev, double lam, double* r_dev, void* stream);   /* Force - lam*Force - rhs(x) */
int kl_check(kl_ctx* ctx, void* stream);          /* sync stream, map device flag to KL_E_* */
double* kl_values_device(kl_ctx* ctx);            /* device K values, length nnz            */

/* Multi-GPU: restrict this context to the element rows [e2_begin, e2_end) of the second
 * parametric direction (a strip).  The strip assembles its partial K / R; owners are
 * completed by the halo reduce in the host layer. */
int kl_set_strip(kl_ctx* ctx, int32_t e2_begin, int32_t e2_end);

/* Two-phase strip assembly for overlapping the halo exchange with the bulk of the strip: kl_strip_begin_device = residual of the
 * strip (r = lam_fext*F_dead + sign_fint*(F_int - P)) + zeroing of the strip's value ranges + Jacobian of its LAST tail_rows element
 * rows (the ones that reach into the next strip); kl_jacobian_rows_device = Jacobian of further rows of the strip from the same
 * per-point records.  Work is enqueued on `stream`, not synchronised. */
int kl_strip_begin_device(kl_ctx* ctx, const double* x_dev, double lam_fext, double sign_fint, double* r_dev, int32_t tail_rows, void* stream);
int kl_jacobian_rows_device(kl_ctx* ctx, int32_t e2_begin, int32_t e2_end, void* stream);

/* Kernel-only timing of the last call in ms (CUDA events on the launch stream). */
int kl_last_timing(const kl_ctx* ctx, float* ms_kernel, float* ms_h2d, float* ms_d2h);
/* One Newton iteration's worth of assembly at ONE state: K(x) into the device values (as kl_jacobian_device) and
 * r = lam_fext*F_ext + sign_fint*F_int (as kl_residual_device); constructSolution runs once and the internal force is
 * integrated by the per-point kernel from the records it has just staged, so the separate residual pass disappears. */
int kl_assemble_device(kl_ctx* ctx, const double* x_dev, double lam_fext, double sign_fint, double* r_dev, void* stream);

/* ---- device-resident linear solve and Newton loop ---------------------------
 * The reference's Newton solver defaults to gsSparseSolver<>::CGDiagonal
 * (src/gsStaticSolvers/gsStaticNewton.hpp:23), i.e. Eigen::ConjugateGradient<SparseMatrix, Lower|Upper,
 * DiagonalPreconditioner> (third-party, Eigen 3.4 as vendored by G+Smo; not in /root/reference): start at
 * x = 0, stop when |r|^2 < max(tol^2 |b|^2, DBL_MIN) or after max_iter iterations; the preconditioner is
 * 1/diag (1 where the diagonal is 0); defaults tol = DBL_EPSILON, max_iter = 2 n (pass tol <= 0 / max_iter <= 0).
 * The matrix is the one left on the device by the last kl_jacobian / kl_jacobian_device / kl_mass call; it must
 * be symmetric (no follower pressure), as Eigen's solver requires.  Only the vectors cross PCIe. */
int kl_cg_solve(kl_ctx* ctx, const double* b_host, double* x_host, double tol, int32_t max_iter,
                int32_t* iters, double* rel_err);
int kl_cg_solve_device(kl_ctx* ctx, const double* b_dev, double* x_dev, double tol, int32_t max_iter,
                       int32_t* iters, double* rel_err, void* stream);
/* y = K x with the device matrix (gather form: exact for the symmetric matrices CG accepts).  Host pointers. */
int kl_spmv(kl_ctx* ctx, const double* x_host, double* y_host);
/* average duration in ms of one CG iteration / of its matrix-vector product in the last solve */
int kl_cg_last_timing(const kl_ctx* ctx, float* ms_total, float* ms_per_iter, float* ms_spmv);

/* gsStaticNewton<T>::_solveNonlinear (src/gsStaticSolvers/gsStaticNewton.hpp:141-196) with the CGDiagonal
 * default, everything device resident: per iteration one Jacobian, one CG solve, one residual; only norms
 * come back to the host.  Option names follow gsStaticBase::defaultOptions (gsStaticBase.h:66-75). */
typedef struct kl_newton_options {
    double tolU, tolF;          /* "tolU" / "tolF" (reference default: "tol" = 1e-6 for both)        */
    double relaxation;          /* "Relaxation" (1)                                                   */
    int32_t max_it;             /* "maxIt" (25)                                                       */
    int32_t linear_start;       /* 1 = DeltaU empty on entry: start from the linear solution K(0) DU = F and U = 0
                                   (gsStaticNewton.hpp:147-152); 0 = start from U with DeltaU = 0     */
    double cg_tol;              /* <= 0: Eigen default DBL_EPSILON                                    */
    int32_t cg_max_iter;        /* <= 0: Eigen default 2 n                                            */
} kl_newton_options;
typedef struct kl_newton_info {
    int32_t status;             /* gsStatus: 0 Success, 1 NotConverged, 2 AssemblyError, 3 SolverError */
    int32_t iterations;         /* m_numIterations                                                    */
    int64_t cg_iterations;      /* summed over all linear solves                                      */
    double residual, residual_ini;   /* |R|, |R0|                                                     */
    double dU_norm, DU_norm;    /* |relax dU|, |DU| of the last iteration                             */
    float ms_assembly, ms_solve;     /* device time spent in assembly kernels / in CG                 */
} kl_newton_info;
int kl_newton_solve(kl_ctx* ctx, double* U_host_inout, const kl_newton_options* opt, kl_newton_info* info);

/* gsALMBase<T>::step() with gsALMCrisfield<T> (src/gsALMSolvers/gsALMBase.hpp:354-416, gsALMCrisfield.hpp:66-226,328-425) and the
 * CGDiagonal solver benchmarks/benchmark_Frustrum_APALM.cpp:435 selects, everything device resident: per corrector iteration one
 * Jacobian, two solves with that matrix (deltaUt = K^-1 Force, deltaUbar = -K^-1 R), one arc-length residual, the quadratic
 * constraint (real, modified and complex-root branches) and the root choice of Ritto-Correa.  Options as in
 * gsALMBase::defaultOptions (:25-52): AngleMethod 0, no quasi-Newton, no stability computation.
 * State in/out (host): U, L = m_U, m_L; DeltaUold (may be NULL = 0), DeltaLold = the previous step (setPrevious: U - Uprev).
 * status 0: converged, state advanced; 1: NotConverged / 2: AssemblyError / 3: SolverError: state untouched, the caller
 * halves arc_length and retries (gsAPALM.hpp:975-983). */
typedef struct kl_alm_options {
    double tolU, tolF;          /* "TolU" (1e-6), "TolF" (1e-3)                                       */
    int32_t max_it;             /* "MaxIter" (100)                                                    */
    double phi;                 /* "Scaling": >= 0 fixed, < 0 automatic (gsALMCrisfield.hpp:26)        */
    double relaxation;          /* "Relaxation" (1)                                                   */
    double cg_tol;              /* <= 0: Eigen default DBL_EPSILON                                    */
    int32_t cg_max_iter;        /* <= 0: Eigen default 2 n                                            */
} kl_alm_options;
typedef struct kl_alm_info {
    int32_t status;             /* gsStatus                                                           */
    int32_t iterations;         /* m_numIterations                                                    */
    int64_t cg_iterations;
    double residueF, residueU, phi, DeltaL;
    float ms_assembly, ms_solve;
} kl_alm_info;
int kl_alm_step(kl_ctx* ctx, double* U_host_inout, double* L_inout, double* DeltaUold_host_inout, double* DeltaLold_inout,
                double arc_length, const kl_alm_options* opt, kl_alm_info* info);

/* ---- stability indicator of the arc-length solvers ---------------------------------------------------------
 * gsALMBase<T>::_computeStability / gsStaticBase<T>::_computeStabilityDet with the "Determinant" method
 * (src/gsALMSolvers/gsALMBase.hpp:546-611, src/gsStaticSolvers/gsStaticBase.h:161-179): m_stabilityVec = SimplicialLDLT::vectorD(),
 * m_negatives = countNegatives(vectorD), m_indicator = min(vectorD), stability = sign(m_indicator).  The matrix is the one the last
 * kl_jacobian / kl_jacobian_device call left on the device; it is factorised there (banded L D L^T without pivoting in a
 * node-major ordering, 8 n (bw + 33) bytes of scratch).  negatives and the SIGN of the indicator do not depend on the ordering
 * (Sylvester's law of inertia) and therefore agree with the reference; the value of the indicator is the smallest pivot of this
 * ordering, as Eigen's is of its AMD ordering.  vectorD_host (n_dofs entries, the pivot of every DoF) may be NULL.
 * Refused for unsymmetric tangents (follower pressure) and on a multi-patch. */
int kl_stability(kl_ctx* ctx, double* indicator, int32_t* negatives, double* vectorD_host);

/* ---- stress / stretch recovery -----------------------------------------------
 * Replaces assembler->constructStress(mp_def, field, stress_type::X) followed by field evaluation
 * (benchmarks/benchmark_Balloon.cpp:381-408, benchmark_TensionWrinkling.cpp:505-540, benchmark_Pillow.cpp:431,484-505),
 * assembler->computePrincipalStretches(pts, mp_def, z) and assembler->boundaryForce(mp_def, patchSide)
 * (unittests/gsStaticSolver_test.cpp:317,321).  The gsKLShell sources that define these quantities are not in the
 * reference tree; the definitions below are the ones this library and its oracle implement (PARITY UNPINNED against
 * upstream; pinned by the reference's own uniaxial-tension test: lambda, S = F_side/(t lambda0 lambda2)).
 * Voigt order of 3-component outputs: (11, 22, 12), tensor components (no engineering factor 2).
 *   DISPLACEMENT              3  u = x_def - x_ori
 *   MEMBRANE_FORCE            3  N^ab, thickness-integrated 2nd Piola-Kirchhoff stress (MaterialOutput::VectorN), curvilinear
 *   FLEXURAL_MOMENT           3  M^ab (MaterialOutput::VectorM), curvilinear
 *   MEMBRANE                  3  Cauchy membrane stress N^ab/(t J) pushed to the local Cartesian frame of the DEFORMED
 *                                mid-surface (e1 = a_1/|a_1|, e2 = a^2/|a^2|), J = J0 lambda3
 *   FLEXURAL                  3  outer-fibre Cauchy bending stress 6 M^ab/(t^2 J) in the same frame
 *   MEMBRANE_STRAIN           3  Green-Lagrange E_ab = (a_ab - A_ab)/2 in the local Cartesian frame of the UNDEFORMED surface
 *   FLEXURAL_STRAIN           3  K_ab = B_ab - b_ab in the same frame
 *   PRINCIPAL_STRETCH         3  lambda(0) <= lambda(1) in-plane at height z, lambda(2) ALWAYS the thickness stretch
 *                                (ordering of unittests/gsStaticSolver_test.cpp:313): 1/J0 for SvK and incompressible laws,
 *                                sqrt(C33) of the plane-stress iteration for compressible laws
 *   PRINCIPAL_STRETCH_DIR     9  spatial unit vectors n_0, n_1 (= F N_i / lambda_i) and the deformed normal; sign arbitrary
 *   PRINCIPAL_STRESS_MEMBRANE 2  eigenvalues of MEMBRANE, ascending;  PRINCIPAL_STRESS_FLEXURAL likewise of FLEXURAL
 *   PRINCIPAL_MEMBRANE_STRAIN 2  eigenvalues of MEMBRANE_STRAIN;      PRINCIPAL_FLEXURAL_STRAIN of FLEXURAL_STRAIN
 *   VON_MISES_MEMBRANE        1  sqrt(s11^2 + s22^2 - s11 s22 + 3 s12^2) of MEMBRANE
 *   TENSION_FIELD             1  1 taut (min principal membrane stress > 0), -1 slack (max principal membrane strain <= 0),
 *                                0 wrinkled otherwise                                                               */
enum {
    KL_STRESS_DISPLACEMENT = 0, KL_STRESS_MEMBRANE_FORCE = 1, KL_STRESS_FLEXURAL_MOMENT = 2, KL_STRESS_MEMBRANE = 3,
    KL_STRESS_FLEXURAL = 4, KL_STRESS_MEMBRANE_STRAIN = 5, KL_STRESS_FLEXURAL_STRAIN = 6, KL_STRESS_PRINCIPAL_STRETCH = 7,
    KL_STRESS_PRINCIPAL_STRETCH_DIR = 8, KL_STRESS_PRINCIPAL_STRESS_MEMBRANE = 9, KL_STRESS_PRINCIPAL_STRESS_FLEXURAL = 10,
    KL_STRESS_PRINCIPAL_MEMBRANE_STRAIN = 11, KL_STRESS_PRINCIPAL_FLEXURAL_STRAIN = 12, KL_STRESS_VON_MISES_MEMBRANE = 13,
    KL_STRESS_TENSION_FIELD = 14, KL_STRESS_NTYPES = 15
};
int kl_stress_dim(int32_t type);                  /* components per point, 0 for an unknown type */
/* Evaluate one quantity at n_pts parametric points uv[2*k..] of the patch for the state x (free DoFs, host); z = height
 * through the thickness used by the stretch outputs (computePrincipalStretches' third argument).  out: [n_pts*dim], point-major. */
int kl_eval_stress(kl_ctx* ctx, const double* x_host, int32_t type, int32_t n_pts, const double* uv_host, double z,
                   double* out_host);
/* assembler->computePrincipalStretches(pts, mp_def, z): out [n_pts*3] */
int kl_principal_stretches(kl_ctx* ctx, const double* x_host, int32_t n_pts, const double* uv_host, double z, double* out_host);
/* assembler->boundaryForce(mp_def, patchSide(0, side)): -F_int (the sign of rhs() = F_ext - F_int; no follower pressure)
 * summed per component over ALL control points of the side, eliminated ones included, so that the Cauchy stress of the
 * reference's test reads S = -sideForce / (thickness lambda(0) lambda(2)) (unittests/gsStaticSolver_test.cpp:321-323).
 * out3: (Fx, Fy, Fz). */
int kl_boundary_force(kl_ctx* ctx, const double* x_host, int32_t side, double* out3_host);

/* ---- multi-patch: several conforming patches glued C0 along whole sides ---------------------------------------------
 * Replaces a gsMultiPatch with computeTopology() / addInterface() handed to gsThinShellAssembler(mp, dbasis, bc, force, mm)
 * (benchmarks/benchmark_Wrinkling.cpp:446-522, benchmark_cylinder_DC.cpp:146): the interfaces are matched inside the common
 * gsDofMapper (gsMultiBasis::matchInterface -> gsDofMapper::matchDofs) and the element loop pushes every patch into
 * ONE matrix.  kl_interface: side[0] of patch[0] is glued to side[1] of patch[1]; the k-th function along side 0 meets the k-th
 * (reversed != 0: the (len-1-k)-th) function along side 1; both sides must carry the same number of functions (conforming).
 * kl_mp_build_dofmap numbers all patches at once (patch-major within a component, otherwise as kl_build_dofmap); dof_map is
 * the concatenation of the per-patch maps: patch q starts at 3 * sum_{r<q} n1[r]*n2[r] and holds [c * ncp_q + i]. */
typedef struct kl_interface { int32_t patch[2]; int32_t side[2]; int32_t reversed; } kl_interface;
int kl_mp_build_dofmap(int32_t n_patches, const int32_t* n1, const int32_t* n2, const kl_bc* bc, int32_t n_interfaces,
                       const kl_interface* interfaces, int32_t* dof_map, int32_t* n_free, int32_t* n_fixed);
/* One kl_problem per patch; every dof_map holds GLOBAL indices, n_free / n_fixed / fixed_values are the global ones (identical in
 * every patch).  Material, loads, degree and knots may differ from patch to patch. */
typedef struct kl_mp kl_mp;
int kl_mp_create(int32_t n_patches, const kl_problem* patches, int device, kl_mp** out);
void kl_mp_destroy(kl_mp* mp);
/* The assembler handle of the whole multi-patch: a kl_ctx accepted by every matrix / vector level entry point of this header
 * (kl_sizes, kl_pattern_*, kl_jacobian[_device|_lower], kl_residual[_device], kl_al_residual[_device], kl_assemble_device,
 * kl_force, kl_mass, kl_check, kl_fetch_values / kl_set_values / kl_pin_values, kl_cg_solve, kl_spmv, kl_newton_solve, kl_alm_step).
 * It is owned by the kl_mp (do not kl_destroy it).  Geometry-level calls (kl_eval_stress, kl_principal_stretches,
 * kl_boundary_force) take the context of ONE patch, kl_mp_patch(mp, q), with the global solution vector. */
kl_ctx* kl_mp_context(kl_mp* mp);
kl_ctx* kl_mp_patch(kl_mp* mp, int32_t q);
int32_t kl_mp_num_patches(const kl_mp* mp);
/* Patch -> GPU partition: active[q] != 0 for the patches THIS process assembles (NULL = all).  The process then
 * produces partial sums: its patches' share of K, of F_int and of the load vectors; the columns of the interface DoFs
 * (kl_mp_interface_dofs: count, then the ascending list) and the vectors are completed by the reduce of the host layer. */
int kl_mp_set_active(kl_mp* mp, const int32_t* active);
int kl_mp_interface_dofs(const kl_mp* mp, int32_t* count, int32_t* dofs);

/* Duration in ms of the last Jacobian kernel launch itself (CUDA events on its stream; syncs). */
int kl_jacobian_kernel_ms(kl_ctx* ctx, float* ms);
/* Same for the per-quadrature-point kernel (geometry + material) that precedes it. */
int kl_points_kernel_ms(kl_ctx* ctx, float* ms);
/* FP64 FMA peak of the current device measured with a register-resident DFMA chain kernel
 * (MEASURED_PEAKS.json has no FP64 figure).  Returns TFLOP/s in *tflops. */
int kl_measure_fp64_peak(int device, double* tflops, float* ms);
const char* kl_last_error(void);
int kl_kernel_launches(const kl_ctx* ctx);        /* number of own kernels launched so far */

#ifdef __cplusplus
}
#endif
#endif /* KL_SHELL_H */
