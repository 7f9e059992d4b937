// kl_internal.h — context and device-side views shared by the translation units of libkl_shell.so
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <string>
#include <vector>
#include "../../include/kl_shell.h"

#define KL_MAXP 4

// device error flag bits (mapped to KL_E_* by kl_check)
#define KLF_NONFINITE 1
#define KLF_JACOBIAN 2
#define KLF_C33 4
#define KLF_METRIC 8      // det of a (through-thickness) metric <= 0 inside the material law

struct PointData;
struct KLSolveWS;   // kl_solve.cu: workspace of the device-resident CG / Newton loop
struct kl_mp;

struct KLMaterial {
    int material, compressible, ngauss, bending, metric_z2;
    double E, nu, t, mu, lam_ps, bulk, c1, c2;
    double zg[12], wg[12];   // thickness Gauss rule on [-1,1]
    double pressure;
};

// Everything a kernel needs (passed by value as a __grid_constant__-friendly POD)
struct KLDev {
    int p;                 // degree (both directions)
    int nq;                // Gauss nodes per direction
    int n1, n2, ncp;       // control points
    int nel1, nel2;        // non-empty elements per direction
    int nfree;
    int rational;
    const int* span1;      // [nel1] knot span index per element
    const int* span2;
    const double* knots1;  // knot vectors (point evaluation of the stress recovery)
    const double* knots2;
    const double* bas1;    // [nel1][nq][3][p+1] values / 1st / 2nd derivative of the p+1 active functions
    const double* bas2;
    const double* wq1;     // [nel1][nq] quadrature weight * half element length
    const double* wq2;
    const double* cp;      // [ncp*3] undeformed control net
    const double* w;       // [ncp] weights or nullptr
    const int* map;        // [3*ncp]
    const double* fixed;   // [nfixed]
    double* disp;          // [ncp*3] displacement control net (constructSolution)
    // sparse matrix
    const int* outer;      // [nfree+1]
    const int* inner;      // [nnz]
    double* values;        // [nnz]
    const int* pos;        // [ncp*3][nst*3] scatter table (-1: eliminated / not coupled)
    int nst;               // (2p+1)^2
    const int* colbase;    // [ncp][4]: outer[map[d][J]] for d=0..2 (-1 if eliminated), [3] = 1 if the column block of J is regular
    int* flag;             // device error flag
    double* lift;          // set-up only: accumulates K(free row, eliminated column) * fixed value (lifting of non-zero Dirichlet values)
    PointData* pd;         // [elements][nq*nq] per-point records written by k_points
    int ablate;            // profiling only (env KL_ABLATE): 1 skip scatter, 2 skip phase 3, 4 skip phase 2
    KLMaterial mat;
};

struct kl_ctx {
    int device = 0;
    KLDev d{};
    kl_problem prob{};               // scalar copy
    std::vector<double> U[2];
    std::vector<int> span[2];
    std::vector<int> flo[2], fhi[2]; // element range of each 1-D function
    int* d_flohi[4] = {nullptr, nullptr, nullptr, nullptr};   // device copies (pattern build): lo1, hi1, lo2, hi2
    int nfixed = 0;
    int64_t nnz = 0;
    int e2_begin = 0, e2_end = 0;    // strip of element rows assembled by this context
    std::vector<int> h_map;          // host copy of the DoF map [3*ncp]
    std::vector<int> h_outer;        // host copy of outer [nfree+1]
    std::vector<std::pair<size_t, size_t>> zero_ranges;   // value ranges the strip contributes to (zeroed per call); empty = whole array
    int cp_row_begin = 0, cp_row_end = 0;                 // control-point rows the strip reads (constructSolution range)
    // owned device buffers
    std::vector<void*> owned;
    double* d_x = nullptr;           // [nfree]
    double* d_r = nullptr;           // [nfree]
    double* d_fext = nullptr;        // [nfree] dead loads: body force, point loads, Neumann tractions
    double* d_force = nullptr;       // [nfree] Force = assemble().rhs(): dead loads + follower pressure on the undeformed surface - Dirichlet lifting
    double* h_pinned_x = nullptr;    // pinned staging for x / r
    double* h_pinned_r = nullptr;
    double* h_stage = nullptr;       // context-owned page-locked staging for copy-outs into pageable caller memory (lazy)
    size_t h_stage_bytes = 0;
    // lower-triangular view (row >= col) for LDLT consumers: built on the first kl_pattern_lower_host / kl_jacobian_lower
    int64_t nnz_lower = 0;
    int* d_outer_lower = nullptr;    // [nfree+1]
    int* d_inner_lower = nullptr;    // [nnz_lower]
    double* d_values_lower = nullptr;// [nnz_lower] packed values
    std::vector<int> h_outer_lower;
    cudaStream_t stream = nullptr;   // own stream for the host-pointer entry points
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev[8]{};
    float ms_kernel = 0, ms_h2d = 0, ms_d2h = 0;
    int launches = 0;
    unsigned attr_done = 0;          // per-context (= per-device) cudaFuncSetAttribute bookkeeping, bit per kernel family
    int n_sm = 0;
    // state cache (same-state fusion behind the separate Residual_t / Jacobian_t closures): d_xstate = the solution vector the per-point
    // records in d.pd were computed for; a Jacobian call at a bit-identical state skips constructSolution + the point kernel
    double* d_xstate = nullptr;      // [nfree]
    int* d_same = nullptr;           // device word: != 0 when the current call's x equals d_xstate
    int pd_valid = 0;                // d.pd / d_xstate hold an evaluation (tangent included) of the element rows [pd_e2b, pd_e2e)
    int pd_e2b = 0, pd_e2e = 0;
    int state_null = 0;              // that state was x == NULL (undeformed configuration, Dirichlet values not applied)
    int spec_on = 1;                 // residual calls also write the per-point records (speculating on a Jacobian at the same state)
    int last_call = 0;               // 1 residual (speculative), 2 jacobian, 0 other
    int spec_allowed = 1;            // env KL_SPECULATE=0 disables the speculation (A/B)
    int jac_shared = 0;              // env KL_JAC_SHARED=1: the shared-memory tile kernel k_jacobian instead of k_jacobian_sw (A/B, P = 3)
    int jac_seg = 0;                 // env KL_SW_SEG: elements per segment of k_jacobian_sw (0 = automatic)
    int n_strips_d2h = 16;           // pipelined D2H granularity (measured 8 / 16 / 32: e2e 23.32 / 23.05 / 23.07 ms per step)
    struct D2HStrip { int e2_begin, e2_end; std::vector<std::pair<size_t, size_t>> ranges; std::vector<std::pair<int, int>> cols; };   // value ranges and column ranges complete after the strip
    std::vector<D2HStrip> d2h_plan;
    std::vector<cudaEvent_t> strip_ev, copy_ev;
    KLSolveWS* solve_ws = nullptr;   // created on the first kl_cg_solve / kl_newton_solve
    // multi-patch (kl_multipatch.cu): a patch context assembles into the matrix / vectors of its kl_mp (d.outer / d.inner / d.values are
    // shared, d.nfree is the global count); the kl_mp's own "matrix context" has no geometry at all (d.ncp == 0)
    int mp_member = 0;
    struct kl_mp* mp = nullptr;      // set on the matrix context of a kl_mp: the assembly entry points dispatch to the patches
};

// Several C0-coupled patches with one gsDofMapper numbering and one sparse matrix (include/kl_shell.h: kl_mp_*)
struct kl_mp {
    int device = 0;
    std::vector<kl_ctx*> patch;      // geometry, tables, scatter maps and per-point records of every patch
    std::vector<int> active;         // patch -> GPU partition: only active patches are assembled by this process
    kl_ctx* g = nullptr;             // matrix context: global pattern / values / vectors / streams (no geometry); accepted by kl_cg_solve,
                                     // kl_spmv, kl_fetch_values, kl_set_values, kl_pattern_host, kl_sizes
    int64_t n_elements = 0, n_qp = 0;
    std::vector<int> coupled_cols;   // global DoFs shared by more than one patch (interface columns), ascending
    // the patches of one assembly run side by side (fork / join around the caller's stream): they only meet in the RED scatter
    std::vector<cudaStream_t> pstream;
    std::vector<cudaEvent_t> pdone;
    cudaEvent_t fork = nullptr;
};

void kl_set_error(const std::string& s);
#define KL_CUDA(call)                                                                          \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess) {                                                               \
            kl_set_error(std::string(#call) + ": " + cudaGetErrorString(e_));                  \
            return KL_E_CUDA;                                                                  \
        }                                                                                      \
    } while (0)

// kl_capi.cu: host-side 1-D basis / quadrature helpers shared with ks_solid.cu
void bspline_span_ders(const std::vector<double>& U, int p, int k, double u, double out[3][KL_MAXP + 1]);
void gauss_rule(int n, double* x, double* w);
int kl_ctx_create_base(const kl_problem* P, int device, kl_ctx** out);   // everything that does not depend on the sparse pattern
int kl_ctx_finish(kl_ctx* ctx, const kl_problem* P);                      // load vectors (+ copy-out plan of a stand-alone patch)
int kl_residual_accumulate(kl_ctx* ctx, const double* x_dev, double* r_dev, cudaStream_t s);   // r += F_int(x) - P(x) of the context's elements
// kl_multipatch.cu: what the matrix context of a kl_mp dispatches to
int kl_mp_jacobian_device(kl_mp* mp, const double* x_dev, cudaStream_t s);
int kl_mp_residual_device(kl_mp* mp, const double* x_dev, double lam_fext, double sign_fint, double* r_dev, cudaStream_t s);
int kl_mp_mass_device(kl_mp* mp, double density, double* values, double* lumped, cudaStream_t s);
int kl_mp_check(kl_mp* mp, cudaStream_t s);
int kl_mp_launches(const kl_mp* mp);
// kl_solve.cu
void kl_solve_free(kl_ctx* ctx);
// kl_pattern.cu
int kl_build_pattern(kl_ctx* ctx);                                  // single patch: keys + compress + tables + value array
long long kl_pattern_key_count(const kl_ctx* ctx);
int kl_pattern_gen_keys(kl_ctx* ctx, unsigned long long* keys_dev);                 // 9 * nst * ncp keys (col << 32 | row) of one patch
int kl_pattern_compress(kl_ctx* owner, unsigned long long* keys, unsigned long long* keys_alt, long long total, int nfree,
                        int** outer, int** inner, long long* nnz);                  // sort + unique -> compressed pattern owned by `owner`
int kl_pattern_tables(kl_ctx* ctx);                                 // scatter table + colbase of one patch against d.outer / d.inner
int kl_lower_tables(kl_ctx* ctx);                                   // lazily builds the lower-triangular pattern + per-strip packed ranges
int kl_launch_pack_lower(kl_ctx* ctx, int col_begin, int col_end, cudaStream_t s);   // packed lower values of the columns [col_begin, col_end)
// kl_assemble.cu
int kl_launch_construct(kl_ctx* ctx, const double* x_dev, cudaStream_t s, const int* skip = nullptr);   // rows [cp_row_begin, cp_row_end)
int kl_launch_state_compare(kl_ctx* ctx, const double* x_dev, cudaStream_t s);   // *d_same = (x == d_xstate), then d_xstate = x
int kl_launch_jacobian(kl_ctx* ctx, int e2_begin, int e2_end, cudaStream_t s);
int kl_launch_points(kl_ctx* ctx, int e2_begin, int e2_end, cudaStream_t s, double* r_dev = nullptr, const int* skip = nullptr);   // r_dev: also r += F_int - F_pressure
int kl_launch_residual(kl_ctx* ctx, double* r_dev, cudaStream_t s, bool full = false);   // r += F_int - F_pressure (atomic); full: r = [3][ncp], internal force at every control point
size_t kl_pointdata_bytes(void);
int kl_launch_bodyforce(kl_ctx* ctx, double* f_dev, const double bf[3], cudaStream_t s);
int kl_launch_mass(kl_ctx* ctx, double rho_t, double* values, double* lumped, cudaStream_t s);
int kl_launch_axpby(kl_ctx* ctx, double* r, const double* fext, double a_r, double b_f, int n, cudaStream_t s);
