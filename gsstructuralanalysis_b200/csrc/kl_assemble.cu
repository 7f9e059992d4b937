// kl_assemble.cu — hand-written sm_100a FP64 kernels of the hot path:
//   k_construct_solution (gsThinShellAssembler::constructSolution, tutorials/nonlinear_shell_static.cpp:123)
//   k_points             geometry + metric + material at every quadrature point  -> PointData (HBM)
//   k_residual           (assembleVector,  tutorials/nonlinear_shell_static.cpp:133)
//   k_jacobian           (assembleMatrix,  tutorials/nonlinear_shell_static.cpp:124)
// Formulation: SURVEY Appendix A.3/A.4; in-reference restatement benchmarks/benchmark_cylinder_DC.cpp:536-555.
//
// Jacobian kernel structure (one CTA works on a group of EPG consecutive elements):
//   phase 2  one thread per (basis function j, point): Z_j = T(point) . d_j   (9 x 5 coefficients) -> smem
//   phase 3  one thread per tile (row of P+1 functions i, function j), upper triangle only:
//            K_ij^{cd} += sum_p d_i[p] Z_j^{p,cd}, sum-factorised over the tensor-product basis
//   scatter  FP64 RED (atomicAdd, no return) into the compressed values through the position table,
//            both (i,j) and the transposed (j,i) entry.
#include <algorithm>
#include <cstddef>
#include <cstdlib>
#include "kl_device.cuh"

size_t kl_pointdata_bytes(void) { return sizeof(PointData); }

// ------------------------------------------------------------------------------------------------
__global__ void k_construct_solution(KLDev d, const double* __restrict__ x, const int* __restrict__ skip, int i_begin, int i_count) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= 3 * i_count || (skip && *skip)) return;
    const int c = k / i_count, i = i_begin + (k - c * i_count);
    const int g = d.map[c * d.ncp + i];
    double v;
    if (!x) v = 0.0;                 // x == NULL: the undeformed configuration (assemble() of the linear system), eliminated DoFs included
    else if (g < d.nfree) v = x[g];
    else v = d.fixed ? d.fixed[g - d.nfree] : 0.0;
    d.disp[3 * i + c] = v;
}

// same-state detection: *same stays non-zero iff x is bit-identical to the state the per-point records were computed for;
// the reference's Newton / arc-length loops call Residual(x) and then Jacobian(x) at one state
// (src/gsStaticSolvers/gsStaticNewton.hpp:160-191), and the closures are separate std::functions
__global__ void k_state_compare(const double* __restrict__ x, const double* __restrict__ xs, int n, int* __restrict__ same) {
    bool diff = false;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x)
        diff |= __double_as_longlong(x ? x[k] : 0.0) != __double_as_longlong(xs[k]);
    if (__any_sync(0xffffffffu, diff) && (threadIdx.x & 31) == 0) *same = 0;
}
__global__ void k_state_store(const double* __restrict__ x, double* __restrict__ xs, int n) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) xs[k] = x ? x[k] : 0.0;
}

__global__ void k_axpby(double* __restrict__ r, const double* __restrict__ f, double a, double b, int n) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) r[k] = a * r[k] + b * f[k];
}

// ------------------------------------------------------------------------------------------------
template <int P>
struct PointCfg {
    static constexpr int NQ2 = (P + 1) * (P + 1);
// p = 3: elements per CTA of the point kernels and the CTAs per SM they are compiled for.  Measured (profiles/r3_ablation.txt): 8 / 3 ->
// 0.782 ms, 6 / 4 -> 0.799, 4 / 6 -> 0.748, 2 / 12 -> 0.745: smaller CTAs stagger their load / compute / store phases better.
#ifndef KL_PTS_EPG
#define KL_PTS_EPG 4
#endif
#ifndef KL_PTS_MINB
#define KL_PTS_MINB 6
#endif
    static constexpr int EPG = (P == 2) ? 14 : (P == 3 ? KL_PTS_EPG : 5);   // elements per CTA
    static constexpr int NT = EPG * NQ2;
};

// phase 1: one thread per quadrature point.  The CTA's records are contiguous in HBM, so they are staged in
// shared memory (q1-major per element) and written back with ONE TMA bulk store instead of 56 strided stores
// per thread.
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

// WITH_RES: also integrate the internal force F_int - F_pressure into r from the staged records (RED), which makes the separate
// k_residual pass unnecessary when Jacobian and residual are wanted at the same state (kl_assemble_device).
template <int P, bool WITH_RES>
__global__ void __launch_bounds__(PointCfg<P>::NT, KL_PTS_MINB) k_points(KLDev d, int e2_begin, int e2_end, double* __restrict__ r, const int* __restrict__ skip) {
    using Cfg = PointCfg<P>;
    if (skip && *skip) return;     // the records of this state are already in d.pd (same-state fusion)
    constexpr int NQ = P + 1, NQ2 = Cfg::NQ2, EPG = Cfg::EPG;
    extern __shared__ __align__(16) unsigned char smem_pts[];
    ElemStage<P>* stage = reinterpret_cast<ElemStage<P>*>(smem_pts);
    PointData* out = reinterpret_cast<PointData*>(smem_pts + ((sizeof(ElemStage<P>) * EPG + 15) / 16) * 16);
    const int tid = threadIdx.x;
    const int le = tid / NQ2, lq = tid - le * NQ2;
    const int nel = d.nel1 * (e2_end - e2_begin);
    const int ebase = blockIdx.x * EPG;
    const int e = ebase + le;
    const bool active = e < nel;
    const int e1 = active ? e % d.nel1 : 0, e2 = active ? e2_begin + e / d.nel1 : e2_begin;
    stage_element<P>(d, e1, e2, stage[le], lq, NQ2);
    __syncthreads();
    const int flag = eval_point<P>(d, stage[le], lq % NQ, lq / NQ, out[le * NQ2 + (lq / NQ) + NQ * (lq % NQ)]);   // [q1][q2]
    if (flag && active) atomicOr(d.flag, flag);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // make the generic-proxy writes visible to the TMA engine
    __syncthreads();
    if (tid == 0) {
        const int ne = min(EPG, nel - ebase);
        PointData* dst = d.pd + (size_t)((ebase % d.nel1) + d.nel1 * (e2_begin + ebase / d.nel1)) * NQ2;
        const unsigned bytes = (unsigned)(ne * NQ2 * sizeof(PointData));
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(out)), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    if (WITH_RES) {
        // internal force, sum-factorised: f_(a,b),c = sum_q2 [ y0 T0_c + y1 T1_c + y2 T2_c ](a,q2) with
        //   T0_c = sum_q1 (x1 N0) a1_c + (x1 N2) a2_c + (x1 Ha1 - x2 Mt0 - x0 p wJ) n_c
        //   T1_c = sum_q1 (x0 N2) a1_c + (x0 N1) a2_c + (x0 Ha2 - x1 Mt2) n_c          T2_c = sum_q1 (-x0 Mt1) n_c
        // (x, y: 1-D values / derivatives in the two directions).  Stage A: thread (a, q2) sums over q1; stage B: thread (a, b) over
        // q2.  137 instead of 368 shared-memory accesses per thread; the partial sums live where the control points were staged.
        const ElemStage<P>& E = stage[le];
        double* T = &stage[le].w1[0];
        static_assert(offsetof(ElemStage<P>, scratch) + sizeof(E.scratch) - offsetof(ElemStage<P>, w1) >= sizeof(double) * 9 * NQ2,
                      "partial sums must fit the dead part of the staging area");
        {
            const int a = lq % (P + 1), q2 = lq / (P + 1);
            double T0[3] = {0, 0, 0}, T1[3] = {0, 0, 0}, T2[3] = {0, 0, 0};
#pragma unroll
            for (int q1 = 0; q1 < NQ; ++q1) {
                const PointData& pd = out[le * NQ2 + q2 + NQ * q1];
                const double x0 = E.b1[q1][0][a], x1 = E.b1[q1][1][a], x2 = E.b1[q1][2][a];
                const double s0a1 = x1 * pd.N[0], s0a2 = x1 * pd.N[2], s0n = x1 * pd.Ha1 - x2 * pd.Mt[0] - x0 * d.mat.pressure * pd.wJ;
                const double s1a1 = x0 * pd.N[2], s1a2 = x0 * pd.N[1], s1n = x0 * pd.Ha2 - x1 * pd.Mt[2];
                const double s2n = -(x0 * pd.Mt[1]);
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    T0[c] += s0a1 * pd.a1[c] + s0a2 * pd.a2[c] + s0n * pd.n[c];
                    T1[c] += s1a1 * pd.a1[c] + s1a2 * pd.a2[c] + s1n * pd.n[c];
                    T2[c] += s2n * pd.n[c];
                }
            }
            double* t = T + 9 * lq;     // [q2][a][9]
#pragma unroll
            for (int c = 0; c < 3; ++c) { t[c] = T0[c]; t[3 + c] = T1[c]; t[6 + c] = T2[c]; }
        }
        __syncthreads();
        if (active) {
            const int a = lq % (P + 1), b = lq / (P + 1);
            double f[3] = {0, 0, 0};
#pragma unroll
            for (int q2 = 0; q2 < NQ; ++q2) {
                const double y0 = E.b2[q2][0][b], y1 = E.b2[q2][1][b], y2 = E.b2[q2][2][b];
                const double* t = T + 9 * (q2 * (P + 1) + a);
#pragma unroll
                for (int c = 0; c < 3; ++c) f[c] += y0 * t[c] + y1 * t[3 + c] + y2 * t[6 + c];
            }
            const int cpi = (d.span1[e1] - P + a) + d.n1 * (d.span2[e2] - P + b);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const int g = d.map[c * d.ncp + cpi];
                if (g < d.nfree) atomicAdd(&r[g], f[c]);
            }
        }
    }
    if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // smem must stay valid until the engine has read it
}

// residual: F_int - F_pressure accumulated into r (atomic); reads PointData
struct ResPoint {   // what the force integrand needs per point
    double q1[3], q2[3], nM[3][3], pn[3];
};
template <int P>
struct BasisStage {
    static constexpr int NQ = P + 1;
    double b1[NQ][3][P + 1];
    double b2[NQ][3][P + 1];
};
template <int P>
__device__ __forceinline__ void stage_basis(const KLDev& d, int e1, int e2, BasisStage<P>& E, int t, int nthr) {
    constexpr int NB = (P + 1) * 3 * (P + 1);
    const double* g1 = d.bas1 + (size_t)e1 * NB;
    const double* g2 = d.bas2 + (size_t)e2 * NB;
    double* s1 = &E.b1[0][0][0];
    double* s2 = &E.b2[0][0][0];
    for (int k = t; k < NB; k += nthr) { s1[k] = g1[k]; s2[k] = g2[k]; }
}

// FULL: internal force only (no follower pressure) at ALL control points, r = [3][ncp] — what boundaryForce sums over a side
template <int P, bool FULL = false>
__global__ void __launch_bounds__(PointCfg<P>::NT, 3) k_residual(KLDev d, double* __restrict__ r, int e2_begin, int e2_end) {
    using Cfg = PointCfg<P>;
    constexpr int NQ = P + 1, NQ2 = Cfg::NQ2, EPG = Cfg::EPG;
    __shared__ ElemStage<P> stage[EPG];
    __shared__ ResPoint rp[EPG][NQ2];
    const int tid = threadIdx.x;
    const int le = tid / NQ2, lq = tid - le * NQ2;
    const int nel = d.nel1 * (e2_end - e2_begin);
    const int e = blockIdx.x * EPG + le;
    const bool active = e < nel;
    const int e1 = active ? e % d.nel1 : 0, e2 = active ? e2_begin + e / d.nel1 : e2_begin;
    stage_element<P>(d, e1, e2, stage[le], lq, NQ2);
    __syncthreads();
    {
        PointData pd;
        const int flag = eval_point<P, false>(d, stage[le], lq % NQ, lq / NQ, pd);
        if (flag && active) atomicOr(d.flag, flag);
        ResPoint& o = rp[le][lq];
        const double pw = FULL ? 0.0 : d.mat.pressure * pd.wJ;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            // coefficient of N_a,1 and N_a,2:  N:dEm  +  n_c * (H . a^gamma)   (the Christoffel part of M:dEf)
            o.q1[c] = pd.N[0] * pd.a1[c] + pd.N[2] * pd.a2[c] + pd.n[c] * pd.Ha1;
            o.q2[c] = pd.N[1] * pd.a2[c] + pd.N[2] * pd.a1[c] + pd.n[c] * pd.Ha2;
            o.nM[0][c] = -pd.n[c] * pd.Mt[0];
            o.nM[1][c] = -pd.n[c] * pd.Mt[1];
            o.nM[2][c] = -pd.n[c] * pd.Mt[2];
            o.pn[c] = -pw * pd.n[c];
        }
    }
    __syncthreads();
    // sum-factorised integration (as in k_points<P, true>): stage A, thread (a, q2): sums over q1; stage B, thread (a, b): over q2
    const ElemStage<P>& E = stage[le];
    double* T = &stage[le].w1[0];     // dead part of the staging area: 9 (P+1)^2 doubles
    {
        const int a = lq % (P + 1), q2 = lq / (P + 1);
        double T0[3] = {0, 0, 0}, T1[3] = {0, 0, 0}, T2[3] = {0, 0, 0};
#pragma unroll
        for (int q1 = 0; q1 < NQ; ++q1) {
            const ResPoint& o = rp[le][q1 + NQ * q2];
            const double x0 = E.b1[q1][0][a], x1 = E.b1[q1][1][a], x2 = E.b1[q1][2][a];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                T0[c] += x1 * o.q1[c] + x2 * o.nM[0][c] + x0 * o.pn[c];     // multiplies N_b(q2)
                T1[c] += x0 * o.q2[c] + x1 * o.nM[2][c];                    // multiplies N_b'(q2)
                T2[c] += x0 * o.nM[1][c];                                   // multiplies N_b''(q2)
            }
        }
        double* t = T + 9 * lq;
#pragma unroll
        for (int c = 0; c < 3; ++c) { t[c] = T0[c]; t[3 + c] = T1[c]; t[6 + c] = T2[c]; }
    }
    __syncthreads();
    if (active) {
        const int a = lq % (P + 1), b = lq / (P + 1);
        double f[3] = {0, 0, 0};
#pragma unroll
        for (int q2 = 0; q2 < NQ; ++q2) {
            const double y0 = E.b2[q2][0][b], y1 = E.b2[q2][1][b], y2 = E.b2[q2][2][b];
            const double* t = T + 9 * (q2 * (P + 1) + a);
#pragma unroll
            for (int c = 0; c < 3; ++c) f[c] += y0 * t[c] + y1 * t[3 + c] + y2 * t[6 + c];
        }
        const int cpi = (d.span1[e1] - P + a) + d.n1 * (d.span2[e2] - P + b);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            if (FULL) { atomicAdd(&r[c * d.ncp + cpi], f[c]); continue; }
            const int g = d.map[c * d.ncp + cpi];
            if (g < d.nfree) atomicAdd(&r[g], f[c]);
        }
    }
}

// external body force (set-up only): f += integral N_a * bf * meas(ori)
template <int P>
__global__ void __launch_bounds__(PointCfg<P>::NT) k_bodyforce(KLDev d, double* __restrict__ f_out, double bfx, double bfy, double bfz) {
    using Cfg = PointCfg<P>;
    constexpr int NQ = P + 1, NQ2 = Cfg::NQ2, EPG = Cfg::EPG;
    __shared__ ElemStage<P> stage[EPG];
    __shared__ double wj[EPG][NQ2];
    const int tid = threadIdx.x;
    const int le = tid / NQ2, lq = tid - le * NQ2;
    const int nel = d.nel1 * d.nel2;
    const int e = blockIdx.x * EPG + le;
    const bool active = e < nel;
    const int e1 = active ? e % d.nel1 : 0, e2 = active ? e / d.nel1 : 0;
    stage_element<P>(d, e1, e2, stage[le], lq, NQ2);
    __syncthreads();
    {
        const int q1 = lq % NQ, q2 = lq / NQ;
        double fo[6][3];
        eval_field3<P, true>(stage[le].X, stage[le].b1[q1], stage[le].b2[q2], fo);
        double A1[3], A2[3], Nn[3];
        if (d.rational) {
            double fw[6];
            eval_field1<P>(stage[le].Wt, stage[le].b1[q1], stage[le].b2[q2], fw);
            const double iw = 1.0 / fw[0];
            for (int c = 0; c < 3; ++c) {
                const double X = fo[0][c] * iw;
                A1[c] = (fo[1][c] - fw[1] * X) * iw;
                A2[c] = (fo[2][c] - fw[2] * X) * iw;
            }
        } else {
            for (int c = 0; c < 3; ++c) { A1[c] = fo[1][c]; A2[c] = fo[2][c]; }
        }
        cross3(A1, A2, Nn);
        wj[le][lq] = stage[le].w1[q1] * stage[le].w2[q2] * sqrt(dot3(Nn, Nn));
    }
    __syncthreads();
    if (active) {
        const int a = lq % (P + 1), b = lq / (P + 1);
        double s = 0.0;
        for (int q2 = 0; q2 < NQ; ++q2)
            for (int q1 = 0; q1 < NQ; ++q1) s += stage[le].b1[q1][0][a] * stage[le].b2[q2][0][b] * wj[le][q1 + NQ * q2];
        const int cpi = (d.span1[e1] - P + a) + d.n1 * (d.span2[e2] - P + b);
        const double bf[3] = {bfx, bfy, bfz};
        for (int c = 0; c < 3; ++c) {
            const int g = d.map[c * d.ncp + cpi];
            if (g < d.nfree && bf[c] != 0.0) atomicAdd(&f_out[g], s * bf[c]);
        }
    }
}

// mass matrix / lumped mass (set-up path, not tuned): M_ab = rho*t * int N_a N_b meas(ori), identical for the 3 components
template <int P>
__global__ void __launch_bounds__(PointCfg<P>::NT) k_mass(KLDev d, double rho_t, double* __restrict__ values, double* __restrict__ lumped) {
    using Cfg = PointCfg<P>;
    constexpr int NQ = P + 1, NQ2 = Cfg::NQ2, EPG = Cfg::EPG, NLOC = (P + 1) * (P + 1), W = 2 * P + 1, S3 = W * W * 3;
    __shared__ ElemStage<P> stage[EPG];
    __shared__ double wj[EPG][NQ2];
    const int tid = threadIdx.x;
    const int le = tid / NQ2, lq = tid - le * NQ2;
    const int nel = d.nel1 * d.nel2;
    const int e = blockIdx.x * EPG + le;
    const bool active = e < nel;
    const int e1 = active ? e % d.nel1 : 0, e2 = active ? e / d.nel1 : 0;
    stage_element<P>(d, e1, e2, stage[le], lq, NQ2);
    __syncthreads();
    {
        const int q1 = lq % NQ, q2 = lq / NQ;
        double fo[6][3];
        eval_field3<P, true>(stage[le].X, stage[le].b1[q1], stage[le].b2[q2], fo);
        double A1[3], A2[3], Nn[3];
        if (d.rational) {
            double fw[6];
            eval_field1<P>(stage[le].Wt, stage[le].b1[q1], stage[le].b2[q2], fw);
            const double iw = 1.0 / fw[0];
            for (int c = 0; c < 3; ++c) {
                const double X = fo[0][c] * iw;
                A1[c] = (fo[1][c] - fw[1] * X) * iw;
                A2[c] = (fo[2][c] - fw[2] * X) * iw;
            }
        } else {
            for (int c = 0; c < 3; ++c) { A1[c] = fo[1][c]; A2[c] = fo[2][c]; }
        }
        cross3(A1, A2, Nn);
        wj[le][lq] = rho_t * stage[le].w1[q1] * stage[le].w2[q2] * sqrt(dot3(Nn, Nn));
    }
    __syncthreads();
    if (!active) return;
    const ElemStage<P>& E = stage[le];
    const int a = lq, a1 = a % (P + 1), a2 = a / (P + 1);
    const int i0 = d.span1[e1] - P, j0 = d.span2[e2] - P;
    const int I1 = i0 + a1, I2 = j0 + a2, Ic = I1 + d.n1 * I2;
    double rowsum = 0.0;
    for (int b = 0; b < NLOC; ++b) {
        const int b1 = b % (P + 1), b2 = b / (P + 1);
        double m = 0.0;
        for (int q2 = 0; q2 < NQ; ++q2)
            for (int q1 = 0; q1 < NQ; ++q1) m += E.b1[q1][0][a1] * E.b2[q2][0][a2] * E.b1[q1][0][b1] * E.b2[q2][0][b2] * wj[le][q1 + NQ * q2];
        rowsum += m;
        if (values) {
            const int J1 = i0 + b1, J2 = j0 + b2, Jc = J1 + d.n1 * J2;
            const int st = (I1 - J1 + P) + W * (I2 - J2 + P);
            for (int c = 0; c < 3; ++c) {
                const int pp = d.pos[(size_t)(Jc * 3 + c) * S3 + st * 3 + c];      // (row (I,c), col (J,c))
                if (pp >= 0) atomicAdd(&values[pp], m);
            }
        }
    }
    if (lumped)
        for (int c = 0; c < 3; ++c) {
            const int g = d.map[c * d.ncp + Ic];
            if (g < d.nfree) atomicAdd(&lumped[g], rowsum);
        }
}

template <int P>
static int launch_mass(kl_ctx* ctx, double rho_t, double* values, double* lumped, cudaStream_t s) {
    using Cfg = PointCfg<P>;
    const int nel = ctx->d.nel1 * ctx->d.nel2;
    k_mass<P><<<(nel + Cfg::EPG - 1) / Cfg::EPG, Cfg::NT, 0, s>>>(ctx->d, rho_t, values, lumped);
    ctx->launches++;
    KL_CUDA(cudaGetLastError());
    return 0;
}
int kl_launch_mass(kl_ctx* ctx, double rho_t, double* values, double* lumped, cudaStream_t s) {
    switch (ctx->d.p) {
        case 2: return launch_mass<2>(ctx, rho_t, values, lumped, s);
        case 3: return launch_mass<3>(ctx, rho_t, values, lumped, s);
        case 4: return launch_mass<4>(ctx, rho_t, values, lumped, s);
    }
    kl_set_error("unsupported degree");
    return KL_E_ARG;
}

// ------------------------------------------------------------------------------------------------
// Jacobian
template <int P>
struct JacCfg {
    static constexpr int NQ = P + 1;
    static constexpr int NQ2 = NQ * NQ;
    static constexpr int NLOC = (P + 1) * (P + 1);
    static constexpr int TILES = (P + 1) * (P + 1) * (P + 2) / 2;      // (i2, j) with i2 <= j2
#ifndef KL_JAC_EPG
#define KL_JAC_EPG 2
#endif
    static constexpr int EPG = (P == 2) ? 4 : (P == 3 ? KL_JAC_EPG : 2);  // elements per CTA
    static constexpr int NTILE = TILES * EPG;                           // threads that own a tile
    static constexpr int NTASK = EPG * NQ * NLOC;                       // phase-2 tasks per chunk
#ifndef KL_JAC_FULLWARPS
#define KL_JAC_FULLWARPS 1
#endif
    // threads: the tiles rounded up to whole warps (96 / 96 / 160).  The extra lanes cost no registers (allocation is per
    // warp) and take phase-2 tasks: 128 tasks run as 4 warp-rounds instead of 5 (measured: more, smaller CTAs beat one task per thread)
    static constexpr int NT = KL_JAC_FULLWARPS ? (NTILE + 31) / 32 * 32 : NTILE;
#ifndef KL_JAC_MINB
#define KL_JAC_MINB 4
#endif
    static constexpr int MINB = (P == 3) ? KL_JAC_MINB : 1;
    static constexpr int QCH = NQ;                                      // points per chunk: fixed q1, all q2
    static constexpr int ZS = 46;                                       // 45 coefficients [cd][p] + 1 pad: stride = 28 banks mod 32
};

template <int P>
struct JacShared {
    using Cfg = JacCfg<P>;
    BasisStage<P> stage[Cfg::EPG];
    double Z[Cfg::EPG][Cfg::QCH][Cfg::NLOC][Cfg::ZS];
    int4 cb[Cfg::EPG][Cfg::NLOC];     // colbase of the element's control points (scatter addressing)
    PointData pd[Cfg::EPG][Cfg::QCH]; // per-point records of the current chunk (TMA bulk copy)
    unsigned long long bar;           // mbarrier of the bulk copies
};

__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// ---- Z = T(point) . d : the 45 coefficients [cd][p] for one input vector d = (N1,N2,N11,N22,N12) at one quadrature point.
//      The F* flags say which inputs are present; absent ones are folded away at compile time (they are written as the
//      additive identity -0.0 so that no multiplication by zero is ever emitted).  All flags true = a basis function.
#define TZ(f, e) ((f) ? (e) : -0.0)
template <bool HASB, bool F1, bool F2, bool F11, bool F22, bool F12>
__device__ __forceinline__ void compute_Zc(const PointData& pd, double N1, double N2, double N11, double N22, double N12, double* Zo) {
    constexpr bool FG = F1 || F2;              // first derivatives present
    constexpr bool FS = F11 || F22 || F12;     // second derivatives present
    constexpr bool H0 = F11 || FG, H1 = F22 || FG, H2 = F12 || FG;
    constexpr bool FSIG = FG || HASB;
    double n[3], a1[3], a2[3], c1[3], c2[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) { n[c] = pd.n[c]; a1[c] = pd.a1[c]; a2[c] = pd.a2[c]; c1[c] = pd.c1[c]; c2[c] = pd.c2[c]; }
    double g[3], hh[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) g[c] = TZ(F1, N1 * c1[c]) + TZ(F2, N2 * c2[c]);
    const double G1[3] = {pd.G1[0], pd.G1[1], pd.G1[2]}, G2[3] = {pd.G2[0], pd.G2[1], pd.G2[2]};
    hh[0] = TZ(F11, N11) + TZ(F1, -(G1[0] * N1)) + TZ(F2, -(G2[0] * N2));
    hh[1] = TZ(F22, N22) + TZ(F1, -(G1[1] * N1)) + TZ(F2, -(G2[1] * N2));
    hh[2] = TZ(H2, 2.0 * (TZ(F12, N12) + TZ(F1, -(G1[2] * N1)) + TZ(F2, -(G2[2] * N2))));
    double AE1[3], AE2[3], BE1[3], BE2[3], Bh[3], Dh[3];
#pragma unroll
    for (int v = 0; v < 3; ++v) {
        AE1[v] = TZ(F1, N1 * pd.A[sidx(v, 0)]) + TZ(F2, N2 * pd.A[sidx(v, 2)]);
        AE2[v] = TZ(F2, N2 * pd.A[sidx(v, 1)]) + TZ(F1, N1 * pd.A[sidx(v, 2)]);
        BE1[v] = TZ(HASB && F1, N1 * pd.B[sidx(v, 0)]) + TZ(HASB && F2, N2 * pd.B[sidx(v, 2)]);
        BE2[v] = TZ(HASB && F2, N2 * pd.B[sidx(v, 1)]) + TZ(HASB && F1, N1 * pd.B[sidx(v, 2)]);
        Bh[v] = TZ(HASB && H0, pd.B[sidx(v, 0)] * hh[0]) + TZ(HASB && H1, pd.B[sidx(v, 1)] * hh[1]) + TZ(HASB && H2, pd.B[sidx(v, 2)] * hh[2]);
        Dh[v] = TZ(H0, pd.D[sidx(v, 0)] * hh[0]) + TZ(H1, pd.D[sidx(v, 1)] * hh[1]) + TZ(H2, pd.D[sidx(v, 2)] * hh[2]);
    }
    const double Mt0 = pd.Mt[0], Mt1 = pd.Mt[1], Mt2 = pd.Mt[2];
    const double Nhat = TZ(F11, Mt0 * N11) + TZ(F22, Mt1 * N22) + TZ(F12, Mt2 * N12);
    const double Ha1 = pd.Ha1, Ha2 = pd.Ha2, Hn = pd.Hn;
    const double eta = TZ(F1, Ha1 * N1) + TZ(F2, Ha2 * N2);
    const double p1 = TZ(F1, pd.N[0] * N1) + TZ(F2, pd.N[2] * N2), p2 = TZ(F2, pd.N[1] * N2) + TZ(F1, pd.N[2] * N1);
    const double ga1 = TZ(F1, pd.acon[0] * N1) + TZ(F2, pd.acon[2] * N2), ga2 = TZ(F1, pd.acon[2] * N1) + TZ(F2, pd.acon[1] * N2);
    const double q[3] = {pd.q[0], pd.q[1], pd.q[2]};
#pragma unroll
    for (int dd = 0; dd < 3; ++dd) {
        double sig[3], mu[3];
#pragma unroll
        for (int v = 0; v < 3; ++v) {
            sig[v] = TZ(FG, AE1[v] * a1[dd] + AE2[v] * a2[dd]) + TZ(HASB, -(n[dd] * Bh[v]));
            mu[v] = TZ(HASB && FG, BE1[v] * a1[dd] + BE2[v] * a2[dd]) + (-(n[dd] * Dh[v]));
        }
        const double s1 = G1[0] * mu[0] + G1[1] * mu[1] + 2.0 * G1[2] * mu[2] + TZ(FS, Nhat * c1[dd]) + TZ(FG, -(Ha1 * g[dd]))
                          + TZ(FG, Hn * n[dd] * ga1);
        const double s2 = G2[0] * mu[0] + G2[1] * mu[1] + 2.0 * G2[2] * mu[2] + TZ(FS, Nhat * c2[dd]) + TZ(FG, -(Ha2 * g[dd]))
                          + TZ(FG, Hn * n[dd] * ga2);
        const double en = TZ(FG, eta * n[dd]);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            double* z = Zo + (c * 3 + dd) * 5;
            double z0 = TZ(FSIG, a1[c] * sig[0] + a2[c] * sig[2]) + n[c] * s1 + TZ(FG, -(en * c1[c]));
            double z1 = TZ(FSIG, a2[c] * sig[1] + a1[c] * sig[2]) + n[c] * s2 + TZ(FG, -(en * c2[c]));
            if (c == dd) {
                z0 += TZ(FG, p1);
                z1 += TZ(FG, p2);
            } else {
                // epsilon_{c dd k} q_k
                const double eq = ((c + 1) % 3 == dd) ? q[(c + 2) % 3] : -q[(dd + 2) % 3];
                z0 += TZ(F2, -(N2 * eq));
                z1 += TZ(F1, N1 * eq);
            }
            const double ng = TZ(FG, n[dd] * g[c]);
            const double z2 = -(n[c] * mu[0]) + TZ(FG, Mt0 * ng);
            const double z3 = -(n[c] * mu[1]) + TZ(FG, Mt1 * ng);
            const double z4 = -(2.0 * n[c] * mu[2]) + TZ(FG, Mt2 * ng);
            z[0] = z0; z[1] = z1; z[2] = z2; z[3] = z3; z[4] = z4;
        }
    }
}
#undef TZ

// Z_j of basis function j at point (q1,q2)
template <int P, bool HASB>
__device__ __forceinline__ void compute_Z(const PointData& pd, const BasisStage<P>& E, int q1, int q2, int j, double* Zo) {
    const int ja = j % (P + 1), jb = j / (P + 1);
    const double x0 = E.b1[q1][0][ja], x1 = E.b1[q1][1][ja], x2 = E.b1[q1][2][ja];
    const double y0 = E.b2[q2][0][jb], y1 = E.b2[q2][1][jb], y2 = E.b2[q2][2][jb];
    compute_Zc<HASB, true, true, true, true, true>(pd, x1 * y0, x0 * y1, x2 * y0, x0 * y2, x1 * y1, Zo);
}

// ---- tile (ti2, tj) over one chunk (fixed q1): V_m^{cd} = sum_{q2} W_m^{cd}(q1,q2) first, then applied once with the
//      first-direction factors X(q1) (sum factorisation).  Zc = Z[element][q2][j][ZS]
template <int P>
__device__ __forceinline__ void tile_chunk(const BasisStage<P>& E, const double (*Zc)[JacCfg<P>::NLOC][JacCfg<P>::ZS], int ch, int ti2, int tj,
                                           double (&acc)[P + 1][9]) {
    constexpr int QCH = JacCfg<P>::QCH;
    double X0[P + 1], X1[P + 1], X2[P + 1];
#pragma unroll
    for (int a = 0; a <= P; ++a) { X0[a] = E.b1[ch][0][a]; X1[a] = E.b1[ch][1][a]; X2[a] = E.b1[ch][2][a]; }
    // (c,d) entries in groups of 4, 4, 1: 20 / 20 / 5(+pad) contiguous coefficients per point; few live registers so that
    // the loads of the next point can be issued ahead of the FMAs of the current one
#pragma unroll
    for (int g = 0; g < 3; ++g) {
        constexpr int NG[3] = {4, 4, 1};
        const int ng = NG[g];
        double V0[4], V1[4], V2[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) { V0[k] = 0.0; V1[k] = 0.0; V2[k] = 0.0; }
#pragma unroll
        for (int qc = 0; qc < QCH; ++qc) {
            const double y0 = E.b2[qc][0][ti2], y1 = E.b2[qc][1][ti2], y2 = E.b2[qc][2][ti2];
            const double2* Zi = reinterpret_cast<const double2*>(Zc[qc][tj]) + g * 10;
            double zz[20];
#pragma unroll
            for (int k = 0; k < (ng == 4 ? 10 : 3); ++k) { const double2 t = Zi[k]; zz[2 * k] = t.x; zz[2 * k + 1] = t.y; }
#pragma unroll
            for (int h = 0; h < ng; ++h) {
                const double z1 = zz[5 * h], z2 = zz[5 * h + 1], z11 = zz[5 * h + 2], z22 = zz[5 * h + 3], z12 = zz[5 * h + 4];
                V0[h] = fma(y2, z22, fma(y1, z2, V0[h]));    // multiplies N_{i1}(q1)
                V1[h] = fma(y1, z12, fma(y0, z1, V1[h]));    // multiplies N'_{i1}(q1)
                V2[h] = fma(y0, z11, V2[h]);                 // multiplies N''_{i1}(q1)
            }
        }
#pragma unroll
        for (int a = 0; a <= P; ++a)
#pragma unroll
            for (int h = 0; h < ng; ++h) acc[a][4 * g + h] = fma(X2[a], V2[h], fma(X1[a], V1[h], fma(X0[a], V0[h], acc[a][4 * g + h])));
    }
}

// set-up only (d.lift != nullptr): an entry whose row is free and whose column is eliminated goes, times the Dirichlet value of the
// column, into the lifting vector  K_L(free, eliminated) g  (gsExprAssembler eliminates the column into the rhs)
__device__ __forceinline__ void lift_entry(const KLDev& d, int cp_row, int c_row, int cp_col, int c_col, double v) {
    const int row = d.map[c_row * d.ncp + cp_row], col = d.map[c_col * d.ncp + cp_col];
    if (row < d.nfree && col >= d.nfree && d.fixed) atomicAdd(&d.lift[row], v * d.fixed[col - d.nfree]);
}

// ---- scatter of one tile (upper triangle i <= j plus the transposed entries).  Regular columns are addressed
//      arithmetically (outer[col] + c*nst + stencil slot); irregular ones (boundary, eliminated or matched DoFs in
//      the stencil) go through the position table.  cb = colbase of the element's control points.
template <int P>
__device__ __forceinline__ void tile_scatter(const KLDev& d, const int4* cb, int e1, int e2, int ti2, int tj, const double (&acc)[P + 1][9]) {
    const int i0 = d.span1[e1] - P, j0 = d.span2[e2] - P;
    const int ja = tj % (P + 1), jb = tj / (P + 1);
    const int J1 = i0 + ja, J2 = j0 + jb, Jc = J1 + d.n1 * J2;
    const int NST = d.nst, S3 = NST * 3, W = 2 * P + 1;
    double* __restrict__ val = d.values;
    const int4 cbJ = cb[tj];
    const int baseJ[3] = {cbJ.x, cbJ.y, cbJ.z};
#pragma unroll
    for (int a = 0; a <= P; ++a) {
        const int i = a + (P + 1) * ti2;
        if (i > tj) continue;
        const int I1 = i0 + a, I2 = j0 + ti2, Ic = I1 + d.n1 * I2;
        const int st_ij = (I1 - J1 + P) + W * (I2 - J2 + P);   // slot of row-function I in the stencil of column-function J
        const int st_ji = (J1 - I1 + P) + W * (J2 - I2 + P);
        const int4 cbI = cb[i];
        const int baseI[3] = {cbI.x, cbI.y, cbI.z};
        int p1[9], p2[9];
        if (cbJ.w) {
#pragma unroll
            for (int c = 0; c < 3; ++c)
#pragma unroll
                for (int dd = 0; dd < 3; ++dd) p1[c * 3 + dd] = baseJ[dd] + c * NST + st_ij;
        } else {
#pragma unroll
            for (int c = 0; c < 3; ++c)
#pragma unroll
                for (int dd = 0; dd < 3; ++dd) p1[c * 3 + dd] = __ldg(&d.pos[(size_t)(Jc * 3 + dd) * S3 + st_ij * 3 + c]);
        }
        if (i != tj) {
            if (cbI.w) {
#pragma unroll
                for (int c = 0; c < 3; ++c)
#pragma unroll
                    for (int dd = 0; dd < 3; ++dd) p2[c * 3 + dd] = baseI[c] + dd * NST + st_ji;
            } else {
#pragma unroll
                for (int c = 0; c < 3; ++c)
#pragma unroll
                    for (int dd = 0; dd < 3; ++dd) p2[c * 3 + dd] = __ldg(&d.pos[(size_t)(Ic * 3 + c) * S3 + st_ji * 3 + dd]);
            }
        } else {
#pragma unroll
            for (int k = 0; k < 9; ++k) p2[k] = -1;
        }
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            const double v = acc[a][k];
            if (p1[k] >= 0) atomicAdd(&val[p1[k]], v);   // entry (row (I,c), col (J,dd))
            if (p2[k] >= 0) atomicAdd(&val[p2[k]], v);   // entry (row (J,dd), col (I,c))
            if (d.lift) {
                lift_entry(d, Ic, k / 3, Jc, k % 3, v);
                if (i != tj) lift_entry(d, Jc, k % 3, Ic, k / 3, v);
            }
        }
    }
}

// phase ablation (profiling only): compiled in with -DKL_ABLATION, selected at run time by the env variable KL_ABLATE
#ifdef KL_ABLATION
#define KL_ABL(d) ((d).ablate)
#else
#define KL_ABL(d) 0
#endif
template <int P, bool HASB>
#ifdef KL_JAC_MAXREG
__global__ void __maxnreg__(KL_JAC_MAXREG) k_jacobian(KLDev d, int e2_begin, int e2_end) {
#else
__global__ void __launch_bounds__(JacCfg<P>::NT, JacCfg<P>::MINB) k_jacobian(KLDev d, int e2_begin, int e2_end) {
#endif
    using Cfg = JacCfg<P>;
    constexpr int NQ = Cfg::NQ, NQ2 = Cfg::NQ2, NLOC = Cfg::NLOC, TILES = Cfg::TILES, EPG = Cfg::EPG, NT = Cfg::NT, QCH = Cfg::QCH;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    JacShared<P>& S = *reinterpret_cast<JacShared<P>*>(smem_raw);
    const int tid = threadIdx.x;
    const int nel = d.nel1 * (e2_end - e2_begin);
    const int ebase = blockIdx.x * EPG;

    // per-point records of element le, chunk ch (= fixed q1) are QCH contiguous records: one TMA bulk copy each
    auto issue_pd = [&](int ch) {
        mbar_expect_tx(&S.bar, (unsigned)(EPG * QCH * sizeof(PointData)));
        for (int le = 0; le < EPG; ++le) {
            int e = ebase + le;
            if (e >= nel) e = nel - 1;
            const size_t ge = (size_t)(e % d.nel1) + (size_t)d.nel1 * (e2_begin + e / d.nel1);
            tma_bulk_g2s(&S.pd[le][0], d.pd + ge * NQ2 + (size_t)ch * QCH, (unsigned)(QCH * sizeof(PointData)), &S.bar);
        }
    };
    if (tid == 0) mbar_init(&S.bar, 1);
    __syncthreads();
    if (tid == 0) issue_pd(0);
    for (int le = 0; le < EPG; ++le) {
        int e = ebase + le;
        if (e >= nel) e = nel - 1;
        stage_basis<P>(d, e % d.nel1, e2_begin + e / d.nel1, S.stage[le], tid, NT);
    }
    for (int k = tid; k < EPG * NLOC; k += NT) {
        const int le = k / NLOC, l = k - le * NLOC;
        int e = ebase + le;
        if (e >= nel) e = nel - 1;
        const int cpi = (d.span1[e % d.nel1] - P + l % (P + 1)) + d.n1 * (d.span2[e2_begin + e / d.nel1] - P + l / (P + 1));
        S.cb[le][l] = reinterpret_cast<const int4*>(d.colbase)[cpi];
    }
    // tile of this thread: tiles enumerated by j ascending, i2 = 0..j2 (threads beyond NTILE only help in phase 2)
    const bool has_tile = tid < Cfg::NTILE;
    const int le_t = has_tile ? tid / TILES : 0, tt = has_tile ? tid - le_t * TILES : 0;
    int tj = 0, ti2 = 0;
    {
        int rem = tt;
        for (int j2 = 0; j2 <= P; ++j2) {
            const int cnt = (P + 1) * (j2 + 1);
            if (rem < cnt) { tj = (P + 1) * j2 + rem / (j2 + 1); ti2 = rem % (j2 + 1); break; }
            rem -= cnt;
        }
    }
    double acc[P + 1][9];
#pragma unroll
    for (int a = 0; a <= P; ++a)
#pragma unroll
        for (int k = 0; k < 9; ++k) acc[a][k] = 0.0;

    for (int ch = 0; ch < NQ2 / QCH; ++ch) {
        __syncthreads();   // basis staged / previous chunk's Z consumed
        mbar_wait(&S.bar, ch & 1);   // this chunk's per-point records have landed in shared memory
        // ---- phase 2: Z_j for the points of this chunk
        if (!(KL_ABL(d) & 4)) {
            constexpr int NTASKS = EPG * QCH * NLOC, NW = (NT + 31) / 32;
            static_assert(NTASKS <= 2 * NT, "at most two phase-2 tasks per thread");
            // one task per thread, the remaining NTASKS - NT go to the threads [shift, shift + NTASKS - NT): with
            // KL_JAC_FULLWARPS == 2 that window starts at a different warp every column
            const int shift = (KL_JAC_FULLWARPS == 2) ? 32 * (ch % NW) : 0;
#pragma unroll 1
            for (int r = 0; r < 2; ++r) {
                int k = tid;
                if (r) { const int x = tid - shift + (tid < shift ? NT : 0); k = NT + x; }
                if (k >= NTASKS) continue;
                const int j = k % NLOC;
                const int qc = (k / NLOC) % QCH;
                const int le = k / (NLOC * QCH);
                compute_Z<P, HASB>(S.pd[le][qc], S.stage[le], ch, qc, j, S.Z[le][qc][j]);
            }
        }
        __syncthreads();
        if (tid == 0 && ch + 1 < NQ2 / QCH) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic reads of S.pd are done; async proxy may overwrite
            issue_pd(ch + 1);
        }
        // ---- phase 3: tile (ti2, tj).  The first-direction factors X(q1) are constant over the chunk, so
        //      V_m^{cd} = sum_{q2} W_m^{cd}(q1,q2) is formed first and applied once per chunk (sum factorisation).
        if (has_tile && !(KL_ABL(d) & 2)) tile_chunk<P>(S.stage[le_t], S.Z[le_t], ch, ti2, tj, acc);
    }
    const int e = ebase + le_t;
    if (has_tile && e < nel && !(KL_ABL(d) & 1)) tile_scatter<P>(d, S.cb[le_t], e % d.nel1, e2_begin + e / d.nel1, ti2, tj, acc);
    if (KL_ABL(d) & 1) { double sink = 0; for (int a = 0; a <= P; ++a) for (int k = 0; k < 9; ++k) sink += acc[a][k]; if (sink == 1.2345e-300) d.values[0] = sink; }
}



// ------------------------------------------------------------------------------------------------
// Jacobian, register-resident sliding-window kernel (P = 3).
//
// One CTA = 64 threads walks a segment of consecutive elements of ONE element row e2 along direction 1.  Thread
// (column function j = (b, b2), lane q2) evaluates Z_j = T . d_j at the four points (q1, q2), q1 = 0..3, itself — Z never
// touches shared memory — keeps only the component pairs c <= d (K^{cd}_{ij} = K^{dc}_{ji}), contracts with the second-direction
// factors of all four row positions i2, and a 4-lane shuffle reduce-scatter over q2 leaves lane i2 with V_m(i2) of its
// column function.  The first-direction factors are applied at once into 24 accumulators acc[a][cd] = K[(a, i2), j][cd].
// Walking along direction 1 the accumulators stay in registers while an (I, J) node pair is still inside the support of
// later elements of the row (the local index a shifts down by one per element; the column function of a thread advances
// by p+1 when it leaves the support), so every pair is written once per element ROW: 1008 RED per element instead of 2304,
// with compile-time shuffles instead of the 846 LDS.128 wavefronts per column of the shared-memory kernel.
template <bool HASB>
__device__ __forceinline__ void sw_point(const PointData& pd, double N1, double N2, double N11, double N22, double N12,
                                         const double (&yk)[4][3], const double (&xa)[3][4], double (&acc)[4][6]) {
    double n[3], a1[3], a2[3], c1[3], c2[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) { n[c] = pd.n[c]; a1[c] = pd.a1[c]; a2[c] = pd.a2[c]; c1[c] = pd.c1[c]; c2[c] = pd.c2[c]; }
    double g[3], hh[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) g[c] = N1 * c1[c] + N2 * c2[c];
    const double G1[3] = {pd.G1[0], pd.G1[1], pd.G1[2]}, G2[3] = {pd.G2[0], pd.G2[1], pd.G2[2]};
    hh[0] = N11 - G1[0] * N1 - G2[0] * N2;
    hh[1] = N22 - G1[1] * N1 - G2[1] * N2;
    hh[2] = 2.0 * (N12 - G1[2] * N1 - G2[2] * N2);
    double AE1[3], AE2[3], BE1[3], BE2[3], Bh[3], Dh[3];
#pragma unroll
    for (int v = 0; v < 3; ++v) {
        AE1[v] = N1 * pd.A[sidx(v, 0)] + N2 * pd.A[sidx(v, 2)];
        AE2[v] = N2 * pd.A[sidx(v, 1)] + N1 * pd.A[sidx(v, 2)];
        if (HASB) {
            BE1[v] = N1 * pd.B[sidx(v, 0)] + N2 * pd.B[sidx(v, 2)];
            BE2[v] = N2 * pd.B[sidx(v, 1)] + N1 * pd.B[sidx(v, 2)];
            Bh[v] = pd.B[sidx(v, 0)] * hh[0] + pd.B[sidx(v, 1)] * hh[1] + pd.B[sidx(v, 2)] * hh[2];
        }
        Dh[v] = pd.D[sidx(v, 0)] * hh[0] + pd.D[sidx(v, 1)] * hh[1] + pd.D[sidx(v, 2)] * hh[2];
    }
    const double Mt0 = pd.Mt[0], Mt1 = pd.Mt[1], Mt2 = pd.Mt[2];
    const double Nhat = Mt0 * N11 + Mt1 * N22 + Mt2 * N12;
    const double Ha1 = pd.Ha1, Ha2 = pd.Ha2, Hn = pd.Hn;
    const double eta = Ha1 * N1 + Ha2 * N2;
    const double p1 = pd.N[0] * N1 + pd.N[2] * N2, p2 = pd.N[1] * N2 + pd.N[2] * N1;
    const double hg1 = Hn * (pd.acon[0] * N1 + pd.acon[2] * N2), hg2 = Hn * (pd.acon[2] * N1 + pd.acon[1] * N2);
    const double q[3] = {pd.q[0], pd.q[1], pd.q[2]};
#pragma unroll
    for (int dd = 0; dd < 3; ++dd) {
        double sig[3], mu[3];
#pragma unroll
        for (int v = 0; v < 3; ++v) {
            sig[v] = AE1[v] * a1[dd] + AE2[v] * a2[dd];
            if (HASB) sig[v] -= n[dd] * Bh[v];
            mu[v] = -(n[dd] * Dh[v]);
            if (HASB) mu[v] += BE1[v] * a1[dd] + BE2[v] * a2[dd];
        }
        const double mu2 = mu[2] + mu[2];
        const double s1 = fma(G1[0], mu[0], fma(G1[1], mu[1], fma(G1[2], mu2, fma(Nhat, c1[dd], fma(hg1, n[dd], -(Ha1 * g[dd]))))));
        const double s2 = fma(G2[0], mu[0], fma(G2[1], mu[1], fma(G2[2], mu2, fma(Nhat, c2[dd], fma(hg2, n[dd], -(Ha2 * g[dd]))))));
        const double en = eta * n[dd];
#pragma unroll
        for (int c = 0; c <= dd; ++c) {
            double z0 = a1[c] * sig[0] + a2[c] * sig[2] + n[c] * s1 - en * c1[c];
            double z1 = a2[c] * sig[1] + a1[c] * sig[2] + n[c] * s2 - en * c2[c];
            if (c == dd) {
                z0 += p1;
                z1 += p2;
            } else {
                const double eq = ((c + 1) % 3 == dd) ? q[(c + 2) % 3] : -q[(dd + 2) % 3];   // epsilon_{c dd k} q_k
                z0 -= N2 * eq;
                z1 += N1 * eq;
            }
            const double ng = n[dd] * g[c];
            const double z2 = Mt0 * ng - n[c] * mu[0];     // multiplies N_i,11
            const double z3 = Mt1 * ng - n[c] * mu[1];     // N_i,22
            const double z4 = Mt2 * ng - n[c] * mu2;           // N_i,12
            // second-direction contraction for the four row positions (slot s = i2 ^ lane) and reduce-scatter over the q2 lanes:
            //   w0 multiplies N_{i1}(q1), w1 N'_{i1}(q1), w2 N''_{i1}(q1)
            double r0, r1, r2, t0, t1, t2;
            r0 = fma(yk[3][2], z3, yk[3][1] * z1); r1 = fma(yk[3][1], z4, yk[3][0] * z0); r2 = yk[3][0] * z2;
            r0 = __shfl_xor_sync(0xffffffffu, r0, 16); r1 = __shfl_xor_sync(0xffffffffu, r1, 16); r2 = __shfl_xor_sync(0xffffffffu, r2, 16);
            r0 = fma(yk[1][2], z3, fma(yk[1][1], z1, r0)); r1 = fma(yk[1][1], z4, fma(yk[1][0], z0, r1)); r2 = fma(yk[1][0], z2, r2);
            r0 = __shfl_xor_sync(0xffffffffu, r0, 8); r1 = __shfl_xor_sync(0xffffffffu, r1, 8); r2 = __shfl_xor_sync(0xffffffffu, r2, 8);
            t0 = fma(yk[2][2], z3, yk[2][1] * z1); t1 = fma(yk[2][1], z4, yk[2][0] * z0); t2 = yk[2][0] * z2;
            t0 = __shfl_xor_sync(0xffffffffu, t0, 16); t1 = __shfl_xor_sync(0xffffffffu, t1, 16); t2 = __shfl_xor_sync(0xffffffffu, t2, 16);
            const double V0 = fma(yk[0][2], z3, fma(yk[0][1], z1, t0)) + r0;
            const double V1 = fma(yk[0][1], z4, fma(yk[0][0], z0, t1)) + r1;
            const double V2 = fma(yk[0][0], z2, t2) + r2;
            const int k = dd * (dd + 1) / 2 + c;
#pragma unroll
            for (int a = 0; a < 4; ++a) acc[a][k] = fma(xa[2][a], V2, fma(xa[1][a], V1, fma(xa[0][a], V0, acc[a][k])));
        }
    }
}

#ifndef KL_SW_MINB
#define KL_SW_MINB 6
#endif
// ---- flush through the TMA engine (EXPERIMENT, off by default: -DKL_SW_BULK=1).  A regular matrix column (J,dd) holds, per row
// component c and row i2 of row functions, the seven row functions I1 = J1-3 .. J1+3 as seven consecutive doubles, and one thread
// produces all seven during the four elements its column function stays in the window.  The thread stages them in shared memory (NK
// runs of 8 doubles) and, when the column function leaves, adds the 16-byte aligned 48 bytes of every run with ONE
// cp.reduce.async.bulk (.add.f64); the odd element and the transposed entries stay RED.F64: 432 RED + 96 bulk operations per element
// instead of 1008 RED.  Measured (profiles/r3_ablation.txt): the flush alone is twice as fast (tools/micro/red_bulk.cu, 1.87 -> 0.94 ms
// at 576 x 576) and the LSU stalls of the kernel vanish (mio_throttle 0.17 -> 0.01, short scoreboard 0.64 -> 0.26 per issue), but
// UBLKRED takes uniform registers, so every bulk operation is issued lane by lane: +27 % warp instructions (1.87 G -> 2.38 G) in a
// kernel that is bound by its issue rate (47 % issue slots, `wait` stalls): 3.51 -> 4.46 ms.  Parity is green in both modes.
#ifndef KL_SW_BULK
#define KL_SW_BULK 0
#endif
struct SwRun {
    double* s;          // this thread's NK runs, 8 doubles each, 16-byte aligned
    unsigned mask;      // run positions (st1 = 0..6) written for the current column function
    bool pending;       // a bulk group of this thread may still be reading s
};
__device__ __forceinline__ void bulk_red_add_f64(double* gdst, const double* ssrc, unsigned bytes) {
    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
}
// component pair k -> (c, dd): k = dd (dd + 1) / 2 + c (c <= dd) for the symmetric part, k = 3 c + dd for the pressure tangent
template <int NK> __device__ __forceinline__ int sw_pair_c(int k) { return NK == 9 ? k / 3 : (k < 1 ? 0 : (k < 3 ? k - 1 : k - 3)); }
template <int NK> __device__ __forceinline__ int sw_pair_d(int k) { return NK == 9 ? k % 3 : (k < 1 ? 0 : (k < 3 ? 1 : 2)); }
// direct entries of one slot: position pos of the NK runs (rowoff = 7 (i2 - b2 + 3): start of the run inside the 49-block)
template <int NK>
__device__ __forceinline__ void sw_stage_slot(double* __restrict__ val, SwRun& rs, const double (&v)[NK], const int4 cbJ, int rowoff, int pos) {
    if (rs.pending) { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); rs.pending = false; }
    const int col[3] = {cbJ.x, cbJ.y, cbJ.z};
#pragma unroll
    for (int k = 0; k < NK; ++k) {
        const int base = col[sw_pair_d<NK>(k)] + rowoff + 49 * sw_pair_c<NK>(k), par = base & 1;
        if (pos == (par ? 0 : 6)) atomicAdd(val + (base + pos), v[k]);
        else rs.s[k * 8 + pos + par] = v[k];
    }
    rs.mask |= 1u << pos;
}
// the column function leaves the window: zero what was never written (segment ends, irregular neighbours), hand the runs to the TMA engine
template <int NK>
__device__ __forceinline__ void sw_drain(double* __restrict__ val, SwRun& rs, const int4 cbJ, int rowoff) {
    if (!rs.mask) return;
    const int col[3] = {cbJ.x, cbJ.y, cbJ.z};
    if (rs.mask != 0x7fu) {
        for (int p = 0; p < 7; ++p)
            if (!((rs.mask >> p) & 1u)) {
#pragma unroll
                for (int k = 0; k < NK; ++k) {
                    const int par = (col[sw_pair_d<NK>(k)] + rowoff + 49 * sw_pair_c<NK>(k)) & 1;
                    rs.s[k * 8 + p + par] = 0.0;
                }
            }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#pragma unroll
    for (int k = 0; k < NK; ++k) {
        const int base = col[sw_pair_d<NK>(k)] + rowoff + 49 * sw_pair_c<NK>(k), par = base & 1;
        bulk_red_add_f64(val + (base + par), rs.s + (k * 8 + 2 * par), 48u);
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    rs.pending = true;
    rs.mask = 0;
}

// flush one window slot: the 3x3 component block (c <= d computed; c < d also written transposed) of node pair (I, J).
// cbJ / cbI = colbase of the two control points; regular columns are addressed arithmetically.
__device__ __forceinline__ void sw_flush_slot(const KLDev& d, double* __restrict__ val, const double (&v)[6], const int4 cbJ, const int4 cbI,
                                              int Ic, int Jc, int st_ij, SwRun& rs, int rowoff, int pos) {
    constexpr int NST = 49, S3 = 147;
    const int st_ji = NST - 1 - st_ij;
    if (KL_SW_BULK && (cbJ.w & cbI.w)) {
        sw_stage_slot<6>(val, rs, v, cbJ, rowoff, pos);
        double* pi0 = val + (cbI.x + st_ji);
        double* pi1 = val + (cbI.y + st_ji);
        atomicAdd(pi0 + NST, v[1]);                                        // transposes of (0,1), (0,2), (1,2): row (J,dd), col (I,c)
        atomicAdd(pi0 + 2 * NST, v[3]);
        atomicAdd(pi1 + 2 * NST, v[4]);
    } else if (cbJ.w & cbI.w) {
        double* pj0 = val + (cbJ.x + st_ij);
        double* pj1 = val + (cbJ.y + st_ij);
        double* pj2 = val + (cbJ.z + st_ij);
        double* pi0 = val + (cbI.x + st_ji);
        double* pi1 = val + (cbI.y + st_ji);
        atomicAdd(pj0, v[0]);                                              // (c,d) = (0,0)
        atomicAdd(pj1, v[1]); atomicAdd(pi0 + NST, v[1]);                  // (0,1) and its transpose: row (J,1), col (I,0)
        atomicAdd(pj1 + NST, v[2]);                                        // (1,1)
        atomicAdd(pj2, v[3]); atomicAdd(pi0 + 2 * NST, v[3]);              // (0,2)
        atomicAdd(pj2 + NST, v[4]); atomicAdd(pi1 + 2 * NST, v[4]);        // (1,2)
        atomicAdd(pj2 + 2 * NST, v[5]);                                    // (2,2)
    } else {
#pragma unroll
        for (int dd = 0; dd < 3; ++dd)
#pragma unroll
            for (int c = 0; c <= dd; ++c) {
                const double x = v[dd * (dd + 1) / 2 + c];
                const int p1 = __ldg(&d.pos[(size_t)(Jc * 3 + dd) * S3 + st_ij * 3 + c]);
                if (p1 >= 0) atomicAdd(&val[p1], x);            // entry (row (I,c), col (J,dd))
                if (d.lift) lift_entry(d, Ic, c, Jc, dd, x);
                if (c < dd) {
                    const int p2 = __ldg(&d.pos[(size_t)(Ic * 3 + c) * S3 + st_ji * 3 + dd]);
                    if (p2 >= 0) atomicAdd(&val[p2], x);        // entry (row (J,dd), col (I,c))
                    if (d.lift) lift_entry(d, Jc, dd, Ic, c, x);
                }
            }
    }
}

// follower-pressure tangent in the same window (PRES instantiation of k_jacobian_sw): K^{c,dd}_{ij} += p wJ R_i n_dd g_j[c],
// g_j = N_j,1 a^1 + N_j,2 a^2.  Unsymmetric, so all nine component pairs are kept (k = 3 c + dd) and nothing is written transposed.
// The integrand is cheap (22 FP64 instructions per point), so lane (j, i2) walks the four q2 points of the q1 column itself instead
// of reduce-scattering over the q2 lanes: the first version with 27 shuffles per point was bound by the shuffle / LSU queue
// (mio_throttle 2.1 + short scoreboard 1.8 stall cycles per issue, 1.96 ms; profiles/r3_pressure_summary.txt).  The four lanes that
// share a column function read the same record at the same time (one broadcast wavefront).
__device__ __forceinline__ void sw_point_pressure(const PointData* pd4, double pressure, double x0, double x1, const double (&yjv)[4],
                                                  const double (&yjd)[4], const double (&yi)[4], const double (&xa)[3][4], double (&acc)[4][9]) {
    double V[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) V[k] = 0.0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const PointData& pd = pd4[q];
        const double N1 = x1 * yjv[q], N2 = x0 * yjd[q];
        const double pw = (pd.wJ * pressure) * yi[q];
        const double pn[3] = {pw * pd.n[0], pw * pd.n[1], pw * pd.n[2]};
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const double g = fma(N1, pd.c1[c], N2 * pd.c2[c]);
#pragma unroll
            for (int dd = 0; dd < 3; ++dd) V[3 * c + dd] = fma(pn[dd], g, V[3 * c + dd]);
        }
    }
#pragma unroll
    for (int k = 0; k < 9; ++k)
#pragma unroll
        for (int a = 0; a < 4; ++a) acc[a][k] = fma(xa[0][a], V[k], acc[a][k]);
}
__device__ __forceinline__ void sw_flush_slot(const KLDev& d, double* __restrict__ val, const double (&v)[9], const int4 cbJ, const int4 /*cbI*/,
                                              int /*Ic*/, int Jc, int st_ij, SwRun& rs, int rowoff, int pos) {
    constexpr int NST = 49, S3 = 147;
    if (KL_SW_BULK && cbJ.w) {
        sw_stage_slot<9>(val, rs, v, cbJ, rowoff, pos);
    } else if (cbJ.w) {
        const int col[3] = {cbJ.x, cbJ.y, cbJ.z};
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int dd = 0; dd < 3; ++dd) atomicAdd(val + (col[dd] + st_ij + c * NST), v[3 * c + dd]);   // row (I,c), column (J,dd)
    } else {
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int dd = 0; dd < 3; ++dd) {
                const int p1 = __ldg(&d.pos[(size_t)(Jc * 3 + dd) * S3 + st_ij * 3 + c]);
                if (p1 >= 0) atomicAdd(&val[p1], v[3 * c + dd]);
            }
    }
}

template <bool HASB, bool PRES = false>
__global__ void __launch_bounds__(64, KL_SW_MINB) k_jacobian_sw(KLDev d, int e2_begin, int e2_end, int seg_len) {
    constexpr int P = 3, NQ2 = 16, NB = 48, W = 7;
    __shared__ __align__(128) PointData s_pd[2][NQ2];
    __shared__ __align__(16) double s_b1[2][NB];        // [q1][m][a] of the element
    __shared__ __align__(16) int4 s_cb[2][4][4];        // colbase of the element's control points [row i2][a]
    __shared__ unsigned long long s_bar[2];
    constexpr int NK = PRES ? 9 : 6;     // component pairs kept per node pair
    constexpr int RS = NK * 8 + 2;       // doubles per thread in the run staging area (+2: threads 2-way instead of 16-way bank-aliased)
    extern __shared__ __align__(16) double s_run[];     // [64][RS] when KL_SW_BULK (dynamic: 54 KB in all for the pressure instantiation)
    const int tid = threadIdx.x;
    const int nrows = e2_end - e2_begin;
    const int row = blockIdx.x % nrows, seg = blockIdx.x / nrows;     // consecutive CTAs take consecutive element rows of one segment column
    const int e2 = e2_begin + row;
    const int e1_begin = seg * seg_len, e1_end = min(d.nel1, e1_begin + seg_len);
    // lane map: the 8 lanes of a quarter-warp share one quadrature point (q2 = lane >> 3): a 128-bit shared-memory load whose address is
    // uniform per quarter-warp costs 2 cycles instead of 4 (tools/micro/lds_shfl.cu); column function j = (lane & 7) + 8 * warp
    const int q2 = (tid >> 3) & 3, jj = (tid & 7) + 8 * (tid >> 5), jcls = jj & 3, b2 = jj >> 2, i2 = q2;
    const int j0 = __ldg(&d.span2[e2]) - P;

    auto issue = [&](int e1, int s) {
        const int i0e = __ldg(&d.span1[e1]) - P;
        mbar_expect_tx(&s_bar[s], (unsigned)(NQ2 * sizeof(PointData) + NB * sizeof(double) + 16 * sizeof(int4)));
        tma_bulk_g2s(&s_pd[s][0], d.pd + ((size_t)e1 + (size_t)d.nel1 * e2) * NQ2, (unsigned)(NQ2 * sizeof(PointData)), &s_bar[s]);
        tma_bulk_g2s(&s_b1[s][0], d.bas1 + (size_t)e1 * NB, (unsigned)(NB * sizeof(double)), &s_bar[s]);
#pragma unroll
        for (int r = 0; r < 4; ++r)
            tma_bulk_g2s(&s_cb[s][r][0], reinterpret_cast<const int4*>(d.colbase) + (i0e + d.n1 * (j0 + r)), (unsigned)(4 * sizeof(int4)), &s_bar[s]);
    };
    if (tid == 0) { mbar_init(&s_bar[0], 1); mbar_init(&s_bar[1], 1); }
    __syncthreads();
    if (tid == 0) {
        issue(e1_begin, 0);
        if (e1_begin + 1 < e1_end) issue(e1_begin + 1, 1);
    }
    // second-direction factors of this thread: of its column function (b2) and of the four row positions, ordered by
    // reduce-scatter slot (slot s <-> i2 = lane ^ s); constant along the walk
    const double* g2 = d.bas2 + (size_t)e2 * NB;
    double yj[3], yk[4][3];
#pragma unroll
    for (int m = 0; m < 3; ++m) {
        yj[m] = __ldg(&g2[(q2 * 3 + m) * 4 + b2]);
#pragma unroll
        for (int s = 0; s < 4; ++s) yk[s][m] = __ldg(&g2[(q2 * 3 + m) * 4 + (q2 ^ s)]);
    }
    // pressure instantiation: values / first derivatives of the column function and values of the row function i2 at all four q2
    double yjv[4], yjd[4], yi[4];
    if constexpr (PRES) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            yjv[q] = __ldg(&g2[(q * 3 + 0) * 4 + b2]);
            yjd[q] = __ldg(&g2[(q * 3 + 1) * 4 + b2]);
            yi[q] = __ldg(&g2[(q * 3 + 0) * 4 + i2]);
        }
    }
    double acc[4][NK];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int k = 0; k < NK; ++k) acc[a][k] = 0.0;
    const double pressure = d.mat.pressure;
    int i0 = __ldg(&d.span1[e1_begin]) - P;
    int b = (jcls - i0) & 3;          // local index of this thread's column function in the current element
    bool live = false;                // the accumulators hold contributions of the current column function
    double* __restrict__ val = d.values;
    const int I2 = j0 + i2, J2 = j0 + b2;
    SwRun rs{s_run + (KL_SW_BULK ? tid * RS : 0), 0u, false};
    const int rowoff = W * (i2 - b2 + P);

    for (int e1 = e1_begin; e1 < e1_end; ++e1) {
        const int le = e1 - e1_begin, s = le & 1;
        const int i0n = (e1 + 1 < e1_end) ? __ldg(&d.span1[e1 + 1]) - P : i0 + P + 1;
        mbar_wait(&s_bar[s], (le >> 1) & 1);
#ifndef KL_SW_UNROLL
#define KL_SW_UNROLL 1
#endif
        constexpr int UNR = KL_SW_UNROLL;
#pragma unroll UNR
        for (int q1 = 0; q1 < 4; ++q1) {
            const double* xb = &s_b1[s][q1 * 12];
            double xa[3][4];
#pragma unroll
            for (int m = 0; m < 3; ++m) {
                const double2 u = *reinterpret_cast<const double2*>(xb + 4 * m), v = *reinterpret_cast<const double2*>(xb + 4 * m + 2);
                xa[m][0] = u.x; xa[m][1] = u.y; xa[m][2] = v.x; xa[m][3] = v.y;
            }
            const double x0 = xb[b], x1 = xb[4 + b], x2 = xb[8 + b];
            if constexpr (PRES) sw_point_pressure(&s_pd[s][q1 * 4], pressure, x0, x1, yjv, yjd, yi, xa, acc);
            else sw_point<HASB>(s_pd[s][q1 * 4 + q2], x1 * yj[0], x0 * yj[1], x2 * yj[0], x0 * yj[2], x1 * yj[1], yk, xa, acc);
        }
        live = true;
        // scatter addressing of this element: column function J and the four row functions of row i2
        int4 cbJ = s_cb[s][b2][b];
        int4 cbI[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) cbI[a] = s_cb[s][i2][a];
        __syncthreads();               // every thread is done with buffer s
        if (tid == 0 && e1 + 2 < e1_end) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            issue(e1 + 2, s);
        }
        // ---- window step: row functions I1 < i0n and column functions J1 < i0n have received their last contribution of this row
        for (int st = i0; st < i0n; ++st) {
#ifdef KL_SW_NOFLUSH        // timing ablation only (profiles/r3_ablation.txt): the flush never executes but the accumulators stay live
            if (live && acc[0][0] == 1.2345e-300) {
#else
            if (live) {
#endif
                const int Jc = (st + b) + d.n1 * J2, Ic0 = st + d.n1 * I2;
                const int st0 = (P - b) + W * (i2 - b2 + P);
                sw_flush_slot(d, val, acc[0], cbJ, cbI[0], Ic0, Jc, st0, rs, rowoff, P - b);
                if (b == 0) {        // the column function leaves the support: all its pairs are complete
#pragma unroll
                    for (int a = 1; a < 4; ++a)
                        if (st + a <= i0 + P) sw_flush_slot(d, val, acc[a], cbJ, cbI[a], Ic0 + a, Jc, st0 + a, rs, rowoff, P + a);
                    if (KL_SW_BULK) sw_drain<NK>(val, rs, cbJ, rowoff);
                }
            }
            // shift the window by one function
            const bool wrap = (b == 0);
#pragma unroll
            for (int k = 0; k < NK; ++k) {
                acc[0][k] = wrap ? 0.0 : acc[1][k];
                acc[1][k] = wrap ? 0.0 : acc[2][k];
                acc[2][k] = wrap ? 0.0 : acc[3][k];
                acc[3][k] = 0.0;
            }
            cbI[0] = cbI[1]; cbI[1] = cbI[2]; cbI[2] = cbI[3];
            if (wrap) live = false;
            b = (b - 1) & 3;
        }
        i0 = i0n;
    }
    if (KL_SW_BULK) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");    // shared memory must outlive the bulk reads
}

// ------------------------------------------------------------------------------------------------
// follower-pressure tangent  -p R_i dn_jd[c] = +p R_i n_d g_j[c]  (unsymmetric, full i x j loop; cheap)
template <int P>
__global__ void __launch_bounds__(256) k_pressure_tangent(KLDev d, int e2_begin, int e2_end) {
    constexpr int NQ = P + 1, NQ2 = NQ * NQ, NLOC = (P + 1) * (P + 1);
    __shared__ BasisStage<P> stage;
    __shared__ double pn[NQ2][3], pc1[NQ2][3], pc2[NQ2][3], pw[NQ2];
    const int tid = threadIdx.x;
    const int e = blockIdx.x;
    const int e1 = e % d.nel1, e2 = e2_begin + e / d.nel1;
    stage_basis<P>(d, e1, e2, stage, tid, blockDim.x);
    if (tid < NQ2) {
        const PointData& pd = d.pd[(size_t)(e1 + d.nel1 * e2) * NQ2 + tid];
        const int lq = (tid / NQ) + NQ * (tid % NQ);
        for (int c = 0; c < 3; ++c) { pn[lq][c] = pd.n[c]; pc1[lq][c] = pd.c1[c]; pc2[lq][c] = pd.c2[c]; }
        pw[lq] = pd.wJ * d.mat.pressure;
    }
    __syncthreads();
    const int i0 = d.span1[e1] - P, j0 = d.span2[e2] - P;
    const int S3 = d.nst * 3, W = 2 * P + 1;
    for (int pr = tid; pr < NLOC * NLOC; pr += blockDim.x) {
        const int i = pr / NLOC, j = pr - i * NLOC;
        const int ia = i % (P + 1), ib = i / (P + 1), ja = j % (P + 1), jb = j / (P + 1);
        double k[3][3] = {{0}};
        for (int lq = 0; lq < NQ2; ++lq) {
            const int q1 = lq % NQ, q2 = lq / NQ;
            const double Ri = stage.b1[q1][0][ia] * stage.b2[q2][0][ib];
            const double N1 = stage.b1[q1][1][ja] * stage.b2[q2][0][jb], N2 = stage.b1[q1][0][ja] * stage.b2[q2][1][jb];
            for (int c = 0; c < 3; ++c) {
                const double gc = N1 * pc1[lq][c] + N2 * pc2[lq][c];
                for (int dd = 0; dd < 3; ++dd) k[c][dd] += pw[lq] * Ri * pn[lq][dd] * gc;
            }
        }
        const int I1 = i0 + ia, I2 = j0 + ib, J1 = i0 + ja, J2 = j0 + jb;
        const int Jc = J1 + d.n1 * J2;
        const int st_ij = (I1 - J1 + P) + W * (I2 - J2 + P);
        for (int c = 0; c < 3; ++c)
            for (int dd = 0; dd < 3; ++dd) {
                const int p1 = d.pos[(size_t)(Jc * 3 + dd) * S3 + st_ij * 3 + c];
                if (p1 >= 0) atomicAdd(&d.values[p1], k[c][dd]);
            }
    }
}

// ------------------------------------------------------------------------------------------------
int kl_launch_construct(kl_ctx* ctx, const double* x_dev, cudaStream_t s, const int* skip) {
    const int i_begin = ctx->cp_row_begin * ctx->d.n1, i_count = (ctx->cp_row_end - ctx->cp_row_begin) * ctx->d.n1;
    const int n = 3 * i_count;
    if (n <= 0) return 0;
    k_construct_solution<<<(n + 255) / 256, 256, 0, s>>>(ctx->d, x_dev, skip, i_begin, i_count);
    ctx->launches++;
    KL_CUDA(cudaGetLastError());
    return 0;
}

int kl_launch_state_compare(kl_ctx* ctx, const double* x_dev, cudaStream_t s) {
    const int n = ctx->d.nfree;
    const bool comparable = ctx->pd_valid && ((x_dev == nullptr) == (ctx->state_null != 0));
    ctx->state_null = x_dev == nullptr;
    KL_CUDA(cudaMemsetAsync(ctx->d_same, comparable ? 1 : 0, sizeof(int), s));
    if (comparable) {
        k_state_compare<<<std::max(1, std::min((n + 255) / 256, 4 * ctx->n_sm)), 256, 0, s>>>(x_dev, ctx->d_xstate, n, ctx->d_same);
        ctx->launches++;
    }
    if (n > 0 && x_dev != ctx->d_xstate) {
        k_state_store<<<(n + 255) / 256, 256, 0, s>>>(x_dev, ctx->d_xstate, n);
        ctx->launches++;
    }
    KL_CUDA(cudaGetLastError());
    return 0;
}

int kl_launch_axpby(kl_ctx* ctx, double* r, const double* fext, double a_r, double b_f, int n, cudaStream_t s) {
    if (n <= 0) return 0;
    k_axpby<<<(n + 255) / 256, 256, 0, s>>>(r, fext, a_r, b_f, n);
    ctx->launches++;
    KL_CUDA(cudaGetLastError());
    return 0;
}

template <int P>
static int launch_points(kl_ctx* ctx, int e2b, int e2e, double* r, cudaStream_t s, const int* skip) {
    using Cfg = PointCfg<P>;
    const int nel = ctx->d.nel1 * (e2e - e2b);
    if (nel <= 0) return 0;
    const size_t smem = ((sizeof(ElemStage<P>) * Cfg::EPG + 15) / 16) * 16 + sizeof(PointData) * Cfg::EPG * Cfg::NQ2;
    if (!(ctx->attr_done & 1u)) {      // function attributes are per device: tracked in the context, not in a static
        KL_CUDA(cudaFuncSetAttribute(k_points<P, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        KL_CUDA(cudaFuncSetAttribute(k_points<P, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ctx->attr_done |= 1u;
    }
    if (r) k_points<P, true><<<(nel + Cfg::EPG - 1) / Cfg::EPG, Cfg::NT, smem, s>>>(ctx->d, e2b, e2e, r, skip);
    else k_points<P, false><<<(nel + Cfg::EPG - 1) / Cfg::EPG, Cfg::NT, smem, s>>>(ctx->d, e2b, e2e, nullptr, skip);
    ctx->launches++;
    KL_CUDA(cudaGetLastError());
    return 0;
}
int kl_launch_points(kl_ctx* ctx, int e2_begin, int e2_end, cudaStream_t s, double* r_dev, const int* skip) {
    switch (ctx->d.p) {
        case 2: return launch_points<2>(ctx, e2_begin, e2_end, r_dev, s, skip);
        case 3: return launch_points<3>(ctx, e2_begin, e2_end, r_dev, s, skip);
        case 4: return launch_points<4>(ctx, e2_begin, e2_end, r_dev, s, skip);
    }
    kl_set_error("unsupported degree");
    return KL_E_ARG;
}

static size_t sw_run_bytes(int nk) { return KL_SW_BULK ? sizeof(double) * 64 * (nk * 8 + 2) : 0; }
template <int P>
static int launch_jac(kl_ctx* ctx, int e2b, int e2e, cudaStream_t s) {
    using Cfg = JacCfg<P>;
    const int nel = ctx->d.nel1 * (e2e - e2b);
    if (nel <= 0) return 0;
    const size_t smem = sizeof(JacShared<P>);
    if (!(ctx->attr_done & 2u)) {
        KL_CUDA(cudaFuncSetAttribute(k_jacobian<P, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        KL_CUDA(cudaFuncSetAttribute(k_jacobian<P, true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        KL_CUDA(cudaFuncSetAttribute(k_jacobian<P, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        KL_CUDA(cudaFuncSetAttribute(k_jacobian<P, false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        ctx->attr_done |= 2u;
    }
    const int grid = (nel + Cfg::EPG - 1) / Cfg::EPG;
    // the membrane-bending coupling block B vanishes identically for the linear (SvK) law and for membranes
    const bool hasB = ctx->d.mat.material != KL_MAT_SVK && ctx->d.mat.bending;
    KL_CUDA(cudaEventRecord(ctx->ev[4], s));
    if (P == 3 && !ctx->jac_shared) {
        // sliding-window kernel: segments of element rows, sized for about 24 waves of resident CTAs, 8 .. 24 elements long
        // (measured at 576 x 576 elements, profiles/r2_ablation.txt: 8 / 12 / 16 / 24 / 32 / 48 / 64 elements per segment give
        // 3.58 / 3.52 / 3.50 / 3.51 / 3.54 / 3.60 / 3.65 ms: short segments keep the tail small, below 10 the window restarts cost more)
        int seg = ctx->jac_seg;
        if (seg <= 0) {
            const long long slots = (long long)ctx->n_sm * KL_SW_MINB * 24;
            seg = (int)std::min<long long>(24, std::max<long long>(8, ((long long)nel + slots - 1) / slots));
        }
        const int nseg = (ctx->d.nel1 + seg - 1) / seg;
        if (!(ctx->attr_done & 4u)) {
            KL_CUDA(cudaFuncSetAttribute(k_jacobian_sw<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sw_run_bytes(6)));
            KL_CUDA(cudaFuncSetAttribute(k_jacobian_sw<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sw_run_bytes(6)));
            KL_CUDA(cudaFuncSetAttribute(k_jacobian_sw<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sw_run_bytes(9)));
            ctx->attr_done |= 4u;
        }
        if (hasB) k_jacobian_sw<true><<<(e2e - e2b) * nseg, 64, sw_run_bytes(6), s>>>(ctx->d, e2b, e2e, seg);
        else k_jacobian_sw<false><<<(e2e - e2b) * nseg, 64, sw_run_bytes(6), s>>>(ctx->d, e2b, e2e, seg);
    } else if (hasB) k_jacobian<P, true><<<grid, Cfg::NT, smem, s>>>(ctx->d, e2b, e2e);
    else k_jacobian<P, false><<<grid, Cfg::NT, smem, s>>>(ctx->d, e2b, e2e);
    KL_CUDA(cudaEventRecord(ctx->ev[5], s));
    ctx->launches++;
    KL_CUDA(cudaGetLastError());
    if (ctx->d.mat.pressure != 0.0 && P == 3 && !ctx->jac_shared) {
        // same walk, pressure term only (9 unsymmetric component pairs): 1008 RED per element at arithmetic addresses
        const int seg = 16, nseg = (ctx->d.nel1 + seg - 1) / seg;
        k_jacobian_sw<false, true><<<(e2e - e2b) * nseg, 64, sw_run_bytes(9), s>>>(ctx->d, e2b, e2e, seg);
        ctx->launches++;
        KL_CUDA(cudaGetLastError());
    } else if (ctx->d.mat.pressure != 0.0) {
        k_pressure_tangent<P><<<nel, 256, 0, s>>>(ctx->d, e2b, e2e);
        ctx->launches++;
        KL_CUDA(cudaGetLastError());
    }
    return 0;
}

int kl_launch_jacobian(kl_ctx* ctx, int e2_begin, int e2_end, cudaStream_t s) {
    switch (ctx->d.p) {
        case 2: return launch_jac<2>(ctx, e2_begin, e2_end, s);
        case 3: return launch_jac<3>(ctx, e2_begin, e2_end, s);
        case 4: return launch_jac<4>(ctx, e2_begin, e2_end, s);
    }
    kl_set_error("unsupported degree");
    return KL_E_ARG;
}

template <int P>
static int launch_res(kl_ctx* ctx, double* r, cudaStream_t s, bool full) {
    using Cfg = PointCfg<P>;
    const int nel = ctx->d.nel1 * (ctx->e2_end - ctx->e2_begin);
    if (nel <= 0) return 0;
    const int grid = (nel + Cfg::EPG - 1) / Cfg::EPG;
    if (full) k_residual<P, true><<<grid, Cfg::NT, 0, s>>>(ctx->d, r, ctx->e2_begin, ctx->e2_end);
    else k_residual<P><<<grid, Cfg::NT, 0, s>>>(ctx->d, r, ctx->e2_begin, ctx->e2_end);
    ctx->launches++;
    KL_CUDA(cudaGetLastError());
    return 0;
}

int kl_launch_residual(kl_ctx* ctx, double* r_dev, cudaStream_t s, bool full) {
    switch (ctx->d.p) {
        case 2: return launch_res<2>(ctx, r_dev, s, full);
        case 3: return launch_res<3>(ctx, r_dev, s, full);
        case 4: return launch_res<4>(ctx, r_dev, s, full);
    }
    kl_set_error("unsupported degree");
    return KL_E_ARG;
}

template <int P>
static int launch_body(kl_ctx* ctx, double* f, const double* bf, cudaStream_t s) {
    using Cfg = PointCfg<P>;
    const int nel = ctx->d.nel1 * ctx->d.nel2;
    k_bodyforce<P><<<(nel + Cfg::EPG - 1) / Cfg::EPG, Cfg::NT, 0, s>>>(ctx->d, f, bf[0], bf[1], bf[2]);
    ctx->launches++;
    KL_CUDA(cudaGetLastError());
    return 0;
}
// f_dev += integral(N_i * bf * meas(ori))
int kl_launch_bodyforce(kl_ctx* ctx, double* f_dev, const double bf[3], cudaStream_t s) {
    switch (ctx->d.p) {
        case 2: return launch_body<2>(ctx, f_dev, bf, s);
        case 3: return launch_body<3>(ctx, f_dev, bf, s);
        case 4: return launch_body<4>(ctx, f_dev, bf, s);
    }
    kl_set_error("unsupported degree");
    return KL_E_ARG;
}

// ------------------------------------------------------------------------------------------------
// FP64 peak microbenchmark: 8 independent DFMA chains per thread, register resident
__global__ void __launch_bounds__(256) k_fp64_peak(double* out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

extern "C" int kl_measure_fp64_peak(int device, double* tflops, float* ms_out) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { kl_set_error("no CUDA device"); return KL_E_NOGPU; }
    if (device >= 0) KL_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    int dev;
    KL_CUDA(cudaGetDevice(&dev));
    KL_CUDA(cudaGetDeviceProperties(&prop, dev));
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 4096;
    double* out;
    KL_CUDA(cudaMalloc(&out, sizeof(double) * blocks * threads));
    cudaEvent_t e0, e1;
    KL_CUDA(cudaEventCreate(&e0)); KL_CUDA(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 6; ++rep) {
        KL_CUDA(cudaEventRecord(e0));
        k_fp64_peak<<<blocks, threads>>>(out, iters, 0.999999, 1e-9);
        KL_CUDA(cudaEventRecord(e1));
        KL_CUDA(cudaEventSynchronize(e1));
        float ms;
        KL_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    const double flops = 2.0 * 64.0 * iters * (double)blocks * threads;
    if (tflops) *tflops = flops / (best * 1e-3) / 1e12;
    if (ms_out) *ms_out = best;
    cudaFree(out); cudaEventDestroy(e0); cudaEventDestroy(e1);
    return KL_OK;
}

extern "C" int kl_points_kernel_ms(kl_ctx* ctx, float* ms) {
    if (!ctx || !ms) return KL_E_ARG;
    KL_CUDA(cudaEventSynchronize(ctx->ev[7]));
    KL_CUDA(cudaEventElapsedTime(ms, ctx->ev[6], ctx->ev[7]));
    return KL_OK;
}

extern "C" int kl_jacobian_kernel_ms(kl_ctx* ctx, float* ms) {
    if (!ctx || !ms) return KL_E_ARG;
    KL_CUDA(cudaEventSynchronize(ctx->ev[5]));
    KL_CUDA(cudaEventElapsedTime(ms, ctx->ev[4], ctx->ev[5]));
    return KL_OK;
}
