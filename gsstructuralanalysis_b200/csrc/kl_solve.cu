// kl_solve.cu — device-resident Jacobi-preconditioned conjugate gradients and the Newton loop around the assembly path.
//
// the reference's Newton solver defaults to gsSparseSolver<>::CGDiagonal
// (src/gsStaticSolvers/gsStaticNewton.hpp:23) = Eigen::ConjugateGradient with DiagonalPreconditioner (Eigen 3.4, vendored by
// G+Smo as gsEigen; third-party, not in the reference tree).  Its published iteration is restated here for the GPU:
//     r = b (x0 = 0), p = D^-1 r, absNew = r.p
//     loop: tmp = A p; alpha = absNew / p.tmp; x += alpha p; r -= alpha tmp; stop if |r|^2 < max(tol^2 |b|^2, DBL_MIN);
//           z = D^-1 r; absOld = absNew; absNew = r.z; p = z + (absNew/absOld) p
// All scalars stay on the device; a batch of iterations is one CUDA graph launch and the host only polls a `done` word.
// Reductions are two-stage with a fixed grid, so the iteration is bit-reproducible run to run.
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include "kl_internal.h"

namespace {

constexpr int CG_THREADS = 256;
constexpr int CG_BATCH = 8;          // iterations per graph launch

struct CGState {
    double pAp, absNew, absOld, rn2, threshold, rhsNorm2, alpha, beta;
    int iters, done, maxit, pad;
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// deterministic block sum (fixed tree); result valid in thread 0
__device__ __forceinline__ double block_sum(double v, double* sm) {
    v = warp_sum(v);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) sm[w] = v;
    __syncthreads();
    double s = 0.0;
    if (threadIdx.x == 0)
        for (int k = 0; k < (int)(blockDim.x >> 5); ++k) s += sm[k];
    return s;
}

// y = A^T x (= A x for the symmetric matrices CG accepts): one warp per compressed column, lanes stride its entries.
// DOT: also the block partial of x.y (p.Ap of the CG iteration) -> part[blockIdx.x].
template <bool DOT>
__global__ void __launch_bounds__(CG_THREADS) k_spmv(const int* __restrict__ outer, const int* __restrict__ inner, const double* __restrict__ val,
                                                     const double* __restrict__ x, double* __restrict__ y, int n, double* __restrict__ part,
                                                     const CGState* __restrict__ st) {
    __shared__ double sm[CG_THREADS / 32];
    if (st && st->done) return;
    const int lane = threadIdx.x & 31;
    const int warp = blockIdx.x * (CG_THREADS / 32) + (threadIdx.x >> 5);
    const int nwarps = gridDim.x * (CG_THREADS / 32);
    double dot = 0.0;
    for (int col = warp; col < n; col += nwarps) {
        const int b = outer[col], e = outer[col + 1];
        double s0 = 0.0, s1 = 0.0;
        int k = b + lane;
        for (; k + 32 < e; k += 64) {
            const double v0 = val[k], v1 = val[k + 32];
            const int i0 = inner[k], i1 = inner[k + 32];
            s0 = fma(v0, x[i0], s0);
            s1 = fma(v1, x[i1], s1);
        }
        if (k < e) s0 = fma(val[k], x[inner[k]], s0);
        const double s = warp_sum(s0 + s1);
        if (lane == 0) {
            y[col] = s;
            if (DOT) dot = fma(s, x[col], dot);
        }
    }
    if (DOT) {
        const double t = block_sum(dot, sm);
        if (threadIdx.x == 0) part[blockIdx.x] = t;
    }
}

// Regular columns (the full (2p+1)^2 x 3 stencil in canonical order, 99 % of a large mesh) do not need the row-index
// array: the rows of column (J,d) are 3 x (2p+1) runs of 2p+1 consecutive DoFs whose first indices depend on J only.
// runbase[J][c][di2] holds them (84 B per control point instead of 588 B of `inner` per column); colinfo[col] = J or -1.
__global__ void k_spmv_tables(int ncp, int nfree, int nst, int W, const int* __restrict__ map, const int* __restrict__ colbase,
                              const int* __restrict__ inner, int* __restrict__ colinfo, int* __restrict__ runbase) {
    const int J = blockIdx.x * blockDim.x + threadIdx.x;
    if (J >= ncp) return;
    const int4 cb = reinterpret_cast<const int4*>(colbase)[J];
    bool ok = cb.w != 0;
    if (ok) {
        const int base[3] = {cb.x, cb.y, cb.z};
        for (int c = 0; c < 3 && ok; ++c)
            for (int r = 0; r < W && ok; ++r) {
                const int start = inner[base[0] + c * nst + r * W];
                runbase[(size_t)J * 3 * W + c * W + r] = start;
                for (int d = 0; d < 3 && ok; ++d)
                    for (int k = 0; k < W; ++k)
                        if (inner[base[d] + c * nst + r * W + k] != start + k) { ok = false; break; }
            }
    }
    for (int d = 0; d < 3; ++d) {
        const int col = map[d * ncp + J];
        if (col < nfree) colinfo[col] = ok ? J : -1;
    }
}

template <int P, bool DOT>
__global__ void __launch_bounds__(CG_THREADS, 4) k_spmv_reg(const int* __restrict__ outer, const int* __restrict__ inner, const double* __restrict__ val,
                                                         const int* __restrict__ colinfo, const int* __restrict__ runbase,
                                                         const double* __restrict__ x, double* __restrict__ y, int n, double* __restrict__ part,
                                                         const CGState* __restrict__ st) {
    constexpr int W = 2 * P + 1, NST = W * W, NE = 3 * NST;
    __shared__ double sm[CG_THREADS / 32];
    if (st && st->done) return;
    const int lane = threadIdx.x & 31;
    const int warp = blockIdx.x * (CG_THREADS / 32) + (threadIdx.x >> 5);
    const int nwarps = gridDim.x * (CG_THREADS / 32);
    double dot = 0.0;
    // two columns per trip: the 2 x 5 value loads and x gathers of both are in flight together (the kernel is bound by
    // memory latency: one column per trip left the HBM pipe at 38 % of its peak)
    constexpr int NM = (NE + 31) / 32;
    // column start and control-point index of the next trip are fetched one trip ahead (first level of the dependent chain
    // outer / colinfo -> run table -> x)
    int nb0 = 0, nb1 = 0, nJ0 = -1, nJ1 = -1;
    if (warp < n) { nb0 = outer[warp]; nJ0 = colinfo[warp]; }
    if (warp + nwarps < n) { nb1 = outer[warp + nwarps]; nJ1 = colinfo[warp + nwarps]; }
    for (int col0 = warp; col0 < n; col0 += 2 * nwarps) {
        const int col1 = col0 + nwarps;
        const bool two = col1 < n;
        const int b0 = nb0, b1 = nb1, J0 = nJ0, J1 = two ? nJ1 : -1;
        {
            const int c2 = col0 + 2 * nwarps, c3 = col1 + 2 * nwarps;
            if (c2 < n) { nb0 = outer[c2]; nJ0 = colinfo[c2]; }
            if (c3 < n) { nb1 = outer[c3]; nJ1 = colinfo[c3]; }
        }
        double s[2] = {0.0, 0.0};
        if (J0 >= 0 && J1 >= 0) {
            const int* rb0 = runbase + (size_t)J0 * 3 * W;
            const int* rb1 = runbase + (size_t)J1 * 3 * W;
            double v0[NM], v1[NM], x0[NM], x1[NM];
#pragma unroll
            for (int m = 0; m < NM; ++m) {
                const int e = lane + 32 * m;
                v0[m] = 0.0; v1[m] = 0.0; x0[m] = 0.0; x1[m] = 0.0;
                if (e < NE) {
                    const int run = e / W, k = e - run * W;      // run = c*W + di2
                    v0[m] = val[b0 + e];
                    v1[m] = val[b1 + e];
                    x0[m] = x[rb0[run] + k];
                    x1[m] = x[rb1[run] + k];
                }
            }
            double a0 = 0.0, a1 = 0.0, c0 = 0.0, c1 = 0.0;
#pragma unroll
            for (int m = 0; m < NM; ++m) {
                if (m & 1) { a1 = fma(v0[m], x0[m], a1); c1 = fma(v1[m], x1[m], c1); }
                else { a0 = fma(v0[m], x0[m], a0); c0 = fma(v1[m], x1[m], c0); }
            }
            s[0] = a0 + a1; s[1] = c0 + c1;
        } else {
#pragma unroll 1
            for (int t = 0; t < 2; ++t) {
                const int col = t ? col1 : col0;
                if (t && !two) break;
                const int b = t ? b1 : b0, J = t ? J1 : J0;
                double s0 = 0.0, s1 = 0.0;
                if (J >= 0) {
                    const int* rb = runbase + (size_t)J * 3 * W;
#pragma unroll
                    for (int m = 0; m < NM; ++m) {
                        const int e = lane + 32 * m;
                        if (e < NE) {
                            const int run = e / W, k = e - run * W;
                            const double v = val[b + e];
                            const int row = rb[run] + k;
                            if (m & 1) s1 = fma(v, x[row], s1); else s0 = fma(v, x[row], s0);
                        }
                    }
                } else {
                    const int e = outer[col + 1];
                    int k = b + lane;
                    for (; k + 32 < e; k += 64) {
                        const double v0 = val[k], v1 = val[k + 32];
                        const int i0 = inner[k], i1 = inner[k + 32];
                        s0 = fma(v0, x[i0], s0);
                        s1 = fma(v1, x[i1], s1);
                    }
                    if (k < e) s0 = fma(val[k], x[inner[k]], s0);
                }
                s[t] = s0 + s1;
            }
        }
        const double r0 = warp_sum(s[0]), r1 = warp_sum(s[1]);
        if (lane == 0) {
            y[col0] = r0;
            if (DOT) dot = fma(r0, x[col0], dot);
            if (two) {
                y[col1] = r1;
                if (DOT) dot = fma(r1, x[col1], dot);
            }
        }
    }
    if (DOT) {
        const double t = block_sum(dot, sm);
        if (threadIdx.x == 0) part[blockIdx.x] = t;
    }
}

// EXPERIMENT (KL_SPMV_CP=1; not the production path).  Control-point form of the regular-column SpMV: the three columns
// (J,0), (J,1), (J,2) of one control point share their row set, so one warp trip gathers x once and streams the three value
// columns against it.  Motivation: ncu of k_spmv_reg shows the L1/LSU pipe at 92 % of its peak with HBM at 57 %
// (profiles/r1_spmv_summary.txt).  Result on B200 at 1M DOF: correct (tests/test_gpu_solve.py green) but slower, 0.330 ms
// against 0.284 ms; the three value streams of a trip lie a third of the matrix apart.  Columns that are not regular
// (colinfo < 0: clipped stencils, coupled DoFs) are swept by a second, generic pass of the same launch.
template <int P, bool DOT>
__global__ void __launch_bounds__(CG_THREADS, 4) k_spmv_cp(const int* __restrict__ outer, const int* __restrict__ inner, const double* __restrict__ val,
                                                        const int* __restrict__ colinfo, const int* __restrict__ runbase,
                                                        const int* __restrict__ map, int ncp,
                                                        const double* __restrict__ x, double* __restrict__ y, int n, double* __restrict__ part,
                                                        const CGState* __restrict__ st) {
    constexpr int W = 2 * P + 1, NST = W * W, NE = 3 * NST, NM = (NE + 31) / 32;
    __shared__ double sm[CG_THREADS / 32];
    if (st && st->done) return;
    const int lane = threadIdx.x & 31;
    const int warp = blockIdx.x * (CG_THREADS / 32) + (threadIdx.x >> 5);
    const int nwarps = gridDim.x * (CG_THREADS / 32);
    double dot = 0.0;
    // pass A: regular columns, one control point per trip; column ids and starts of the next trip are fetched one trip ahead
    int nc0 = -1, nc1 = -1, nc2 = -1, nb0 = 0, nb1 = 0, nb2 = 0;
#define KL_CP_FETCH(J)                                                          \
    {                                                                           \
        const int c0 = map[(J)], c1 = map[ncp + (J)], c2 = map[2 * ncp + (J)];   \
        nc0 = (c0 < n && colinfo[c0] == (J)) ? c0 : -1;                          \
        nc1 = (c1 < n && colinfo[c1] == (J)) ? c1 : -1;                          \
        nc2 = (c2 < n && colinfo[c2] == (J)) ? c2 : -1;                          \
        nb0 = nc0 >= 0 ? outer[nc0] : 0;                                         \
        nb1 = nc1 >= 0 ? outer[nc1] : 0;                                         \
        nb2 = nc2 >= 0 ? outer[nc2] : 0;                                         \
    }
    if (warp < ncp) KL_CP_FETCH(warp)
    for (int J = warp; J < ncp; J += nwarps) {
        const int c0 = nc0, c1 = nc1, c2 = nc2, b0 = nb0, b1 = nb1, b2 = nb2;
        if (J + nwarps < ncp) KL_CP_FETCH(J + nwarps)
        if (c0 < 0 && c1 < 0 && c2 < 0) continue;
        const int* rb = runbase + (size_t)J * 3 * W;
        double v0[NM], v1[NM], v2[NM], xx[NM];
#pragma unroll
        for (int m = 0; m < NM; ++m) {
            const int e = lane + 32 * m;
            v0[m] = 0.0; v1[m] = 0.0; v2[m] = 0.0; xx[m] = 0.0;
            if (e < NE) {
                const int run = e / W, k = e - run * W;      // run = c*W + di2
                if (c0 >= 0) v0[m] = val[b0 + e];
                if (c1 >= 0) v1[m] = val[b1 + e];
                if (c2 >= 0) v2[m] = val[b2 + e];
                xx[m] = x[rb[run] + k];
            }
        }
        double s0 = 0.0, s1 = 0.0, s2 = 0.0;
#pragma unroll
        for (int m = 0; m < NM; ++m) {
            s0 = fma(v0[m], xx[m], s0);
            s1 = fma(v1[m], xx[m], s1);
            s2 = fma(v2[m], xx[m], s2);
        }
        const double r0 = warp_sum(s0), r1 = warp_sum(s1), r2 = warp_sum(s2);
        if (lane == 0) {
            if (c0 >= 0) { y[c0] = r0; if (DOT) dot = fma(r0, x[c0], dot); }
            if (c1 >= 0) { y[c1] = r1; if (DOT) dot = fma(r1, x[c1], dot); }
            if (c2 >= 0) { y[c2] = r2; if (DOT) dot = fma(r2, x[c2], dot); }
        }
    }
#undef KL_CP_FETCH
    // pass B: the remaining columns through the row-index array; 32 column flags per warp load
    for (int cbase = warp * 32; cbase < n; cbase += nwarps * 32) {
        const int cmine = cbase + lane;
        unsigned todo = __ballot_sync(0xffffffffu, cmine < n && colinfo[cmine] < 0);
        while (todo) {
            const int col = cbase + __ffs(todo) - 1;
            todo &= todo - 1;
            const int b = outer[col], e = outer[col + 1];
            double s0 = 0.0, s1 = 0.0;
            int k = b + lane;
            for (; k + 32 < e; k += 64) {
                const double a0 = val[k], a1 = val[k + 32];
                const int i0 = inner[k], i1 = inner[k + 32];
                s0 = fma(a0, x[i0], s0);
                s1 = fma(a1, x[i1], s1);
            }
            if (k < e) s0 = fma(val[k], x[inner[k]], s0);
            const double r = warp_sum(s0 + s1);
            if (lane == 0) {
                y[col] = r;
                if (DOT) dot = fma(r, x[col], dot);
            }
        }
    }
    if (DOT) {
        const double t = block_sum(dot, sm);
        if (threadIdx.x == 0) part[blockIdx.x] = t;
    }
}

// 1 / diagonal (1 where it is zero or absent) — Eigen::DiagonalPreconditioner
__global__ void k_cg_invdiag(const int* __restrict__ outer, const int* __restrict__ inner, const double* __restrict__ val, double* __restrict__ invdiag, int n) {
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= n) return;
    int lo = outer[col], hi = outer[col + 1] - 1;
    double dg = 0.0;
    while (lo <= hi) {   // inner indices of a column are sorted
        const int mid = (lo + hi) >> 1;
        const int r = inner[mid];
        if (r == col) { dg = val[mid]; break; }
        if (r < col) lo = mid + 1; else hi = mid - 1;
    }
    invdiag[col] = dg != 0.0 ? 1.0 / dg : 1.0;
}

// x = 0, r = b, p = D^-1 r; partials of |b|^2 and r.p
__global__ void __launch_bounds__(CG_THREADS) k_cg_init(const double* __restrict__ b, const double* __restrict__ invdiag, double* __restrict__ x,
                                                        double* __restrict__ r, double* __restrict__ p, int n, double* __restrict__ part, int nb) {
    __shared__ double sm[CG_THREADS / 32];
    double s0 = 0.0, s1 = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double bi = b[i], pi = invdiag[i] * bi;
        x[i] = 0.0; r[i] = bi; p[i] = pi;
        s0 = fma(bi, bi, s0);
        s1 = fma(bi, pi, s1);
    }
    const double t0 = block_sum(s0, sm);
    const double t1 = block_sum(s1, sm);
    if (threadIdx.x == 0) { part[blockIdx.x] = t0; part[nb + blockIdx.x] = t1; }
}

__device__ __forceinline__ double sum_partials(const double* part, int nb, double* sm) {
    double s = 0.0;
    for (int k = threadIdx.x; k < nb; k += blockDim.x) s += part[k];
    return block_sum(s, sm);
}

__global__ void __launch_bounds__(CG_THREADS) k_cg_scal_init(CGState* st, const double* __restrict__ part, int nb, double tol, int maxit) {
    __shared__ double sm[CG_THREADS / 32];
    const double b2 = sum_partials(part, nb, sm);
    const double rp = sum_partials(part + nb, nb, sm);
    if (threadIdx.x == 0) {
        st->rhsNorm2 = b2; st->rn2 = b2; st->absNew = rp; st->absOld = rp;
        st->threshold = fmax(tol * tol * b2, DBL_MIN);
        st->iters = 0; st->maxit = maxit; st->alpha = st->beta = st->pAp = 0.0;
        st->done = (b2 == 0.0 || b2 < st->threshold || maxit <= 0) ? 1 : 0;
    }
}

__global__ void __launch_bounds__(CG_THREADS) k_cg_scal_alpha(CGState* st, const double* __restrict__ part, int nb) {
    __shared__ double sm[CG_THREADS / 32];
    if (st->done) return;
    const double pAp = sum_partials(part, nb, sm);
    if (threadIdx.x == 0) { st->pAp = pAp; st->alpha = st->absNew / pAp; }
}

// x += alpha p; r -= alpha tmp; z = D^-1 r; partials of |r|^2 and r.z
__global__ void __launch_bounds__(CG_THREADS) k_cg_update(const CGState* __restrict__ st, const double* __restrict__ p, const double* __restrict__ tmp,
                                                          const double* __restrict__ invdiag, double* __restrict__ x, double* __restrict__ r,
                                                          double* __restrict__ z, int n, double* __restrict__ part, int nb) {
    __shared__ double sm[CG_THREADS / 32];
    if (st->done) return;
    const double alpha = st->alpha;
    double s0 = 0.0, s1 = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        x[i] = fma(alpha, p[i], x[i]);
        const double ri = fma(-alpha, tmp[i], r[i]);
        const double zi = invdiag[i] * ri;
        r[i] = ri; z[i] = zi;
        s0 = fma(ri, ri, s0);
        s1 = fma(ri, zi, s1);
    }
    const double t0 = block_sum(s0, sm);
    const double t1 = block_sum(s1, sm);
    if (threadIdx.x == 0) { part[blockIdx.x] = t0; part[nb + blockIdx.x] = t1; }
}

__global__ void __launch_bounds__(CG_THREADS) k_cg_scal_beta(CGState* st, const double* __restrict__ part, int nb) {
    __shared__ double sm[CG_THREADS / 32];
    if (st->done) return;
    const double rn2 = sum_partials(part, nb, sm);
    const double rz = sum_partials(part + nb, nb, sm);
    if (threadIdx.x == 0) {
        st->rn2 = rn2;
        if (rn2 < st->threshold) { st->done = 1; return; }   // Eigen leaves the loop before counting this iteration
        st->absOld = st->absNew; st->absNew = rz; st->beta = rz / st->absOld;
        st->iters += 1;
        if (st->iters >= st->maxit) st->done = 1;
    }
}

__global__ void __launch_bounds__(CG_THREADS) k_cg_dir(const CGState* __restrict__ st, const double* __restrict__ z, double* __restrict__ p, int n) {
    if (st->done) return;
    const double beta = st->beta;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) p[i] = fma(beta, p[i], z[i]);
}

// Newton helpers: out = a + s*b; partial of |v|^2
__global__ void __launch_bounds__(CG_THREADS) k_vec_axpy_out(double* out, const double* a, const double* __restrict__ b, double s, int n) {   // out may alias a
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = fma(s, b[i], a[i]);
}
__global__ void __launch_bounds__(CG_THREADS) k_vec_norm2(const double* __restrict__ v, int n, double* __restrict__ part) {
    __shared__ double sm[CG_THREADS / 32];
    double s = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) s = fma(v[i], v[i], s);
    const double t = block_sum(s, sm);
    if (threadIdx.x == 0) part[blockIdx.x] = t;
}
__global__ void __launch_bounds__(CG_THREADS) k_vec_norm_final(const double* __restrict__ part, int nb, double* out) {
    __shared__ double sm[CG_THREADS / 32];
    const double s = sum_partials(part, nb, sm);
    if (threadIdx.x == 0) *out = sqrt(s);
}

// the seven inner products of one Crisfield corrector iteration in one pass (gsALMCrisfield::computeLambdasSimple / computeLambdaDOT,
// src/gsALMSolvers/gsALMCrisfield.hpp:206-226,404-425): part[k][block], k = Ut.Ut, Ut.DU, Ubar.Ut, DU.DU, DU.Ubar, Ubar.Ubar, DUold.Ut
__global__ void __launch_bounds__(CG_THREADS) k_alm_dots(const double* __restrict__ Ut, const double* __restrict__ Ubar, const double* __restrict__ DU,
                                                         const double* __restrict__ DUold, int n, double* __restrict__ part) {
    __shared__ double sm[CG_THREADS / 32];
    double a[7] = {0, 0, 0, 0, 0, 0, 0};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double ut = Ut[i], ub = Ubar[i], du = DU[i], dold = DUold[i];
        a[0] = fma(ut, ut, a[0]); a[1] = fma(ut, du, a[1]); a[2] = fma(ub, ut, a[2]); a[3] = fma(du, du, a[3]);
        a[4] = fma(du, ub, a[4]); a[5] = fma(ub, ub, a[5]); a[6] = fma(dold, ut, a[6]);
    }
#pragma unroll
    for (int k = 0; k < 7; ++k) {
        const double t = block_sum(a[k], sm);
        if (threadIdx.x == 0) part[k * gridDim.x + blockIdx.x] = t;
    }
}
__global__ void __launch_bounds__(CG_THREADS) k_alm_dots_final(const double* __restrict__ part, int nb, double* __restrict__ out) {
    __shared__ double sm[CG_THREADS / 32];
    for (int k = 0; k < 7; ++k) {
        const double s = sum_partials(part + k * nb, nb, sm);
        if (threadIdx.x == 0) out[k] = s;
        __syncthreads();
    }
}
// out = a*x + b*y (y may be null)
__global__ void __launch_bounds__(CG_THREADS) k_vec_lin2(double* out, const double* x, double a, const double* y, double b, int n) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = y ? fma(a, x[i], b * y[i]) : a * x[i];
}

}  // namespace

struct KLSolveWS {
    double *aU = nullptr, *aDU = nullptr, *adU = nullptr, *aUt = nullptr, *aUbar = nullptr, *aR = nullptr, *aDUold = nullptr, *aX = nullptr;   // arc-length vectors (lazy)
    double* dots = nullptr;       // device [8], pinned twin dots_host
    double* dots_host = nullptr;
    int n = 0, nb = 0, nb_spmv = 0;
    double *p = nullptr, *tmp = nullptr, *z = nullptr, *r = nullptr, *x = nullptr, *invdiag = nullptr, *part = nullptr, *b = nullptr;
    int *colinfo = nullptr, *runbase = nullptr;   // index-free SpMV of regular columns
    CGState* st = nullptr;        // device
    CGState* st_host = nullptr;   // pinned
    double* scal = nullptr;       // device scratch scalar (norms)
    double* scal_host = nullptr;  // pinned
    cudaGraphExec_t graph = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr, e2 = nullptr, e3 = nullptr;
    double *nU = nullptr, *nDU = nullptr, *ndU = nullptr, *nR = nullptr, *nX = nullptr;   // Newton vectors
    float ms_total = 0, ms_iter = 0, ms_spmv = 0;
};

static int ws_get(kl_ctx* ctx, KLSolveWS** out) {
    if (ctx->solve_ws) { *out = ctx->solve_ws; return KL_OK; }
    KLSolveWS* w = new KLSolveWS();
    ctx->solve_ws = w;
    const int n = ctx->d.nfree;
    w->n = n;
    int dev = 0, nsm = 0;
    KL_CUDA(cudaGetDevice(&dev));
    KL_CUDA(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
    const int want = (n + CG_THREADS - 1) / CG_THREADS;
    w->nb = want < nsm * 4 ? (want > 0 ? want : 1) : nsm * 4;
    const int wantw = (n + CG_THREADS / 32 - 1) / (CG_THREADS / 32);
    // grid-stride kernel: launch exactly the blocks that are resident at once (4 per SM, __launch_bounds__ of k_spmv_reg); the
    // former 8 per SM ran a partial second wave at a third of the occupancy
    w->nb_spmv = wantw < nsm * 4 ? (wantw > 0 ? wantw : 1) : nsm * 4;
    const size_t vb = sizeof(double) * (size_t)(n > 0 ? n : 1);
    double** vecs[] = {&w->p, &w->tmp, &w->z, &w->r, &w->x, &w->invdiag, &w->b, &w->nU, &w->nDU, &w->ndU, &w->nR, &w->nX};
    for (double** v : vecs) KL_CUDA(cudaMalloc((void**)v, vb));
    const int npart = 8 * (w->nb > w->nb_spmv ? w->nb : w->nb_spmv);
    KL_CUDA(cudaMalloc((void**)&w->part, sizeof(double) * npart));
    KL_CUDA(cudaMalloc((void**)&w->st, sizeof(CGState)));
    KL_CUDA(cudaMalloc((void**)&w->scal, sizeof(double) * 4));
    KL_CUDA(cudaMallocHost((void**)&w->st_host, sizeof(CGState)));
    KL_CUDA(cudaMallocHost((void**)&w->scal_host, sizeof(double) * 4));
    static const bool no_fast = getenv("KL_SPMV_GENERIC") != nullptr;
    if (!no_fast && n > 0 && ctx->d.ncp > 0) {      // the matrix context of a kl_mp has no single tensor grid: generic SpMV
        const int W = 2 * ctx->d.p + 1;
        KL_CUDA(cudaMalloc((void**)&w->colinfo, sizeof(int) * (size_t)n));
        KL_CUDA(cudaMalloc((void**)&w->runbase, sizeof(int) * (size_t)ctx->d.ncp * 3 * W));
        KL_CUDA(cudaMemset(w->colinfo, 0xff, sizeof(int) * (size_t)n));
        k_spmv_tables<<<(ctx->d.ncp + 127) / 128, 128>>>(ctx->d.ncp, n, ctx->d.nst, W, ctx->d.map, ctx->d.colbase, ctx->d.inner, w->colinfo, w->runbase);
        KL_CUDA(cudaGetLastError());
        KL_CUDA(cudaDeviceSynchronize());
        ctx->launches++;
    }
    KL_CUDA(cudaEventCreate(&w->e0));
    KL_CUDA(cudaEventCreate(&w->e1));
    KL_CUDA(cudaEventCreate(&w->e2));
    KL_CUDA(cudaEventCreate(&w->e3));
    *out = w;
    return KL_OK;
}

void kl_solve_free(kl_ctx* ctx) {
    KLSolveWS* w = ctx->solve_ws;
    if (!w) return;
    double* vecs[] = {w->p, w->tmp, w->z, w->r, w->x, w->invdiag, w->b, w->nU, w->nDU, w->ndU, w->nR, w->nX, w->part, w->scal,
                      w->aU, w->aDU, w->adU, w->aUt, w->aUbar, w->aR, w->aDUold, w->aX, w->dots};
    if (w->dots_host) cudaFreeHost(w->dots_host);
    for (double* v : vecs) if (v) cudaFree(v);
    if (w->st) cudaFree(w->st);
    if (w->colinfo) cudaFree(w->colinfo);
    if (w->runbase) cudaFree(w->runbase);
    if (w->st_host) cudaFreeHost(w->st_host);
    if (w->scal_host) cudaFreeHost(w->scal_host);
    if (w->graph) cudaGraphExecDestroy(w->graph);
    for (cudaEvent_t e : {w->e0, w->e1, w->e2, w->e3}) if (e) cudaEventDestroy(e);
    delete w;
    ctx->solve_ws = nullptr;
}

template <bool DOT>
static int launch_spmv(kl_ctx* ctx, KLSolveWS* w, const double* x, double* y, double* part, const CGState* st, cudaStream_t s) {
    const KLDev& d = ctx->d;
    // experiment, off by default: measured slower than k_spmv_reg (SpMV 0.330 vs 0.284 ms, CG iteration 0.484 vs 0.329 ms at
    // 1M DOF, profiles/r1_ablation.txt)
    static const bool by_cp = getenv("KL_SPMV_CP") != nullptr;
    if (w->colinfo && by_cp) {
        switch (d.p) {
            case 2: k_spmv_cp<2, DOT><<<w->nb_spmv, CG_THREADS, 0, s>>>(d.outer, d.inner, d.values, w->colinfo, w->runbase, d.map, d.ncp, x, y, w->n, part, st); break;
            case 3: k_spmv_cp<3, DOT><<<w->nb_spmv, CG_THREADS, 0, s>>>(d.outer, d.inner, d.values, w->colinfo, w->runbase, d.map, d.ncp, x, y, w->n, part, st); break;
            case 4: k_spmv_cp<4, DOT><<<w->nb_spmv, CG_THREADS, 0, s>>>(d.outer, d.inner, d.values, w->colinfo, w->runbase, d.map, d.ncp, x, y, w->n, part, st); break;
            default: kl_set_error("unsupported degree"); return KL_E_ARG;
        }
    } else if (w->colinfo) {
        switch (d.p) {
            case 2: k_spmv_reg<2, DOT><<<w->nb_spmv, CG_THREADS, 0, s>>>(d.outer, d.inner, d.values, w->colinfo, w->runbase, x, y, w->n, part, st); break;
            case 3: k_spmv_reg<3, DOT><<<w->nb_spmv, CG_THREADS, 0, s>>>(d.outer, d.inner, d.values, w->colinfo, w->runbase, x, y, w->n, part, st); break;
            case 4: k_spmv_reg<4, DOT><<<w->nb_spmv, CG_THREADS, 0, s>>>(d.outer, d.inner, d.values, w->colinfo, w->runbase, x, y, w->n, part, st); break;
            default: kl_set_error("unsupported degree"); return KL_E_ARG;
        }
    } else {
        k_spmv<DOT><<<w->nb_spmv, CG_THREADS, 0, s>>>(d.outer, d.inner, d.values, x, y, w->n, part, st);
    }
    KL_CUDA(cudaGetLastError());
    return KL_OK;
}

static int cg_iteration_launch(kl_ctx* ctx, KLSolveWS* w, cudaStream_t s) {
    const KLDev& d = ctx->d;
    const int n = w->n;
    if (int rc = launch_spmv<true>(ctx, w, w->p, w->tmp, w->part, w->st, s)) return rc;
    k_cg_scal_alpha<<<1, CG_THREADS, 0, s>>>(w->st, w->part, w->nb_spmv);
    k_cg_update<<<w->nb, CG_THREADS, 0, s>>>(w->st, w->p, w->tmp, w->invdiag, w->x, w->r, w->z, n, w->part, w->nb);
    k_cg_scal_beta<<<1, CG_THREADS, 0, s>>>(w->st, w->part, w->nb);
    k_cg_dir<<<w->nb, CG_THREADS, 0, s>>>(w->st, w->z, w->p, n);
    KL_CUDA(cudaGetLastError());
    return KL_OK;
}

// solves K x = b with b in w->b, result in w->x (both device, context-owned so that the iteration graph is captured once)
static int cg_run(kl_ctx* ctx, KLSolveWS* w, double tol, int max_iter, int* iters, double* rel_err, cudaStream_t s) {
    const KLDev& d = ctx->d;
    const int n = w->n;
    if (ctx->d.mat.pressure != 0.0) {
        kl_set_error("kl_cg_solve: the follower-pressure tangent is unsymmetric; conjugate gradients need a symmetric matrix");
        return KL_E_ARG;
    }
    if (tol <= 0.0) tol = DBL_EPSILON;
    if (max_iter <= 0) max_iter = 2 * n;
    KL_CUDA(cudaEventRecord(w->e0, s));
    k_cg_invdiag<<<(n + 255) / 256, 256, 0, s>>>(d.outer, d.inner, d.values, w->invdiag, n);
    k_cg_init<<<w->nb, CG_THREADS, 0, s>>>(w->b, w->invdiag, w->x, w->r, w->p, n, w->part, w->nb);
    k_cg_scal_init<<<1, CG_THREADS, 0, s>>>(w->st, w->part, w->nb, tol, max_iter);
    KL_CUDA(cudaGetLastError());
    ctx->launches += 3;
    const bool use_graph = s != nullptr && s != cudaStreamLegacy;   // the legacy default stream cannot be captured
    if (use_graph && !w->graph) {
        cudaGraph_t g = nullptr;
        KL_CUDA(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
        int rc = KL_OK;
        for (int k = 0; k < CG_BATCH && rc == KL_OK; ++k) rc = cg_iteration_launch(ctx, w, s);
        cudaError_t ce = cudaStreamEndCapture(s, &g);
        if (rc != KL_OK) { if (g) cudaGraphDestroy(g); return rc; }
        KL_CUDA(ce);
        KL_CUDA(cudaGraphInstantiate(&w->graph, g, 0));
        KL_CUDA(cudaGraphDestroy(g));
    }
    long batches = 0;
    for (;;) {
        KL_CUDA(cudaMemcpyAsync(w->st_host, w->st, sizeof(CGState), cudaMemcpyDeviceToHost, s));
        KL_CUDA(cudaStreamSynchronize(s));
        if (w->st_host->done) break;
        {   // an unassembled (p.Ap = 0 -> alpha = inf) or broken matrix makes the iteration non-finite: the stop test |r|^2 < threshold can
            // then never fire, so leave at once.  A finite p.Ap <= 0 is NOT an error: Eigen's ConjugateGradient has no such test, and the
            // arc-length solvers do solve with an indefinite tangent past a limit point (gsALMBase.hpp:249-258) — the iteration goes on
            // until the tolerance or max_iter, exactly as in the reference.
            const CGState& c = *w->st_host;
            if (batches > 0 && (!std::isfinite(c.rn2) || !std::isfinite(c.pAp))) break;
        }
        if (use_graph) KL_CUDA(cudaGraphLaunch(w->graph, s));
        else
            for (int k = 0; k < CG_BATCH; ++k)
                if (int rc = cg_iteration_launch(ctx, w, s)) return rc;
        ++batches;
    }
    KL_CUDA(cudaEventRecord(w->e1, s));
    KL_CUDA(cudaStreamSynchronize(s));
    const CGState& st = *w->st_host;
    ctx->launches += (int)(batches * CG_BATCH * 5);
    if (iters) *iters = st.iters;
    if (rel_err) *rel_err = st.rhsNorm2 > 0.0 ? std::sqrt(st.rn2 / st.rhsNorm2) : 0.0;
    if (!std::isfinite(st.rn2) || !std::isfinite(st.pAp)) {
        kl_set_error("kl_cg_solve: non-finite value in the iteration (matrix not assembled?)");
        return KL_E_NONFINITE;
    }
    cudaEventElapsedTime(&w->ms_total, w->e0, w->e1);
    const long done_iters = st.iters + (st.rn2 < st.threshold && st.rhsNorm2 > 0.0 ? 1 : 0);
    w->ms_iter = done_iters > 0 ? w->ms_total / (float)done_iters : 0.f;
    return KL_OK;
}

extern "C" int kl_cg_solve_device(kl_ctx* ctx, const double* b_dev, double* x_dev, double tol, int32_t max_iter, int32_t* iters,
                                  double* rel_err, void* stream) {
    if (!ctx || !b_dev || !x_dev) { kl_set_error("kl_cg_solve_device: null argument"); return KL_E_ARG; }
    KL_CUDA(cudaSetDevice(ctx->device));
    KLSolveWS* w = nullptr;
    int rc = ws_get(ctx, &w);
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    const size_t vb = sizeof(double) * (size_t)w->n;
    KL_CUDA(cudaMemcpyAsync(w->b, b_dev, vb, cudaMemcpyDeviceToDevice, s));
    if ((rc = cg_run(ctx, w, tol, max_iter, iters, rel_err, s))) return rc;
    KL_CUDA(cudaMemcpyAsync(x_dev, w->x, vb, cudaMemcpyDeviceToDevice, s));
    KL_CUDA(cudaStreamSynchronize(s));
    return KL_OK;
}

extern "C" int kl_cg_solve(kl_ctx* ctx, const double* b_host, double* x_host, double tol, int32_t max_iter, int32_t* iters, double* rel_err) {
    if (!ctx || !b_host || !x_host) { kl_set_error("kl_cg_solve: null argument"); return KL_E_ARG; }
    KL_CUDA(cudaSetDevice(ctx->device));
    KLSolveWS* w = nullptr;
    int rc = ws_get(ctx, &w);
    if (rc) return rc;
    cudaStream_t s = ctx->stream;
    const size_t vb = sizeof(double) * (size_t)w->n;
    std::memcpy(ctx->h_pinned_x, b_host, vb);
    KL_CUDA(cudaMemcpyAsync(w->b, ctx->h_pinned_x, vb, cudaMemcpyHostToDevice, s));
    if ((rc = cg_run(ctx, w, tol, max_iter, iters, rel_err, s))) return rc;
    KL_CUDA(cudaMemcpyAsync(ctx->h_pinned_r, w->x, vb, cudaMemcpyDeviceToHost, s));
    KL_CUDA(cudaStreamSynchronize(s));
    std::memcpy(x_host, ctx->h_pinned_r, vb);
    return KL_OK;
}

extern "C" int kl_spmv(kl_ctx* ctx, const double* x_host, double* y_host) {
    if (!ctx || !x_host || !y_host) { kl_set_error("kl_spmv: null argument"); return KL_E_ARG; }
    KL_CUDA(cudaSetDevice(ctx->device));
    KLSolveWS* w = nullptr;
    int rc = ws_get(ctx, &w);
    if (rc) return rc;
    cudaStream_t s = ctx->stream;
    const size_t vb = sizeof(double) * (size_t)w->n;
    std::memcpy(ctx->h_pinned_x, x_host, vb);
    KL_CUDA(cudaMemcpyAsync(w->p, ctx->h_pinned_x, vb, cudaMemcpyHostToDevice, s));
    static const bool twice = getenv("KL_SPMV_TWICE") != nullptr;   // timing aid: time a second, warm launch
    if (twice && (rc = launch_spmv<false>(ctx, w, w->p, w->tmp, nullptr, nullptr, s))) return rc;
    KL_CUDA(cudaEventRecord(w->e2, s));
    if ((rc = launch_spmv<false>(ctx, w, w->p, w->tmp, nullptr, nullptr, s))) return rc;
    KL_CUDA(cudaEventRecord(w->e3, s));
    KL_CUDA(cudaGetLastError());
    ctx->launches++;
    KL_CUDA(cudaMemcpyAsync(ctx->h_pinned_r, w->tmp, vb, cudaMemcpyDeviceToHost, s));
    KL_CUDA(cudaStreamSynchronize(s));
    cudaEventElapsedTime(&w->ms_spmv, w->e2, w->e3);
    std::memcpy(y_host, ctx->h_pinned_r, vb);
    return KL_OK;
}

extern "C" int kl_cg_last_timing(const kl_ctx* ctx, float* ms_total, float* ms_per_iter, float* ms_spmv) {
    if (!ctx || !ctx->solve_ws) return KL_E_ARG;
    if (ms_total) *ms_total = ctx->solve_ws->ms_total;
    if (ms_per_iter) *ms_per_iter = ctx->solve_ws->ms_iter;
    if (ms_spmv) *ms_spmv = ctx->solve_ws->ms_spmv;
    return KL_OK;
}

// ---- Newton loop (gsStaticNewton<T>::_solveNonlinear, src/gsStaticSolvers/gsStaticNewton.hpp:141-196; _start :282-325) ----
static int dev_norm(kl_ctx* ctx, KLSolveWS* w, const double* v, double* out, cudaStream_t s) {
    k_vec_norm2<<<w->nb, CG_THREADS, 0, s>>>(v, w->n, w->part);
    k_vec_norm_final<<<1, CG_THREADS, 0, s>>>(w->part, w->nb, w->scal);
    KL_CUDA(cudaGetLastError());
    ctx->launches += 2;
    KL_CUDA(cudaMemcpyAsync(w->scal_host, w->scal, sizeof(double), cudaMemcpyDeviceToHost, s));
    KL_CUDA(cudaStreamSynchronize(s));
    *out = w->scal_host[0];
    return KL_OK;
}

extern "C" int kl_newton_solve(kl_ctx* ctx, double* U_host, const kl_newton_options* opt, kl_newton_info* info) {
    if (!ctx || !U_host || !opt || !info) { kl_set_error("kl_newton_solve: null argument"); return KL_E_ARG; }
    KL_CUDA(cudaSetDevice(ctx->device));
    KLSolveWS* w = nullptr;
    int rc = ws_get(ctx, &w);
    if (rc) return rc;
    cudaStream_t s = ctx->stream;
    const int n = w->n;
    const size_t vb = sizeof(double) * (size_t)n;
    const double relax = opt->relaxation != 0.0 ? opt->relaxation : 1.0;
    std::memset(info, 0, sizeof(*info));
    float ms_asm = 0.f, ms_sol = 0.f, ms = 0.f;
    cudaEvent_t ea = ctx->ev[0], eb = ctx->ev[1];
    int it_cg = 0;
    double err_cg = 0.0;
#define NW_ASM(call)                                                                                   \
    do {                                                                                               \
        KL_CUDA(cudaEventRecord(ea, s));                                                               \
        int rc_ = (call);                                                                              \
        if (rc_ == KL_OK) rc_ = kl_check(ctx, s);                                                      \
        KL_CUDA(cudaEventRecord(eb, s));                                                               \
        KL_CUDA(cudaStreamSynchronize(s));                                                             \
        cudaEventElapsedTime(&ms, ea, eb); ms_asm += ms;                                               \
        if (rc_ == KL_E_CUDA || rc_ == KL_E_ARG) return rc_;                                           \
        if (rc_ != KL_OK) { info->status = 2; info->ms_assembly = ms_asm; info->ms_solve = ms_sol; return KL_OK; } \
    } while (0)
#define NW_CG()                                                                                        \
    do {                                                                                               \
        int rc_ = cg_run(ctx, w, opt->cg_tol, opt->cg_max_iter, &it_cg, &err_cg, s);                   \
        ms_sol += w->ms_total; info->cg_iterations += it_cg;                                           \
        if (rc_ == KL_E_CUDA || rc_ == KL_E_ARG) return rc_;                                           \
        if (rc_ != KL_OK) { info->status = 3; info->ms_assembly = ms_asm; info->ms_solve = ms_sol; return KL_OK; } \
    } while (0)

    // m_U
    std::memcpy(ctx->h_pinned_x, U_host, vb);
    KL_CUDA(cudaMemcpyAsync(w->nU, ctx->h_pinned_x, vb, cudaMemcpyHostToDevice, s));
    double residual = 0.0, residualIni = 0.0;
    if (opt->linear_start) {
        // deltaU = DeltaU = K(0)^-1 F; U = 0; relative residual based on the linear solution (headstart)
        NW_ASM(kl_jacobian_device(ctx, nullptr, s));
        KL_CUDA(cudaMemcpyAsync(w->b, ctx->d_force, vb, cudaMemcpyDeviceToDevice, s));   // K(0) DU = Force (lifting and undeformed pressure load included)
        NW_CG();
        KL_CUDA(cudaMemcpyAsync(w->nDU, w->x, vb, cudaMemcpyDeviceToDevice, s));
        KL_CUDA(cudaMemsetAsync(w->nU, 0, vb, s));
        k_vec_axpy_out<<<w->nb, CG_THREADS, 0, s>>>(w->nX, w->nU, w->nDU, 1.0, n);
        NW_ASM(kl_residual_device(ctx, w->nX, 1.0, -1.0, w->nR, s));
        if ((rc = dev_norm(ctx, w, w->nR, &residual, s))) return rc;
        if (residual == 0.0) residual = 1.0;
        NW_ASM(kl_residual_device(ctx, w->nU, 1.0, -1.0, w->ndU, s));     // Residual(U) only for its norm
        if ((rc = dev_norm(ctx, w, w->ndU, &residualIni, s))) return rc;
        if (residualIni == 0.0) residualIni = 1.0;
    } else {
        KL_CUDA(cudaMemsetAsync(w->nDU, 0, vb, s));
        NW_ASM(kl_residual_device(ctx, w->nU, 1.0, -1.0, w->nR, s));
        if ((rc = dev_norm(ctx, w, w->nR, &residual, s))) return rc;
        if (residual == 0.0) residual = 1.0;
        residualIni = residual;
    }
    info->residual_ini = residualIni;
    info->status = 1;
    const int maxit = opt->max_it > 0 ? opt->max_it : 25;
    int k = 0;
    for (; k != maxit; ++k) {
        k_vec_axpy_out<<<w->nb, CG_THREADS, 0, s>>>(w->nX, w->nU, w->nDU, 1.0, n);
        NW_ASM(kl_jacobian_device(ctx, w->nX, s));
        KL_CUDA(cudaMemcpyAsync(w->b, w->nR, vb, cudaMemcpyDeviceToDevice, s));
        NW_CG();
        KL_CUDA(cudaMemcpyAsync(w->ndU, w->x, vb, cudaMemcpyDeviceToDevice, s));
        k_vec_axpy_out<<<w->nb, CG_THREADS, 0, s>>>(w->nDU, w->nDU, w->ndU, relax, n);
        k_vec_axpy_out<<<w->nb, CG_THREADS, 0, s>>>(w->nX, w->nU, w->nDU, 1.0, n);
        ctx->launches += 3;
        NW_ASM(kl_residual_device(ctx, w->nX, 1.0, -1.0, w->nR, s));
        double ndU = 0.0, nDU = 0.0;
        if ((rc = dev_norm(ctx, w, w->nR, &residual, s))) return rc;
        if ((rc = dev_norm(ctx, w, w->ndU, &ndU, s))) return rc;
        if ((rc = dev_norm(ctx, w, w->nDU, &nDU, s))) return rc;
        info->residual = residual; info->dU_norm = relax * ndU; info->DU_norm = nDU;
        if (relax * ndU / nDU < opt->tolU && residual / residualIni < opt->tolF) { info->status = 0; break; }
    }
    info->iterations = k;      // m_numIterations: index of the converging iteration, or maxIt
    // U += DeltaU in both outcomes (gsStaticNewton.hpp:180,185)
    k_vec_axpy_out<<<w->nb, CG_THREADS, 0, s>>>(w->nU, w->nU, w->nDU, 1.0, n);
    KL_CUDA(cudaGetLastError());
    KL_CUDA(cudaMemcpyAsync(ctx->h_pinned_r, w->nU, vb, cudaMemcpyDeviceToHost, s));
    KL_CUDA(cudaStreamSynchronize(s));
    std::memcpy(U_host, ctx->h_pinned_r, vb);
    info->ms_assembly = ms_asm;
    info->ms_solve = ms_sol;
    return KL_OK;
#undef NW_ASM
#undef NW_CG
}


// ---- Crisfield arc-length step (gsALMBase<T>::_step, src/gsALMSolvers/gsALMBase.hpp:354-416, with gsALMCrisfield<T>,
//      src/gsALMSolvers/gsALMCrisfield.hpp:66-226,328-425), device resident: per corrector iteration one Jacobian, two CGDiagonal
//      solves with the same matrix (deltaUt = K^-1 F, deltaUbar = -K^-1 R), one arc-length residual and the constraint update; only
//      the solution vectors cross PCIe, once per step.  Options follow gsALMBase::defaultOptions (:25-52) / gsALMCrisfield (:26-27):
//      AngleMethod = 0 (previous step), no quasi-Newton, no stability computation.
static int alm_ws(kl_ctx* ctx, KLSolveWS* w) {
    if (w->aU) return KL_OK;
    const size_t vb = sizeof(double) * (size_t)(w->n > 0 ? w->n : 1);
    double** vecs[] = {&w->aU, &w->aDU, &w->adU, &w->aUt, &w->aUbar, &w->aR, &w->aDUold, &w->aX};
    for (double** v : vecs) KL_CUDA(cudaMalloc((void**)v, vb));
    KL_CUDA(cudaMalloc((void**)&w->dots, sizeof(double) * 8));
    KL_CUDA(cudaMallocHost((void**)&w->dots_host, sizeof(double) * 8));
    (void)ctx;
    return KL_OK;
}

static inline int sgn(double v) { return (v > 0) - (v < 0); }

extern "C" int kl_alm_step(kl_ctx* ctx, double* U_host, double* L_io, double* DUold_host, double* DLold_io, double arc_length,
                           const kl_alm_options* opt, kl_alm_info* info) {
    if (!ctx || !U_host || !L_io || !DLold_io || !opt || !info) { kl_set_error("kl_alm_step: null argument"); return KL_E_ARG; }
    KL_CUDA(cudaSetDevice(ctx->device));
    KLSolveWS* w = nullptr;
    int rc = ws_get(ctx, &w);
    if (rc) return rc;
    if ((rc = alm_ws(ctx, w))) return rc;
    cudaStream_t s = ctx->stream;
    const int n = w->n;
    const size_t vb = sizeof(double) * (size_t)n;
    std::memset(info, 0, sizeof(*info));
    float ms_asm = 0.f, ms_sol = 0.f, ms = 0.f;
    cudaEvent_t ea = ctx->ev[0], eb = ctx->ev[1];
    int it_cg = 0;
    double err_cg = 0.0;
    const double* F = ctx->d_force;
#define AL_FAIL(code) do { info->status = (code); info->ms_assembly = ms_asm; info->ms_solve = ms_sol; return KL_OK; } while (0)
#define AL_ASM(call)                                                                                   \
    do {                                                                                               \
        KL_CUDA(cudaEventRecord(ea, s));                                                               \
        int rc_ = (call);                                                                              \
        if (rc_ == KL_OK) rc_ = kl_check(ctx, s);                                                      \
        KL_CUDA(cudaEventRecord(eb, s));                                                               \
        KL_CUDA(cudaStreamSynchronize(s));                                                             \
        cudaEventElapsedTime(&ms, ea, eb); ms_asm += ms;                                               \
        if (rc_ == KL_E_CUDA || rc_ == KL_E_ARG) return rc_;                                           \
        if (rc_ != KL_OK) AL_FAIL(2);                                                                  \
    } while (0)
#define AL_SOLVE(rhs_expr, out)                                                                        \
    do {                                                                                               \
        rhs_expr;                                                                                      \
        int rc_ = cg_run(ctx, w, opt->cg_tol, opt->cg_max_iter, &it_cg, &err_cg, s);                   \
        ms_sol += w->ms_total; info->cg_iterations += it_cg;                                           \
        if (rc_ == KL_E_CUDA || rc_ == KL_E_ARG) return rc_;                                           \
        if (rc_ != KL_OK) AL_FAIL(3);                                                                  \
        KL_CUDA(cudaMemcpyAsync((out), w->x, vb, cudaMemcpyDeviceToDevice, s));                        \
    } while (0)
    auto dots = [&](double* d7) -> int {
        k_alm_dots<<<w->nb, CG_THREADS, 0, s>>>(w->aUt, w->aUbar, w->aDU, w->aDUold, n, w->part);
        k_alm_dots_final<<<1, CG_THREADS, 0, s>>>(w->part, w->nb, w->dots);
        KL_CUDA(cudaGetLastError());
        ctx->launches += 2;
        KL_CUDA(cudaMemcpyAsync(w->dots_host, w->dots, sizeof(double) * 7, cudaMemcpyDeviceToHost, s));
        KL_CUDA(cudaStreamSynchronize(s));
        for (int k = 0; k < 7; ++k) d7[k] = w->dots_host[k];
        return KL_OK;
    };
    auto lin2 = [&](double* out, const double* x, double a, const double* y, double b) {
        k_vec_lin2<<<w->nb, CG_THREADS, 0, s>>>(out, x, a, y, b, n);
        ctx->launches++;
    };

    // state in: m_U, m_L, m_DeltaUold, m_DeltaLold (setSolution / setPrevious, gsALMBase.h:185-194)
    double L = *L_io, DLold = *DLold_io;
    std::memcpy(ctx->h_pinned_x, U_host, vb);
    KL_CUDA(cudaMemcpyAsync(w->aU, ctx->h_pinned_x, vb, cudaMemcpyHostToDevice, s));
    KL_CUDA(cudaStreamSynchronize(s));
    if (DUold_host) {
        std::memcpy(ctx->h_pinned_x, DUold_host, vb);
        KL_CUDA(cudaMemcpyAsync(w->aDUold, ctx->h_pinned_x, vb, cudaMemcpyHostToDevice, s));
        KL_CUDA(cudaStreamSynchronize(s));
    } else {
        KL_CUDA(cudaMemsetAsync(w->aDUold, 0, vb, s));
    }
    double nF = 0, nU = 0, nDUold = 0;
    if ((rc = dev_norm(ctx, w, F, &nF, s))) return rc;
    if ((rc = dev_norm(ctx, w, w->aU, &nU, s))) return rc;
    if ((rc = dev_norm(ctx, w, w->aDUold, &nDUold, s))) return rc;
    const double FF = nF * nF;
    const bool phi_user = opt->phi >= 0.0;
    double phi = phi_user ? opt->phi : 0.0;
    const double relax = opt->relaxation != 0.0 ? opt->relaxation : 1.0;
    const int maxit = opt->max_it > 0 ? opt->max_it : 100;
    // initiateStep
    KL_CUDA(cudaMemsetAsync(w->aDU, 0, vb, s));
    KL_CUDA(cudaMemsetAsync(w->aUbar, 0, vb, s));
    KL_CUDA(cudaMemsetAsync(w->adU, 0, vb, s));
    double DL = 0.0, dL = 0.0, eta = 1.0, d7[7];
    // predictor (gsALMCrisfield.hpp:111-164)
    AL_ASM(kl_jacobian_device(ctx, w->aU, s));
    AL_SOLVE(KL_CUDA(cudaMemcpyAsync(w->b, F, vb, cudaMemcpyDeviceToDevice, s)), w->aUt);
    if ((rc = dots(d7))) return rc;
    if (nDUold * nDUold == 0.0 && DLold * DLold == 0.0) {       // no information about a previous step
        dL = arc_length / std::sqrt(2.0 * d7[0]);
        if (!phi_user) phi = std::sqrt(d7[0] / FF);
    } else {
        if (!phi_user) phi = std::sqrt(nU * nU / (L * L * FF));
        const double A0 = phi * phi * FF;                       // computeLambdaMU (:386-401)
        const int dir = sgn(d7[6] + A0 * DLold);
        const double denum = std::sqrt(d7[0] + A0);
        dL = dir * (denum == 0.0 ? arc_length : arc_length / denum);
    }
    lin2(w->adU, w->aUt, dL, nullptr, 0.0);                     // deltaU = deltaL * deltaUt (deltaUbar = 0 here)
    lin2(w->aDU, w->aDU, 1.0, w->adU, 1.0);
    DL += dL;
    const double A0 = phi * phi * FF;
    // residual and its reference norms (gsALMBase.hpp:151-171)
    lin2(w->aX, w->aU, 1.0, w->aDU, 1.0);
    AL_ASM(kl_al_residual_device(ctx, w->aX, L + DL, w->aR, s));
    double nR = 0, ndU = 0, nDU = 0;
    if ((rc = dev_norm(ctx, w, w->aR, &nR, s))) return rc;
    if ((rc = dev_norm(ctx, w, w->aDU, &nDU, s))) return rc;
    const double basisF = std::fabs(L + DL) * nF, basisU = nDU;
    info->residueF = nR / basisF; info->residueU = 1.0;
    info->status = 1;
    int it = 1;
    for (; it < maxit; ++it) {
        // quasiNewtonIteration: new tangent, deltaUt (:67-72)
        AL_ASM(kl_jacobian_device(ctx, w->aX, s));
        AL_SOLVE(KL_CUDA(cudaMemcpyAsync(w->b, F, vb, cudaMemcpyDeviceToDevice, s)), w->aUt);
        // iteration: deltaUbar = K^-1 (-R), constraint (:75-99)
        AL_SOLVE(lin2(w->b, w->aR, -1.0, nullptr, 0.0), w->aUbar);
        if ((rc = dots(d7))) return rc;
        eta = 1.0;
        const double lamold = dL;
        const double a0 = d7[0] + A0, b0 = 2.0 * (d7[1] + DL * A0), b1 = 2.0 * d7[2];
        const double c0 = d7[3] + DL * DL * A0 - arc_length * arc_length, c1 = 2.0 * d7[4], c2 = d7[5];
        double al1 = a0, al2 = b0 + eta * b1, al3 = c0 + eta * c1 + eta * eta * c2;
        double disc = al2 * al2 - 4.0 * al1 * al3;
        double dLs[2] = {0, 0};
        bool complex_root = false;
        if (disc >= 0.0) {
            dLs[0] = (-al2 + std::sqrt(disc)) / (2.0 * al1);
            dLs[1] = (-al2 - std::sqrt(disc)) / (2.0 * al1);
        } else {
            // computeLambdasModified (Lam 1992) -> eta; computeLambdasEta (Zhou 1995)
            const double m1 = b1 * b1 - 4.0 * a0 * c2, m2 = 2.0 * b0 * b1 - 4.0 * a0 * c1, m3 = b0 * b0 - 4.0 * a0 * c0;
            disc = m2 * m2 - 4.0 * m1 * m3;
            if (disc >= 0.0) {
                const double e1 = (-m2 + std::sqrt(disc)) / (2.0 * m1), e2 = (-m2 - std::sqrt(disc)) / (2.0 * m1);
                const double eta1 = std::min(e1, e2), eta2 = std::max(e1, e2), xi = 0.05 * std::fabs(eta2 - eta1);
                if (eta2 < 1.0) eta = eta2 - xi;
                else if (eta2 > 1.0 && -m2 / m1 < 1.0) eta = eta2 + xi;
                else if (eta1 < 1.0 && -m2 / m1 > 1.0) eta = eta1 - xi;
                else if (eta1 > 1.0) eta = eta1 + xi;
            }
            if (disc >= 0.0 && eta > 0.05) {
                al2 = b0 + eta * b1;
                dLs[0] = dLs[1] = -al2 / (2.0 * al1);
            } else {
                eta = 1.0;
                complex_root = true;
            }
        }
        if (!complex_root) {
            // computeLambdaDOT (Ritto-Correa 2008)
            const double t = d7[6] + phi * phi * DLold;
            const double DOT1 = dLs[0] * t, DOT2 = dLs[1] * t;
            dL = (DOT1 < DOT2) ? dLs[1] : dLs[0];
            lin2(w->adU, w->aUbar, eta, w->aUt, dL);
        } else {
            // computeLambdasComplex (Lam 1992, eq. 13-17): Fint = K (U + DeltaU), scaled back onto the constraint
            lin2(w->adU, w->aDU, 1.0, w->aUbar, 1.0);                      // DeltaUcr
            lin2(w->p, w->aU, 1.0, w->aDU, 1.0);
            if ((rc = launch_spmv<false>(ctx, w, w->p, w->tmp, nullptr, nullptr, s))) return rc;
            ctx->launches++;
            double nCr = 0, FintF = 0;
            // Fint.F via |Fint + F|^2 - |Fint|^2 - |F|^2
            double nA = 0, nB = 0;
            if ((rc = dev_norm(ctx, w, w->tmp, &nA, s))) return rc;
            lin2(w->tmp, w->tmp, 1.0, F, 1.0);
            if ((rc = dev_norm(ctx, w, w->tmp, &nB, s))) return rc;
            FintF = 0.5 * (nB * nB - nA * nA - FF);
            if ((rc = dev_norm(ctx, w, w->adU, &nCr, s))) return rc;
            const double DLcr = FintF / FF - L;
            const double mu = arc_length / std::sqrt(nCr * nCr + A0 * DLcr * DLcr);
            dL = mu * DLcr - DL;
            lin2(w->adU, w->adU, mu, w->aDU, -1.0);
        }
        // relaxation against an oscillating load factor
        if ((lamold * dL < 0) && (std::fabs(dL) <= std::fabs(lamold)) && relax != 1.0) {
            lin2(w->adU, w->aUt, relax * dL, w->aUbar, relax * eta);
            dL = relax * dL;
        }
        lin2(w->aDU, w->aDU, 1.0, w->adU, 1.0);
        DL += dL;
        lin2(w->aX, w->aU, 1.0, w->aDU, 1.0);
        AL_ASM(kl_al_residual_device(ctx, w->aX, L + DL, w->aR, s));
        if ((rc = dev_norm(ctx, w, w->aR, &nR, s))) return rc;
        if ((rc = dev_norm(ctx, w, w->adU, &ndU, s))) return rc;
        info->residueF = nR / basisF; info->residueU = ndU / basisU;
        if (info->residueF < opt->tolF && info->residueU < opt->tolU) { info->status = 0; break; }
    }
    info->iterations = it;
    info->phi = phi; info->DeltaL = DL;
    info->ms_assembly = ms_asm; info->ms_solve = ms_sol;
    if (info->status != 0) return KL_OK;                         // NotConverged: the caller's state is untouched (it bisects)
    // iterationFinish: U += DeltaU, L += DeltaL, DeltaUold = DeltaU (AngleMethod = step)
    KL_CUDA(cudaMemcpyAsync(ctx->h_pinned_r, w->aX, vb, cudaMemcpyDeviceToHost, s));
    KL_CUDA(cudaStreamSynchronize(s));
    std::memcpy(U_host, ctx->h_pinned_r, vb);
    if (DUold_host) {
        KL_CUDA(cudaMemcpyAsync(ctx->h_pinned_r, w->aDU, vb, cudaMemcpyDeviceToHost, s));
        KL_CUDA(cudaStreamSynchronize(s));
        std::memcpy(DUold_host, ctx->h_pinned_r, vb);
    }
    *L_io = L + DL;
    *DLold_io = DL;
    KL_CUDA(cudaGetLastError());
    return KL_OK;
#undef AL_FAIL
#undef AL_ASM
#undef AL_SOLVE
}
