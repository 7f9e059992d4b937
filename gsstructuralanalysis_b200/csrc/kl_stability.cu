// kl_stability.cu — stability indicator of the arc-length solvers on the device.
//
// Reference: gsALMBase<T>::_computeStability (src/gsALMSolvers/gsALMBase.hpp:546-611) and gsStaticBase<T>::_computeStabilityDet
// (src/gsStaticSolvers/gsStaticBase.h:161-179), bifurcation method "Determinant": factorise the tangent with
// gsSparseSolver<>::SimplicialLDLT (Eigen 3.4, third-party, vendored by G+Smo as gsEigen; not in /root/reference), take
// m_stabilityVec = vectorD(), m_negatives = countNegatives(vectorD), m_indicator = min(vectorD), stability = sign(indicator).
//
// Here: K = L D L^T without pivoting (as SimplicialLDLT: a fill-reducing permutation, no numerical pivoting) of the matrix the
// last kl_jacobian* call left on the device.  The fill-reducing permutation of a tensor-product patch is the node-major ordering
// along the shorter direction, which makes K a band matrix of half-width 3 (p n_short + p + 1) - 1; the band is factorised by a
// blocked right-looking algorithm in LAPACK band storage (AB[(i - j) + j ldab] = K_ij, sub-blocks addressed with leading
// dimension ldab - 1 as dpbtrf does): per block column one CTA factorises the NB x NB diagonal block, one kernel solves the
// panel below it (W = A21 L11^-T, L21 = W D^-1), one kernel applies the symmetric rank-NB update C -= L21 W^T to the lower
// triangle of the trailing window.  By Sylvester's law of inertia the number of negative pivots and the sign of the smallest
// pivot do not depend on the permutation, so stability() / stabilityChange() agree with the reference; the VALUE of the
// indicator is the smallest pivot of this ordering (Eigen's depends on its AMD ordering in the same way).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <numeric>
#include "kl_internal.h"

namespace {

constexpr int NB = 32;       // block size of the factorisation
constexpr int TU = 64;       // tile of the trailing update

// scatter the lower triangle of the permuted matrix into band storage
__global__ void k_band_fill(const int* __restrict__ outer, const int* __restrict__ inner, const double* __restrict__ val, const int* __restrict__ perm,
                            int n, long long ldab, double* __restrict__ AB) {
    const int lane = threadIdx.x & 31;
    const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (j >= n) return;
    const int pj = perm[j];
    for (int k = outer[j] + lane; k < outer[j + 1]; k += 32) {
        const int pi = perm[inner[k]];
        if (pi >= pj) AB[(long long)(pi - pj) + (long long)pj * ldab] = val[k];
    }
}
__global__ void k_band_width(const int* __restrict__ outer, const int* __restrict__ inner, const int* __restrict__ perm, int n, int* __restrict__ bw) {
    const int lane = threadIdx.x & 31;
    const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (j >= n) return;
    const int pj = perm[j];
    int m = 0;
    for (int k = outer[j] + lane; k < outer[j + 1]; k += 32) m = max(m, abs(perm[inner[k]] - pj));
    for (int o = 16; o; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) atomicMax(bw, m);
}

// (a) unblocked LDL^T of the nb x nb diagonal block, one CTA of NB*NB/4 threads... kept simple: NB x NB threads' worth of work
//     done by 256 threads with the block in shared memory
__global__ void __launch_bounds__(256) k_ldl_diag(double* __restrict__ A, long long ld, int nb, double* __restrict__ dvec) {
    __shared__ double s[NB][NB + 1];
    const int tid = threadIdx.x;
    for (int e = tid; e < nb * nb; e += 256) {
        const int a = e % nb, b = e / nb;
        s[a][b] = a >= b ? A[a + (long long)b * ld] : 0.0;
    }
    __syncthreads();
    for (int k = 0; k < nb; ++k) {
        const double d = s[k][k];
        // column k of L
        for (int a = k + 1 + tid; a < nb; a += 256) s[a][k] = s[a][k] / d;
        __syncthreads();
        // trailing update: s[a][b] -= l_ak d l_bk for a >= b > k
        const int r = nb - k - 1;
        for (int e = tid; e < r * r; e += 256) {
            const int a = k + 1 + e % r, b = k + 1 + e / r;
            if (a >= b) s[a][b] -= s[a][k] * d * s[b][k];
        }
        __syncthreads();
    }
    for (int e = tid; e < nb * nb; e += 256) {
        const int a = e % nb, b = e / nb;
        if (a >= b) A[a + (long long)b * ld] = s[a][b];
    }
    if (tid < nb) dvec[tid] = s[tid][tid];
}

// (b) panel: row a of A21 (m x nb):  w = a21 L11^-T (forward substitution), l = w / d; L21 overwrites A21, W goes to Wbuf (ld = m)
__global__ void __launch_bounds__(128) k_ldl_panel(const double* __restrict__ A11, double* __restrict__ A21, long long ld, int m, int nb,
                                                   double* __restrict__ Wbuf) {
    __shared__ double L[NB][NB + 1];
    __shared__ double dinv[NB];
    for (int e = threadIdx.x; e < nb * nb; e += blockDim.x) {
        const int a = e % nb, b = e / nb;
        L[a][b] = a > b ? A11[a + (long long)b * ld] : 0.0;
    }
    if (threadIdx.x < nb) dinv[threadIdx.x] = 1.0 / A11[threadIdx.x + (long long)threadIdx.x * ld];
    __syncthreads();
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= m) return;
    double w[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) w[b] = b < nb ? A21[a + (long long)b * ld] : 0.0;
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        double v = w[b];
#pragma unroll
        for (int c = 0; c < b; ++c) v = fma(-w[c], L[b][c], v);
        w[b] = v;
    }
#pragma unroll
    for (int b = 0; b < NB; ++b)
        if (b < nb) {
            Wbuf[a + (size_t)b * m] = w[b];
            A21[a + (long long)b * ld] = w[b] * dinv[b];
        }
}

// (c) C[a][b] -= sum_c L21[a][c] W[b][c] on the lower triangle (tiles with tile row >= tile column), TU x TU outputs per CTA,
//     256 threads x (4 x 4) outputs; L21 is read from the band (leading dimension ld), W from Wbuf (leading dimension m)
__global__ void __launch_bounds__(256) k_ldl_update(double* __restrict__ C, const double* __restrict__ L21, long long ld, const double* __restrict__ Wbuf,
                                                    int m, int nb) {
    // linear tile index -> (ti >= tj)
    int t = blockIdx.x;
    int ti = (int)((sqrt(8.0 * t + 1.0) - 1.0) * 0.5);
    while ((long long)(ti + 1) * (ti + 2) / 2 <= t) ++ti;
    while ((long long)ti * (ti + 1) / 2 > t) --ti;
    const int tj = t - ti * (ti + 1) / 2;
    __shared__ double sL[NB][TU + 1];      // [c][a]
    __shared__ double sW[NB][TU + 1];      // [c][b]
    const int a0 = ti * TU, b0 = tj * TU;
    for (int e = threadIdx.x; e < TU * NB; e += 256) {
        const int r = e % TU, c = e / TU;
        sL[c][r] = (c < nb && a0 + r < m) ? L21[(a0 + r) + (long long)c * ld] : 0.0;
        sW[c][r] = (c < nb && b0 + r < m) ? Wbuf[(b0 + r) + (size_t)c * m] : 0.0;
    }
    __syncthreads();
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;     // a = a0 + tx + 16 i, b = b0 + ty + 16 j
    double acc[4][4] = {};
#pragma unroll 8
    for (int c = 0; c < NB; ++c) {
        double la[4], wb[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { la[i] = sL[c][tx + 16 * i]; wb[i] = sW[c][ty + 16 * i]; }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fma(la[i], wb[j], acc[i][j]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int b = b0 + ty + 16 * j;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int a = a0 + tx + 16 * i;
            if (a < m && b < m && a >= b) C[a + (long long)b * ld] -= acc[i][j];
        }
    }
}

}  // namespace

// node-major ordering along the shorter direction: perm[g] = position of free DoF g; matched DoFs take their first position
static int band_ordering(const kl_ctx* ctx, std::vector<int>& perm) {
    const KLDev& d = ctx->d;
    const int n = d.nfree;
    std::vector<long long> key((size_t)n, -1);
    const bool row_major = d.n1 <= d.n2;           // i1 fastest when the first direction is the shorter one
    for (int c = 0; c < 3; ++c)
        for (int i = 0; i < d.ncp; ++i) {
            const int g = ctx->h_map[(size_t)c * d.ncp + i];
            if (g >= n) continue;
            const int i1 = i % d.n1, i2 = i / d.n1;
            const long long node = row_major ? (long long)i2 * d.n1 + i1 : (long long)i1 * d.n2 + i2;
            const long long k = node * 3 + c;
            if (key[g] < 0 || k < key[g]) key[g] = k;
        }
    std::vector<int> idx((size_t)n);
    std::iota(idx.begin(), idx.end(), 0);
    std::sort(idx.begin(), idx.end(), [&](int a, int b) { return key[a] < key[b] || (key[a] == key[b] && a < b); });
    perm.assign((size_t)n, 0);
    for (int k = 0; k < n; ++k) perm[idx[k]] = k;
    return 0;
}

extern "C" int kl_stability(kl_ctx* ctx, double* indicator, int32_t* negatives, double* vectorD_host) {
    if (!ctx) { kl_set_error("kl_stability: null context"); return KL_E_ARG; }
    if (ctx->mp || ctx->mp_member) { kl_set_error("kl_stability: the band ordering is built for one tensor-product patch; not available on a multi-patch"); return KL_E_ARG; }
    if (ctx->d.mat.pressure != 0.0) { kl_set_error("kl_stability: LDL^T needs a symmetric matrix (follower pressure makes the tangent unsymmetric)"); return KL_E_ARG; }
    KL_CUDA(cudaSetDevice(ctx->device));
    const KLDev& d = ctx->d;
    const int n = d.nfree;
    if (n <= 0) { if (indicator) *indicator = 0; if (negatives) *negatives = 0; return KL_OK; }
    cudaStream_t s = ctx->stream;
    KL_CUDA(cudaStreamSynchronize(s));
    std::vector<int> perm;
    band_ordering(ctx, perm);
    int *d_perm = nullptr, *d_bw = nullptr;
    double *AB = nullptr, *Wbuf = nullptr, *dvec = nullptr;
    int rc = KL_OK;
    auto cleanup = [&]() { cudaFree(d_perm); cudaFree(d_bw); cudaFree(AB); cudaFree(Wbuf); cudaFree(dvec); };
#define ST_CUDA(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { kl_set_error(std::string("kl_stability: " #call ": ") + cudaGetErrorString(e_)); cleanup(); return KL_E_CUDA; } } while (0)
    ST_CUDA(cudaMalloc((void**)&d_perm, sizeof(int) * (size_t)n));
    ST_CUDA(cudaMalloc((void**)&d_bw, sizeof(int)));
    ST_CUDA(cudaMemcpyAsync(d_perm, perm.data(), sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, s));
    ST_CUDA(cudaMemsetAsync(d_bw, 0, sizeof(int), s));
    k_band_width<<<(n + 7) / 8, 256, 0, s>>>(d.outer, d.inner, d_perm, n, d_bw);
    int bw = 0;
    ST_CUDA(cudaMemcpyAsync(&bw, d_bw, sizeof(int), cudaMemcpyDeviceToHost, s));
    ST_CUDA(cudaStreamSynchronize(s));
    const long long ldab = (long long)bw + NB + 1;
    const size_t ab_bytes = sizeof(double) * (size_t)ldab * (size_t)n;
    size_t free_b = 0, total_b = 0;
    ST_CUDA(cudaMemGetInfo(&free_b, &total_b));
    if (ab_bytes + (size_t)(bw + NB) * NB * 8 + (size_t)n * 8 + (64u << 20) > free_b) {
        kl_set_error("kl_stability: the band factor (" + std::to_string(ab_bytes >> 20) + " MiB, half-width " + std::to_string(bw) + ") does not fit the free device memory");
        cleanup();
        return KL_E_ARG;
    }
    ST_CUDA(cudaMalloc((void**)&AB, ab_bytes));
    ST_CUDA(cudaMalloc((void**)&Wbuf, sizeof(double) * (size_t)(bw + NB) * NB));
    ST_CUDA(cudaMalloc((void**)&dvec, sizeof(double) * (size_t)n));
    ST_CUDA(cudaMemsetAsync(AB, 0, ab_bytes, s));
    k_band_fill<<<(n + 7) / 8, 256, 0, s>>>(d.outer, d.inner, d.values, d_perm, n, ldab, AB);
    ctx->launches += 2;
    const long long ld = ldab - 1;
    for (int k0 = 0; k0 < n; k0 += NB) {
        const int nb = std::min(NB, n - k0);
        double* A11 = AB + (long long)k0 * ldab;
        k_ldl_diag<<<1, 256, 0, s>>>(A11, ld, nb, dvec + k0);
        const int k1 = k0 + nb;
        const int m = std::min(bw, n - k1);       // rows below the diagonal block that the block column reaches
        ctx->launches++;
        if (m <= 0) continue;
        double* A21 = AB + (long long)nb + (long long)k0 * ldab;        // element (k1, k0)
        k_ldl_panel<<<(m + 127) / 128, 128, 0, s>>>(A11, A21, ld, m, nb, Wbuf);
        double* C = AB + (long long)k1 * ldab;                          // element (k1, k1)
        const int nt = (m + TU - 1) / TU;
        k_ldl_update<<<(unsigned)((long long)nt * (nt + 1) / 2), 256, 0, s>>>(C, A21, ld, Wbuf, m, nb);
        ctx->launches += 2;
    }
    ST_CUDA(cudaGetLastError());
    std::vector<double> D((size_t)n);
    ST_CUDA(cudaMemcpyAsync(D.data(), dvec, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, s));
    ST_CUDA(cudaStreamSynchronize(s));
#undef ST_CUDA
    cleanup();
    double mn = D[0];
    int neg = 0;
    bool finite = true;
    for (int k = 0; k < n; ++k) {
        finite &= std::isfinite(D[k]);
        if (D[k] < 0.0) ++neg;
        if (D[k] < mn) mn = D[k];
    }
    if (!finite) { kl_set_error("kl_stability: zero pivot / non-finite value in the LDL^T factorisation (singular leading block or unassembled matrix)"); rc = KL_E_NONFINITE; }
    if (indicator) *indicator = mn;
    if (negatives) *negatives = neg;
    if (vectorD_host)      // in the ORIGINAL DoF order: D of the pivot that eliminated DoF g
        for (int g = 0; g < n; ++g) vectorD_host[g] = D[perm[g]];
    return rc;
}
