// kl_stress.cu — stress / stretch recovery on the device:
//   kl_eval_stress          assembler->constructStress(mp_def, field, stress_type::X) + field evaluation
//                           (benchmarks/benchmark_Balloon.cpp:381-408, benchmark_TensionWrinkling.cpp:505-540,
//                            benchmark_Pillow.cpp:431,484-505)
//   kl_principal_stretches  assembler->computePrincipalStretches(pts, mp_def, z)   (unittests/gsStaticSolver_test.cpp:317)
//   kl_boundary_force       assembler->boundaryForce(mp_def, patchSide)            (unittests/gsStaticSolver_test.cpp:321)
// Definitions of the quantities: include/kl_shell.h.  One thread per evaluation point: it finds its knot spans by bisection,
// evaluates the 1-D basis functions and their first two derivatives (triangular Cox-de Boor table, as the host tables of
// kl_capi.cu) and gathers its (p+1)^2 control points; only the parametric coordinates are uploaded.  All in-plane
// tensors are handled in curvilinear components with closed-form 2x2 eigen-decompositions (the oracle goes through 3-D
// tensors instead, oracle/kl_oracle.c: klo_eval_stress).
#include <algorithm>
#include <cstring>
#include "kl_device.cuh"

struct StressPoint {
    int s1, s2;
    double b1[3][KL_MAXP + 1], b2[3][KL_MAXP + 1];
};

// knot span of u: the last non-empty span whose first knot is <= u (span[] lists the non-empty spans in ascending order)
__device__ __forceinline__ int find_span_dev(const double* __restrict__ U, const int* __restrict__ span, int nel, double u) {
    int lo = 0, hi = nel - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (U[span[mid]] <= u) lo = mid; else hi = mid - 1;
    }
    return span[lo];
}
// values, first and second derivatives of the p+1 functions that are non-zero on span k (same recurrence as
// bspline_span_ders in kl_capi.cu: T[m][q][j] = m-th derivative of the j-th function of degree q)
__device__ __forceinline__ void span_ders_dev(const double* __restrict__ U, int p, int k, double u, double out[3][KL_MAXP + 1]) {
    double T[3][KL_MAXP + 1][KL_MAXP + 1];
#pragma unroll
    for (int m = 0; m < 3; ++m)
#pragma unroll
        for (int q = 0; q <= KL_MAXP; ++q)
#pragma unroll
            for (int j = 0; j <= KL_MAXP; ++j) T[m][q][j] = 0.0;
    T[0][0][0] = 1.0;
#pragma unroll
    for (int q = 1; q <= KL_MAXP; ++q) {
        if (q > p) break;
#pragma unroll
        for (int j = 0; j <= KL_MAXP; ++j) {
            if (j > q) break;
            const int i = k - q + j;
            const double dl = U[i + q] - U[i], dr = U[i + q + 1] - U[i + 1];
#pragma unroll
            for (int m = 0; m <= 2; ++m) {
                double v = 0.0;
                if (m == 0) {
                    if (j >= 1) v += (u - U[i]) / dl * T[0][q - 1][j - 1];
                    if (j <= q - 1) v += (U[i + q + 1] - u) / dr * T[0][q - 1][j];
                } else {
                    if (j >= 1) v += T[m - 1][q - 1][j - 1] / dl;
                    if (j <= q - 1) v -= T[m - 1][q - 1][j] / dr;
                    v *= q;
                }
                T[m][q][j] = v;
            }
        }
    }
#pragma unroll
    for (int m = 0; m < 3; ++m)
#pragma unroll
        for (int j = 0; j <= KL_MAXP; ++j) out[m][j] = (j <= p) ? T[m][p < 1 ? 0 : p][j] : 0.0;
}

__host__ __device__ inline int stress_dim(int type) {
    switch (type) {
        case KL_STRESS_PRINCIPAL_STRETCH_DIR: return 9;
        case KL_STRESS_PRINCIPAL_STRESS_MEMBRANE: case KL_STRESS_PRINCIPAL_STRESS_FLEXURAL:
        case KL_STRESS_PRINCIPAL_MEMBRANE_STRAIN: case KL_STRESS_PRINCIPAL_FLEXURAL_STRAIN: return 2;
        case KL_STRESS_VON_MISES_MEMBRANE: case KL_STRESS_TENSION_FIELD: return 1;
    }
    return (type >= 0 && type < KL_STRESS_NTYPES) ? 3 : 0;
}
extern "C" int kl_stress_dim(int32_t type) { return stress_dim(type); }

// ascending eigenvalues of the symmetric 2x2 (s11, s22, s12)
__device__ __forceinline__ void eig2_values(const double s[3], double w[2]) {
    const double m = 0.5 * (s[0] + s[1]), r = hypot(0.5 * (s[0] - s[1]), s[2]);
    w[0] = m - r; w[1] = m + r;
}
// T S T^T for the change of frame (curvilinear contravariant components -> local Cartesian): T = [[t11, t12], [0, t22]]
__device__ __forceinline__ void push2(const double S[3], double t11, double t12, double t22, double scale, double out[3]) {
    out[0] = scale * (t11 * t11 * S[0] + 2.0 * t11 * t12 * S[2] + t12 * t12 * S[1]);
    out[1] = scale * (t22 * t22 * S[1]);
    out[2] = scale * (t22 * (t11 * S[2] + t12 * S[1]));
}

__global__ void __launch_bounds__(128) k_eval_stress(KLDev d, const double* __restrict__ uv, int npts, int type, double z,
                                                     double* __restrict__ out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= npts) return;
    const int p = d.p, dim = stress_dim(type);
    StressPoint sp;
    {
        const double u = uv[2 * k], v = uv[2 * k + 1];
        sp.s1 = find_span_dev(d.knots1, d.span1, d.nel1, u);
        sp.s2 = find_span_dev(d.knots2, d.span2, d.nel2, v);
        span_ders_dev(d.knots1, p, sp.s1, u, sp.b1);
        span_ders_dev(d.knots2, p, sp.s2, v, sp.b2);
    }
    double* res = out + (size_t)dim * k;
    // geometry: fo[m][c] (weighted if rational), fw[m], fu[m][c];  m = (val, d1, d2, d11, d22, d12)
    double fo[6][3] = {}, fu[6][3] = {}, fw[6] = {};
    for (int b = 0; b <= p; ++b)
        for (int a = 0; a <= p; ++a) {
            const int i = (sp.s1 - p + a) + d.n1 * (sp.s2 - p + b);
            const double R[6] = {sp.b1[0][a] * sp.b2[0][b], sp.b1[1][a] * sp.b2[0][b], sp.b1[0][a] * sp.b2[1][b],
                                 sp.b1[2][a] * sp.b2[0][b], sp.b1[0][a] * sp.b2[2][b], sp.b1[1][a] * sp.b2[1][b]};
            const double w = d.rational ? d.w[i] : 1.0;
#pragma unroll
            for (int m = 0; m < 6; ++m) {
                fw[m] += R[m] * w;
#pragma unroll
                for (int c = 0; c < 3; ++c) { fo[m][c] += R[m] * w * d.cp[3 * i + c]; fu[m][c] += R[m] * d.disp[3 * i + c]; }
            }
        }
    if (type == KL_STRESS_DISPLACEMENT) {
#pragma unroll
        for (int c = 0; c < 3; ++c) res[c] = fu[0][c];
        return;
    }
    double A1[3], A2[3], H[3][3], a1[3], a2[3], h[3][3];
    {
        const double iw = 1.0 / fw[0];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const double X = fo[0][c] * iw;
            A1[c] = (fo[1][c] - fw[1] * X) * iw;
            A2[c] = (fo[2][c] - fw[2] * X) * iw;
            H[0][c] = (fo[3][c] - fw[3] * X - 2.0 * fw[1] * A1[c]) * iw;
            H[1][c] = (fo[4][c] - fw[4] * X - 2.0 * fw[2] * A2[c]) * iw;
            H[2][c] = (fo[5][c] - fw[5] * X - fw[1] * A2[c] - fw[2] * A1[c]) * iw;
            a1[c] = A1[c] + fu[1][c]; a2[c] = A2[c] + fu[2][c];
            h[0][c] = H[0][c] + fu[3][c]; h[1][c] = H[1][c] + fu[4][c]; h[2][c] = H[2][c] + fu[5][c];
        }
    }
    double Nn[3], nn[3];
    cross3(A1, A2, Nn); cross3(a1, a2, nn);
    const double JA = sqrt(dot3(Nn, Nn)), Ja = sqrt(dot3(nn, nn));
    if (!(JA > 0.0) || !(Ja > 0.0) || !isfinite(Ja)) { atomicOr(d.flag, KLF_JACOBIAN); return; }
#pragma unroll
    for (int c = 0; c < 3; ++c) { Nn[c] /= JA; nn[c] /= Ja; }
    const double Ac[3] = {dot3(A1, A1), dot3(A2, A2), dot3(A1, A2)}, ac[3] = {dot3(a1, a1), dot3(a2, a2), dot3(a1, a2)};
    double Bc[3], bc[3];
#pragma unroll
    for (int m = 0; m < 3; ++m) { Bc[m] = d.mat.bending ? dot3(H[m], Nn) : 0.0; bc[m] = d.mat.bending ? dot3(h[m], nn) : 0.0; }
    double Ai[3], ai[3], dA, da;
    inv2s(Ac, Ai, dA); inv2s(ac, ai, da);
    const bool comp = d.mat.material != KL_MAT_SVK && d.mat.compressible;

    if (type == KL_STRESS_PRINCIPAL_STRETCH || type == KL_STRESS_PRINCIPAL_STRETCH_DIR) {
        // metric at height z exactly as the material law sees it
        double Gc[3], gc[3];
#pragma unroll
        for (int v = 0; v < 3; ++v) { Gc[v] = Ac[v] - 2.0 * z * Bc[v]; gc[v] = ac[v] - 2.0 * z * bc[v]; }
        if (d.mat.metric_z2) {
            const double Bm[2][2] = {{Bc[0], Bc[2]}, {Bc[2], Bc[1]}}, bm[2][2] = {{bc[0], bc[2]}, {bc[2], bc[1]}};
            const double AI[2][2] = {{Ai[0], Ai[2]}, {Ai[2], Ai[1]}}, aI[2][2] = {{ai[0], ai[2]}, {ai[2], ai[1]}};
            const int vi[3] = {0, 1, 0}, vj[3] = {0, 1, 1};
            for (int v = 0; v < 3; ++v)
                for (int c = 0; c < 2; ++c)
                    for (int e = 0; e < 2; ++e) {
                        Gc[v] += z * z * Bm[vi[v]][c] * AI[c][e] * Bm[e][vj[v]];
                        gc[v] += z * z * bm[vi[v]][c] * aI[c][e] * bm[e][vj[v]];
                    }
        }
        double Gi[3], gi[3], dG, dg;
        inv2s(Gc, Gi, dG); inv2s(gc, gi, dg);
        if (!(dG > 0.0) || !(dg > 0.0)) { atomicOr(d.flag, KLF_METRIC); return; }
        const double J0sq = dg / dG;
        // generalised eigenproblem g v = lambda^2 G v: trace and determinant of G^-1 g
        const double trs = gc[0] * Gi[0] + gc[1] * Gi[1] + 2.0 * gc[2] * Gi[2];
        // (difference of the roots)^2 written as a sum of squares: no cancellation for nearly equal stretches
        const double m11 = Gi[0] * gc[0] + Gi[2] * gc[2], m12 = Gi[0] * gc[2] + Gi[2] * gc[1];
        const double m21 = Gi[2] * gc[0] + Gi[1] * gc[2], m22 = Gi[2] * gc[2] + Gi[1] * gc[1];
        const double disc = sqrt(fmax((m11 - m22) * (m11 - m22) + 4.0 * m12 * m21, 0.0));
        const double l1 = 0.5 * (trs - disc), l2 = 0.5 * (trs + disc);
        double c33 = 1.0 / J0sq;
        if (comp) {
            double S[3], C[6];
            if (!hyper_comp<false>(d.mat, Gi, gc, gi, J0sq, S, C, &c33)) { atomicOr(d.flag, KLF_C33); return; }
        }
        if (type == KL_STRESS_PRINCIPAL_STRETCH) { res[0] = sqrt(l1); res[1] = sqrt(l2); res[2] = sqrt(c33); return; }
        // eigenvector of the smaller stretch from the better conditioned row of (g - l1 G), the other one G-orthogonal to it
        double r1[2] = {gc[0] - l1 * Gc[0], gc[2] - l1 * Gc[2]}, r2[2] = {gc[2] - l1 * Gc[2], gc[1] - l1 * Gc[1]};
        double v[2];
        if (r1[0] * r1[0] + r1[1] * r1[1] >= r2[0] * r2[0] + r2[1] * r2[1]) { v[0] = -r1[1]; v[1] = r1[0]; }
        else { v[0] = -r2[1]; v[1] = r2[0]; }
        if (!(disc > 1e-14 * trs)) { v[0] = 1.0; v[1] = 0.0; }   // equal stretches: any direction is principal
        const double nv = sqrt(v[0] * v[0] * Gc[0] + 2.0 * v[0] * v[1] * Gc[2] + v[1] * v[1] * Gc[1]);
        v[0] /= nv; v[1] /= nv;
        const double wv[2] = {Gc[0] * v[0] + Gc[2] * v[1], Gc[2] * v[0] + Gc[1] * v[1]};
        const double v2[2] = {-wv[1] / sqrt(dG), wv[0] / sqrt(dG)};
        // g_a(z) = a_a - z b_a^c a_c
        const double bm00 = bc[0] * ai[0] + bc[2] * ai[2], bm01 = bc[0] * ai[2] + bc[2] * ai[1];
        const double bm10 = bc[2] * ai[0] + bc[1] * ai[2], bm11 = bc[2] * ai[2] + bc[1] * ai[1];
        double d1[3], d2[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const double g1 = a1[c] - z * (bm00 * a1[c] + bm01 * a2[c]), g2 = a2[c] - z * (bm10 * a1[c] + bm11 * a2[c]);
            d1[c] = v[0] * g1 + v[1] * g2;
            d2[c] = v2[0] * g1 + v2[1] * g2;
        }
        const double n1 = sqrt(dot3(d1, d1)), n2 = sqrt(dot3(d2, d2));
#pragma unroll
        for (int c = 0; c < 3; ++c) { res[c] = d1[c] / n1; res[3 + c] = d2[c] / n2; res[6 + c] = nn[c]; }
        return;
    }

    // strains in the local Cartesian frame of the undeformed surface (E1 = A_1/|A_1|, E2 = A^2/|A^2|): covariant components
    // transform with Q = [[A^1.E1, A^2.E1], [A^1.E2, A^2.E2]] = [[1/|A_1|, 0], [A^12/sqrt(A^22), sqrt(A^22)]]
    double Em[3], Ef[3];
    {
        const double q11 = 1.0 / sqrt(Ac[0]), q21 = Ai[2] / sqrt(Ai[1]), q22 = sqrt(Ai[1]);
        const double e[3] = {0.5 * (ac[0] - Ac[0]), 0.5 * (ac[1] - Ac[1]), 0.5 * (ac[2] - Ac[2])};
        const double kp[3] = {Bc[0] - bc[0], Bc[1] - bc[1], Bc[2] - bc[2]};
        Em[0] = q11 * q11 * e[0];
        Em[1] = q21 * q21 * e[0] + 2.0 * q21 * q22 * e[2] + q22 * q22 * e[1];
        Em[2] = q11 * (q21 * e[0] + q22 * e[2]);
        Ef[0] = q11 * q11 * kp[0];
        Ef[1] = q21 * q21 * kp[0] + 2.0 * q21 * q22 * kp[2] + q22 * q22 * kp[1];
        Ef[2] = q11 * (q21 * kp[0] + q22 * kp[2]);
    }
    double A[6], B[6], D[6], N[3], M[3];
    const int flag = material_point<false>(d.mat, Ac, Bc, ac, bc, A, B, D, N, M);
    if (flag) { atomicOr(d.flag, flag); return; }
    if (!d.mat.bending) M[0] = M[1] = M[2] = 0.0;
    // Cauchy stress in the local Cartesian frame of the deformed surface: sigma^ab = S^ab / J on the basis a_a,
    // T = [[a_1.e1, a_2.e1], [a_1.e2, a_2.e2]] = [[|a_1|, a_12/|a_1|], [0, 1/|a^2|]]
    double sm[3], sf[3];
    {
        const double J0sq = da / dA;
        double c33 = 1.0 / J0sq;
        if (comp) {
            double S[3], C[6];
            if (!hyper_comp<false>(d.mat, Ai, ac, ai, J0sq, S, C, &c33)) { atomicOr(d.flag, KLF_C33); return; }
        }
        const double J = sqrt(J0sq * c33);
        const double t11 = sqrt(ac[0]), t12 = ac[2] / t11, t22 = 1.0 / sqrt(ai[1]);
        const double t = d.mat.t;
        push2(N, t11, t12, t22, 1.0 / (t * J), sm);
        push2(M, t11, t12, t22, 6.0 / (t * t * J), sf);
    }
    double w[2];
    switch (type) {
        case KL_STRESS_MEMBRANE_FORCE: res[0] = N[0]; res[1] = N[1]; res[2] = N[2]; break;
        case KL_STRESS_FLEXURAL_MOMENT: res[0] = M[0]; res[1] = M[1]; res[2] = M[2]; break;
        case KL_STRESS_MEMBRANE: res[0] = sm[0]; res[1] = sm[1]; res[2] = sm[2]; break;
        case KL_STRESS_FLEXURAL: res[0] = sf[0]; res[1] = sf[1]; res[2] = sf[2]; break;
        case KL_STRESS_MEMBRANE_STRAIN: res[0] = Em[0]; res[1] = Em[1]; res[2] = Em[2]; break;
        case KL_STRESS_FLEXURAL_STRAIN: res[0] = Ef[0]; res[1] = Ef[1]; res[2] = Ef[2]; break;
        case KL_STRESS_PRINCIPAL_STRESS_MEMBRANE: eig2_values(sm, w); res[0] = w[0]; res[1] = w[1]; break;
        case KL_STRESS_PRINCIPAL_STRESS_FLEXURAL: eig2_values(sf, w); res[0] = w[0]; res[1] = w[1]; break;
        case KL_STRESS_PRINCIPAL_MEMBRANE_STRAIN: eig2_values(Em, w); res[0] = w[0]; res[1] = w[1]; break;
        case KL_STRESS_PRINCIPAL_FLEXURAL_STRAIN: eig2_values(Ef, w); res[0] = w[0]; res[1] = w[1]; break;
        case KL_STRESS_VON_MISES_MEMBRANE: res[0] = sqrt(sm[0] * sm[0] + sm[1] * sm[1] - sm[0] * sm[1] + 3.0 * sm[2] * sm[2]); break;
        case KL_STRESS_TENSION_FIELD: {
            double we[2];
            eig2_values(sm, w); eig2_values(Em, we);
            res[0] = w[0] > 0.0 ? 1.0 : (we[1] <= 0.0 ? -1.0 : 0.0);
        } break;
    }
}

// minus the sum of the full internal force over the control points of one side: one block, fixed order (deterministic)
__global__ void __launch_bounds__(256) k_side_sum(const double* __restrict__ ffull, int ncp, int first, int stride, int count, double* __restrict__ out3) {
    __shared__ double sh[256];
    for (int c = 0; c < 3; ++c) {
        double s = 0.0;
        for (int i = threadIdx.x; i < count; i += 256) s += ffull[c * ncp + first + stride * i];
        sh[threadIdx.x] = s;
        __syncthreads();
        for (int w = 128; w > 0; w >>= 1) { if (threadIdx.x < w) sh[threadIdx.x] += sh[threadIdx.x + w]; __syncthreads(); }
        if (threadIdx.x == 0) out3[c] = -sh[0];   // sign of rhs() = F_ext - F_int
        __syncthreads();
    }
}

static int upload_state(kl_ctx* ctx, const double* x_host, cudaStream_t s) {
    const double* xd = nullptr;
    if (x_host) {
        std::memcpy(ctx->h_pinned_x, x_host, sizeof(double) * ctx->d.nfree);
        KL_CUDA(cudaMemcpyAsync(ctx->d_x, ctx->h_pinned_x, sizeof(double) * ctx->d.nfree, cudaMemcpyHostToDevice, s));
        xd = ctx->d_x;
    }
    return kl_launch_construct(ctx, xd, s);
}

extern "C" int kl_eval_stress(kl_ctx* ctx, const double* x_host, int32_t type, int32_t n_pts, const double* uv_host, double z,
                              double* out_host) {
    if (!ctx || n_pts < 0 || (n_pts > 0 && (!uv_host || !out_host))) { kl_set_error("kl_eval_stress: bad argument"); return KL_E_ARG; }
    if (ctx->mp) { kl_set_error("kl_eval_stress: fields are evaluated per patch: pass kl_mp_patch(mp, q), not the matrix context"); return KL_E_ARG; }
    const int dim = stress_dim(type);
    if (!dim) { kl_set_error("kl_eval_stress: unknown stress type"); return KL_E_ARG; }
    if (n_pts == 0) return KL_OK;
    KL_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    // only the domain check happens on the host; spans and basis functions are evaluated by the kernel
    for (int k = 0; k < n_pts; ++k)
        for (int dir = 0; dir < 2; ++dir) {
            const double u = uv_host[2 * k + dir];
            if (!(u >= ctx->U[dir].front() && u <= ctx->U[dir].back())) { kl_set_error("kl_eval_stress: point outside the parametric domain"); return KL_E_ARG; }
        }
    double* d_pts = nullptr;
    double* d_out = nullptr;
    KL_CUDA(cudaMalloc((void**)&d_pts, sizeof(double) * 2 * (size_t)n_pts));
    if (cudaMalloc((void**)&d_out, sizeof(double) * (size_t)n_pts * dim) != cudaSuccess) { cudaFree(d_pts); kl_set_error("kl_eval_stress: cudaMalloc"); return KL_E_CUDA; }
    int rc = KL_OK;
    do {
        if (cudaMemcpyAsync(d_pts, uv_host, sizeof(double) * 2 * (size_t)n_pts, cudaMemcpyHostToDevice, s) != cudaSuccess) { rc = KL_E_CUDA; break; }
        if ((rc = upload_state(ctx, x_host, s))) break;
        k_eval_stress<<<(n_pts + 127) / 128, 128, 0, s>>>(ctx->d, d_pts, n_pts, type, z, d_out);
        ctx->launches++;
        if (cudaGetLastError() != cudaSuccess) { rc = KL_E_CUDA; break; }
        if (cudaMemcpyAsync(out_host, d_out, sizeof(double) * (size_t)n_pts * dim, cudaMemcpyDeviceToHost, s) != cudaSuccess) { rc = KL_E_CUDA; break; }
        rc = kl_check(ctx, s);
    } while (0);
    if (rc == KL_E_CUDA) kl_set_error(std::string("kl_eval_stress: ") + cudaGetErrorString(cudaGetLastError()));
    cudaStreamSynchronize(s);
    cudaFree(d_pts);
    cudaFree(d_out);
    return rc;
}

extern "C" int kl_principal_stretches(kl_ctx* ctx, const double* x_host, int32_t n_pts, const double* uv_host, double z, double* out_host) {
    return kl_eval_stress(ctx, x_host, KL_STRESS_PRINCIPAL_STRETCH, n_pts, uv_host, z, out_host);
}

extern "C" int kl_boundary_force(kl_ctx* ctx, const double* x_host, int32_t side, double* out3_host) {
    if (!ctx || !out3_host || side < 0 || side > 3) { kl_set_error("kl_boundary_force: bad argument"); return KL_E_ARG; }
    if (ctx->mp) { kl_set_error("kl_boundary_force: patch sides belong to a patch: pass kl_mp_patch(mp, q), not the matrix context"); return KL_E_ARG; }
    KL_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const int ncp = ctx->d.ncp, n1 = ctx->d.n1, n2 = ctx->d.n2;
    double* d_full = nullptr;
    KL_CUDA(cudaMalloc((void**)&d_full, sizeof(double) * (3 * (size_t)ncp + 3)));
    int rc = KL_OK;
    do {
        if (cudaMemsetAsync(d_full, 0, sizeof(double) * (3 * (size_t)ncp + 3), s) != cudaSuccess) { rc = KL_E_CUDA; break; }
        if ((rc = upload_state(ctx, x_host, s))) break;
        // the whole patch contributes, whatever strip this context assembles
        const int b = ctx->e2_begin, e = ctx->e2_end;
        ctx->e2_begin = 0; ctx->e2_end = ctx->d.nel2;
        rc = kl_launch_residual(ctx, d_full, s, true);
        ctx->e2_begin = b; ctx->e2_end = e;
        if (rc) break;
        int first, stride, count;
        if (side == KL_WEST || side == KL_EAST) { first = side == KL_WEST ? 0 : n1 - 1; stride = n1; count = n2; }
        else { first = side == KL_SOUTH ? 0 : n1 * (n2 - 1); stride = 1; count = n1; }
        k_side_sum<<<1, 256, 0, s>>>(d_full, ncp, first, stride, count, d_full + 3 * (size_t)ncp);
        ctx->launches++;
        if (cudaGetLastError() != cudaSuccess) { rc = KL_E_CUDA; break; }
        if (cudaMemcpyAsync(out3_host, d_full + 3 * (size_t)ncp, 3 * sizeof(double), cudaMemcpyDeviceToHost, s) != cudaSuccess) { rc = KL_E_CUDA; break; }
        rc = kl_check(ctx, s);
    } while (0);
    if (rc == KL_E_CUDA) kl_set_error(std::string("kl_boundary_force: ") + cudaGetErrorString(cudaGetLastError()));
    cudaStreamSynchronize(s);
    cudaFree(d_full);
    return rc;
}
