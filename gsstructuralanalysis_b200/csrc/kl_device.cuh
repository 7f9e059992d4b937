// kl_device.cuh — per-quadrature-point device code: tensor-product B-spline evaluation from
// shared-memory staged 1-D tables, KL kinematics, material laws with through-thickness
// integration.  Replaces (per point) what gsExprEvaluator + gsMaterialMatrixIntegrate do
// inside gsThinShellAssembler::assembleMatrix / assembleVector
// (reference call sites: tutorials/nonlinear_shell_static.cpp:124,133; formulation:
//  benchmarks/benchmark_cylinder_DC.cpp:536-555).
#pragma once
#include "kl_internal.h"

// symmetric Voigt 3x3 storage: (0,0)=0 (1,1)=1 (2,2)=2 (0,1)=3 (0,2)=4 (1,2)=5
__device__ __forceinline__ int sidx(int v, int u) {
    return v == u ? v : (v + u == 1 ? 3 : (v + u == 2 ? 4 : 5));
}
__device__ __forceinline__ double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
__device__ __forceinline__ void cross3(const double* a, const double* b, double* c) {
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
}

// Per-element staging area in shared memory
template <int P>
struct ElemStage {
    static constexpr int NQ = P + 1;
    static constexpr int NLOC = (P + 1) * (P + 1);
    double b1[NQ][3][P + 1];   // 1-D values / derivatives of direction 1 at the element's nodes
    double b2[NQ][3][P + 1];
    double w1[NQ], w2[NQ];
    double X[NLOC][3];         // undeformed control points (times weight if rational)
    double U[NLOC][3];         // displacement control points
    double Wt[NLOC];           // weights (rational geometry only)
    double scratch[2 * P * (P + 1)];   // with w1..Wt (dead after the point evaluation) 9 (P+1)^2 doubles: the partial sums of the
                                       // sum-factorised internal force (k_points<P, true>)
};

// cooperative load of one element's staging data by (P+1)^2 threads with lane id `t` (one control point per thread).
// All global loads are issued before the first shared-memory store, so the thread pays one memory latency instead of one
// per loop trip (the staging prologue was 41 % of the stall samples of k_points with 12 resident warps per SM).
template <int P>
__device__ __forceinline__ void stage_element(const KLDev& d, int e1, int e2, ElemStage<P>& E, int t, int /*nthr = (P+1)^2*/) {
    constexpr int NQ = P + 1, NLOC = (P + 1) * (P + 1), NB = NQ * 3 * (P + 1), NT = NLOC, RB = (NB + NT - 1) / NT;
    const double* g1 = d.bas1 + (size_t)e1 * NB;
    const double* g2 = d.bas2 + (size_t)e2 * NB;
    const int i0 = __ldg(&d.span1[e1]) - P, j0 = __ldg(&d.span2[e2]) - P;
    double vb1[RB], vb2[RB];
#pragma unroll
    for (int r = 0; r < RB; ++r) {
        const int k = t + r * NT;
        vb1[r] = k < NB ? __ldg(g1 + k) : 0.0;
        vb2[r] = k < NB ? __ldg(g2 + k) : 0.0;
    }
    const double wq1 = t < NQ ? __ldg(&d.wq1[e1 * NQ + t]) : 0.0, wq2 = t < NQ ? __ldg(&d.wq2[e2 * NQ + t]) : 0.0;
    const int a = t % (P + 1), b = t / (P + 1);
    const int cpi = (i0 + a) + d.n1 * (j0 + b);
    const double wv = d.rational ? __ldg(&d.w[cpi]) : 1.0;
    double x[3], u[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) { x[c] = __ldg(&d.cp[3 * cpi + c]); u[c] = d.disp[3 * cpi + c]; }
    double* s1 = &E.b1[0][0][0];
    double* s2 = &E.b2[0][0][0];
#pragma unroll
    for (int r = 0; r < RB; ++r) {
        const int k = t + r * NT;
        if (k < NB) { s1[k] = vb1[r]; s2[k] = vb2[r]; }
    }
    if (t < NQ) { E.w1[t] = wq1; E.w2[t] = wq2; }
#pragma unroll
    for (int c = 0; c < 3; ++c) { E.X[t][c] = x[c] * wv; E.U[t][c] = u[c]; }
    E.Wt[t] = wv;
}

// sum-factorised evaluation of a 3-component field: out[k][c], k = (val, d1, d2, d11, d22, d12)
template <int P, bool VAL>
__device__ __forceinline__ void eval_field3(const double (*F)[3], const double (*B1)[P + 1], const double (*B2)[P + 1],
                                            double out[6][3]) {
#pragma unroll
    for (int k = 0; k < 6; ++k) out[k][0] = out[k][1] = out[k][2] = 0.0;
#pragma unroll
    for (int b = 0; b <= P; ++b) {
        double t0[3] = {0, 0, 0}, t1[3] = {0, 0, 0}, t2[3] = {0, 0, 0};
#pragma unroll
        for (int a = 0; a <= P; ++a) {
            const double* f = F[a + (P + 1) * b];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                t0[c] = fma(B1[0][a], f[c], t0[c]);
                t1[c] = fma(B1[1][a], f[c], t1[c]);
                t2[c] = fma(B1[2][a], f[c], t2[c]);
            }
        }
        const double y0 = B2[0][b], y1 = B2[1][b], y2 = B2[2][b];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            if (VAL) out[0][c] = fma(y0, t0[c], out[0][c]);
            out[1][c] = fma(y0, t1[c], out[1][c]);
            out[2][c] = fma(y1, t0[c], out[2][c]);
            out[3][c] = fma(y0, t2[c], out[3][c]);
            out[4][c] = fma(y2, t0[c], out[4][c]);
            out[5][c] = fma(y1, t1[c], out[5][c]);
        }
    }
}
template <int P>
__device__ __forceinline__ void eval_field1(const double* F, const double (*B1)[P + 1], const double (*B2)[P + 1], double out[6]) {
#pragma unroll
    for (int k = 0; k < 6; ++k) out[k] = 0.0;
#pragma unroll
    for (int b = 0; b <= P; ++b) {
        double t0 = 0, t1 = 0, t2 = 0;
#pragma unroll
        for (int a = 0; a <= P; ++a) {
            const double f = F[a + (P + 1) * b];
            t0 = fma(B1[0][a], f, t0); t1 = fma(B1[1][a], f, t1); t2 = fma(B1[2][a], f, t2);
        }
        const double y0 = B2[0][b], y1 = B2[1][b], y2 = B2[2][b];
        out[0] = fma(y0, t0, out[0]); out[1] = fma(y0, t1, out[1]); out[2] = fma(y1, t0, out[2]);
        out[3] = fma(y0, t2, out[3]); out[4] = fma(y2, t0, out[4]); out[5] = fma(y1, t1, out[5]);
    }
}

// ---------------------------------------------------------------------------------------------
// material laws
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void inv2s(const double m[3], double inv[3], double& det) {
    det = m[0] * m[1] - m[2] * m[2];
    const double r = 1.0 / det;
    inv[0] = m[1] * r; inv[1] = m[0] * r; inv[2] = -m[2] * r;
}
// sym(v,u) = g^{ac} g^{bd} + g^{ad} g^{bc} in symmetric Voigt storage
__device__ __forceinline__ void symprod(const double g[3], double s[6]) {
    s[0] = 2.0 * g[0] * g[0];
    s[1] = 2.0 * g[1] * g[1];
    s[2] = g[0] * g[1] + g[2] * g[2];
    s[3] = 2.0 * g[2] * g[2];
    s[4] = 2.0 * g[0] * g[2];
    s[5] = 2.0 * g[1] * g[2];
}
__device__ __forceinline__ void outer_s(const double a[3], const double b[3], double o[6]) {   // a_v b_u symmetrised storage (a_v b_u, exact only if symmetric)
    o[0] = a[0] * b[0]; o[1] = a[1] * b[1]; o[2] = a[2] * b[2]; o[3] = a[0] * b[1]; o[4] = a[0] * b[2]; o[5] = a[1] * b[2];
}

// incompressible NH / MR at one thickness point (Kiendl et al. 2015, static condensation C33 = J0^-2)
template <bool TANGENT>
__device__ __forceinline__ void hyper_incomp(const KLMaterial& m, const double Gi[3], const double gc[3], const double gi[3],
                                             double J0sq, double S[3], double C[6]) {
    const double c33 = 1.0 / J0sq;
    const double trs = gc[0] * Gi[0] + gc[1] * Gi[1] + 2.0 * gc[2] * Gi[2];
    const double dpsi33 = 0.5 * m.c1 + 0.5 * m.c2 * trs;
    const double sa = m.c1 + m.c2 * c33, sb = m.c2 * J0sq - 2.0 * dpsi33 * c33;
#pragma unroll
    for (int v = 0; v < 3; ++v) S[v] = sa * Gi[v] + sb * gi[v];
    if (!TANGENT) return;
    double sy[6];
    symprod(gi, sy);
    const double k1 = 2.0 * m.c2 * J0sq + 4.0 * dpsi33 * c33, k2 = -m.c2 * J0sq + 2.0 * dpsi33 * c33, k3 = -2.0 * m.c2 * c33;
#pragma unroll
    for (int v = 0; v < 3; ++v)
#pragma unroll
        for (int u = v; u < 3; ++u) {
            const int s = sidx(v, u);
            C[s] = k1 * gi[v] * gi[u] + k2 * sy[s] + k3 * (Gi[v] * gi[u] + Gi[u] * gi[v]);
        }
}

// compressible NH / MR: psi = c1/2 (J^-2/3 I1 - 3) + c2/2 (J^-4/3 I2 - 3) + K/4 (J^2 - 1 - 2 ln J);
// Newton on C33 until S33 = 0, then static condensation.  Returns false if not converged.
template <bool TANGENT>
__device__ __forceinline__ bool hyper_comp(const KLMaterial& m, const double Gi[3], const double gc[3], const double gi[3],
                                           double J0sq, double S[3], double C[6], double* c33_out = nullptr) {
    // contravariant push of the in-plane C: Cup = Gi * gc * Gi  (2x2, symmetric)
    const double t00 = Gi[0] * gc[0] + Gi[2] * gc[2], t01 = Gi[0] * gc[2] + Gi[2] * gc[1];
    const double t10 = Gi[2] * gc[0] + Gi[1] * gc[2], t11 = Gi[2] * gc[2] + Gi[1] * gc[1];
    const double Cup[3] = {t00 * Gi[0] + t01 * Gi[2], t10 * Gi[2] + t11 * Gi[1], t00 * Gi[2] + t01 * Gi[1]};
    const double trs = t00 + t11;
    const double tr2s = Cup[0] * gc[0] + Cup[1] * gc[1] + 2.0 * Cup[2] * gc[2];
    const double K = m.bulk, c1 = m.c1, c2 = m.c2;
    double c33 = 1.0 / J0sq, I1, I2, Jsq, j23, j43, ci;   // start from the incompressible solution C33 = J0^-2
    bool conv = false;
    for (int it = 0; it < 100; ++it) {
        I1 = trs + c33;
        I2 = 0.5 * (I1 * I1 - tr2s - c33 * c33);
        Jsq = J0sq * c33;
        j23 = 1.0 / cbrt(Jsq);
        j43 = j23 * j23;
        ci = 1.0 / c33;
        if (conv) break;
        const double dI2 = I1 - c33;
        const double S33 = c1 * j23 * (1.0 - I1 / 3.0 * ci) + c2 * j43 * (dI2 - 2.0 / 3.0 * I2 * ci) + 0.5 * K * (Jsq - 1.0) * ci;
        const double C3333 = 2.0 * c1 * j23 * (-2.0 / 3.0 * ci + 4.0 / 9.0 * I1 * ci * ci)
                           + 2.0 * c2 * j43 * (-4.0 / 3.0 * dI2 * ci + 10.0 / 9.0 * I2 * ci * ci) + K * ci * ci;
        const double dc = -2.0 * S33 / C3333;
        c33 += dc;
        if (!(c33 > 0.0) || !isfinite(c33)) return false;
        if (fabs(dc) <= 1e-14 * fabs(c33)) conv = true;
    }
    if (!conv) return false;
    if (c33_out) *c33_out = c33;   // thickness stretch^2 (stress recovery, kl_stress.cu)
    // in-plane stress and the tensor components needed for condensation
    double dI2v[3], Sv[3], Cab33[3];
#pragma unroll
    for (int v = 0; v < 3; ++v) {
        dI2v[v] = I1 * Gi[v] - Cup[v];
        Sv[v] = c1 * j23 * (Gi[v] - I1 / 3.0 * gi[v]) + c2 * j43 * (dI2v[v] - 2.0 / 3.0 * I2 * gi[v]) + 0.5 * K * (Jsq - 1.0) * gi[v];
        S[v] = Sv[v];
        if (!TANGENT) continue;
        // C^{ab33}: Ci^{33}=ci, G^{33}=1, Ic^{ab33}=0, dI2^{33}=I1-c33, d2I2^{ab33}=G^{ab}
        Cab33[v] = 2.0 * c1 * j23 * (-1.0 / 3.0 * ci * Gi[v] - 1.0 / 3.0 * gi[v] + I1 / 9.0 * gi[v] * ci)
                 + 2.0 * c2 * j43 * (-2.0 / 3.0 * ci * (dI2v[v] - 2.0 / 3.0 * I2 * gi[v]) + Gi[v] - 2.0 / 3.0 * (I1 - c33) * gi[v])
                 + K * Jsq * gi[v] * ci;
    }
    if (!TANGENT) return true;
    const double dI2_33 = I1 - c33;
    const double C3333 = 2.0 * c1 * j23 * (-2.0 / 3.0 * ci + 4.0 / 9.0 * I1 * ci * ci)
                       + 2.0 * c2 * j43 * (-4.0 / 3.0 * dI2_33 * ci + 10.0 / 9.0 * I2 * ci * ci) + K * ci * ci;
    double syg[6], syG[6];
    symprod(gi, syg);   // 2*Ic^{abcd}
    symprod(Gi, syG);
#pragma unroll
    for (int v = 0; v < 3; ++v)
#pragma unroll
        for (int u = v; u < 3; ++u) {
            const int s = sidx(v, u);
            const double Ic = 0.5 * syg[s];
            const double d2I2 = Gi[v] * Gi[u] - 0.5 * syG[s];
            const double iso1 = 2.0 * c1 * j23 * (-1.0 / 3.0 * gi[u] * Gi[v] - 1.0 / 3.0 * Gi[u] * gi[v] + I1 / 9.0 * gi[v] * gi[u] + I1 / 3.0 * Ic);
            const double iso2 = 2.0 * c2 * j43 * (-2.0 / 3.0 * gi[u] * (dI2v[v] - 2.0 / 3.0 * I2 * gi[v]) + d2I2
                                                 - 2.0 / 3.0 * dI2v[u] * gi[v] + 2.0 / 3.0 * I2 * Ic);
            const double vol = K * (Jsq * gi[v] * gi[u] - (Jsq - 1.0) * Ic);
            C[s] = iso1 + iso2 + vol - Cab33[v] * Cab33[u] / C3333;
        }
    return true;
}

// A,B,D (sym Voigt, 6 each), N, M (3 each) from the covariant metrics / curvatures [11,22,12].
// Returns a KLF_* flag word (0 = ok).
template <bool TANGENT>
__device__ __forceinline__ int material_point(const KLMaterial& m, const double Ac[3], const double Bc[3], const double ac[3],
                                              const double bc[3], double A[6], double B[6], double D[6], double N[3], double M[3]) {
    int flag = 0;
#pragma unroll
    for (int k = 0; k < 6; ++k) { A[k] = 0; B[k] = 0; D[k] = 0; }
#pragma unroll
    for (int k = 0; k < 3; ++k) { N[k] = 0; M[k] = 0; }
    double Ai[3], ai[3], dA, da;
    inv2s(Ac, Ai, dA);
    inv2s(ac, ai, da);
    if (!(dA > 0.0) || !(da > 0.0)) return KLF_METRIC;
    if (m.material == KL_MAT_SVK) {
        double sy[6], Cm[6];
        symprod(Ai, sy);
#pragma unroll
        for (int v = 0; v < 3; ++v)
#pragma unroll
            for (int u = v; u < 3; ++u) {
                const int s = sidx(v, u);
                Cm[s] = m.lam_ps * Ai[v] * Ai[u] + m.mu * sy[s];
            }
        const double eps[3] = {0.5 * (ac[0] - Ac[0]), 0.5 * (ac[1] - Ac[1]), ac[2] - Ac[2]};
        const double kap[3] = {Bc[0] - bc[0], Bc[1] - bc[1], 2.0 * (Bc[2] - bc[2])};
        const double t = m.t, t3 = m.t * m.t * m.t / 12.0;
#pragma unroll
        for (int k = 0; k < 6; ++k) { A[k] = t * Cm[k]; D[k] = t3 * Cm[k]; }
#pragma unroll
        for (int v = 0; v < 3; ++v) {
            double sn = 0, sm = 0;
#pragma unroll
            for (int u = 0; u < 3; ++u) { sn += A[sidx(v, u)] * eps[u]; sm += D[sidx(v, u)] * kap[u]; }
            N[v] = sn; M[v] = sm;
        }
        return 0;
    }
    double BAB[3] = {0, 0, 0}, bab[3] = {0, 0, 0};
    if (m.metric_z2) {
        // n,a . n,b = b_ag a^gd b_db
        const double Bm[2][2] = {{Bc[0], Bc[2]}, {Bc[2], Bc[1]}}, bm[2][2] = {{bc[0], bc[2]}, {bc[2], bc[1]}};
        const double AI[2][2] = {{Ai[0], Ai[2]}, {Ai[2], Ai[1]}}, aI[2][2] = {{ai[0], ai[2]}, {ai[2], ai[1]}};
        const int vi[3] = {0, 1, 0}, vj[3] = {0, 1, 1};
#pragma unroll
        for (int v = 0; v < 3; ++v)
#pragma unroll
            for (int c = 0; c < 2; ++c)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    BAB[v] += Bm[vi[v]][c] * AI[c][e] * Bm[e][vj[v]];
                    bab[v] += bm[vi[v]][c] * aI[c][e] * bm[e][vj[v]];
                }
    }
    for (int k = 0; k < m.ngauss; ++k) {
        const double z = 0.5 * m.t * m.zg[k], wz = 0.5 * m.t * m.wg[k];
        double Gc[3], gc[3], Gi[3], gi[3], dG, dg, S[3], C[6];
#pragma unroll
        for (int v = 0; v < 3; ++v) {
            Gc[v] = Ac[v] - 2.0 * z * Bc[v] + z * z * BAB[v];
            gc[v] = ac[v] - 2.0 * z * bc[v] + z * z * bab[v];
        }
        inv2s(Gc, Gi, dG);
        inv2s(gc, gi, dg);
        if (!(dG > 0.0) || !(dg > 0.0)) { flag |= KLF_METRIC; break; }
        const double J0sq = dg / dG;
        if (m.compressible) {
            if (!hyper_comp<TANGENT>(m, Gi, gc, gi, J0sq, S, C)) { flag |= KLF_C33; break; }
        } else {
            hyper_incomp<TANGENT>(m, Gi, gc, gi, J0sq, S, C);
        }
        const double wz1 = wz * z, wz2 = wz * z * z;
#pragma unroll
        for (int v = 0; v < 3; ++v) { N[v] = fma(wz, S[v], N[v]); M[v] = fma(wz1, S[v], M[v]); }
#pragma unroll
        if (TANGENT)
#pragma unroll
            for (int s = 0; s < 6; ++s) { A[s] = fma(wz, C[s], A[s]); B[s] = fma(wz1, C[s], B[s]); D[s] = fma(wz2, C[s], D[s]); }
    }
    return flag;
}

// ---------------------------------------------------------------------------------------------
// Per-point record written by phase 1 of the assembly kernels (all stress-like quantities are
// pre-multiplied by the quadrature weight times meas(ori)).
// ---------------------------------------------------------------------------------------------
struct PointData {
    double a1[3], a2[3], n[3], c1[3], c2[3];   // covariant / normal / contravariant vectors of the deformed surface
    double G1[3], G2[3];                        // Christoffel symbols Gamma^1_ab, Gamma^2_ab (ab = 11,22,12)
    double A[6], B[6], D[6];                    // thickness-integrated tangent (sym Voigt) * wJ
    double N[3];                                // membrane forces * wJ
    double Mt[3];                               // (M1, M2, 2*M3) * wJ
    double Ha1, Ha2, Hn;                        // H.a^1, H.a^2, H.n   with H = Mt_ab x,ab
    double q[3];                                // (H - n Hn)/|a1 x a2|
    double acon[3];                             // a^11, a^22, a^12
    double wJ;                                  // weight * meas(ori)
    double pad[3];                              // 58 doubles = 464 B: a multiple of 16 (TMA) with a stride of 20 banks mod 32, so the
                                                // thread-strided 128-bit stores of k_points are conflict-free (56 doubles: 8-way conflicts)
};
static_assert(sizeof(PointData) % 8 == 0, "PointData must be a whole number of doubles");

// TANGENT = false skips the material tangent (A,B,D): all the residual needs are the stress resultants
template <int P, bool TANGENT = true>
__device__ __forceinline__ int eval_point(const KLDev& d, const ElemStage<P>& E, int q1, int q2, PointData& o) {
    double fo[6][3], fu[6][3];
    eval_field3<P, true>(E.X, E.b1[q1], E.b2[q2], fo);
    eval_field3<P, false>(E.U, E.b1[q1], E.b2[q2], fu);
    double A1[3], A2[3], H[3][3];
    if (d.rational) {
        double fw[6];
        eval_field1<P>(E.Wt, E.b1[q1], E.b2[q2], fw);
        const double iw = 1.0 / fw[0];
        double X[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            X[c] = fo[0][c] * iw;
            A1[c] = (fo[1][c] - fw[1] * X[c]) * iw;
            A2[c] = (fo[2][c] - fw[2] * X[c]) * iw;
            H[0][c] = (fo[3][c] - fw[3] * X[c] - 2.0 * fw[1] * A1[c]) * iw;
            H[1][c] = (fo[4][c] - fw[4] * X[c] - 2.0 * fw[2] * A2[c]) * iw;
            H[2][c] = (fo[5][c] - fw[5] * X[c] - fw[1] * A2[c] - fw[2] * A1[c]) * iw;
        }
    } else {
#pragma unroll
        for (int c = 0; c < 3; ++c) { A1[c] = fo[1][c]; A2[c] = fo[2][c]; H[0][c] = fo[3][c]; H[1][c] = fo[4][c]; H[2][c] = fo[5][c]; }
    }
    double h[3][3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        o.a1[c] = A1[c] + fu[1][c];
        o.a2[c] = A2[c] + fu[2][c];
        h[0][c] = H[0][c] + fu[3][c];
        h[1][c] = H[1][c] + fu[4][c];
        h[2][c] = H[2][c] + fu[5][c];
    }
    double Nn[3], nn[3];
    cross3(A1, A2, Nn);
    cross3(o.a1, o.a2, nn);
    const double JA = sqrt(dot3(Nn, Nn)), Ja = sqrt(dot3(nn, nn));
    int flag = 0;
    if (!(JA > 0.0) || !(Ja > 0.0) || !isfinite(Ja)) flag |= KLF_JACOBIAN;
    const double iJA = 1.0 / JA, iJa = 1.0 / Ja;
#pragma unroll
    for (int c = 0; c < 3; ++c) { Nn[c] *= iJA; nn[c] *= iJa; o.n[c] = nn[c]; }
    double Ac[3] = {dot3(A1, A1), dot3(A2, A2), dot3(A1, A2)};
    double ac[3] = {dot3(o.a1, o.a1), dot3(o.a2, o.a2), dot3(o.a1, o.a2)};
    double Bc[3], bc[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) { Bc[k] = dot3(H[k], Nn); bc[k] = dot3(h[k], nn); }
    if (!d.mat.bending) {
#pragma unroll
        for (int k = 0; k < 3; ++k) { Bc[k] = 0; bc[k] = 0; }
    }
    double Mm[3];
    flag |= material_point<TANGENT>(d.mat, Ac, Bc, ac, bc, o.A, o.B, o.D, o.N, Mm);
    const double wJ = E.w1[q1] * E.w2[q2] * JA;
    o.wJ = wJ;
#pragma unroll
    for (int k = 0; k < 6; ++k) { o.A[k] *= wJ; o.B[k] *= wJ; o.D[k] *= wJ; }
    if (!d.mat.bending) {
#pragma unroll
        for (int k = 0; k < 6; ++k) { o.B[k] = 0; o.D[k] = 0; }
        Mm[0] = Mm[1] = Mm[2] = 0.0;
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) o.N[k] *= wJ;
    o.Mt[0] = Mm[0] * wJ; o.Mt[1] = Mm[1] * wJ; o.Mt[2] = 2.0 * Mm[2] * wJ;
    // contravariant basis, Christoffel symbols
    double ai[3], da;
    inv2s(ac, ai, da);
    o.acon[0] = ai[0]; o.acon[1] = ai[1]; o.acon[2] = ai[2];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        o.c1[c] = ai[0] * o.a1[c] + ai[2] * o.a2[c];
        o.c2[c] = ai[2] * o.a1[c] + ai[1] * o.a2[c];
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) { o.G1[k] = dot3(h[k], o.c1); o.G2[k] = dot3(h[k], o.c2); }
    double Hv[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) Hv[c] = o.Mt[0] * h[0][c] + o.Mt[1] * h[1][c] + o.Mt[2] * h[2][c];
    o.Ha1 = dot3(Hv, o.c1); o.Ha2 = dot3(Hv, o.c2); o.Hn = dot3(Hv, nn);
#pragma unroll
    for (int c = 0; c < 3; ++c) o.q[c] = (Hv[c] - nn[c] * o.Hn) * iJa;
    // non-finite guard on a few representative outputs
    const double chk = o.A[0] + o.D[0] + o.N[0] + o.Mt[0] + o.Hn + o.G1[0];
    if (!isfinite(chk)) flag |= KLF_NONFINITE;
    return flag;
}
