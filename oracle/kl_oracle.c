/*
 * kl_oracle.c — CPU ORACLE for the Kirchhoff–Love shell Jacobian/residual assembly path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT THE PRODUCT.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.  The product path
 * (gsstructuralanalysis_b200/csrc) never links or calls anything in oracle/.
 *
 * PARITY UNPINNED: the arithmetic of this path lives in gismo/gismo@dev, gismo/gsKLShell@main
 * (un-pinned branch heads, .github/workflows/ci.yml:26-41 of the reference) which are NOT
 * present in /root/reference and cannot be built here (no Eigen, no network).  This file
 * restates the published algorithm (Kiendl et al. 2015, "Isogeometric Kirchhoff–Love shell
 * formulations for general hyperelastic materials"; the gsKLShell expression set-up as it is
 * re-stated INSIDE the reference at benchmarks/benchmark_cylinder_DC.cpp:536-555) in a slow,
 * brute-force style, and is anchored by
 *   - the reference's own known answers (unittests/gsStaticSolver_test.cpp:358-385,415;
 *     filedata/pde/kirchhoff_shell_scordelis.xml:104-107),
 *   - K = d(F_int)/du by finite differences, F_int = dW/du of the discrete energy
 *     (tests/test_oracle_*.py).
 *
 * Reference call sites followed (reference file:line):
 *   constructSolution  tutorials/nonlinear_shell_static.cpp:123,132  (DoF-map rule:
 *                      benchmarks/benchmark_cylinder_DC.cpp:507-526)
 *   assembleMatrix     tutorials/nonlinear_shell_static.cpp:124
 *   assembleVector     tutorials/nonlinear_shell_static.cpp:133      (virtual work:
 *                      benchmarks/benchmark_cylinder_DC.cpp:541-554)
 *   quadrature options filedata/options/solver_options.xml:11-16 (quA, quB, quRule=1)
 *   material options   tutorials/nonlinear_shell_static.cpp:101-105,
 *                      unittests/gsStaticSolver_test.cpp:189-218
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "../include/kl_shell.h"

#define MAXP 5
#define MAXLOC ((MAXP + 1) * (MAXP + 1))
#define MAXQ 12

typedef struct klo {
    int p[2], nk[2], n[2];
    double* U[2];
    int nel[2];           /* non-empty knot spans per direction            */
    int* span[2];         /* knot span index of each element               */
    int ncp;
    double* cp;           /* [ncp*3] */
    double* w;            /* [ncp] or NULL */
    int* map;             /* [3*ncp] */
    int nfree, nfixed;
    double* fixed;        /* [nfixed] */
    kl_problem P;         /* scalar fields only (pointers invalid)         */
    int npl;
    double *pl_uv, *pl_val;
    /* pattern (column-compressed; structurally symmetric) */
    int* outer;
    int* inner;
    long nnz;
    double* fext;         /* [nfree] dead loads */
    double* force;        /* [nfree] Force = assemble().rhs(): dead loads + pressure on the undeformed surface - Dirichlet lifting */
    double* lift;         /* set-up only: K_L(free, eliminated) g */
    int nneu; int* neu_side; double* neu_val;
    int nthreads;
    int e2_begin, e2_end;  /* element rows assembled (strip partition tests) */
} klo;

#include "bspline_common.h"

/* small vector helpers */
static inline double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static inline void cross3(const double* a, const double* b, double* c) {
    c[0] = a[1] * b[2] - a[2] * b[1]; c[1] = a[2] * b[0] - a[0] * b[2]; c[2] = a[0] * b[1] - a[1] * b[0];
}

/* ------------------------------------------------------------------------------------ */
/* Material laws.  Voigt order (11,22,12).  Input: covariant metric/curvature of the
 * undeformed (Ac,Bc) and deformed (ac,bc) mid-surface as [11,22,12].  Output: A,B,D (3x3),
 * N,M (3): thickness-integrated tangent and stress resultants (SURVEY A.5;
 * gsMaterialMatrixIntegrate<MatrixA..D / VectorN / VectorM>).  Returns 0 or KL_E_*.    */
static const int VI[3] = {0, 1, 0}, VJ[3] = {0, 1, 1};

static void inv2(const double m[3], double inv[3], double* det) {
    *det = m[0] * m[1] - m[2] * m[2];
    inv[0] = m[1] / *det; inv[1] = m[0] / *det; inv[2] = -m[2] / *det;
}
static inline double s2(const double m[3], int i, int j) { return i == j ? m[i] : m[2]; }

/* 3D hyperelastic point evaluation with plane-stress treatment.  G,g: covariant in-plane
 * metrics at thickness coordinate z (Voigt).  Out: S[3], C[3][3] (condensed).           */
static int hyper_point(const kl_problem* P, const double Gc[3], const double gc[3], double S[3], double C[3][3], double* c33_out) {
    double Gi[3], gi[3], detG, detg;
    inv2(Gc, Gi, &detG);
    inv2(gc, gi, &detg);
    if (!(detg > 0.0) || !(detG > 0.0)) return KL_E_JACOBIAN;
    double J0sq = detg / detG;
    double mu = P->E / (2.0 * (1.0 + P->nu));
    double c1 = mu, c2 = 0.0;
    if (P->material == KL_MAT_MR) { c2 = mu / (P->mr_ratio + 1.0); c1 = P->mr_ratio * c2; }
    if (!P->compressible) {
        /* Kiendl et al. 2015, static condensation with C33 = J0^-2; psi = c1/2 (I1-3) + c2/2 (I2-3),
         * I1 = trs + C33, I2 = C33*trs + J0^2,  trs = g_ab G^ab                                   */
        double trs = 0.0;
        for (int a = 0; a < 2; ++a) for (int b = 0; b < 2; ++b) trs += s2(gc, a, b) * s2(Gi, a, b);
        double c33 = 1.0 / J0sq;
        if (c33_out) *c33_out = c33;
        double dpsi33 = 0.5 * c1 + 0.5 * c2 * trs;
        for (int v = 0; v < 3; ++v) {
            int a = VI[v], b = VJ[v];
            double dpsi_ab = 0.5 * c1 * s2(Gi, a, b) + 0.5 * c2 * (c33 * s2(Gi, a, b) + J0sq * s2(gi, a, b));
            S[v] = 2.0 * dpsi_ab - 2.0 * dpsi33 * c33 * s2(gi, a, b);
            for (int u = 0; u < 3; ++u) {
                int c = VI[u], d = VJ[u];
                double gab = s2(gi, a, b), gcd = s2(gi, c, d);
                double sym = s2(gi, a, c) * s2(gi, b, d) + s2(gi, a, d) * s2(gi, b, c);
                double d2psi_abcd = 0.5 * c2 * J0sq * (gab * gcd - 0.5 * sym);
                double d2psi_33ab = 0.5 * c2 * s2(Gi, a, b);
                double d2psi_33cd = 0.5 * c2 * s2(Gi, c, d);
                C[v][u] = 4.0 * d2psi_abcd - 4.0 * d2psi_33ab * c33 * gcd - 4.0 * d2psi_33cd * c33 * gab
                          + 2.0 * dpsi33 * c33 * (2.0 * gab * gcd + sym);
            }
        }
        return 0;
    }
    /* compressible: psi = c1/2 (J^-2/3 I1 - 3) + c2/2 (J^-4/3 I2 - 3) + K/4 (J^2 - 1 - 2 ln J);
     * full 3D tensors on block-diagonal C = diag(g_ab, C33); Newton on C33 until S33 = 0;
     * then C^abcd - C^ab33 C^33cd / C^3333.                                                */
    double K = 2.0 * mu * (1.0 + P->nu) / (3.0 - 6.0 * P->nu);
    double G3[3][3] = {{0}}, Cc[3][3] = {{0}}, Ci[3][3] = {{0}}, Cup[3][3];
    for (int a = 0; a < 2; ++a) for (int b = 0; b < 2; ++b) { G3[a][b] = s2(Gi, a, b); Cc[a][b] = s2(gc, a, b); Ci[a][b] = s2(gi, a, b); }
    G3[2][2] = 1.0;
    double c33 = 1.0 / J0sq;   /* start from the incompressible solution C33 = J0^-2 */
    double S3[3][3], C4[3][3][3][3];
    int converged = 0;
    for (int it = 0; it < 100; ++it) {
        Cc[2][2] = c33; Ci[2][2] = 1.0 / c33;
        double Jsq = J0sq * c33, J = sqrt(Jsq);
        double I1 = 0.0, trC2 = 0.0;
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
            I1 += Cc[i][j] * G3[i][j];
            double s = 0.0;
            for (int k = 0; k < 3; ++k) for (int l = 0; l < 3; ++l) s += G3[i][k] * Cc[k][l] * G3[l][j];
            Cup[i][j] = s;
        }
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) trC2 += Cup[i][j] * Cc[i][j];
        double I2 = 0.5 * (I1 * I1 - trC2);
        double j23 = pow(J, -2.0 / 3.0), j43 = j23 * j23;
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
            double dI2 = I1 * G3[i][j] - Cup[i][j];
            S3[i][j] = c1 * j23 * (G3[i][j] - I1 / 3.0 * Ci[i][j]) + c2 * j43 * (dI2 - 2.0 / 3.0 * I2 * Ci[i][j])
                       + 0.5 * K * (Jsq - 1.0) * Ci[i][j];
            for (int k = 0; k < 3; ++k) for (int l = 0; l < 3; ++l) {
                double Ic = 0.5 * (Ci[i][k] * Ci[j][l] + Ci[i][l] * Ci[j][k]);
                double dI2kl = I1 * G3[k][l] - Cup[k][l];
                double d2I2 = G3[i][j] * G3[k][l] - 0.5 * (G3[i][k] * G3[j][l] + G3[i][l] * G3[j][k]);
                double iso1 = 2.0 * c1 * j23 * (-1.0 / 3.0 * Ci[k][l] * G3[i][j] - 1.0 / 3.0 * G3[k][l] * Ci[i][j]
                                                + I1 / 9.0 * Ci[i][j] * Ci[k][l] + I1 / 3.0 * Ic);
                double iso2 = 2.0 * c2 * j43 * (-2.0 / 3.0 * Ci[k][l] * (dI2 - 2.0 / 3.0 * I2 * Ci[i][j]) + d2I2
                                                - 2.0 / 3.0 * dI2kl * Ci[i][j] + 2.0 / 3.0 * I2 * Ic);
                double vol = K * (Jsq * Ci[i][j] * Ci[k][l] - (Jsq - 1.0) * Ic);
                C4[i][j][k][l] = iso1 + iso2 + vol;
            }
        }
        if (converged) break;
        double dc33 = -2.0 * S3[2][2] / C4[2][2][2][2];
        c33 += dc33;
        if (!(c33 > 0.0) || !isfinite(c33)) return KL_E_C33;
        if (fabs(dc33) <= 1e-14 * fabs(c33)) converged = 1;   /* one more pass evaluates at the root */
    }
    if (!converged) return KL_E_C33;
    if (c33_out) *c33_out = c33;
    for (int v = 0; v < 3; ++v) {
        int a = VI[v], b = VJ[v];
        S[v] = S3[a][b];
        for (int u = 0; u < 3; ++u) {
            int c = VI[u], d = VJ[u];
            C[v][u] = C4[a][b][c][d] - C4[a][b][2][2] * C4[2][2][c][d] / C4[2][2][2][2];
        }
    }
    return 0;
}

static int material_eval(const kl_problem* P, const double Ac[3], const double Bc[3], const double ac[3], const double bc[3],
                         double A[3][3], double B[3][3], double D[3][3], double N[3], double M[3]) {
    memset(A, 0, 72); memset(B, 0, 72); memset(D, 0, 72); memset(N, 0, 24); memset(M, 0, 24);
    double t = P->thickness;
    if (P->material == KL_MAT_SVK) {
        /* gsMaterialMatrixLinear: C^abcd = lam_ps A^ab A^cd + mu (A^ac A^bd + A^ad A^bc), A = t C, D = t^3/12 C */
        double Ai[3], det;
        inv2(Ac, Ai, &det);
        if (!(det > 0.0)) return KL_E_JACOBIAN;
        double mu = P->E / (2.0 * (1.0 + P->nu));
        double lam = P->E * P->nu / ((1.0 + P->nu) * (1.0 - 2.0 * P->nu));
        double lps = 2.0 * lam * mu / (lam + 2.0 * mu);
        double Cm[3][3];
        for (int v = 0; v < 3; ++v) for (int u = 0; u < 3; ++u) {
            int a = VI[v], b = VJ[v], c = VI[u], d = VJ[u];
            Cm[v][u] = lps * s2(Ai, a, b) * s2(Ai, c, d) + mu * (s2(Ai, a, c) * s2(Ai, b, d) + s2(Ai, a, d) * s2(Ai, b, c));
        }
        double eps[3] = {0.5 * (ac[0] - Ac[0]), 0.5 * (ac[1] - Ac[1]), (ac[2] - Ac[2])};       /* [e11,e22,2e12] */
        double kap[3] = {Bc[0] - bc[0], Bc[1] - bc[1], 2.0 * (Bc[2] - bc[2])};                  /* [k11,k22,2k12] */
        for (int v = 0; v < 3; ++v) for (int u = 0; u < 3; ++u) {
            A[v][u] = t * Cm[v][u];
            D[v][u] = t * t * t / 12.0 * Cm[v][u];
            N[v] += A[v][u] * eps[u];
            M[v] += D[v][u] * kap[u];
        }
        return 0;
    }
    /* hyperelastic: NumGauss points through the thickness */
    int ng = P->num_gauss_thickness;
    double xg[MAXQ], wg[MAXQ];
    gauss_legendre(ng, xg, wg);
    double Ai[3], ai[3], dA, da;
    inv2(Ac, Ai, &dA);
    inv2(ac, ai, &da);
    if (!(dA > 0.0) || !(da > 0.0)) return KL_E_JACOBIAN;
    for (int k = 0; k < ng; ++k) {
        double z = 0.5 * t * xg[k], wz = 0.5 * t * wg[k];
        double Gc[3], gc[3];
        for (int v = 0; v < 3; ++v) { Gc[v] = Ac[v] - 2.0 * z * Bc[v]; gc[v] = ac[v] - 2.0 * z * bc[v]; }
        if (P->metric_z2) {
            /* n,a . n,b = b_ag a^gd b_db */
            for (int v = 0; v < 3; ++v) {
                int a = VI[v], b = VJ[v];
                double sG = 0.0, sg = 0.0;
                for (int c = 0; c < 2; ++c) for (int d = 0; d < 2; ++d) {
                    sG += s2(Bc, a, c) * s2(Ai, c, d) * s2(Bc, d, b);
                    sg += s2(bc, a, c) * s2(ai, c, d) * s2(bc, d, b);
                }
                Gc[v] += z * z * sG; gc[v] += z * z * sg;
            }
        }
        double S[3], C[3][3];
        int rc = hyper_point(P, Gc, gc, S, C, NULL);
        if (rc) return rc;
        for (int v = 0; v < 3; ++v) {
            N[v] += wz * S[v];
            M[v] += wz * z * S[v];
            for (int u = 0; u < 3; ++u) {
                A[v][u] += wz * C[v][u];
                B[v][u] += wz * z * C[v][u];
                D[v][u] += wz * z * z * C[v][u];
            }
        }
    }
    return 0;
}

/* exported for unit tests of the material laws */
int klo_material(const kl_problem* P, const double* Ac, const double* Bc, const double* ac, const double* bc,
                 double* A9, double* B9, double* D9, double* N3, double* M3) {
    double A[3][3], B[3][3], D[3][3];
    int rc = material_eval(P, Ac, Bc, ac, bc, A, B, D, N3, M3);
    memcpy(A9, A, 72); memcpy(B9, B, 72); memcpy(D9, D, 72);
    return rc;
}

/* ------------------------------------------------------------------------------------ */
/* DoF numbering (gsFeSpace::setupMapper + gsDofMapper::finalize)             */
static int uf_find(int* parent, int i) { while (parent[i] != i) { parent[i] = parent[parent[i]]; i = parent[i]; } return i; }
static void uf_union(int* parent, int a, int b) { a = uf_find(parent, a); b = uf_find(parent, b); if (a != b) { if (a < b) parent[b] = a; else parent[a] = b; } }

int klo_build_dofmap(int n1, int n2, const kl_bc* bc, int* map, int* n_free, int* n_fixed) {
    int ncp = n1 * n2;
    int* parent = (int*)malloc(sizeof(int) * ncp);
    char* elim = (char*)malloc(ncp);
    char* coupled = (char*)malloc(ncp);
    int* gid = (int*)malloc(sizeof(int) * ncp);
    int free_off = 0;
    int* nfree_c = (int*)calloc(4, sizeof(int));
    /* pass 1: per component, decide state; free numbering; remember eliminated groups */
    int** elim_gid = (int**)malloc(3 * sizeof(int*));
    int nelim_total = 0;
    for (int c = 0; c < 3; ++c) {
        for (int i = 0; i < ncp; ++i) { parent[i] = i; elim[i] = 0; coupled[i] = 0; }
        for (int s = 0; s < 4; ++s) {
            int kind = bc->side[s][c];
            if (kind == KL_BC_FREE) continue;
            int len = (s < 2) ? n2 : n1;
            for (int k = 0; k < len; ++k) {
                int b0, b1;
                if (s == KL_WEST)       { b0 = 0 + n1 * k;        b1 = 1 + n1 * k; }
                else if (s == KL_EAST)  { b0 = n1 - 1 + n1 * k;   b1 = n1 - 2 + n1 * k; }
                else if (s == KL_SOUTH) { b0 = k;                 b1 = k + n1; }
                else                    { b0 = k + n1 * (n2 - 1); b1 = k + n1 * (n2 - 2); }
                if (kind == KL_BC_DIRICHLET) elim[b0] = 1;
                else if (kind == KL_BC_CLAMPED) { uf_union(parent, b0, b1); coupled[b0] = coupled[b1] = 1; }
                else if (kind == KL_BC_COLLAPSED) {
                    int first = (s == KL_WEST) ? 0 : (s == KL_EAST) ? n1 - 1 : (s == KL_SOUTH) ? 0 : n1 * (n2 - 1);
                    if (b0 != first) { uf_union(parent, first, b0); coupled[b0] = coupled[first] = 1; }
                }
            }
        }
        const int corner_idx[4] = {0, n1 - 1, n1 * (n2 - 1), n1 * n2 - 1};
        for (int k = 0; k < 4; ++k) if (bc->corner[k][c]) elim[corner_idx[k]] = 1;
        /* a coupled group is eliminated if any member is */
        for (int i = 0; i < ncp; ++i) if (elim[i]) elim[uf_find(parent, i)] = 1;
        for (int i = 0; i < ncp; ++i) if (elim[uf_find(parent, i)]) elim[i] = 1;
        /* numbering: plain free, then coupled groups (first appearance), eliminated later */
        int cnt = 0;
        for (int i = 0; i < ncp; ++i) gid[i] = -1;
        for (int i = 0; i < ncp; ++i) if (!elim[i] && !coupled[i]) map[c * ncp + i] = free_off + cnt++;
        for (int i = 0; i < ncp; ++i) if (!elim[i] && coupled[i]) {
            int r = uf_find(parent, i);
            if (gid[r] < 0) gid[r] = free_off + cnt++;
            map[c * ncp + i] = gid[r];
        }
        nfree_c[c] = cnt;
        free_off += cnt;
        /* eliminated: group ids in order of first appearance; stored negative for now */
        elim_gid[c] = (int*)malloc(sizeof(int) * ncp);
        for (int i = 0; i < ncp; ++i) gid[i] = -1;
        for (int i = 0; i < ncp; ++i) if (elim[i]) {
            int r = uf_find(parent, i);
            if (gid[r] < 0) gid[r] = nelim_total++;
            elim_gid[c][i] = gid[r];
        } else elim_gid[c][i] = -1;
    }
    for (int c = 0; c < 3; ++c) {
        for (int i = 0; i < ncp; ++i) if (elim_gid[c][i] >= 0) map[c * ncp + i] = free_off + elim_gid[c][i];
        free(elim_gid[c]);
    }
    *n_free = free_off; *n_fixed = nelim_total;
    free(elim_gid); free(parent); free(elim); free(coupled); free(gid); free(nfree_c);
    return 0;
}

/* ------------------------------------------------------------------------------------ */
static int cmp_int(const void* a, const void* b) { int x = *(const int*)a, y = *(const int*)b; return (x > y) - (x < y); }

static void build_pattern(klo* o) {
    /* structural pattern: (row,col) for all free DoF pairs whose basis functions share an element */
    int n1 = o->n[0], n2 = o->n[1], ncp = o->ncp, nf = o->nfree;
    /* 1-D element ranges per function */
    int* lo[2]; int* hi[2];
    for (int d = 0; d < 2; ++d) {
        lo[d] = (int*)malloc(sizeof(int) * o->n[d]); hi[d] = (int*)malloc(sizeof(int) * o->n[d]);
        for (int i = 0; i < o->n[d]; ++i) { lo[d][i] = 1 << 30; hi[d][i] = -1; }
        for (int e = 0; e < o->nel[d]; ++e) for (int a = 0; a <= o->p[d]; ++a) {
            int i = o->span[d][e] - o->p[d] + a;
            if (e < lo[d][i]) lo[d][i] = e;
            if (e > hi[d][i]) hi[d][i] = e;
        }
    }
    int* cnt = (int*)calloc(nf + 1, sizeof(int));
    for (int pass = 0; pass < 2; ++pass) {
        int** lists = NULL; int* fill = NULL;
        if (pass == 1) {
            lists = (int**)malloc(sizeof(int*) * nf); fill = (int*)calloc(nf, sizeof(int));
            for (int r = 0; r < nf; ++r) lists[r] = (int*)malloc(sizeof(int) * (cnt[r] > 0 ? cnt[r] : 1));
        }
        for (int i2 = 0; i2 < n2; ++i2) for (int i1 = 0; i1 < n1; ++i1) {
            int I = i1 + n1 * i2;
            for (int j2 = 0; j2 < n2; ++j2) {
                if (hi[1][j2] < lo[1][i2] || lo[1][j2] > hi[1][i2]) continue;
                for (int j1 = 0; j1 < n1; ++j1) {
                    if (hi[0][j1] < lo[0][i1] || lo[0][j1] > hi[0][i1]) continue;
                    int J = j1 + n1 * j2;
                    for (int c = 0; c < 3; ++c) {
                        int col = o->map[c * ncp + I];
                        if (col >= nf) continue;
                        for (int d = 0; d < 3; ++d) {
                            int row = o->map[d * ncp + J];
                            if (row >= nf) continue;
                            if (pass == 0) cnt[col]++; else lists[col][fill[col]++] = row;
                        }
                    }
                }
            }
        }
        if (pass == 1) {
            o->outer = (int*)malloc(sizeof(int) * (nf + 1));
            o->outer[0] = 0;
            for (int r = 0; r < nf; ++r) {
                qsort(lists[r], fill[r], sizeof(int), cmp_int);
                int u = 0;
                for (int k = 0; k < fill[r]; ++k) if (k == 0 || lists[r][k] != lists[r][k - 1]) lists[r][u++] = lists[r][k];
                fill[r] = u;
                o->outer[r + 1] = o->outer[r] + u;
            }
            o->nnz = o->outer[nf];
            o->inner = (int*)malloc(sizeof(int) * (o->nnz > 0 ? o->nnz : 1));
            for (int r = 0; r < nf; ++r) { memcpy(o->inner + o->outer[r], lists[r], sizeof(int) * fill[r]); free(lists[r]); }
            free(lists); free(fill);
        }
    }
    free(cnt);
    for (int d = 0; d < 2; ++d) { free(lo[d]); free(hi[d]); }
}

static inline long find_pos(const klo* o, int row, int col) {
    int lo = o->outer[col], hi = o->outer[col + 1] - 1;
    while (lo <= hi) { int mid = (lo + hi) >> 1; int v = o->inner[mid]; if (v == row) return mid; if (v < row) lo = mid + 1; else hi = mid - 1; }
    return -1;
}

/* ------------------------------------------------------------------------------------ */
/* Evaluate everything the integrands need at one quadrature point of one element.       */
typedef struct qpdata {
    int nloc;
    int cpidx[MAXLOC];
    double R[MAXLOC], dR[MAXLOC][2], ddR[MAXLOC][3];   /* ddR: 11,22,12 */
    double X[3], A1[3], A2[3], H[3][3];                /* ori: point, tangents, second derivs (11,22,12) */
    double a1[3], a2[3], h[3][3];                      /* def */
} qpdata;

static void eval_qp(const klo* o, int e1, int e2, double u, double v, const double* disp /* [ncp*3] */, qpdata* q) {
    int p1 = o->p[0], p2 = o->p[1];
    int s1 = o->span[0][e1], s2_ = o->span[1][e2];
    double d1[3][MAXP + 1], d2[3][MAXP + 1];
    ders_basis(s1, u, p1, 2, o->U[0], d1);
    ders_basis(s2_, v, p2, 2, o->U[1], d2);
    q->nloc = (p1 + 1) * (p2 + 1);
    for (int b = 0; b <= p2; ++b) for (int a = 0; a <= p1; ++a) {
        int l = a + (p1 + 1) * b;
        q->cpidx[l] = (s1 - p1 + a) + o->n[0] * (s2_ - p2 + b);
        q->R[l] = d1[0][a] * d2[0][b];
        q->dR[l][0] = d1[1][a] * d2[0][b];
        q->dR[l][1] = d1[0][a] * d2[1][b];
        q->ddR[l][0] = d1[2][a] * d2[0][b];
        q->ddR[l][1] = d1[0][a] * d2[2][b];
        q->ddR[l][2] = d1[1][a] * d2[1][b];
    }
    /* undeformed geometry (rational if weights are present) */
    double W = 0, Wd[2] = {0, 0}, Wdd[3] = {0, 0, 0};
    double P[3] = {0, 0, 0}, Pd[2][3] = {{0}}, Pdd[3][3] = {{0}};
    for (int l = 0; l < q->nloc; ++l) {
        int i = q->cpidx[l];
        double w = o->w ? o->w[i] : 1.0;
        W += q->R[l] * w;
        for (int k = 0; k < 2; ++k) Wd[k] += q->dR[l][k] * w;
        for (int k = 0; k < 3; ++k) Wdd[k] += q->ddR[l][k] * w;
        for (int c = 0; c < 3; ++c) {
            double xw = o->cp[3 * i + c] * w;
            P[c] += q->R[l] * xw;
            for (int k = 0; k < 2; ++k) Pd[k][c] += q->dR[l][k] * xw;
            for (int k = 0; k < 3; ++k) Pdd[k][c] += q->ddR[l][k] * xw;
        }
    }
    double Xd[2][3];
    for (int c = 0; c < 3; ++c) {
        q->X[c] = P[c] / W;
        for (int k = 0; k < 2; ++k) Xd[k][c] = (Pd[k][c] - Wd[k] * q->X[c]) / W;
        q->A1[c] = Xd[0][c]; q->A2[c] = Xd[1][c];
        q->H[0][c] = (Pdd[0][c] - Wdd[0] * q->X[c] - 2.0 * Wd[0] * Xd[0][c]) / W;
        q->H[1][c] = (Pdd[1][c] - Wdd[1] * q->X[c] - 2.0 * Wd[1] * Xd[1][c]) / W;
        q->H[2][c] = (Pdd[2][c] - Wdd[2] * q->X[c] - Wd[0] * Xd[1][c] - Wd[1] * Xd[0][c]) / W;
    }
    /* deformed = undeformed + displacement field in the polynomial basis */
    for (int c = 0; c < 3; ++c) { q->a1[c] = q->A1[c]; q->a2[c] = q->A2[c]; for (int k = 0; k < 3; ++k) q->h[k][c] = q->H[k][c]; }
    for (int l = 0; l < q->nloc; ++l) {
        int i = q->cpidx[l];
        for (int c = 0; c < 3; ++c) {
            double uu = disp[3 * i + c];
            q->a1[c] += q->dR[l][0] * uu; q->a2[c] += q->dR[l][1] * uu;
            for (int k = 0; k < 3; ++k) q->h[k][c] += q->ddR[l][k] * uu;
        }
    }
}

/* constructSolution: displacement control net from the DoF vector */
static void construct_disp(const klo* o, const double* x, double* disp) {
    for (int c = 0; c < 3; ++c) for (int i = 0; i < o->ncp; ++i) {
        int g = o->map[c * o->ncp + i];
        disp[3 * i + c] = !x ? 0.0 : ((g < o->nfree) ? x[g] : (o->fixed ? o->fixed[g - o->nfree] : 0.0));   /* x == NULL: undeformed */
    }
}

/* mode bit 1: matrix, bit 2: internal force vector.  fint gets F_int - F_pressure (free rows). */
static int assemble_full(const klo* o, const double* x, double* Kval, double* fint, double* ffull /* [3*ncp] internal force at ALL control points, no pressure */) {
    int err = 0;
    double* disp = (double*)malloc(sizeof(double) * 3 * o->ncp);
    construct_disp(o, x, disp);
    if (Kval) memset(Kval, 0, sizeof(double) * o->nnz);
    if (fint) memset(fint, 0, sizeof(double) * o->nfree);
    int nq1 = o->P.quA * o->p[0] + o->P.quB, nq2 = o->P.quA * o->p[1] + o->P.quB;
    double xq1[MAXQ], wq1[MAXQ], xq2[MAXQ], wq2[MAXQ];
    gauss_legendre(nq1, xq1, wq1);
    gauss_legendre(nq2, xq2, wq2);
    int nel = o->nel[0] * o->nel[1];
    int ncp = o->ncp, nf = o->nfree;
#pragma omp parallel for schedule(dynamic, 16) num_threads(o->nthreads)
    for (int e = 0; e < nel; ++e) {
        if (err) continue;
        int e1 = e % o->nel[0], e2 = e / o->nel[0];
        if (e2 < o->e2_begin || e2 >= o->e2_end) continue;
        double ua = o->U[0][o->span[0][e1]], ub = o->U[0][o->span[0][e1] + 1];
        double va = o->U[1][o->span[1][e2]], vb = o->U[1][o->span[1][e2] + 1];
        qpdata q;
        int nloc = (o->p[0] + 1) * (o->p[1] + 1), nd = 3 * nloc;
        double* Ke = Kval ? (double*)calloc((size_t)nd * nd, sizeof(double)) : NULL;
        double fe[3 * MAXLOC], fi[3 * MAXLOC];
        memset(fe, 0, sizeof(fe));
        memset(fi, 0, sizeof(fi));
        for (int q2 = 0; q2 < nq2 && !err; ++q2) for (int q1 = 0; q1 < nq1 && !err; ++q1) {
            double u = 0.5 * (ua + ub) + 0.5 * (ub - ua) * xq1[q1];
            double v = 0.5 * (va + vb) + 0.5 * (vb - va) * xq2[q2];
            double wt = 0.25 * (ub - ua) * (vb - va) * wq1[q1] * wq2[q2];
            eval_qp(o, e1, e2, u, v, disp, &q);
            double Nn[3], nn[3], JA, Ja;
            cross3(q.A1, q.A2, Nn); JA = sqrt(dot3(Nn, Nn));
            cross3(q.a1, q.a2, nn); Ja = sqrt(dot3(nn, nn));
            if (!(JA > 0.0) || !(Ja > 0.0) || !isfinite(Ja)) { err = KL_E_JACOBIAN; break; }
            for (int c = 0; c < 3; ++c) { Nn[c] /= JA; nn[c] /= Ja; }
            double Ac[3] = {dot3(q.A1, q.A1), dot3(q.A2, q.A2), dot3(q.A1, q.A2)};
            double ac[3] = {dot3(q.a1, q.a1), dot3(q.a2, q.a2), dot3(q.a1, q.a2)};
            double Bc[3], bc[3];
            for (int k = 0; k < 3; ++k) { Bc[k] = dot3(q.H[k], Nn); bc[k] = dot3(q.h[k], nn); }
            if (!o->P.bending) { for (int k = 0; k < 3; ++k) { Bc[k] = 0; bc[k] = 0; } }
            double A[3][3], B[3][3], D[3][3], N[3], M[3];
            int rc = material_eval(&o->P, Ac, Bc, ac, bc, A, B, D, N, M);
            if (rc) { err = rc; break; }
            double wJ = wt * JA;   /* meas(ori) */
            /* first variations per local dof (l,c): dEm[3], dEf[3] (Voigt, m2 applied), dn[3], m = dñ/J */
            double dEm[3 * MAXLOC][3], dEf[3 * MAXLOC][3], dn[3 * MAXLOC][3], mm[3 * MAXLOC][3];
            for (int l = 0; l < nloc; ++l) for (int c = 0; c < 3; ++c) {
                int r = 3 * l + c;
                double ec[3] = {0, 0, 0}; ec[c] = 1.0;
                dEm[r][0] = q.dR[l][0] * q.a1[c];
                dEm[r][1] = q.dR[l][1] * q.a2[c];
                dEm[r][2] = q.dR[l][0] * q.a2[c] + q.dR[l][1] * q.a1[c];
                double t1[3], t2[3], dnt[3];
                cross3(ec, q.a2, t1); cross3(q.a1, ec, t2);
                for (int k = 0; k < 3; ++k) dnt[k] = (q.dR[l][0] * t1[k] + q.dR[l][1] * t2[k]) / Ja;
                double nd_ = dot3(nn, dnt);
                for (int k = 0; k < 3; ++k) { mm[r][k] = dnt[k]; dn[r][k] = dnt[k] - nn[k] * nd_; }
                if (o->P.bending) {
                    dEf[r][0] = -(q.ddR[l][0] * nn[c] + dot3(q.h[0], dn[r]));
                    dEf[r][1] = -(q.ddR[l][1] * nn[c] + dot3(q.h[1], dn[r]));
                    dEf[r][2] = -2.0 * (q.ddR[l][2] * nn[c] + dot3(q.h[2], dn[r]));
                } else { dEf[r][0] = dEf[r][1] = dEf[r][2] = 0.0; }
            }
            double pr = o->P.pressure;
            /* internal force (minus follower pressure) */
            for (int r = 0; r < nd; ++r) {
                double s = 0.0;
                for (int k = 0; k < 3; ++k) s += N[k] * dEm[r][k] + M[k] * dEf[r][k];
                fe[r] += wJ * s;
                fi[r] += wJ * s;
                if (pr != 0.0) fe[r] -= wJ * pr * q.R[r / 3] * nn[r % 3];
            }
            if (!Ke) continue;
            for (int r = 0; r < nd; ++r) {
                int l = r / 3, c = r % 3;
                double Nr[3], Mr[3];   /* dEm_r A + dEf_r B ;  dEm_r C + dEf_r D  (C = B) */
                for (int k = 0; k < 3; ++k) {
                    Nr[k] = 0; Mr[k] = 0;
                    for (int j = 0; j < 3; ++j) { Nr[k] += dEm[r][j] * A[j][k] + dEf[r][j] * B[j][k]; Mr[k] += dEm[r][j] * B[j][k] + dEf[r][j] * D[j][k]; }
                }
                for (int s = 0; s < nd; ++s) {
                    int m = s / 3, d = s % 3;
                    double val = 0.0;
                    for (int k = 0; k < 3; ++k) val += Nr[k] * dEm[s][k] + Mr[k] * dEf[s][k];
                    /* N : d2Em */
                    if (c == d)
                        val += N[0] * q.dR[l][0] * q.dR[m][0] + N[1] * q.dR[l][1] * q.dR[m][1]
                             + N[2] * (q.dR[l][0] * q.dR[m][1] + q.dR[l][1] * q.dR[m][0]);
                    if (o->P.bending || pr != 0.0) {
                        /* second variation of the unit normal */
                        double ec[3] = {0, 0, 0}, ed[3] = {0, 0, 0}, cx[3], d2n[3];
                        ec[c] = 1.0; ed[d] = 1.0;
                        cross3(ec, ed, cx);
                        double om = (q.dR[l][0] * q.dR[m][1] - q.dR[l][1] * q.dR[m][0]) / Ja;
                        double ncx = dot3(nn, cx);
                        double nmr = dot3(nn, mm[r]), nms = dot3(nn, mm[s]), dd = dot3(dn[r], dn[s]);
                        for (int k = 0; k < 3; ++k)
                            d2n[k] = om * (cx[k] - nn[k] * ncx) - nms * dn[r][k] - nmr * dn[s][k] - nn[k] * dd;
                        if (o->P.bending) {
                            double d2b[3];
                            for (int k = 0; k < 3; ++k) d2b[k] = q.ddR[l][k] * dn[s][c] + q.ddR[m][k] * dn[r][d] + dot3(q.h[k], d2n);
                            val += -(M[0] * d2b[0] + M[1] * d2b[1] + 2.0 * M[2] * d2b[2]);
                        }
                        /* follower pressure tangent: - p R_l dn_s[c] */
                        if (pr != 0.0) val -= pr * q.R[l] * dn[s][c];
                    }
                    Ke[(size_t)r * nd + s] += wJ * val;
                }
            }
        }
        if (!err) {
            for (int r = 0; r < nd; ++r) {
                if (ffull) {
#pragma omp atomic
                    ffull[(r % 3) * ncp + q.cpidx[r / 3]] += fi[r];
                }
                int gr = o->map[(r % 3) * ncp + q.cpidx[r / 3]];
                if (gr >= nf) continue;
                if (fint) {
#pragma omp atomic
                    fint[gr] += fe[r];
                }
                if (Ke) for (int s = 0; s < nd; ++s) {
                    int gs = o->map[(s % 3) * ncp + q.cpidx[s / 3]];
                    if (gs >= nf) {      /* eliminated column: its Dirichlet value times the entry goes to the lifting vector (set-up) */
                        if (o->lift && o->fixed) {
#pragma omp atomic
                            o->lift[gr] += Ke[(size_t)r * nd + s] * o->fixed[gs - nf];
                        }
                        continue;
                    }
                    long pos = find_pos(o, gr, gs);
#pragma omp atomic
                    Kval[pos] += Ke[(size_t)r * nd + s];
                }
            }
        }
        free(Ke);
    }
    free(disp);
    return err;
}

static int assemble(const klo* o, const double* x, double* Kval, double* fint) { return assemble_full(o, x, Kval, fint, NULL); }

/* external force: constant body force * N_i * meas(ori) + point loads (setPointLoads) */
static void build_fext(klo* o) {
    o->fext = (double*)calloc(o->nfree > 0 ? o->nfree : 1, sizeof(double));
    double* disp = (double*)calloc(3 * o->ncp, sizeof(double));
    int nq1 = o->P.quA * o->p[0] + o->P.quB, nq2 = o->P.quA * o->p[1] + o->P.quB;
    double xq1[MAXQ], wq1[MAXQ], xq2[MAXQ], wq2[MAXQ];
    gauss_legendre(nq1, xq1, wq1);
    gauss_legendre(nq2, xq2, wq2);
    const double* bf = o->P.body_force;
    qpdata q;
    if (bf[0] != 0.0 || bf[1] != 0.0 || bf[2] != 0.0)
        for (int e2 = 0; e2 < o->nel[1]; ++e2) for (int e1 = 0; e1 < o->nel[0]; ++e1) {
            double ua = o->U[0][o->span[0][e1]], ub = o->U[0][o->span[0][e1] + 1];
            double va = o->U[1][o->span[1][e2]], vb = o->U[1][o->span[1][e2] + 1];
            for (int q2 = 0; q2 < nq2; ++q2) for (int q1 = 0; q1 < nq1; ++q1) {
                double u = 0.5 * (ua + ub) + 0.5 * (ub - ua) * xq1[q1];
                double v = 0.5 * (va + vb) + 0.5 * (vb - va) * xq2[q2];
                double wt = 0.25 * (ub - ua) * (vb - va) * wq1[q1] * wq2[q2];
                eval_qp(o, e1, e2, u, v, disp, &q);
                double Nn[3]; cross3(q.A1, q.A2, Nn);
                double wJ = wt * sqrt(dot3(Nn, Nn));
                for (int l = 0; l < q.nloc; ++l) for (int c = 0; c < 3; ++c) {
                    int g = o->map[c * o->ncp + q.cpidx[l]];
                    if (g < o->nfree) o->fext[g] += wJ * q.R[l] * bf[c];
                }
            }
        }
    for (int k = 0; k < o->npl; ++k) {
        double u = o->pl_uv[2 * k], v = o->pl_uv[2 * k + 1];
        int s1 = find_span(o->n[0], o->p[0], u, o->U[0]), s2_ = find_span(o->n[1], o->p[1], v, o->U[1]);
        int e1 = 0, e2 = 0;
        for (int e = 0; e < o->nel[0]; ++e) if (o->span[0][e] == s1) e1 = e;
        for (int e = 0; e < o->nel[1]; ++e) if (o->span[1][e] == s2_) e2 = e;
        eval_qp(o, e1, e2, u, v, disp, &q);
        for (int l = 0; l < q.nloc; ++l) for (int c = 0; c < 3; ++c) {
            int g = o->map[c * o->ncp + q.cpidx[l]];
            if (g < o->nfree) o->fext[g] += q.R[l] * o->pl_val[3 * k + c];
        }
    }
    /* Neumann sides (BCs.addCondition(side, condition_type::neumann, &neuData), benchmarks/benchmark_Frustrum_APALM.cpp:236-242,267;
     * benchmark_Cylinder.cpp:118-129): F[i,c] += int_side N_i t_c |dX/dxi| dxi on the undeformed edge; the surface point evaluation
     * on the side supplies basis values and the tangent (A1 along south/north, A2 along west/east). */
    for (int k = 0; k < o->nneu; ++k) {
        int side = o->neu_side[k];
        int dir = (side == KL_WEST || side == KL_EAST) ? 1 : 0;
        int nqs = dir == 0 ? nq1 : nq2;
        const double* xq = dir == 0 ? xq1 : xq2;
        const double* wq = dir == 0 ? wq1 : wq2;
        double fixedpar = (side == KL_WEST || side == KL_SOUTH) ? o->U[1 - dir][o->p[1 - dir]] : o->U[1 - dir][o->n[1 - dir]];
        int efix = (side == KL_WEST || side == KL_SOUTH) ? 0 : o->nel[1 - dir] - 1;
        for (int e = 0; e < o->nel[dir]; ++e) {
            double ua = o->U[dir][o->span[dir][e]], ub = o->U[dir][o->span[dir][e] + 1];
            for (int qq = 0; qq < nqs; ++qq) {
                double t = 0.5 * (ua + ub) + 0.5 * (ub - ua) * xq[qq];
                double u = dir == 0 ? t : fixedpar, v = dir == 0 ? fixedpar : t;
                eval_qp(o, dir == 0 ? e : efix, dir == 0 ? efix : e, u, v, disp, &q);
                const double* tan = dir == 0 ? q.A1 : q.A2;
                double wJ = 0.5 * (ub - ua) * wq[qq] * sqrt(dot3(tan, tan));
                for (int l = 0; l < q.nloc; ++l) for (int c = 0; c < 3; ++c) {
                    int g = o->map[c * o->ncp + q.cpidx[l]];
                    if (g < o->nfree) o->fext[g] += wJ * q.R[l] * o->neu_val[3 * k + c];
                }
            }
        }
    }
    free(disp);
}

static int assemble_full(const klo* o, const double* x, double* Kval, double* fint, double* ffull);
/* Force = assemble(); rhs() of the LINEAR system at the undeformed geometry (benchmarks/benchmark_Balloon.cpp:262-263):
 * dead loads + follower pressure on the undeformed surface - lifting of non-zero Dirichlet values */
static void build_force(klo* o) {
    int n = o->nfree > 0 ? o->nfree : 1;
    o->force = (double*)malloc(sizeof(double) * n);
    memcpy(o->force, o->fext, sizeof(double) * n);
    int lifting = 0;
    if (o->fixed) for (int k = 0; k < o->nfixed; ++k) lifting |= o->fixed[k] != 0.0;
    if (o->P.pressure != 0.0) {
        double* r = (double*)calloc(n, sizeof(double));
        assemble_full(o, NULL, NULL, r, NULL);           /* F_int(0) - P(0) = -P(0) at the undeformed configuration */
        for (int i = 0; i < o->nfree; ++i) o->force[i] -= r[i];
        free(r);
    }
    if (lifting) {
        double* K = (double*)malloc(sizeof(double) * (o->nnz > 0 ? o->nnz : 1));
        double pr = o->P.pressure;
        o->P.pressure = 0.0;
        o->lift = (double*)calloc(n, sizeof(double));
        assemble_full(o, NULL, K, NULL, NULL);
        for (int i = 0; i < o->nfree; ++i) o->force[i] -= o->lift[i];
        free(o->lift); o->lift = NULL;
        o->P.pressure = pr;
        free(K);
    }
}

/* ------------------------------------------------------------------------------------ */
static double* dupd(const double* s, size_t n) { if (!s) return NULL; double* d = (double*)malloc(sizeof(double) * (n ? n : 1)); memcpy(d, s, sizeof(double) * n); return d; }

klo* klo_create(const kl_problem* P) {
    klo* o = (klo*)calloc(1, sizeof(klo));
    o->P = *P;
    if (o->P.quA == 0 && o->P.quB == 0) { o->P.quA = 1; o->P.quB = 1; }
    if (o->P.num_gauss_thickness <= 0) o->P.num_gauss_thickness = 4;
    for (int d = 0; d < 2; ++d) {
        o->p[d] = P->degree[d]; o->nk[d] = P->n_knots[d]; o->n[d] = o->nk[d] - o->p[d] - 1;
        o->U[d] = dupd(P->knots[d], o->nk[d]);
        o->span[d] = (int*)malloc(sizeof(int) * o->nk[d]);
        o->nel[d] = 0;
        for (int k = o->p[d]; k < o->n[d]; ++k) if (o->U[d][k + 1] > o->U[d][k]) o->span[d][o->nel[d]++] = k;
    }
    o->ncp = o->n[0] * o->n[1];
    o->cp = dupd(P->cp, 3 * (size_t)o->ncp);
    o->w = dupd(P->weights, o->ncp);
    o->map = (int*)malloc(sizeof(int) * 3 * o->ncp);
    memcpy(o->map, P->dof_map, sizeof(int) * 3 * o->ncp);
    o->nfree = P->n_free; o->nfixed = P->n_fixed;
    o->fixed = dupd(P->fixed_values, P->n_fixed);
    o->npl = P->n_point_loads;
    o->pl_uv = dupd(P->point_load_uv, 2 * (size_t)o->npl);
    o->pl_val = dupd(P->point_load_val, 3 * (size_t)o->npl);
    o->nneu = P->n_neumann;
    o->neu_side = NULL; o->neu_val = NULL;
    if (o->nneu > 0) {
        o->neu_side = (int*)malloc(sizeof(int) * o->nneu);
        memcpy(o->neu_side, P->neumann_side, sizeof(int) * o->nneu);
        o->neu_val = dupd(P->neumann_val, 3 * (size_t)o->nneu);
    }
    o->nthreads = 1;
#ifdef _OPENMP
    o->nthreads = omp_get_max_threads();
#endif
    o->e2_begin = 0; o->e2_end = o->nel[1];
    build_pattern(o);
    build_fext(o);
    build_force(o);
    return o;
}

void klo_destroy(klo* o) {
    if (!o) return;
    for (int d = 0; d < 2; ++d) { free(o->U[d]); free(o->span[d]); }
    free(o->cp); free(o->w); free(o->map); free(o->fixed); free(o->pl_uv); free(o->pl_val);
    free(o->outer); free(o->inner); free(o->fext); free(o->force); free(o->neu_side); free(o->neu_val); free(o);
}

void klo_set_strip(klo* o, int b, int e) { o->e2_begin = b; o->e2_end = e; }
void klo_set_threads(klo* o, int n) { o->nthreads = n > 0 ? n : 1; }
int klo_get_threads(const klo* o) { return o->nthreads; }

int klo_sizes(const klo* o, int* n_dofs, long* nnz, long* n_elements, long* n_qp) {
    *n_dofs = o->nfree; *nnz = o->nnz;
    *n_elements = (long)o->nel[0] * o->nel[1];
    *n_qp = *n_elements * (o->P.quA * o->p[0] + o->P.quB) * (o->P.quA * o->p[1] + o->P.quB);
    return 0;
}
int klo_pattern(const klo* o, int* outer, int* inner) {
    memcpy(outer, o->outer, sizeof(int) * (o->nfree + 1));
    memcpy(inner, o->inner, sizeof(int) * o->nnz);
    return 0;
}
int klo_jacobian(const klo* o, const double* x, double* values) { return assemble(o, x, values, NULL); }
/* dead loads only (body force, point loads, Neumann tractions): what rhs(x) = F_dead - (F_int - P)(x) is built from */
int klo_dead_force(const klo* o, double* f) { memcpy(f, o->fext, sizeof(double) * o->nfree); return 0; }
int klo_force(const klo* o, double* f) { memcpy(f, o->force, sizeof(double) * o->nfree); return 0; }
/* r = F_ext - F_int (assembleVector; rhs()) */
int klo_residual(const klo* o, const double* x, double* r) {
    int rc = assemble(o, x, NULL, r);
    for (int i = 0; i < o->nfree; ++i) r[i] = o->fext[i] - r[i];
    return rc;
}
/* r = Force - lam*Force - rhs(x)  (benchmarks/benchmark_Roof.cpp:335-344, benchmark_Balloon.cpp:285) */
int klo_al_residual(const klo* o, const double* x, double lam, double* r) {
    int rc = assemble(o, x, NULL, r);
    for (int i = 0; i < o->nfree; ++i) r[i] = (1.0 - lam) * o->force[i] - (o->fext[i] - r[i]);
    return rc;
}
/* both in one sweep (used by the CPU baseline timing: Jacobian + residual per "step") */
int klo_jacobian_residual(const klo* o, const double* x, double* values, double* r) {
    int rc = assemble(o, x, values, r);
    for (int i = 0; i < o->nfree; ++i) r[i] = o->fext[i] - r[i];
    return rc;
}

/* mass matrix (pattern of K; only c == d blocks are non-zero) and lumped mass = row sums over ALL basis functions
 * (assembleMass / assembleMass(true), unittests/gsStaticSolver_test.cpp:249-253) */
int klo_mass(const klo* o, double density, double* values, double* lumped) {
    double* disp = (double*)calloc(3 * o->ncp, sizeof(double));
    if (values) memset(values, 0, sizeof(double) * o->nnz);
    if (lumped) memset(lumped, 0, sizeof(double) * o->nfree);
    int nq1 = o->P.quA * o->p[0] + o->P.quB, nq2 = o->P.quA * o->p[1] + o->P.quB;
    double xq1[MAXQ], wq1[MAXQ], xq2[MAXQ], wq2[MAXQ];
    gauss_legendre(nq1, xq1, wq1);
    gauss_legendre(nq2, xq2, wq2);
    qpdata q;
    for (int e2 = 0; e2 < o->nel[1]; ++e2) for (int e1 = 0; e1 < o->nel[0]; ++e1) {
        double ua = o->U[0][o->span[0][e1]], ub = o->U[0][o->span[0][e1] + 1];
        double va = o->U[1][o->span[1][e2]], vb = o->U[1][o->span[1][e2] + 1];
        for (int q2 = 0; q2 < nq2; ++q2) for (int q1 = 0; q1 < nq1; ++q1) {
            double u = 0.5 * (ua + ub) + 0.5 * (ub - ua) * xq1[q1];
            double v = 0.5 * (va + vb) + 0.5 * (vb - va) * xq2[q2];
            double wt = 0.25 * (ub - ua) * (vb - va) * wq1[q1] * wq2[q2];
            eval_qp(o, e1, e2, u, v, disp, &q);
            double Nn[3]; cross3(q.A1, q.A2, Nn);
            double wJ = wt * sqrt(dot3(Nn, Nn)) * density * o->P.thickness;
            for (int a = 0; a < q.nloc; ++a) for (int c = 0; c < 3; ++c) {
                int gr = o->map[c * o->ncp + q.cpidx[a]];
                if (gr >= o->nfree) continue;
                for (int b = 0; b < q.nloc; ++b) {
                    double m = wJ * q.R[a] * q.R[b];
                    if (lumped) lumped[gr] += m;
                    int gc = o->map[c * o->ncp + q.cpidx[b]];
                    if (values && gc < o->nfree) values[find_pos(o, gr, gc)] += m;
                }
            }
        }
    }
    free(disp);
    return 0;
}

/* basis evaluation exported for tests against scipy.interpolate.BSpline */
int klo_basis_ders(int p, int nk, const double* U, double u, int* span_out, double* ders /* 3*(p+1) */) {
    int n = nk - p - 1;
    int s = find_span(n, p, u, U);
    double d[3][MAXP + 1];
    ders_basis(s, u, p, 2, U, d);
    for (int k = 0; k < 3; ++k) for (int j = 0; j <= p; ++j) ders[k * (p + 1) + j] = d[k][j];
    *span_out = s;
    return 0;
}
int klo_gauss(int n, double* x, double* w) { gauss_legendre(n, x, w); return 0; }

/* ---- linear solve of the Newton loop --------------------------------------------------------------
 * The reference's gsStaticNewton defaults to gsSparseSolver<>::CGDiagonal (src/gsStaticSolvers/gsStaticNewton.hpp:23), i.e.
 * Eigen::ConjugateGradient<SparseMatrix, Lower|Upper, DiagonalPreconditioner> of Eigen 3.4 (third-party; vendored by G+Smo
 * as gsEigen, absent from /root/reference).  This is a scalar restatement of Eigen's published iteration
 * (Eigen/src/IterativeLinearSolvers/ConjugateGradient.h, "conjugate_gradient"): x0 = 0, threshold
 * max(tol^2 |b|^2, DBL_MIN) on |r|^2, preconditioner 1/diag (1 where the diagonal vanishes), the iteration count is the
 * number of completed loop bodies (the converging one is not counted), error = sqrt(|r|^2/|b|^2).
 * The matrix is compressed-column (outer/inner/values); A*p is formed column-wise (the true product, no symmetry assumed). */
#include <float.h>
static void csc_mult(int n, const int* outer, const int* inner, const double* val, const double* p, double* y) {
    for (int i = 0; i < n; ++i) y[i] = 0.0;
    for (int j = 0; j < n; ++j) {
        const double pj = p[j];
        for (int k = outer[j]; k < outer[j + 1]; ++k) y[inner[k]] += val[k] * pj;
    }
}
int klo_cg_solve(int n, const int* outer, const int* inner, const double* val, const double* b, double* x, double tol, int max_iter,
                 int* iters, double* rel_err) {
    if (tol <= 0.0) tol = DBL_EPSILON;
    if (max_iter <= 0) max_iter = 2 * n;
    double* w = (double*)malloc(sizeof(double) * 5 * (size_t)(n > 0 ? n : 1));
    double *r = w, *p = w + n, *z = w + 2 * n, *tmp = w + 3 * n, *invd = w + 4 * n;
    for (int j = 0; j < n; ++j) {
        double dg = 0.0;
        for (int k = outer[j]; k < outer[j + 1]; ++k) if (inner[k] == j) { dg = val[k]; break; }
        invd[j] = dg != 0.0 ? 1.0 / dg : 1.0;
    }
    double rhs2 = 0.0;
    for (int i = 0; i < n; ++i) { x[i] = 0.0; r[i] = b[i]; rhs2 += b[i] * b[i]; }
    int i = 0;
    double rn2 = rhs2;
    if (rhs2 == 0.0) { *iters = 0; *rel_err = 0.0; free(w); return 0; }
    const double threshold = fmax(tol * tol * rhs2, DBL_MIN);
    if (rn2 >= threshold) {
        double absNew = 0.0;
        for (int k = 0; k < n; ++k) { p[k] = invd[k] * r[k]; absNew += r[k] * p[k]; }
        while (i < max_iter) {
            csc_mult(n, outer, inner, val, p, tmp);
            double pAp = 0.0;
            for (int k = 0; k < n; ++k) pAp += p[k] * tmp[k];
            const double alpha = absNew / pAp;
            rn2 = 0.0;
            for (int k = 0; k < n; ++k) { x[k] += alpha * p[k]; r[k] -= alpha * tmp[k]; rn2 += r[k] * r[k]; }
            if (rn2 < threshold) break;
            const double absOld = absNew;
            absNew = 0.0;
            for (int k = 0; k < n; ++k) { z[k] = invd[k] * r[k]; absNew += r[k] * z[k]; }
            const double beta = absNew / absOld;
            for (int k = 0; k < n; ++k) p[k] = z[k] + beta * p[k];
            ++i;
        }
    }
    *iters = i;
    *rel_err = sqrt(rn2 / rhs2);
    free(w);
    return 0;
}

/* ---- stress / stretch recovery --------------------------------------------------------------------
 * Restates what constructStress / computePrincipalStretches / boundaryForce are used for in the reference
 * (unittests/gsStaticSolver_test.cpp:313-324; benchmarks/benchmark_Balloon.cpp:381-408; benchmark_Pillow.cpp:431,484-505)
 * with the definitions written in include/kl_shell.h.  Deliberately a different route than the device code: everything
 * goes through full 3-D tensors (deformation gradient F = a_i (x) A^i with a_3 = lambda3 n, sigma = F S F^T / det F,
 * E = (F^T F - I)/2) projected on explicit orthonormal frames, and the principal values come from a Cholesky
 * transformation + Jacobi rotation instead of the closed-form roots of the characteristic polynomial. */
static double det3(double M[3][3]) {
    return M[0][0] * (M[1][1] * M[2][2] - M[1][2] * M[2][1]) - M[0][1] * (M[1][0] * M[2][2] - M[1][2] * M[2][0])
         + M[0][2] * (M[1][0] * M[2][1] - M[1][1] * M[2][0]);
}
/* eigen decomposition of the symmetric 2x2 (s11,s22,s12): ascending values w, orthonormal vectors V[k] */
static void eig2(double s11, double s22, double s12, double w[2], double V[2][2]) {
    double th = 0.5 * atan2(2.0 * s12, s11 - s22), c = cos(th), s = sin(th);
    double wa = c * c * s11 + 2 * c * s * s12 + s * s * s22, wb = s * s * s11 - 2 * c * s * s12 + c * c * s22;
    if (wa <= wb) { w[0] = wa; w[1] = wb; V[0][0] = c; V[0][1] = s; V[1][0] = -s; V[1][1] = c; }
    else { w[0] = wb; w[1] = wa; V[0][0] = -s; V[0][1] = c; V[1][0] = c; V[1][1] = s; }
}
int klo_stress_dim(int type) {
    static const int dim[KL_STRESS_NTYPES] = {3, 3, 3, 3, 3, 3, 3, 3, 9, 2, 2, 2, 2, 1, 1};
    return (type >= 0 && type < KL_STRESS_NTYPES) ? dim[type] : 0;
}
int klo_eval_stress(const klo* o, const double* x, int type, int npts, const double* uv, double z, double* out) {
    int dim = klo_stress_dim(type);
    if (!dim) return KL_E_ARG;
    double* disp = (double*)malloc(sizeof(double) * 3 * o->ncp);
    construct_disp(o, x, disp);
    int err = 0;
    for (int k = 0; k < npts && !err; ++k) {
        double u = uv[2 * k], v = uv[2 * k + 1], *res = out + (size_t)dim * k;
        int s1 = find_span(o->n[0], o->p[0], u, o->U[0]), s2_ = find_span(o->n[1], o->p[1], v, o->U[1]);
        int e1 = 0, e2 = 0;
        for (int e = 0; e < o->nel[0]; ++e) if (o->span[0][e] == s1) e1 = e;
        for (int e = 0; e < o->nel[1]; ++e) if (o->span[1][e] == s2_) e2 = e;
        qpdata q;
        eval_qp(o, e1, e2, u, v, disp, &q);
        if (type == KL_STRESS_DISPLACEMENT) {
            for (int c = 0; c < 3; ++c) { double s = 0; for (int l = 0; l < q.nloc; ++l) s += q.R[l] * disp[3 * q.cpidx[l] + c]; res[c] = s; }
            continue;
        }
        double Nn[3], nn[3];
        cross3(q.A1, q.A2, Nn); cross3(q.a1, q.a2, nn);
        double JA = sqrt(dot3(Nn, Nn)), Ja = sqrt(dot3(nn, nn));
        if (!(JA > 0.0) || !(Ja > 0.0)) { err = KL_E_JACOBIAN; break; }
        for (int c = 0; c < 3; ++c) { Nn[c] /= JA; nn[c] /= Ja; }
        double Ac[3] = {dot3(q.A1, q.A1), dot3(q.A2, q.A2), dot3(q.A1, q.A2)};
        double ac[3] = {dot3(q.a1, q.a1), dot3(q.a2, q.a2), dot3(q.a1, q.a2)};
        double Bc[3], bc[3];
        for (int i = 0; i < 3; ++i) { Bc[i] = dot3(q.H[i], Nn); bc[i] = dot3(q.h[i], nn); }
        if (!o->P.bending) for (int i = 0; i < 3; ++i) { Bc[i] = 0; bc[i] = 0; }
        double Ai[3], ai[3], dA, da;
        inv2(Ac, Ai, &dA); inv2(ac, ai, &da);
        /* contravariant vectors of the undeformed mid-surface, orthonormal frames (E1,E2,N), (e1,e2,n) */
        double Au[2][3], E1[3], E2[3], e1v[3], e2v[3];
        for (int c = 0; c < 3; ++c) { Au[0][c] = Ai[0] * q.A1[c] + Ai[2] * q.A2[c]; Au[1][c] = Ai[2] * q.A1[c] + Ai[1] * q.A2[c]; }
        for (int c = 0; c < 3; ++c) { E1[c] = q.A1[c] / sqrt(Ac[0]); e1v[c] = q.a1[c] / sqrt(ac[0]); }
        cross3(Nn, E1, E2); cross3(nn, e1v, e2v);
        /* thickness stretch at the mid-surface */
        double J0sq0 = da / dA, c33mid = 1.0 / J0sq0;
        if (o->P.material != KL_MAT_SVK && o->P.compressible) {
            double S_[3], C_[3][3];
            int rc = hyper_point(&o->P, Ac, ac, S_, C_, &c33mid);
            if (rc) { err = rc; break; }
        }
        double lam3mid = sqrt(c33mid);
        /* F = a_1 (x) A^1 + a_2 (x) A^2 + lambda3 n (x) N */
        double F[3][3];
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) F[i][j] = q.a1[i] * Au[0][j] + q.a2[i] * Au[1][j] + lam3mid * nn[i] * Nn[j];
        double J = det3(F);
        if (type == KL_STRESS_PRINCIPAL_STRETCH || type == KL_STRESS_PRINCIPAL_STRETCH_DIR) {
            /* metric at height z exactly as the material law sees it (g_ab - 2 z b_ab [+ z^2 b_ac a^cd b_db]) */
            double Gc[3], gc[3];
            for (int i = 0; i < 3; ++i) { Gc[i] = Ac[i] - 2.0 * z * Bc[i]; gc[i] = ac[i] - 2.0 * z * bc[i]; }
            if (o->P.metric_z2)
                for (int i = 0; i < 3; ++i) {
                    int a = VI[i], b = VJ[i];
                    double sG = 0, sg = 0;
                    for (int c = 0; c < 2; ++c) for (int d = 0; d < 2; ++d) {
                        sG += s2(Bc, a, c) * s2(Ai, c, d) * s2(Bc, d, b);
                        sg += s2(bc, a, c) * s2(ai, c, d) * s2(bc, d, b);
                    }
                    Gc[i] += z * z * sG; gc[i] += z * z * sg;
                }
            double dG = Gc[0] * Gc[1] - Gc[2] * Gc[2], dg = gc[0] * gc[1] - gc[2] * gc[2];
            if (!(dG > 0.0) || !(dg > 0.0)) { err = KL_E_JACOBIAN; break; }
            /* Cholesky G = L L^T;  Chat = L^-1 g L^-T is the right Cauchy-Green tensor in an orthonormal frame */
            double l11 = sqrt(Gc[0]), l21 = Gc[2] / l11, l22 = sqrt(Gc[1] - l21 * l21);
            double m11 = gc[0] / (l11 * l11);
            double m12 = (gc[2] - l21 * gc[0] / l11) / (l11 * l22);
            double m22 = (gc[1] - 2.0 * l21 * gc[2] / l11 + l21 * l21 * gc[0] / (l11 * l11)) / (l22 * l22);
            double w[2], V[2][2];
            eig2(m11, m22, m12, w, V);
            double c33 = dG / dg;
            if (o->P.material != KL_MAT_SVK && o->P.compressible) {
                double S_[3], C_[3][3];
                int rc = hyper_point(&o->P, Gc, gc, S_, C_, &c33);
                if (rc) { err = rc; break; }
            }
            if (type == KL_STRESS_PRINCIPAL_STRETCH) { res[0] = sqrt(w[0]); res[1] = sqrt(w[1]); res[2] = sqrt(c33); continue; }
            /* contravariant components v = L^-T what; spatial direction n_i = v^a g_a(z) / lambda_i, g_a(z) = a_a - z b_a^c a_c */
            double gz[2][3];
            {
                double bm[2][2] = {{bc[0] * ai[0] + bc[2] * ai[2], bc[0] * ai[2] + bc[2] * ai[1]},
                                   {bc[2] * ai[0] + bc[1] * ai[2], bc[2] * ai[2] + bc[1] * ai[1]}};   /* b_a^c */
                for (int c = 0; c < 3; ++c) {
                    gz[0][c] = q.a1[c] - z * (bm[0][0] * q.a1[c] + bm[0][1] * q.a2[c]);
                    gz[1][c] = q.a2[c] - z * (bm[1][0] * q.a1[c] + bm[1][1] * q.a2[c]);
                }
            }
            for (int i = 0; i < 2; ++i) {
                double v2 = V[i][1] / l22, v1 = (V[i][0] - l21 * v2) / l11;
                double d[3], nrm;
                for (int c = 0; c < 3; ++c) d[c] = v1 * gz[0][c] + v2 * gz[1][c];
                nrm = sqrt(dot3(d, d));
                for (int c = 0; c < 3; ++c) res[3 * i + c] = d[c] / nrm;
            }
            for (int c = 0; c < 3; ++c) res[6 + c] = nn[c];
            continue;
        }
        /* strains: 3-D Green-Lagrange tensor and the curvature change tensor, projected on (E1,E2) */
        double Em[3] = {0, 0, 0}, Ef[3] = {0, 0, 0};
        {
            double E3[3][3], K3[3][3];
            for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
                double s = 0;
                for (int m = 0; m < 3; ++m) s += F[m][i] * F[m][j];
                E3[i][j] = 0.5 * (s - (i == j ? 1.0 : 0.0));
                double kk = 0;
                for (int a = 0; a < 2; ++a) for (int b = 0; b < 2; ++b) kk += (s2(Bc, a, b) - s2(bc, a, b)) * Au[a][i] * Au[b][j];
                K3[i][j] = kk;
            }
            const double* Fr[2] = {E1, E2};
            for (int vv = 0; vv < 3; ++vv) {
                const double *p_ = Fr[VI[vv]], *q_ = Fr[VJ[vv]];
                for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { Em[vv] += p_[i] * E3[i][j] * q_[j]; Ef[vv] += p_[i] * K3[i][j] * q_[j]; }
            }
        }
        double A[3][3], B[3][3], D[3][3], N[3], M[3];
        int rc = material_eval(&o->P, Ac, Bc, ac, bc, A, B, D, N, M);
        if (rc) { err = rc; break; }
        if (!o->P.bending) M[0] = M[1] = M[2] = 0.0;
        /* Cauchy tensors sigma = F S F^T / J with S = N^ab/t A_a (x) A_b, projected on (e1,e2) */
        double sm[3] = {0, 0, 0}, sf[3] = {0, 0, 0};
        {
            const double* Acov[2] = {q.A1, q.A2};
            double S3[3][3], M3[3][3], t = o->P.thickness;
            for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
                double a_ = 0, b_ = 0;
                for (int a = 0; a < 2; ++a) for (int b = 0; b < 2; ++b) { a_ += s2(N, a, b) * Acov[a][i] * Acov[b][j]; b_ += s2(M, a, b) * Acov[a][i] * Acov[b][j]; }
                S3[i][j] = a_ / t; M3[i][j] = 6.0 * b_ / (t * t);
            }
            const double* fr[2] = {e1v, e2v};
            for (int vv = 0; vv < 3; ++vv) {
                double Fp[3] = {0, 0, 0}, Fq[3] = {0, 0, 0};   /* F^T e */
                for (int i = 0; i < 3; ++i) for (int m = 0; m < 3; ++m) { Fp[i] += F[m][i] * fr[VI[vv]][m]; Fq[i] += F[m][i] * fr[VJ[vv]][m]; }
                for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { sm[vv] += Fp[i] * S3[i][j] * Fq[j] / J; sf[vv] += Fp[i] * M3[i][j] * Fq[j] / J; }
            }
        }
        double w[2], V[2][2];
        switch (type) {
            case KL_STRESS_MEMBRANE_FORCE: for (int i = 0; i < 3; ++i) res[i] = N[i]; break;
            case KL_STRESS_FLEXURAL_MOMENT: for (int i = 0; i < 3; ++i) res[i] = M[i]; break;
            case KL_STRESS_MEMBRANE: for (int i = 0; i < 3; ++i) res[i] = sm[i]; break;
            case KL_STRESS_FLEXURAL: for (int i = 0; i < 3; ++i) res[i] = sf[i]; break;
            case KL_STRESS_MEMBRANE_STRAIN: for (int i = 0; i < 3; ++i) res[i] = Em[i]; break;
            case KL_STRESS_FLEXURAL_STRAIN: for (int i = 0; i < 3; ++i) res[i] = Ef[i]; break;
            case KL_STRESS_PRINCIPAL_STRESS_MEMBRANE: eig2(sm[0], sm[1], sm[2], w, V); res[0] = w[0]; res[1] = w[1]; break;
            case KL_STRESS_PRINCIPAL_STRESS_FLEXURAL: eig2(sf[0], sf[1], sf[2], w, V); res[0] = w[0]; res[1] = w[1]; break;
            case KL_STRESS_PRINCIPAL_MEMBRANE_STRAIN: eig2(Em[0], Em[1], Em[2], w, V); res[0] = w[0]; res[1] = w[1]; break;
            case KL_STRESS_PRINCIPAL_FLEXURAL_STRAIN: eig2(Ef[0], Ef[1], Ef[2], w, V); res[0] = w[0]; res[1] = w[1]; break;
            case KL_STRESS_VON_MISES_MEMBRANE: res[0] = sqrt(sm[0] * sm[0] + sm[1] * sm[1] - sm[0] * sm[1] + 3.0 * sm[2] * sm[2]); break;
            case KL_STRESS_TENSION_FIELD: {
                double ws[2], we[2];
                eig2(sm[0], sm[1], sm[2], ws, V);
                eig2(Em[0], Em[1], Em[2], we, V);
                res[0] = ws[0] > 0.0 ? 1.0 : (we[1] <= 0.0 ? -1.0 : 0.0);
            } break;
            default: err = KL_E_ARG;
        }
    }
    free(disp);
    return err;
}
/* boundaryForce(mp_def, patchSide(0, side)): MINUS the internal force summed per component over all control points of the side */
int klo_boundary_force(const klo* o, const double* x, int side, double* out3) {
    double* ff = (double*)calloc((size_t)3 * o->ncp, sizeof(double));
    int rc = assemble_full(o, x, NULL, NULL, ff);
    int n1 = o->n[0], n2 = o->n[1];
    for (int c = 0; c < 3; ++c) {
        double s = 0.0;
        if (side == KL_WEST || side == KL_EAST) { int i1 = side == KL_WEST ? 0 : n1 - 1; for (int i2 = 0; i2 < n2; ++i2) s += ff[c * o->ncp + i1 + n1 * i2]; }
        else { int i2 = side == KL_SOUTH ? 0 : n2 - 1; for (int i1 = 0; i1 < n1; ++i1) s += ff[c * o->ncp + i1 + n1 * i2]; }
        out3[c] = -s;   /* sign of rhs() = F_ext - F_int: the reference forms S = -sideForce / area (unittests/gsStaticSolver_test.cpp:323) */
    }
    free(ff);
    return rc;
}
