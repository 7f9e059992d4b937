/*
 * ks_oracle.c — CPU ORACLE for the gsElasticity solid assembly path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT THE PRODUCT.  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may load it.
 *
 * PARITY UNPINNED: the arithmetic lives in gismo/gsElasticity@master (un-pinned branch head,
 * .github/workflows/ci.yml:40-41 of the reference), absent from /root/reference and not buildable here.  This file
 * restates the textbook total-Lagrangian formulation that module implements (gsElasticityAssembler::assemble(x, fixedDofs)
 * with gsVisitorNonLinearElasticity, [UPSTREAM-RECALLED]): per Gauss point F = I + grad u, C = F^T F, E = (C - I)/2,
 *   saint_venant_kirchhoff: S = lambda tr(E) I + 2 mu E
 *   neo_hooke_ln:           S = (lambda ln J - mu) C^-1 + mu I
 *   neo_hooke_quad:         S = (lambda (J^2-1)/2 - mu) C^-1 + mu I
 * K_ab = int B_a^T CC B_b + (grad N_a . S grad N_b) I,  rhs_a = F_ext - int B_a^T S, in the brute-force Voigt/B-matrix form
 * (the GPU path uses a different, tensor-contracted form).  Anchors: K = -d(rhs)/du by finite differences, F_int = dW/du of
 * an independently coded discrete energy, the cantilever tip deflection of beam theory (tests/test_oracle_solid.py).
 *
 * Reference call sites followed: tutorials/nonlinear_solid_static.cpp:92-121 (assembler, options, closures, assemble()),
 * benchmarks/benchmark_Elasticity_Beam_APALM.cpp:226-236 (Dirichlet side + Neumann traction), :307-325 (AL residual).
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "../include/ks_solid.h"
#define MAXP 5
#include "bspline_common.h"

#define KS_MAXP 4
#define KS_MAXLOC ((KS_MAXP + 1) * (KS_MAXP + 1) * (KS_MAXP + 1))

typedef struct kso {
    int p[3], nk[3], n[3], nel[3];
    double* U[3];
    int* span[3];
    int ncp;
    double* cp;
    int* map;
    int nfree, nfixed;
    double* fixed;
    int law;
    double lambda, mu;
    int* outer;
    int* inner;
    long nnz;
    double* fext;
    int nthreads;
} kso;

int kso_build_dofmap(int n1, int n2, int n3, const ks_bc* bc, int* map, int* n_free, int* n_fixed) {
    const int ncp = n1 * n2 * n3;
    const int n[3] = {n1, n2, n3};
    char* elim = (char*)calloc((size_t)3 * ncp, 1);
    for (int c = 0; c < 3; ++c)
        for (int i3 = 0; i3 < n3; ++i3)
            for (int i2 = 0; i2 < n2; ++i2)
                for (int i1 = 0; i1 < n1; ++i1) {
                    const int id[3] = {i1, i2, i3};
                    int e = 0;
                    for (int d = 0; d < 3; ++d) {
                        if (id[d] == 0 && bc->side[2 * d][c]) e = 1;
                        if (id[d] == n[d] - 1 && bc->side[2 * d + 1][c]) e = 1;
                    }
                    int corner = -1;
                    if ((i1 == 0 || i1 == n1 - 1) && (i2 == 0 || i2 == n2 - 1) && (i3 == 0 || i3 == n3 - 1))
                        corner = (i1 ? 1 : 0) | (i2 ? 2 : 0) | (i3 ? 4 : 0);
                    if (corner >= 0 && bc->corner[corner][c]) e = 1;
                    elim[c * ncp + i1 + n1 * (i2 + n2 * i3)] = (char)e;
                }
    int nf = 0;
    for (int k = 0; k < 3 * ncp; ++k) if (!elim[k]) map[k] = nf++;
    int ne = 0;
    for (int k = 0; k < 3 * ncp; ++k) if (elim[k]) map[k] = nf + ne++;
    *n_free = nf; *n_fixed = ne;
    free(elim);
    return 0;
}

static int cmp_int(const void* a, const void* b) { int x = *(const int*)a, y = *(const int*)b; return (x > y) - (x < y); }

static void build_pattern(kso* o) {
    /* node J couples with every node that shares a non-empty element with it */
    int *lo[3], *hi[3];
    for (int d = 0; d < 3; ++d) {
        lo[d] = (int*)malloc(sizeof(int) * o->n[d]);
        hi[d] = (int*)malloc(sizeof(int) * o->n[d]);
        for (int i = 0; i < o->n[d]; ++i) { lo[d][i] = o->n[d]; hi[d][i] = -1; }
        for (int e = 0; e < o->nel[d]; ++e) {
            const int s = o->span[d][e];
            for (int i = s - o->p[d]; i <= s; ++i) {       /* nodes active on this element see nodes s-p..s */
                if (s - o->p[d] < lo[d][i]) lo[d][i] = s - o->p[d];
                if (s > hi[d][i]) hi[d][i] = s;
            }
        }
    }
    o->outer = (int*)calloc((size_t)o->nfree + 1, sizeof(int));
    const int n1 = o->n[0], n2 = o->n[1];
    for (int pass = 0; pass < 2; ++pass) {
        for (int J = 0; J < o->ncp; ++J) {
            const int j1 = J % n1, j2 = (J / n1) % n2, j3 = J / (n1 * n2);
            for (int d = 0; d < 3; ++d) {
                const int col = o->map[d * o->ncp + J];
                if (col >= o->nfree) continue;
                long cnt = 0;
                int* dst = pass ? o->inner + o->outer[col] : NULL;
                for (int c = 0; c < 3; ++c)
                    for (int i3 = lo[2][j3]; i3 <= hi[2][j3]; ++i3)
                        for (int i2 = lo[1][j2]; i2 <= hi[1][j2]; ++i2)
                            for (int i1 = lo[0][j1]; i1 <= hi[0][j1]; ++i1) {
                                const int row = o->map[c * o->ncp + i1 + n1 * (i2 + n2 * i3)];
                                if (row >= o->nfree) continue;
                                if (pass) dst[cnt] = row;
                                ++cnt;
                            }
                if (!pass) o->outer[col + 1] = (int)cnt;
                else qsort(dst, (size_t)cnt, sizeof(int), cmp_int);
            }
        }
        if (!pass) {
            for (int k = 0; k < o->nfree; ++k) o->outer[k + 1] += o->outer[k];
            o->nnz = o->outer[o->nfree];
            o->inner = (int*)malloc(sizeof(int) * (size_t)(o->nnz > 0 ? o->nnz : 1));
        }
    }
    for (int d = 0; d < 3; ++d) { free(lo[d]); free(hi[d]); }
}

static inline long find_pos(const kso* o, int row, int col) {
    long lo = o->outer[col], hi = o->outer[col + 1] - 1;
    while (lo <= hi) {
        const long mid = (lo + hi) / 2;
        if (o->inner[mid] == row) return mid;
        if (o->inner[mid] < row) lo = mid + 1; else hi = mid - 1;
    }
    return -1;
}

static double det3(const double A[3][3]) {
    return A[0][0] * (A[1][1] * A[2][2] - A[1][2] * A[2][1]) - A[0][1] * (A[1][0] * A[2][2] - A[1][2] * A[2][0]) +
           A[0][2] * (A[1][0] * A[2][1] - A[1][1] * A[2][0]);
}
static void inv3(const double A[3][3], double det, double B[3][3]) {
    B[0][0] = (A[1][1] * A[2][2] - A[1][2] * A[2][1]) / det;
    B[0][1] = (A[0][2] * A[2][1] - A[0][1] * A[2][2]) / det;
    B[0][2] = (A[0][1] * A[1][2] - A[0][2] * A[1][1]) / det;
    B[1][0] = (A[1][2] * A[2][0] - A[1][0] * A[2][2]) / det;
    B[1][1] = (A[0][0] * A[2][2] - A[0][2] * A[2][0]) / det;
    B[1][2] = (A[0][2] * A[1][0] - A[0][0] * A[1][2]) / det;
    B[2][0] = (A[1][0] * A[2][1] - A[1][1] * A[2][0]) / det;
    B[2][1] = (A[0][1] * A[2][0] - A[0][0] * A[2][1]) / det;
    B[2][2] = (A[0][0] * A[1][1] - A[0][1] * A[1][0]) / det;
}

static const int VI[6] = {0, 1, 2, 0, 1, 0}, VJ[6] = {0, 1, 2, 1, 2, 2};   /* Voigt pairs 11 22 33 12 23 13 */

/* second Piola-Kirchhoff stress and material tangent (Voigt 6x6, tensor components) at deformation gradient F */
static int material(const kso* o, const double F[3][3], double S[3][3], double CC[6][6]) {
    const double lam = o->lambda, mu = o->mu;
    double C[3][3], I3[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) { C[i][j] = 0; for (int k = 0; k < 3; ++k) C[i][j] += F[k][i] * F[k][j]; }
    if (o->law == KS_LAW_SVK || o->law == KS_LAW_HOOKE) {
        double E[3][3], tr = 0;
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j)
                E[i][j] = o->law == KS_LAW_HOOKE ? 0.5 * (F[i][j] + F[j][i]) - I3[i][j] : 0.5 * (C[i][j] - I3[i][j]);
        for (int i = 0; i < 3; ++i) tr += E[i][i];
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) S[i][j] = lam * tr * I3[i][j] + 2 * mu * E[i][j];
        for (int v = 0; v < 6; ++v)
            for (int w = 0; w < 6; ++w) {
                const int i = VI[v], j = VJ[v], k = VI[w], l = VJ[w];
                CC[v][w] = lam * I3[i][j] * I3[k][l] + mu * (I3[i][k] * I3[j][l] + I3[i][l] * I3[j][k]);
            }
        return 0;
    }
    const double J = det3(F);
    if (!(J > 0.0)) return 1;
    double Ci[3][3];
    inv3(C, det3(C), Ci);
    double a, b;    /* S = a C^-1 + mu I; CC = b C^-1 x C^-1 - a (C^-1 . C^-1) */
    if (o->law == KS_LAW_NEO_HOOKE_LN) { a = lam * log(J) - mu; b = lam; }
    else { a = lam * (J * J - 1) / 2 - mu; b = lam * J * J; }
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) S[i][j] = a * Ci[i][j] + mu * I3[i][j];
    for (int v = 0; v < 6; ++v)
        for (int w = 0; w < 6; ++w) {
            const int i = VI[v], j = VJ[v], k = VI[w], l = VJ[w];
            CC[v][w] = b * Ci[i][j] * Ci[k][l] - a * (Ci[i][k] * Ci[j][l] + Ci[i][l] * Ci[j][k]);
        }
    return 0;
}

/* basis values and parametric gradients of the active functions at (u,v,w); returns the node indices */
static int eval_basis(const kso* o, const int s[3], const double uvw[3], double* N, double (*dN)[3], int* node) {
    double d[3][3][MAXP + 1];
    for (int k = 0; k < 3; ++k) ders_basis(s[k], uvw[k], o->p[k], 1, o->U[k], d[k]);
    int a = 0;
    for (int a3 = 0; a3 <= o->p[2]; ++a3)
        for (int a2 = 0; a2 <= o->p[1]; ++a2)
            for (int a1 = 0; a1 <= o->p[0]; ++a1, ++a) {
                N[a] = d[0][0][a1] * d[1][0][a2] * d[2][0][a3];
                dN[a][0] = d[0][1][a1] * d[1][0][a2] * d[2][0][a3];
                dN[a][1] = d[0][0][a1] * d[1][1][a2] * d[2][0][a3];
                dN[a][2] = d[0][0][a1] * d[1][0][a2] * d[2][1][a3];
                node[a] = (s[0] - o->p[0] + a1) + o->n[0] * ((s[1] - o->p[1] + a2) + o->n[1] * (s[2] - o->p[2] + a3));
            }
    return a;
}

static void build_fext(kso* o, const ks_problem* P) {
    o->fext = (double*)calloc((size_t)(o->nfree > 0 ? o->nfree : 1), sizeof(double));
    double xq[3][MAXP + 2], wq[3][MAXP + 2];
    for (int d = 0; d < 3; ++d) gauss_legendre(o->p[d] + 1, xq[d], wq[d]);
    double N[KS_MAXLOC], dN[KS_MAXLOC][3];
    int node[KS_MAXLOC];
    const double* bf = P->body_force;
    if (bf[0] != 0 || bf[1] != 0 || bf[2] != 0)
        for (int e3 = 0; e3 < o->nel[2]; ++e3)
            for (int e2 = 0; e2 < o->nel[1]; ++e2)
                for (int e1 = 0; e1 < o->nel[0]; ++e1) {
                    const int s[3] = {o->span[0][e1], o->span[1][e2], o->span[2][e3]};
                    double a[3], h[3];
                    for (int d = 0; d < 3; ++d) { a[d] = o->U[d][s[d]]; h[d] = o->U[d][s[d] + 1] - a[d]; }
                    for (int q3 = 0; q3 <= o->p[2]; ++q3)
                        for (int q2 = 0; q2 <= o->p[1]; ++q2)
                            for (int q1 = 0; q1 <= o->p[0]; ++q1) {
                                const double uvw[3] = {a[0] + 0.5 * h[0] * (xq[0][q1] + 1), a[1] + 0.5 * h[1] * (xq[1][q2] + 1),
                                                       a[2] + 0.5 * h[2] * (xq[2][q3] + 1)};
                                const int nl = eval_basis(o, s, uvw, N, dN, node);
                                double Jg[3][3] = {{0}};
                                for (int b = 0; b < nl; ++b)
                                    for (int k = 0; k < 3; ++k) for (int l = 0; l < 3; ++l) Jg[k][l] += o->cp[3 * node[b] + k] * dN[b][l];
                                const double w = wq[0][q1] * wq[1][q2] * wq[2][q3] * 0.125 * h[0] * h[1] * h[2] * fabs(det3(Jg));
                                for (int b = 0; b < nl; ++b)
                                    for (int c = 0; c < 3; ++c) {
                                        const int g = o->map[c * o->ncp + node[b]];
                                        if (g < o->nfree) o->fext[g] += w * N[b] * bf[c];
                                    }
                            }
                }
    for (int t = 0; t < P->n_tractions; ++t) {
        const int side = P->traction_side[t], dn = side / 2, hiSide = side & 1;
        const int da = (dn + 1) % 3, db = (dn + 2) % 3;
        const double* tv = P->traction_val + 3 * t;
        const int en = hiSide ? o->nel[dn] - 1 : 0;
        for (int ea = 0; ea < o->nel[da]; ++ea)
            for (int eb = 0; eb < o->nel[db]; ++eb) {
                int s[3];
                s[dn] = o->span[dn][en]; s[da] = o->span[da][ea]; s[db] = o->span[db][eb];
                const double aa = o->U[da][s[da]], ha = o->U[da][s[da] + 1] - aa, ab = o->U[db][s[db]], hb = o->U[db][s[db] + 1] - ab;
                for (int qa = 0; qa <= o->p[da]; ++qa)
                    for (int qb = 0; qb <= o->p[db]; ++qb) {
                        double uvw[3];
                        uvw[dn] = hiSide ? o->U[dn][o->nk[dn] - 1] : o->U[dn][0];
                        uvw[da] = aa + 0.5 * ha * (xq[da][qa] + 1);
                        uvw[db] = ab + 0.5 * hb * (xq[db][qb] + 1);
                        const int nl = eval_basis(o, s, uvw, N, dN, node);
                        double ta[3] = {0, 0, 0}, tb[3] = {0, 0, 0};
                        for (int b = 0; b < nl; ++b)
                            for (int k = 0; k < 3; ++k) { ta[k] += o->cp[3 * node[b] + k] * dN[b][da]; tb[k] += o->cp[3 * node[b] + k] * dN[b][db]; }
                        const double nx = ta[1] * tb[2] - ta[2] * tb[1], ny = ta[2] * tb[0] - ta[0] * tb[2], nz = ta[0] * tb[1] - ta[1] * tb[0];
                        const double w = wq[da][qa] * wq[db][qb] * 0.25 * ha * hb * sqrt(nx * nx + ny * ny + nz * nz);
                        for (int b = 0; b < nl; ++b)
                            for (int c = 0; c < 3; ++c) {
                                const int g = o->map[c * o->ncp + node[b]];
                                if (g < o->nfree) o->fext[g] += w * N[b] * tv[c];
                            }
                    }
            }
    }
}

kso* kso_create(const ks_problem* P) {
    if (P->weights) return NULL;
    kso* o = (kso*)calloc(1, sizeof(kso));
    o->ncp = 1;
    for (int d = 0; d < 3; ++d) {
        o->p[d] = P->degree[d]; o->nk[d] = P->n_knots[d]; o->n[d] = o->nk[d] - o->p[d] - 1;
        if (o->p[d] < 1 || o->p[d] > KS_MAXP) { free(o); return NULL; }
        o->U[d] = (double*)malloc(sizeof(double) * o->nk[d]);
        memcpy(o->U[d], P->knots[d], sizeof(double) * o->nk[d]);
        o->span[d] = (int*)malloc(sizeof(int) * o->nk[d]);
        o->nel[d] = 0;
        for (int k = o->p[d]; k < o->n[d]; ++k) if (o->U[d][k + 1] > o->U[d][k]) o->span[d][o->nel[d]++] = k;
        o->ncp *= o->n[d];
    }
    o->cp = (double*)malloc(sizeof(double) * 3 * o->ncp);
    memcpy(o->cp, P->cp, sizeof(double) * 3 * o->ncp);
    o->map = (int*)malloc(sizeof(int) * 3 * o->ncp);
    memcpy(o->map, P->dof_map, sizeof(int) * 3 * o->ncp);
    o->nfree = P->n_free; o->nfixed = P->n_fixed;
    o->fixed = (double*)calloc((size_t)(o->nfixed > 0 ? o->nfixed : 1), sizeof(double));
    if (P->fixed_values) memcpy(o->fixed, P->fixed_values, sizeof(double) * o->nfixed);
    o->law = P->material_law;
    o->lambda = P->E * P->nu / ((1 + P->nu) * (1 - 2 * P->nu));
    o->mu = P->E / (2 * (1 + P->nu));
    o->nthreads = 1;
#ifdef _OPENMP
    o->nthreads = omp_get_max_threads();
#endif
    build_pattern(o);
    build_fext(o, P);
    return o;
}

void kso_destroy(kso* o) {
    if (!o) return;
    for (int d = 0; d < 3; ++d) { free(o->U[d]); free(o->span[d]); }
    free(o->cp); free(o->map); free(o->fixed); free(o->outer); free(o->inner); free(o->fext);
    free(o);
}
void kso_set_threads(kso* o, int n) { o->nthreads = n > 0 ? n : 1; }
int kso_get_threads(const kso* o) { return o->nthreads; }
int kso_sizes(const kso* o, int* n_dofs, long* nnz, long* n_elements, long* n_qp) {
    const long ne = (long)o->nel[0] * o->nel[1] * o->nel[2];
    *n_dofs = o->nfree; *nnz = o->nnz; *n_elements = ne;
    *n_qp = ne * (o->p[0] + 1) * (o->p[1] + 1) * (o->p[2] + 1);
    return 0;
}
int kso_pattern(const kso* o, int* outer, int* inner) {
    memcpy(outer, o->outer, sizeof(int) * ((size_t)o->nfree + 1));
    memcpy(inner, o->inner, sizeof(int) * (size_t)o->nnz);
    return 0;
}
int kso_force(const kso* o, double* f) { memcpy(f, o->fext, sizeof(double) * o->nfree); return 0; }

/* assemble(x, fixedDofs): values (may be NULL) = K(x); r (may be NULL) = F_ext - F_int(x); energy (may be NULL) = stored energy */
int kso_assemble(const kso* o, const double* x, double* values, double* r, double* energy) {
    if (values) memset(values, 0, sizeof(double) * (size_t)o->nnz);
    if (r) memset(r, 0, sizeof(double) * (size_t)o->nfree);
    double* disp = (double*)malloc(sizeof(double) * 3 * o->ncp);
    for (int c = 0; c < 3; ++c)
        for (int i = 0; i < o->ncp; ++i) {
            const int g = o->map[c * o->ncp + i];
            disp[3 * i + c] = g < o->nfree ? (x ? x[g] : 0.0) : o->fixed[g - o->nfree];
        }
    double xq[3][MAXP + 2], wq[3][MAXP + 2];
    for (int d = 0; d < 3; ++d) gauss_legendre(o->p[d] + 1, xq[d], wq[d]);
    const long nel = (long)o->nel[0] * o->nel[1] * o->nel[2];
    int bad = 0;
    double etot = 0.0;
#pragma omp parallel for schedule(dynamic, 4) num_threads(o->nthreads) reduction(| : bad) reduction(+ : etot)
    for (long e = 0; e < nel; ++e) {
        const int e1 = (int)(e % o->nel[0]), e2 = (int)((e / o->nel[0]) % o->nel[1]), e3 = (int)(e / ((long)o->nel[0] * o->nel[1]));
        const int s[3] = {o->span[0][e1], o->span[1][e2], o->span[2][e3]};
        double a[3], h[3];
        for (int d = 0; d < 3; ++d) { a[d] = o->U[d][s[d]]; h[d] = o->U[d][s[d] + 1] - a[d]; }
        const int nl = (o->p[0] + 1) * (o->p[1] + 1) * (o->p[2] + 1);
        double* Ke = values ? (double*)calloc((size_t)9 * nl * nl, sizeof(double)) : NULL;
        double Fe[KS_MAXLOC][3];
        memset(Fe, 0, sizeof(Fe));
        double N[KS_MAXLOC], dN[KS_MAXLOC][3], gN[KS_MAXLOC][3];
        int node[KS_MAXLOC];
        for (int q3 = 0; q3 <= o->p[2]; ++q3)
            for (int q2 = 0; q2 <= o->p[1]; ++q2)
                for (int q1 = 0; q1 <= o->p[0]; ++q1) {
                    const double uvw[3] = {a[0] + 0.5 * h[0] * (xq[0][q1] + 1), a[1] + 0.5 * h[1] * (xq[1][q2] + 1),
                                           a[2] + 0.5 * h[2] * (xq[2][q3] + 1)};
                    eval_basis(o, s, uvw, N, dN, node);
                    double Jg[3][3] = {{0}}, Ji[3][3];
                    for (int b = 0; b < nl; ++b)
                        for (int k = 0; k < 3; ++k) for (int l = 0; l < 3; ++l) Jg[k][l] += o->cp[3 * node[b] + k] * dN[b][l];
                    const double dJ = det3(Jg);
                    inv3(Jg, dJ, Ji);
                    const double w = wq[0][q1] * wq[1][q2] * wq[2][q3] * 0.125 * h[0] * h[1] * h[2] * fabs(dJ);
                    double F[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
                    for (int b = 0; b < nl; ++b) {
                        for (int k = 0; k < 3; ++k) gN[b][k] = dN[b][0] * Ji[0][k] + dN[b][1] * Ji[1][k] + dN[b][2] * Ji[2][k];
                        for (int c = 0; c < 3; ++c) for (int k = 0; k < 3; ++k) F[c][k] += disp[3 * node[b] + c] * gN[b][k];
                    }
                    double S[3][3], CC[6][6];
                    if (material(o, F, S, CC)) { bad |= 1; continue; }
                    const int linear = o->law == KS_LAW_HOOKE;
                    double Fb[3][3];
                    for (int c = 0; c < 3; ++c) for (int k = 0; k < 3; ++k) Fb[c][k] = linear ? (c == k) : F[c][k];
                    if (energy) {
                        double C[3][3], E[3][3], trE = 0, EE = 0;
                        for (int i = 0; i < 3; ++i)
                            for (int j = 0; j < 3; ++j) { C[i][j] = 0; for (int k = 0; k < 3; ++k) C[i][j] += F[k][i] * F[k][j]; }
                        for (int i = 0; i < 3; ++i)
                            for (int j = 0; j < 3; ++j) {
                                E[i][j] = linear ? 0.5 * (F[i][j] + F[j][i]) - (i == j) : 0.5 * (C[i][j] - (i == j));
                                EE += E[i][j] * E[i][j];
                            }
                        trE = E[0][0] + E[1][1] + E[2][2];
                        const double Jd = det3(F), I1 = C[0][0] + C[1][1] + C[2][2];
                        double psi;
                        if (o->law == KS_LAW_SVK || linear) psi = 0.5 * o->lambda * trE * trE + o->mu * EE;
                        else if (o->law == KS_LAW_NEO_HOOKE_LN) psi = 0.5 * o->mu * (I1 - 3) - o->mu * log(Jd) + 0.5 * o->lambda * log(Jd) * log(Jd);
                        else psi = 0.5 * o->mu * (I1 - 3) - o->mu * log(Jd) + 0.25 * o->lambda * (Jd * Jd - 1) - 0.5 * o->lambda * log(Jd);
                        etot += w * psi;
                    }
                    /* B_b (6x3): dE_v = sum_c B[v][c] du_c */
                    double B[KS_MAXLOC][6][3];
                    for (int b = 0; b < nl; ++b)
                        for (int v = 0; v < 6; ++v)
                            for (int c = 0; c < 3; ++c)
                                B[b][v][c] = VI[v] == VJ[v] ? Fb[c][VI[v]] * gN[b][VI[v]] : Fb[c][VI[v]] * gN[b][VJ[v]] + Fb[c][VJ[v]] * gN[b][VI[v]];
                    const double Sv[6] = {S[0][0], S[1][1], S[2][2], S[0][1], S[1][2], S[0][2]};
                    for (int b = 0; b < nl; ++b)
                        for (int c = 0; c < 3; ++c) {
                            double f = 0;
                            for (int v = 0; v < 6; ++v) f += B[b][v][c] * Sv[v];
                            Fe[b][c] += w * f;
                        }
                    if (Ke)
                        for (int b = 0; b < nl; ++b) {
                            double CB[6][3];
                            for (int v = 0; v < 6; ++v)
                                for (int d = 0; d < 3; ++d) { CB[v][d] = 0; for (int ww = 0; ww < 6; ++ww) CB[v][d] += CC[v][ww] * B[b][ww][d]; }
                            double Sg[3];
                            for (int k = 0; k < 3; ++k) Sg[k] = S[k][0] * gN[b][0] + S[k][1] * gN[b][1] + S[k][2] * gN[b][2];
                            for (int aa = 0; aa < nl; ++aa) {
                                const double geo = linear ? 0.0 : gN[aa][0] * Sg[0] + gN[aa][1] * Sg[1] + gN[aa][2] * Sg[2];
                                for (int c = 0; c < 3; ++c)
                                    for (int d = 0; d < 3; ++d) {
                                        double m = 0;
                                        for (int v = 0; v < 6; ++v) m += B[aa][v][c] * CB[v][d];
                                        if (c == d) m += geo;
                                        Ke[((size_t)(aa * 3 + c) * nl + b) * 3 + d] += w * m;
                                    }
                            }
                        }
                }
        /* push (gsSparseSystem::push: free rows/columns only; eliminated columns act through F_int of the current state) */
        eval_basis(o, s, a, N, dN, node);
        for (int aa = 0; aa < nl; ++aa)
            for (int c = 0; c < 3; ++c) {
                const int gr = o->map[c * o->ncp + node[aa]];
                if (gr >= o->nfree) continue;
                if (r) {
#pragma omp atomic
                    r[gr] -= Fe[aa][c];
                }
                if (Ke)
                    for (int b = 0; b < nl; ++b)
                        for (int d = 0; d < 3; ++d) {
                            const int gc = o->map[d * o->ncp + node[b]];
                            if (gc >= o->nfree) continue;
                            const long pos = find_pos(o, gr, gc);
#pragma omp atomic
                            values[pos] += Ke[((size_t)(aa * 3 + c) * nl + b) * 3 + d];
                        }
            }
        free(Ke);
    }
    if (r) for (int k = 0; k < o->nfree; ++k) r[k] += o->fext[k];
    if (energy) *energy = etot;
    free(disp);
    return bad ? -4 : 0;
}

/* gsMassAssembler (tutorials/nonlinear_solid_dynamic.cpp:98-109): M_ab^{cd} = delta_cd density int N_a N_b, on the pattern of K */
int kso_mass(const kso* o, double density, double* values) {
    memset(values, 0, sizeof(double) * (size_t)o->nnz);
    double xq[3][MAXP + 2], wq[3][MAXP + 2];
    for (int d = 0; d < 3; ++d) gauss_legendre(o->p[d] + 1, xq[d], wq[d]);
    double N[KS_MAXLOC], dN[KS_MAXLOC][3];
    int node[KS_MAXLOC];
    for (int e3 = 0; e3 < o->nel[2]; ++e3)
        for (int e2 = 0; e2 < o->nel[1]; ++e2)
            for (int e1 = 0; e1 < o->nel[0]; ++e1) {
                const int s[3] = {o->span[0][e1], o->span[1][e2], o->span[2][e3]};
                double a[3], h[3];
                for (int d = 0; d < 3; ++d) { a[d] = o->U[d][s[d]]; h[d] = o->U[d][s[d] + 1] - a[d]; }
                for (int q3 = 0; q3 <= o->p[2]; ++q3)
                    for (int q2 = 0; q2 <= o->p[1]; ++q2)
                        for (int q1 = 0; q1 <= o->p[0]; ++q1) {
                            const double uvw[3] = {a[0] + 0.5 * h[0] * (xq[0][q1] + 1), a[1] + 0.5 * h[1] * (xq[1][q2] + 1),
                                                   a[2] + 0.5 * h[2] * (xq[2][q3] + 1)};
                            const int nl = eval_basis(o, s, uvw, N, dN, node);
                            double Jg[3][3] = {{0}};
                            for (int b = 0; b < nl; ++b)
                                for (int k = 0; k < 3; ++k) for (int l = 0; l < 3; ++l) Jg[k][l] += o->cp[3 * node[b] + k] * dN[b][l];
                            const double w = density * wq[0][q1] * wq[1][q2] * wq[2][q3] * 0.125 * h[0] * h[1] * h[2] * fabs(det3(Jg));
                            for (int aa = 0; aa < nl; ++aa)
                                for (int b = 0; b < nl; ++b)
                                    for (int c = 0; c < 3; ++c) {
                                        const int gr = o->map[c * o->ncp + node[aa]], gc = o->map[c * o->ncp + node[b]];
                                        if (gr < o->nfree && gc < o->nfree) values[find_pos(o, gr, gc)] += w * N[aa] * N[b];
                                    }
                        }
            }
    return 0;
}
