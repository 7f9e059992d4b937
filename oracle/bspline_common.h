/*
 * bspline_common.h — 1-D B-spline basis evaluation and Gauss rule shared by the CPU oracles (kl_oracle.c shells,
 * ks_oracle.c solids).  TEST INFRASTRUCTURE, NOT THE PRODUCT (see kl_oracle.c).
 */
#ifndef ORACLE_BSPLINE_COMMON_H
#define ORACLE_BSPLINE_COMMON_H
#include <math.h>
#ifndef MAXP
#define MAXP 5
#endif

/* ------------------------------------------------------------------------------------ */
/* B-spline basis: Piegl & Tiller, The NURBS Book, A2.1 (FindSpan) and A2.3 (DersBasisFuns)
 * = what gsBSplineBasis::evalAllDers_into computes (values, 1st, 2nd derivatives).       */
static __attribute__((unused)) int find_span(int n, int p, double u, const double* U) {
    /* n = number of basis functions */
    if (u >= U[n]) {
        int s = n - 1;
        while (s > p && U[s] == U[s + 1]) --s;
        return s;
    }
    int lo = p, hi = n, mid = (lo + hi) / 2;
    while (u < U[mid] || u >= U[mid + 1]) {
        if (u < U[mid]) hi = mid; else lo = mid;
        mid = (lo + hi) / 2;
    }
    return mid;
}

static void ders_basis(int span, double u, int p, int nd, const double* U, double ders[3][MAXP + 1]) {
    double ndu[MAXP + 1][MAXP + 1], left[MAXP + 1], right[MAXP + 1], a[2][MAXP + 1];
    ndu[0][0] = 1.0;
    for (int j = 1; j <= p; ++j) {
        left[j] = u - U[span + 1 - j];
        right[j] = U[span + j] - u;
        double saved = 0.0;
        for (int r = 0; r < j; ++r) {
            ndu[j][r] = right[r + 1] + left[j - r];
            double temp = ndu[r][j - 1] / ndu[j][r];
            ndu[r][j] = saved + right[r + 1] * temp;
            saved = left[j - r] * temp;
        }
        ndu[j][j] = saved;
    }
    for (int j = 0; j <= p; ++j) ders[0][j] = ndu[j][p];
    for (int r = 0; r <= p; ++r) {
        int s1 = 0, s2 = 1;
        a[0][0] = 1.0;
        for (int k = 1; k <= nd; ++k) {
            double d = 0.0;
            int rk = r - k, pk = p - k;
            if (k > p) { ders[k][r] = 0.0; continue; }
            if (r >= k) { a[s2][0] = a[s1][0] / ndu[pk + 1][rk]; d = a[s2][0] * ndu[rk][pk]; }
            int j1 = (rk >= -1) ? 1 : -rk;
            int j2 = (r - 1 <= pk) ? k - 1 : p - r;
            for (int j = j1; j <= j2; ++j) {
                a[s2][j] = (a[s1][j] - a[s1][j - 1]) / ndu[pk + 1][rk + j];
                d += a[s2][j] * ndu[rk + j][pk];
            }
            if (r <= pk) { a[s2][k] = -a[s1][k - 1] / ndu[pk + 1][r]; d += a[s2][k] * ndu[r][pk]; }
            ders[k][r] = d;
            int t = s1; s1 = s2; s2 = t;
        }
    }
    double f = p;
    for (int k = 1; k <= nd; ++k) {
        for (int j = 0; j <= p; ++j) ders[k][j] *= f;
        f *= (p - k);
    }
}

/* Gauss–Legendre on [-1,1] (quRule=1), Newton on P_n */
static void gauss_legendre(int n, double* x, double* w) {
    for (int i = 0; i < n; ++i) {
        double z = cos(M_PI * (i + 0.75) / (n + 0.5)), pp = 1.0;
        for (int it = 0; it < 100; ++it) {
            double p1 = 1.0, p2 = 0.0;
            for (int j = 1; j <= n; ++j) { double p3 = p2; p2 = p1; p1 = ((2.0 * j - 1.0) * z * p2 - (j - 1.0) * p3) / j; }
            pp = n * (z * p1 - p2) / (z * z - 1.0);
            double dz = p1 / pp;
            z -= dz;
            if (fabs(dz) < 1e-16) break;
        }
        {   /* final derivative at converged z */
            double p1 = 1.0, p2 = 0.0;
            for (int j = 1; j <= n; ++j) { double p3 = p2; p2 = p1; p1 = ((2.0 * j - 1.0) * z * p2 - (j - 1.0) * p3) / j; }
            pp = n * (z * p1 - p2) / (z * z - 1.0);
        }
        x[n - 1 - i] = z;
        w[n - 1 - i] = 2.0 / ((1.0 - z * z) * pp * pp);
    }
}

#endif
