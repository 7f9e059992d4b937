// kl_capi.cu — the C ABI (include/kl_shell.h) and the host-side set-up of the device context:
// DoF numbering, knot-span / quadrature / 1-D basis tables, external force vector, streams.
// There is NO CPU fallback: every compute entry point fails with KL_E_NOGPU / KL_E_CUDA when no
// device is present.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include "kl_internal.h"

static thread_local std::string g_err;
void kl_set_error(const std::string& s) { g_err = s; }
extern "C" const char* kl_last_error(void) { return g_err.c_str(); }

// ------------------------------------------------------------------------------------------------
// DoF numbering (gsFeSpace::setupMapper + gsDofMapper::finalize)
//   dirichlet side  -> eliminate the boundary DoFs of that component
//   clamped side    -> match each boundary DoF with its neighbour in the second row
//   collapsed side  -> match all boundary DoFs of the side with the first one
//   corner value    -> eliminate that DoF
// A matched group is eliminated as a whole when any member is.  Numbering is component-major:
// plain free DoFs in tensor order, then matched groups by first appearance; eliminated DoFs follow
// after ALL free ones (global_to_bindex = index - freeSize).
// Several patches: the DoF universe of a component is the concatenation of the patches (patch-major, gsDofMapper offsets);
// interfaces add match pairs between the two sides (gsMultiBasis::matchInterface -> gsDofMapper::matchDofs), k-th function of
// side 0 with the k-th (or, reversed, the (len-1-k)-th) function of side 1.
static int side_dof(int n1, int n2, int s, int k, int layer) {
    switch (s) {
        case KL_WEST: return layer + n1 * k;
        case KL_EAST: return (n1 - 1 - layer) + n1 * k;
        case KL_SOUTH: return k + n1 * layer;
        default: return k + n1 * (n2 - 1 - layer);
    }
}
static int build_dofmap_mp(int np, const int32_t* n1, const int32_t* n2, const kl_bc* bc, int nif, const kl_interface* ifs,
                           int32_t* dof_map, int32_t* n_free, int32_t* n_fixed) {
    std::vector<int> off(np + 1, 0);
    for (int q = 0; q < np; ++q) {
        if (n1[q] < 2 || n2[q] < 2) { kl_set_error("kl_build_dofmap: a patch needs at least 2 x 2 control points"); return KL_E_ARG; }
        off[q + 1] = off[q] + n1[q] * n2[q];
    }
    const int N = off[np];
    for (int k = 0; k < nif; ++k) {
        const kl_interface& f = ifs[k];
        for (int e = 0; e < 2; ++e)
            if (f.patch[e] < 0 || f.patch[e] >= np || f.side[e] < KL_WEST || f.side[e] > KL_NORTH) { kl_set_error("kl_mp_build_dofmap: bad interface"); return KL_E_ARG; }
        const int l0 = f.side[0] <= KL_EAST ? n2[f.patch[0]] : n1[f.patch[0]], l1 = f.side[1] <= KL_EAST ? n2[f.patch[1]] : n1[f.patch[1]];
        if (l0 != l1) { kl_set_error("kl_mp_build_dofmap: the two sides of an interface have different numbers of functions (non-conforming)"); return KL_E_ARG; }
    }
    std::vector<int> parent(N), state(N);   // state: 0 plain, 1 matched, 2 eliminated
    std::vector<std::vector<int>> elim_slot(3, std::vector<int>(N, -1));
    std::vector<std::pair<int, int>> pairs;
    int free_total = 0, elim_total = 0;
    auto find = [&](int i) { while (parent[i] != i) { parent[i] = parent[parent[i]]; i = parent[i]; } return i; };
    for (int c = 0; c < 3; ++c) {
        pairs.clear();
        for (int i = 0; i < N; ++i) { parent[i] = i; state[i] = 0; }
        for (int q = 0; q < np; ++q) {
            const int m1 = n1[q], m2 = n2[q], o = off[q];
            for (int s = 0; s < 4; ++s) {
                const int kind = bc[q].side[s][c];
                const int len = (s == KL_WEST || s == KL_EAST) ? m2 : m1;
                for (int k = 0; k < len; ++k) {
                    const int b0 = o + side_dof(m1, m2, s, k, 0);
                    if (kind == KL_BC_DIRICHLET) state[b0] = 2;
                    else if (kind == KL_BC_CLAMPED) pairs.emplace_back(b0, o + side_dof(m1, m2, s, k, 1));
                    else if (kind == KL_BC_COLLAPSED && k > 0) pairs.emplace_back(o + side_dof(m1, m2, s, 0, 0), b0);
                }
            }
            const int corners[4] = {0, m1 - 1, m1 * (m2 - 1), m1 * m2 - 1};
            for (int k = 0; k < 4; ++k) if (bc[q].corner[k][c]) state[o + corners[k]] = 2;
        }
        for (int k = 0; k < nif; ++k) {
            const kl_interface& f = ifs[k];
            const int qa = f.patch[0], qb = f.patch[1];
            const int len = f.side[0] <= KL_EAST ? n2[qa] : n1[qa];
            for (int i = 0; i < len; ++i)
                pairs.emplace_back(off[qa] + side_dof(n1[qa], n2[qa], f.side[0], i, 0),
                                   off[qb] + side_dof(n1[qb], n2[qb], f.side[1], f.reversed ? len - 1 - i : i, 0));
        }
        // matched groups (root = smallest member) and the spread of an elimination over a group
        for (auto& pr : pairs) {
            if (state[pr.first] != 2) state[pr.first] = 1;
            if (state[pr.second] != 2) state[pr.second] = 1;
            const int a = find(pr.first), b = find(pr.second);
            if (a != b) { if (a < b) parent[b] = a; else parent[a] = b; }
        }
        std::vector<char> group_elim(N, 0);
        for (auto& pr : pairs) if (state[pr.first] == 2 || state[pr.second] == 2) group_elim[find(pr.first)] = 1;
        for (auto& pr : pairs) if (group_elim[find(pr.first)]) state[pr.first] = state[pr.second] = 2;
        // numbering: plain free DoFs in (patch, tensor) order, then matched groups by first appearance
        int cnt = 0;
        std::vector<int> val(N, -1), gid(N, -1);
        for (int i = 0; i < N; ++i) if (state[i] == 0) val[i] = free_total + cnt++;
        for (int i = 0; i < N; ++i) if (state[i] == 1) {
            int& g = gid[find(i)];
            if (g < 0) g = free_total + cnt++;
            val[i] = g;
        }
        free_total += cnt;
        std::fill(gid.begin(), gid.end(), -1);
        for (int i = 0; i < N; ++i) if (state[i] == 2) {
            int& g = gid[find(i)];
            if (g < 0) g = elim_total++;
            elim_slot[c][i] = g;
        }
        for (int q = 0; q < np; ++q) {
            const int ncp = n1[q] * n2[q];
            for (int i = 0; i < ncp; ++i) dof_map[(size_t)3 * off[q] + (size_t)c * ncp + i] = val[off[q] + i];
        }
    }
    for (int c = 0; c < 3; ++c)
        for (int q = 0; q < np; ++q) {
            const int ncp = n1[q] * n2[q];
            for (int i = 0; i < ncp; ++i)
                if (elim_slot[c][off[q] + i] >= 0) dof_map[(size_t)3 * off[q] + (size_t)c * ncp + i] = free_total + elim_slot[c][off[q] + i];
        }
    *n_free = free_total;
    *n_fixed = elim_total;
    return KL_OK;
}

extern "C" int kl_build_dofmap(int32_t n1, int32_t n2, const kl_bc* bc, int32_t* dof_map, int32_t* n_free, int32_t* n_fixed) {
    if (n1 < 2 || n2 < 2 || !bc || !dof_map || !n_free || !n_fixed) { kl_set_error("kl_build_dofmap: bad argument"); return KL_E_ARG; }
    return build_dofmap_mp(1, &n1, &n2, bc, 0, nullptr, dof_map, n_free, n_fixed);
}
extern "C" int kl_mp_build_dofmap(int32_t n_patches, const int32_t* n1, const int32_t* n2, const kl_bc* bc, int32_t n_interfaces,
                                  const kl_interface* ifs, int32_t* dof_map, int32_t* n_free, int32_t* n_fixed) {
    if (n_patches < 1 || !n1 || !n2 || !bc || !dof_map || !n_free || !n_fixed || n_interfaces < 0 || (n_interfaces > 0 && !ifs)) {
        kl_set_error("kl_mp_build_dofmap: bad argument");
        return KL_E_ARG;
    }
    return build_dofmap_mp(n_patches, n1, n2, bc, n_interfaces, ifs, dof_map, n_free, n_fixed);
}

// ------------------------------------------------------------------------------------------------
// 1-D B-spline values and derivatives on a knot span by the Cox–de Boor triangle
// (gsBSplineBasis::evalAllDers_into):  T[m][q][j] = m-th derivative of N_{k-q+j, q}(u)
void bspline_span_ders(const std::vector<double>& U, int p, int k, double u, double out[3][KL_MAXP + 1]) {
    double T[3][KL_MAXP + 1][KL_MAXP + 1];
    std::memset(T, 0, sizeof(T));
    T[0][0][0] = 1.0;
    for (int q = 1; q <= p; ++q)
        for (int j = 0; j <= q; ++j) {
            const int i = k - q + j;
            for (int m = 0; m <= 2; ++m) {
                double v = 0.0;
                if (m == 0) {
                    if (j >= 1) v += (u - U[i]) / (U[i + q] - U[i]) * T[0][q - 1][j - 1];
                    if (j <= q - 1) v += (U[i + q + 1] - u) / (U[i + q + 1] - U[i + 1]) * T[0][q - 1][j];
                } else {
                    if (j >= 1) v += T[m - 1][q - 1][j - 1] / (U[i + q] - U[i]);
                    if (j <= q - 1) v -= T[m - 1][q - 1][j] / (U[i + q + 1] - U[i + 1]);
                    v *= q;
                }
                T[m][q][j] = v;
            }
        }
    for (int m = 0; m < 3; ++m)
        for (int j = 0; j <= p; ++j) out[m][j] = T[m][p][j];
}

// Gauss–Legendre nodes/weights on [-1,1] (quRule = 1), Newton iteration in extended precision
void gauss_rule(int n, double* x, double* w) {
    for (int i = 0; i < (n + 1) / 2; ++i) {
        long double z = cosl(3.141592653589793238462643383279502884L * (i + 0.75L) / (n + 0.5L)), pp = 0;
        for (int it = 0; it < 60; ++it) {
            long double p1 = 1, p2 = 0;
            for (int j = 1; j <= n; ++j) { long double p3 = p2; p2 = p1; p1 = ((2 * j - 1) * z * p2 - (j - 1) * p3) / j; }
            pp = n * (z * p1 - p2) / (z * z - 1);
            z -= p1 / pp;
        }
        {
            long double p1 = 1, p2 = 0;
            for (int j = 1; j <= n; ++j) { long double p3 = p2; p2 = p1; p1 = ((2 * j - 1) * z * p2 - (j - 1) * p3) / j; }
            pp = n * (z * p1 - p2) / (z * z - 1);
        }
        x[i] = (double)(-z); x[n - 1 - i] = (double)z;
        w[i] = w[n - 1 - i] = (double)(2 / ((1 - z * z) * pp * pp));
    }
}

template <class T>
static int upload(kl_ctx* ctx, const T** dst, const T* src, size_t n) {
    T* p = nullptr;
    KL_CUDA(cudaMalloc((void**)&p, sizeof(T) * (n ? n : 1)));
    ctx->owned.push_back((void*)p);
    if (n) KL_CUDA(cudaMemcpy(p, src, sizeof(T) * n, cudaMemcpyHostToDevice));
    *dst = p;
    return 0;
}

static int build_tables(kl_ctx* ctx) {
    KLDev& d = ctx->d;
    const int p = d.p, nq = d.nq;
    double xg[16], wg[16];
    gauss_rule(nq, xg, wg);
    for (int dir = 0; dir < 2; ++dir) {
        const std::vector<double>& U = ctx->U[dir];
        const int n = (int)U.size() - p - 1;
        std::vector<int>& span = ctx->span[dir];
        span.clear();
        for (int k = p; k < n; ++k) if (U[k + 1] > U[k]) span.push_back(k);
        const int nel = (int)span.size();
        ctx->flo[dir].assign(n, 1 << 30);
        ctx->fhi[dir].assign(n, -1);
        std::vector<double> bas((size_t)nel * nq * 3 * (p + 1)), wq((size_t)nel * nq);
        for (int e = 0; e < nel; ++e) {
            const int k = span[e];
            for (int a = 0; a <= p; ++a) {
                ctx->flo[dir][k - p + a] = std::min(ctx->flo[dir][k - p + a], e);
                ctx->fhi[dir][k - p + a] = std::max(ctx->fhi[dir][k - p + a], e);
            }
            const double ua = U[k], ub = U[k + 1];
            for (int q = 0; q < nq; ++q) {
                const double u = 0.5 * (ua + ub) + 0.5 * (ub - ua) * xg[q];
                double ders[3][KL_MAXP + 1];
                bspline_span_ders(U, p, k, u, ders);
                for (int m = 0; m < 3; ++m)
                    for (int a = 0; a <= p; ++a) bas[(((size_t)e * nq + q) * 3 + m) * (p + 1) + a] = ders[m][a];
                wq[(size_t)e * nq + q] = 0.5 * (ub - ua) * wg[q];
            }
        }
        const int* dspan; const double *dbas, *dwq, *dknots;
        if (int rc = upload(ctx, &dknots, U.data(), U.size())) return rc;
        if (dir == 0) d.knots1 = dknots; else d.knots2 = dknots;
        if (int rc = upload(ctx, &dspan, span.data(), span.size())) return rc;
        if (int rc = upload(ctx, &dbas, bas.data(), bas.size())) return rc;
        if (int rc = upload(ctx, &dwq, wq.data(), wq.size())) return rc;
        if (dir == 0) { d.span1 = dspan; d.bas1 = dbas; d.wq1 = dwq; d.nel1 = nel; }
        else { d.span2 = dspan; d.bas2 = dbas; d.wq2 = dwq; d.nel2 = nel; }
    }
    return 0;
}

static int build_fext(kl_ctx* ctx, const kl_problem* P) {
    KLDev& d = ctx->d;
    KL_CUDA(cudaMalloc((void**)&ctx->d_fext, sizeof(double) * std::max(d.nfree, 1)));
    ctx->owned.push_back(ctx->d_fext);
    KL_CUDA(cudaMemset(ctx->d_fext, 0, sizeof(double) * std::max(d.nfree, 1)));
    const double* bf = P->body_force;
    if (bf[0] != 0.0 || bf[1] != 0.0 || bf[2] != 0.0) {
        if (int rc = kl_launch_bodyforce(ctx, ctx->d_fext, bf, 0)) return rc;
        KL_CUDA(cudaDeviceSynchronize());
    }
    if (P->n_point_loads > 0) {
        // point loads: F[i,c] += N_i(u,v) * load[c]   (gsThinShellAssembler::setPointLoads)
        std::vector<double> f(std::max(d.nfree, 1), 0.0), fdev(std::max(d.nfree, 1));
        const int p = d.p;
        for (int k = 0; k < P->n_point_loads; ++k) {
            const double uv[2] = {P->point_load_uv[2 * k], P->point_load_uv[2 * k + 1]};
            int sp[2];
            double val[2][KL_MAXP + 1];
            for (int dir = 0; dir < 2; ++dir) {
                const std::vector<double>& U = ctx->U[dir];
                const std::vector<int>& span = ctx->span[dir];
                int s = span.back();
                for (size_t e = 0; e < span.size(); ++e) if (uv[dir] >= U[span[e]] && uv[dir] < U[span[e] + 1]) { s = span[e]; break; }
                sp[dir] = s;
                double ders[3][KL_MAXP + 1];
                bspline_span_ders(U, p, s, uv[dir], ders);
                for (int a = 0; a <= p; ++a) val[dir][a] = ders[0][a];
            }
            for (int b = 0; b <= p; ++b)
                for (int a = 0; a <= p; ++a) {
                    const int cpi = (sp[0] - p + a) + d.n1 * (sp[1] - p + b);
                    for (int c = 0; c < 3; ++c) {
                        const int g = P->dof_map[c * d.ncp + cpi];
                        if (g < d.nfree) f[g] += val[0][a] * val[1][b] * P->point_load_val[3 * k + c];
                    }
                }
        }
        KL_CUDA(cudaMemcpy(fdev.data(), ctx->d_fext, sizeof(double) * d.nfree, cudaMemcpyDeviceToHost));
        for (int i = 0; i < d.nfree; ++i) fdev[i] += f[i];
        KL_CUDA(cudaMemcpy(ctx->d_fext, fdev.data(), sizeof(double) * d.nfree, cudaMemcpyHostToDevice));
    }
    if (P->n_neumann > 0) {
        // Neumann sides: F[i,c] += int_side N_i t_c |dX/dxi| dxi over the UNDEFORMED edge, Gauss rule of the assembly per boundary
        // element.  On a side of an open knot vector only the edge row of functions is non-zero and the edge curve depends on the
        // edge control points alone (rational: quotient rule on the 1-D NURBS curve).
        if (!P->neumann_side || !P->neumann_val) { kl_set_error("kl_create: null Neumann array"); return KL_E_ARG; }
        std::vector<double> f(std::max(d.nfree, 1), 0.0), fdev(std::max(d.nfree, 1));
        const int p = d.p, nq = d.nq;
        double xg[16], wg[16];
        gauss_rule(nq, xg, wg);
        for (int k = 0; k < P->n_neumann; ++k) {
            const int side = P->neumann_side[k];
            if (side < KL_WEST || side > KL_NORTH) { kl_set_error("kl_create: bad Neumann side"); return KL_E_ARG; }
            const int dir = (side == KL_WEST || side == KL_EAST) ? 1 : 0;      // the parametric direction that runs along the side
            const std::vector<double>& U = ctx->U[dir];
            const std::vector<int>& span = ctx->span[dir];
            const int fixed_idx = (side == KL_WEST || side == KL_SOUTH) ? 0 : ((dir == 1 ? d.n1 : d.n2) - 1);
            auto cpi_of = [&](int i) { return dir == 0 ? i + d.n1 * fixed_idx : fixed_idx + d.n1 * i; };
            for (size_t e = 0; e < span.size(); ++e) {
                const int s = span[e];
                const double ua = U[s], ub = U[s + 1];
                for (int q = 0; q < nq; ++q) {
                    const double u = 0.5 * (ua + ub) + 0.5 * (ub - ua) * xg[q];
                    double ders[3][KL_MAXP + 1];
                    bspline_span_ders(U, p, s, u, ders);
                    double X[3] = {0, 0, 0}, dX[3] = {0, 0, 0}, W0 = 0, W1 = 0;
                    for (int a = 0; a <= p; ++a) {
                        const int ci = cpi_of(s - p + a);
                        const double w = P->weights ? P->weights[ci] : 1.0;
                        W0 += ders[0][a] * w; W1 += ders[1][a] * w;
                        for (int c = 0; c < 3; ++c) { X[c] += ders[0][a] * w * P->cp[3 * ci + c]; dX[c] += ders[1][a] * w * P->cp[3 * ci + c]; }
                    }
                    double t2 = 0;
                    for (int c = 0; c < 3; ++c) { const double tc = (dX[c] - W1 * X[c] / W0) / W0; t2 += tc * tc; }
                    const double wJ = 0.5 * (ub - ua) * wg[q] * std::sqrt(t2);
                    for (int a = 0; a <= p; ++a) {
                        const int ci = cpi_of(s - p + a);
                        for (int c = 0; c < 3; ++c) {
                            const int g = P->dof_map[c * d.ncp + ci];
                            if (g < d.nfree) f[g] += wJ * ders[0][a] * P->neumann_val[3 * k + c];
                        }
                    }
                }
            }
        }
        KL_CUDA(cudaMemcpy(fdev.data(), ctx->d_fext, sizeof(double) * d.nfree, cudaMemcpyDeviceToHost));
        for (int i = 0; i < d.nfree; ++i) fdev[i] += f[i];
        KL_CUDA(cudaMemcpy(ctx->d_fext, fdev.data(), sizeof(double) * d.nfree, cudaMemcpyHostToDevice));
    }
    return 0;
}

// Force = assemble().rhs(): dead loads + the follower pressure on the undeformed surface - the lifting of non-zero Dirichlet values
static int build_force(kl_ctx* ctx, const kl_problem* P) {
    KLDev& d = ctx->d;
    const size_t vb = sizeof(double) * std::max(d.nfree, 1);
    KL_CUDA(cudaMalloc((void**)&ctx->d_force, vb));
    ctx->owned.push_back(ctx->d_force);
    KL_CUDA(cudaMemcpy(ctx->d_force, ctx->d_fext, vb, cudaMemcpyDeviceToDevice));
    bool lifting = false;
    if (P->fixed_values) for (int k = 0; k < P->n_fixed; ++k) lifting |= P->fixed_values[k] != 0.0;
    if (d.mat.pressure == 0.0 && !lifting) return 0;
    cudaStream_t s = 0;
    int rc;
    // the undeformed configuration: every displacement coefficient zero, the eliminated ones included
    KL_CUDA(cudaMemsetAsync(d.disp, 0, sizeof(double) * 3 * d.ncp, s));
    if (d.mat.pressure != 0.0) {
        KL_CUDA(cudaMemsetAsync(ctx->d_r, 0, vb, s));
        if ((rc = kl_launch_residual(ctx, ctx->d_r, s))) return rc;                          // F_int(0) - P(0) = -P(0)
        if ((rc = kl_launch_axpby(ctx, ctx->d_force, ctx->d_r, 1.0, -1.0, d.nfree, s))) return rc;
    }
    if (lifting) {
        double* lift;
        KL_CUDA(cudaMalloc((void**)&lift, vb));
        KL_CUDA(cudaMemsetAsync(lift, 0, vb, s));
        if ((rc = kl_launch_points(ctx, 0, d.nel2, s))) return rc;
        KL_CUDA(cudaMemsetAsync(d.values, 0, sizeof(double) * (size_t)ctx->nnz, s));
        const double pr = d.mat.pressure;
        d.lift = lift;
        d.mat.pressure = 0.0;                   // K_L of the lifting is the material + geometric stiffness at u = 0
        rc = kl_launch_jacobian(ctx, 0, d.nel2, s);
        d.lift = nullptr;
        d.mat.pressure = pr;
        if (rc) { cudaFree(lift); return rc; }
        if ((rc = kl_launch_axpby(ctx, ctx->d_force, lift, 1.0, -1.0, d.nfree, s))) { cudaFree(lift); return rc; }
        KL_CUDA(cudaStreamSynchronize(s));
        cudaFree(lift);
        KL_CUDA(cudaMemsetAsync(d.values, 0, sizeof(double) * (size_t)ctx->nnz, s));
    }
    KL_CUDA(cudaStreamSynchronize(s));
    return kl_check(ctx, s);
}

// Pipelined copy-out plan: the element rows are cut into strips; a column of K is final once every element in the
// support of every control point mapped to it has been assembled, so after each strip a few contiguous value ranges
// can already travel to the host while the next strip is being assembled.
static int build_d2h_plan(kl_ctx* ctx, const kl_problem* P) {
    const KLDev& d = ctx->d;
    const int nel2 = d.nel2, n1 = d.n1, n2 = d.n2, nf = d.nfree;
    int S = ctx->n_strips_d2h;
    if (const char* e = getenv("KL_D2H_STRIPS")) S = atoi(e);
    S = std::max(1, std::min(S, nel2 / std::max(1, 2 * d.p)));
    std::vector<int>& outer = ctx->h_outer;
    outer.resize((size_t)nf + 1);
    KL_CUDA(cudaMemcpy(outer.data(), d.outer, sizeof(int) * ((size_t)nf + 1), cudaMemcpyDeviceToHost));
    std::vector<int> last_row(nf, -1);
    for (int c = 0; c < 3; ++c)
        for (int i = 0; i < n1 * n2; ++i) {
            const int g = P->dof_map[c * d.ncp + i];
            if (g < nf) last_row[g] = std::max(last_row[g], i / n1);
        }
    ctx->d2h_plan.clear();
    int prev_done = 0;
    std::vector<int> strip_of_row(n2, S - 1);
    for (int s = 0; s < S; ++s) {
        kl_ctx::D2HStrip st;
        st.e2_begin = (int)((long long)nel2 * s / S);
        st.e2_end = (int)((long long)nel2 * (s + 1) / S);
        int done = prev_done;
        while (done < n2 && ctx->fhi[1][done] < st.e2_end) ++done;   // rows whose support ends inside the assembled part
        for (int r = prev_done; r < done; ++r) strip_of_row[r] = s;
        prev_done = done;
        ctx->d2h_plan.push_back(st);
    }
    for (int g = 0; g < nf; ++g) {
        const int s = last_row[g] >= 0 ? strip_of_row[last_row[g]] : S - 1;
        auto& r = ctx->d2h_plan[s].ranges;
        const size_t a = (size_t)outer[g], b = (size_t)outer[g + 1];
        if (!r.empty() && r.back().second == a) r.back().second = b; else r.emplace_back(a, b);
        auto& c = ctx->d2h_plan[s].cols;
        if (!c.empty() && c.back().second == g) c.back().second = g + 1; else c.emplace_back(g, g + 1);
    }
    ctx->strip_ev.resize(S);
    for (auto& e : ctx->strip_ev) KL_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    ctx->copy_ev.resize(S);
    for (auto& e : ctx->copy_ev) KL_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    return 0;
}

// ------------------------------------------------------------------------------------------------
// everything of a context that does not depend on the sparse pattern
int kl_ctx_create_base(const kl_problem* P, int device, kl_ctx** out) {
    if (!P || !out) { kl_set_error("kl_create: null argument"); return KL_E_ARG; }
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        kl_set_error("no CUDA device: libkl_shell has no CPU fallback");
        return KL_E_NOGPU;
    }
    if (device >= 0) KL_CUDA(cudaSetDevice(device));
    if (P->degree[0] != P->degree[1] || P->degree[0] < 2 || P->degree[0] > KL_MAXP) {
        kl_set_error("kl_create: degrees must be equal and in [2,4]");
        return KL_E_ARG;
    }
    const int p = P->degree[0];
    const int quA = (P->quA == 0 && P->quB == 0) ? 1 : P->quA, quB = (P->quA == 0 && P->quB == 0) ? 1 : P->quB;
    if (quA * p + quB != p + 1) { kl_set_error("kl_create: only quA*p+quB == p+1 Gauss nodes are supported on the device"); return KL_E_ARG; }
    if (P->material != KL_MAT_SVK && P->material != KL_MAT_NH && P->material != KL_MAT_MR) { kl_set_error("kl_create: unsupported material"); return KL_E_ARG; }
    if (!P->knots[0] || !P->knots[1] || !P->cp || !P->dof_map) { kl_set_error("kl_create: null array"); return KL_E_ARG; }
    kl_ctx* ctx = new kl_ctx();
#define KL_CUDA_CTX(call)                                                                      \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess) {                                                               \
            kl_set_error(std::string(#call) + ": " + cudaGetErrorString(e_));                  \
            kl_destroy(ctx);                                                                   \
            return KL_E_CUDA;                                                                  \
        }                                                                                      \
    } while (0)
    KL_CUDA_CTX(cudaGetDevice(&ctx->device));
    ctx->prob = *P;
    KLDev& d = ctx->d;
    d.p = p; d.nq = p + 1;
    for (int dir = 0; dir < 2; ++dir) ctx->U[dir].assign(P->knots[dir], P->knots[dir] + P->n_knots[dir]);
    d.n1 = P->n_knots[0] - p - 1; d.n2 = P->n_knots[1] - p - 1; d.ncp = d.n1 * d.n2;
    d.nfree = P->n_free; ctx->nfixed = P->n_fixed;
    d.nst = (2 * p + 1) * (2 * p + 1);
    d.rational = P->weights != nullptr;
    int rc;
    if ((rc = build_tables(ctx))) { kl_destroy(ctx); return rc; }
    if ((rc = upload(ctx, &d.cp, P->cp, (size_t)3 * d.ncp))) { kl_destroy(ctx); return rc; }
    d.w = nullptr;
    if (P->weights && (rc = upload(ctx, &d.w, P->weights, (size_t)d.ncp))) { kl_destroy(ctx); return rc; }
    if ((rc = upload(ctx, &d.map, P->dof_map, (size_t)3 * d.ncp))) { kl_destroy(ctx); return rc; }
    d.fixed = nullptr;
    if (P->fixed_values && P->n_fixed > 0 && (rc = upload(ctx, &d.fixed, P->fixed_values, (size_t)P->n_fixed))) { kl_destroy(ctx); return rc; }
    KL_CUDA_CTX(cudaMalloc((void**)&d.disp, sizeof(double) * 3 * d.ncp)); ctx->owned.push_back(d.disp);
    KL_CUDA_CTX(cudaMemset(d.disp, 0, sizeof(double) * 3 * d.ncp));
    {
        const size_t npts = (size_t)d.nel1 * d.nel2 * d.nq * d.nq;
        void* pdbuf = nullptr;
        KL_CUDA_CTX(cudaMalloc(&pdbuf, kl_pointdata_bytes() * (npts ? npts : 1)));
        ctx->owned.push_back(pdbuf);
        d.pd = (PointData*)pdbuf;
    }
    KL_CUDA_CTX(cudaMalloc((void**)&d.flag, sizeof(int))); ctx->owned.push_back(d.flag);
    KL_CUDA_CTX(cudaMemset(d.flag, 0, sizeof(int)));
    KL_CUDA_CTX(cudaMalloc((void**)&ctx->d_x, sizeof(double) * std::max(d.nfree, 1))); ctx->owned.push_back(ctx->d_x);
    KL_CUDA_CTX(cudaMalloc((void**)&ctx->d_r, sizeof(double) * std::max(d.nfree, 1))); ctx->owned.push_back(ctx->d_r);
    KL_CUDA_CTX(cudaMalloc((void**)&ctx->d_xstate, sizeof(double) * std::max(d.nfree, 1))); ctx->owned.push_back(ctx->d_xstate);
    KL_CUDA_CTX(cudaMalloc((void**)&ctx->d_same, sizeof(int))); ctx->owned.push_back(ctx->d_same);
    KL_CUDA_CTX(cudaMemset(ctx->d_same, 0, sizeof(int)));
    ctx->spec_allowed = getenv("KL_SPECULATE") ? atoi(getenv("KL_SPECULATE")) : 1;
    KL_CUDA_CTX(cudaMallocHost((void**)&ctx->h_pinned_x, sizeof(double) * std::max(d.nfree, 1)));
    KL_CUDA_CTX(cudaMallocHost((void**)&ctx->h_pinned_r, sizeof(double) * std::max(d.nfree, 1)));
    // material constants
    KLMaterial& m = d.mat;
    m.material = P->material; m.compressible = P->compressible; m.bending = P->bending; m.metric_z2 = P->metric_z2;
    m.ngauss = P->num_gauss_thickness > 0 ? P->num_gauss_thickness : 4;
    if (m.ngauss > 12) { kl_set_error("kl_create: NumGauss > 12"); kl_destroy(ctx); return KL_E_ARG; }
    m.E = P->E; m.nu = P->nu; m.t = P->thickness;
    m.mu = P->E / (2.0 * (1.0 + P->nu));
    {
        const double lam = P->E * P->nu / ((1.0 + P->nu) * (1.0 - 2.0 * P->nu));
        m.lam_ps = 2.0 * lam * m.mu / (lam + 2.0 * m.mu);
    }
    m.bulk = 2.0 * m.mu * (1.0 + P->nu) / (3.0 - 6.0 * P->nu);
    m.c1 = m.mu; m.c2 = 0.0;
    if (P->material == KL_MAT_MR) { m.c2 = m.mu / (P->mr_ratio + 1.0); m.c1 = P->mr_ratio * m.c2; }
    gauss_rule(m.ngauss, m.zg, m.wg);
    m.pressure = P->pressure;
    ctx->e2_begin = 0; ctx->e2_end = d.nel2;
    ctx->cp_row_begin = 0; ctx->cp_row_end = d.n2;
    ctx->h_map.assign(P->dof_map, P->dof_map + (size_t)3 * d.ncp);
    d.ablate = getenv("KL_ABLATE") ? atoi(getenv("KL_ABLATE")) : 0;
    KL_CUDA_CTX(cudaDeviceGetAttribute(&ctx->n_sm, cudaDevAttrMultiProcessorCount, ctx->device));
    ctx->jac_shared = getenv("KL_JAC_SHARED") ? atoi(getenv("KL_JAC_SHARED")) : 0;
    ctx->jac_seg = getenv("KL_SW_SEG") ? atoi(getenv("KL_SW_SEG")) : 0;
    KL_CUDA_CTX(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    KL_CUDA_CTX(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    for (auto& e : ctx->ev) KL_CUDA_CTX(cudaEventCreate(&e));
    *out = ctx;
    return KL_OK;
#undef KL_CUDA_CTX
}

// the pattern-dependent rest: load vectors, copy-out plan (not for members of a kl_mp: their columns are shared with other patches)
int kl_ctx_finish(kl_ctx* ctx, const kl_problem* P) {
    int rc;
    if ((rc = build_fext(ctx, P))) return rc;
    if (!ctx->mp_member && (rc = build_d2h_plan(ctx, P))) return rc;
    return build_force(ctx, P);
}

extern "C" int kl_create(const kl_problem* P, int device, kl_ctx** out) {
    if (out) *out = nullptr;
    kl_ctx* ctx = nullptr;
    int rc = kl_ctx_create_base(P, device, &ctx);
    if (rc) return rc;
    if ((rc = kl_build_pattern(ctx)) || (rc = kl_ctx_finish(ctx, P))) { kl_destroy(ctx); return rc; }
    *out = ctx;
    return KL_OK;
}

extern "C" void kl_destroy(kl_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    kl_solve_free(ctx);
    if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
    for (auto& e : ctx->copy_ev) if (e) cudaEventDestroy(e);
    for (void* p : ctx->owned) cudaFree(p);
    if (ctx->h_pinned_x) cudaFreeHost(ctx->h_pinned_x);
    if (ctx->h_pinned_r) cudaFreeHost(ctx->h_pinned_r);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    for (auto& e : ctx->ev) if (e) cudaEventDestroy(e);
    for (auto& e : ctx->strip_ev) if (e) cudaEventDestroy(e);
    delete ctx;
}

extern "C" int kl_sizes(const kl_ctx* ctx, int32_t* n_dofs, int64_t* nnz, int64_t* n_elements, int64_t* n_qp) {
    if (!ctx) return KL_E_ARG;
    if (n_dofs) *n_dofs = ctx->d.nfree;
    if (nnz) *nnz = ctx->nnz;
    const int64_t ne = ctx->mp ? ctx->mp->n_elements : (int64_t)ctx->d.nel1 * ctx->d.nel2;
    if (n_elements) *n_elements = ne;
    if (n_qp) *n_qp = ctx->mp ? ctx->mp->n_qp : ne * ctx->d.nq * ctx->d.nq;
    return KL_OK;
}

extern "C" int kl_pattern_host(const kl_ctx* ctx, int32_t* outer, int32_t* inner) {
    if (!ctx || !outer || !inner) return KL_E_ARG;
    KL_CUDA(cudaSetDevice(ctx->device));
    KL_CUDA(cudaMemcpy(outer, ctx->d.outer, sizeof(int) * ((size_t)ctx->d.nfree + 1), cudaMemcpyDeviceToHost));
    KL_CUDA(cudaMemcpy(inner, ctx->d.inner, sizeof(int) * (size_t)ctx->nnz, cudaMemcpyDeviceToHost));
    return KL_OK;
}
extern "C" int kl_pattern_device(const kl_ctx* ctx, const int32_t** outer_dev, const int32_t** inner_dev) {
    if (!ctx) return KL_E_ARG;
    if (outer_dev) *outer_dev = ctx->d.outer;
    if (inner_dev) *inner_dev = ctx->d.inner;
    return KL_OK;
}
extern "C" double* kl_values_device(kl_ctx* ctx) { return ctx ? ctx->d.values : nullptr; }
extern "C" int kl_kernel_launches(const kl_ctx* ctx) { return ctx ? ctx->launches + (ctx->mp ? kl_mp_launches(ctx->mp) : 0) : 0; }

extern "C" int kl_set_strip(kl_ctx* ctx, int32_t e2_begin, int32_t e2_end) {
    if (ctx && ctx->mp_member) { kl_set_error("kl_set_strip: a patch of a kl_mp is partitioned by patches (kl_mp_set_active), not by strips"); return KL_E_ARG; }
    if (!ctx || e2_begin < 0 || e2_end > ctx->d.nel2 || e2_begin > e2_end) { kl_set_error("kl_set_strip: bad range"); return KL_E_ARG; }
    ctx->e2_begin = e2_begin; ctx->e2_end = e2_end;
    ctx->pd_valid = 0;
    // the strip reads the control-point rows of its elements and contributes only to their columns: zero just those value ranges
    const KLDev& d = ctx->d;
    ctx->zero_ranges.clear();
    if (e2_begin == 0 && e2_end == d.nel2) { ctx->cp_row_begin = 0; ctx->cp_row_end = d.n2; return KL_OK; }
    if (e2_end == e2_begin) { ctx->cp_row_begin = ctx->cp_row_end = 0; ctx->zero_ranges.emplace_back(0, 0); return KL_OK; }
    ctx->cp_row_begin = ctx->span[1][e2_begin] - d.p;
    ctx->cp_row_end = ctx->span[1][e2_end - 1] + 1;
    std::vector<char> touched((size_t)d.nfree, 0);
    for (int c = 0; c < 3; ++c)
        for (int i = ctx->cp_row_begin * d.n1; i < ctx->cp_row_end * d.n1; ++i) {
            const int g = ctx->h_map[(size_t)c * d.ncp + i];
            if (g < d.nfree) touched[g] = 1;
        }
    for (int g = 0; g < d.nfree; ++g) {
        if (!touched[g]) continue;
        const size_t a = (size_t)ctx->h_outer[g], b = (size_t)ctx->h_outer[g + 1];
        if (!ctx->zero_ranges.empty() && ctx->zero_ranges.back().second == a) ctx->zero_ranges.back().second = b;
        else ctx->zero_ranges.emplace_back(a, b);
    }
    return KL_OK;
}

extern "C" int kl_check(kl_ctx* ctx, void* stream) {
    if (!ctx) return KL_E_ARG;
    if (ctx->mp) {
        KL_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
        return kl_mp_check(ctx->mp, (cudaStream_t)stream);
    }
    int flag = 0;
    KL_CUDA(cudaMemcpyAsync(&flag, ctx->d.flag, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    KL_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    if (flag) {
        KL_CUDA(cudaMemsetAsync(ctx->d.flag, 0, sizeof(int), (cudaStream_t)stream));
        const std::string bits = " (device flag " + std::to_string(flag) + ")";
        if (flag & KLF_JACOBIAN) { kl_set_error("inverted element: |a1 x a2| <= 0" + bits); return KL_E_JACOBIAN; }
        if (flag & KLF_METRIC) { kl_set_error("det of the (through-thickness) metric <= 0" + bits); return KL_E_JACOBIAN; }
        if (flag & KLF_C33) { kl_set_error("plane-stress iteration on C33 did not converge"); return KL_E_C33; }
        kl_set_error("non-finite value at a quadrature point");
        return KL_E_NONFINITE;
    }
    return KL_OK;
}

// ---- device-resident entry points -----------------------------------------------------------------
// Same-state fusion.  The reference's solvers ask for Residual(x) and then Jacobian(x) at ONE state through two separate
// std::functions (gsStaticNewton.hpp:160-191, gsALMBase.hpp:214-258).  A residual call therefore runs the per-point kernel
// with the material tangent and leaves the records of the whole mesh in d.pd (speculation); a Jacobian call first compares
// its x with the state of those records ON THE DEVICE and lets constructSolution + the point kernel return at once when they
// match.  Nothing is synchronised and a miss costs one extra launch.  Speculation switches itself off while the caller only
// asks for residuals (explicit dynamics, dynamic relaxation, line searches) and on again at the next Jacobian call.
static int zero_values(kl_ctx* ctx, cudaStream_t s) {
    if (ctx->zero_ranges.empty()) { KL_CUDA(cudaMemsetAsync(ctx->d.values, 0, sizeof(double) * (size_t)ctx->nnz, s)); return 0; }
    for (const auto& r : ctx->zero_ranges)
        if (r.second > r.first) KL_CUDA(cudaMemsetAsync(ctx->d.values + r.first, 0, sizeof(double) * (r.second - r.first), s));
    return 0;
}
static int points_and_jacobian(kl_ctx* ctx, const double* x_dev, cudaStream_t s, int e2b, int e2e, bool launch_jac) {
    int rc;
    if (ctx->pd_valid && (ctx->pd_e2b != e2b || ctx->pd_e2e != e2e)) ctx->pd_valid = 0;     // records of another set of element rows
    if ((rc = kl_launch_state_compare(ctx, x_dev, s))) return rc;
    if ((rc = kl_launch_construct(ctx, x_dev, s, ctx->d_same))) return rc;
    KL_CUDA(cudaEventRecord(ctx->ev[6], s));
    if ((rc = kl_launch_points(ctx, e2b, e2e, s, nullptr, ctx->d_same))) return rc;
    KL_CUDA(cudaEventRecord(ctx->ev[7], s));
    ctx->pd_valid = 1; ctx->pd_e2b = e2b; ctx->pd_e2e = e2e;
    ctx->last_call = 2;
    ctx->spec_on = 1;
    if ((rc = zero_values(ctx, s))) return rc;
    return launch_jac ? kl_launch_jacobian(ctx, e2b, e2e, s) : 0;
}

extern "C" int kl_jacobian_device(kl_ctx* ctx, const double* x_dev, void* stream) {
    if (!ctx) return KL_E_ARG;
    if (ctx->mp) return kl_mp_jacobian_device(ctx->mp, x_dev, (cudaStream_t)stream);
    return points_and_jacobian(ctx, x_dev, (cudaStream_t)stream, ctx->e2_begin, ctx->e2_end, true);
}

// r += F_int(x) - P(x) over the element rows of the context (no zeroing, no load vector): shared by the single-patch entry point
// below and by the multi-patch assembler, which sums the patches into one vector
int kl_residual_accumulate(kl_ctx* ctx, const double* x_dev, double* r_dev, cudaStream_t s) {
    int rc;
    if (ctx->last_call == 1) ctx->spec_on = 0;          // two residuals in a row: the caller is not running a Newton-type loop
    if ((rc = kl_launch_construct(ctx, x_dev, s))) return rc;
    if (ctx->spec_on && ctx->spec_allowed) {
        // per-point records with the tangent + internal force from the staged records; remember the state they belong to
        ctx->pd_valid = 0;
        if ((rc = kl_launch_state_compare(ctx, x_dev, s))) return rc;      // pd_valid == 0: only stores the state
        KL_CUDA(cudaEventRecord(ctx->ev[6], s));
        if ((rc = kl_launch_points(ctx, ctx->e2_begin, ctx->e2_end, s, r_dev))) return rc;
        KL_CUDA(cudaEventRecord(ctx->ev[7], s));
        ctx->pd_valid = 1; ctx->pd_e2b = ctx->e2_begin; ctx->pd_e2e = ctx->e2_end;
        ctx->last_call = 1;
    } else {
        if ((rc = kl_launch_residual(ctx, r_dev, s))) return rc;
        ctx->last_call = 0;
    }
    return 0;
}

extern "C" int kl_residual_device(kl_ctx* ctx, const double* x_dev, double lam_fext, double sign_fint, double* r_dev, void* stream) {
    if (!ctx || !r_dev) return KL_E_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    if (ctx->mp) return kl_mp_residual_device(ctx->mp, x_dev, lam_fext, sign_fint, r_dev, s);
    KL_CUDA(cudaMemsetAsync(r_dev, 0, sizeof(double) * ctx->d.nfree, s));
    if (int rc = kl_residual_accumulate(ctx, x_dev, r_dev, s)) return rc;
    return kl_launch_axpby(ctx, r_dev, ctx->d_fext, sign_fint, lam_fext, ctx->d.nfree, s);
}

// Strip assembly in two phases so that the halo exchange overlaps the bulk of the work: phase 0 evaluates the points of
// the whole strip (records + internal force), zeroes the strip's value ranges and assembles the LAST `tail_rows` element rows, whose
// contributions reach the next strip; the caller starts the exchange and calls kl_jacobian_rows_device for the remaining rows.
extern "C" int kl_strip_begin_device(kl_ctx* ctx, const double* x_dev, double lam_fext, double sign_fint, double* r_dev, int32_t tail_rows, void* stream) {
    if (!ctx || !r_dev || tail_rows < 0) return KL_E_ARG;
    if (ctx->mp || ctx->mp_member) { kl_set_error("strips are not available on a multi-patch assembler"); return KL_E_ARG; }
    cudaStream_t s = (cudaStream_t)stream;
    int rc;
    const int spec = ctx->spec_on, last = ctx->last_call;
    ctx->spec_on = 1; ctx->last_call = 0;                     // the per-point records of the strip are needed by the Jacobian rows
    rc = kl_residual_device(ctx, x_dev, lam_fext, sign_fint, r_dev, stream);
    ctx->spec_on = spec; ctx->last_call = last;
    if (rc) return rc;
    if (!ctx->spec_allowed) { kl_set_error("kl_strip_begin_device needs the per-point records (KL_SPECULATE=0 is set)"); return KL_E_ARG; }
    if ((rc = zero_values(ctx, s))) return rc;
    const int b = std::max(ctx->e2_begin, ctx->e2_end - tail_rows);
    return kl_launch_jacobian(ctx, b, ctx->e2_end, s);
}
extern "C" int kl_jacobian_rows_device(kl_ctx* ctx, int32_t e2_begin, int32_t e2_end, void* stream) {
    if (!ctx || e2_begin < ctx->e2_begin || e2_end > ctx->e2_end) { kl_set_error("kl_jacobian_rows_device: rows outside the strip"); return KL_E_ARG; }
    if (!ctx->pd_valid || ctx->pd_e2b != ctx->e2_begin || ctx->pd_e2e != ctx->e2_end) { kl_set_error("kl_jacobian_rows_device: no per-point records of this strip"); return KL_E_ARG; }
    return kl_launch_jacobian(ctx, e2_begin, e2_end, (cudaStream_t)stream);
}

extern "C" int kl_al_residual_device(kl_ctx* ctx, const double* x_dev, double lam, double* r_dev, void* stream) {
    // Force - lam*Force - rhs(x) with rhs(x) = F_dead - (F_int(x) - P(x))
    if (int rc = kl_residual_device(ctx, x_dev, -1.0, 1.0, r_dev, stream)) return rc;
    return kl_launch_axpby(ctx, r_dev, ctx->d_force, 1.0, 1.0 - lam, ctx->d.nfree, (cudaStream_t)stream);
}

extern "C" int kl_assemble_device(kl_ctx* ctx, const double* x_dev, double lam_fext, double sign_fint, double* r_dev, void* stream) {
    if (!ctx || !r_dev) return KL_E_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    int rc;
    if (ctx->mp) {      // the patches speculate on the Jacobian in their residual pass: the same fusion, patch by patch
        if ((rc = kl_mp_residual_device(ctx->mp, x_dev, lam_fext, sign_fint, r_dev, s))) return rc;
        return kl_mp_jacobian_device(ctx->mp, x_dev, s);
    }
    if ((rc = kl_launch_construct(ctx, x_dev, s))) return rc;
    KL_CUDA(cudaMemsetAsync(r_dev, 0, sizeof(double) * ctx->d.nfree, s));
    ctx->pd_valid = 0;
    if ((rc = kl_launch_state_compare(ctx, x_dev, s))) return rc;
    KL_CUDA(cudaEventRecord(ctx->ev[6], s));
    // the point kernel integrates the internal force from the records it has just staged: no separate residual pass
    if ((rc = kl_launch_points(ctx, ctx->e2_begin, ctx->e2_end, s, r_dev))) return rc;
    KL_CUDA(cudaEventRecord(ctx->ev[7], s));
    ctx->pd_valid = 1; ctx->pd_e2b = ctx->e2_begin; ctx->pd_e2e = ctx->e2_end;
    ctx->last_call = 2;
    if ((rc = kl_launch_axpby(ctx, r_dev, ctx->d_fext, sign_fint, lam_fext, ctx->d.nfree, s))) return rc;
    if ((rc = zero_values(ctx, s))) return rc;
    return kl_launch_jacobian(ctx, ctx->e2_begin, ctx->e2_end, s);
}

// ---- host-pointer entry points (what the Jacobian_t / Residual_t closures call) -------------------
// page-locked (cudaMallocHost / cudaHostRegister'ed) caller memory is copied from / to directly; pageable memory goes through
// the context's pinned staging buffers
static bool host_is_pinned(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) == cudaSuccess && a.type == cudaMemoryTypeHost) return true;
    cudaGetLastError();
    return false;
}
static const double* stage_x(kl_ctx* ctx, const double* x_host, int n) {
    if (host_is_pinned(x_host)) return x_host;
    std::memcpy(ctx->h_pinned_x, x_host, sizeof(double) * n);
    return ctx->h_pinned_x;
}

static int run_residual(kl_ctx* ctx, const double* x_host, double lam_fext, double sign_fint, double* r_host, bool al = false) {
    if (!ctx || !r_host) { kl_set_error("null argument"); return KL_E_ARG; }
    KL_CUDA(cudaSetDevice(ctx->device));
    const int n = ctx->d.nfree;
    cudaStream_t s = ctx->stream;
    KL_CUDA(cudaEventRecord(ctx->ev[0], s));
    const double* xd = nullptr;
    if (x_host) {
        KL_CUDA(cudaMemcpyAsync(ctx->d_x, stage_x(ctx, x_host, n), sizeof(double) * n, cudaMemcpyHostToDevice, s));
        xd = ctx->d_x;
    }
    KL_CUDA(cudaEventRecord(ctx->ev[1], s));
    int rc = al ? kl_al_residual_device(ctx, xd, lam_fext, ctx->d_r, s) : kl_residual_device(ctx, xd, lam_fext, sign_fint, ctx->d_r, s);
    if (rc) return rc;
    KL_CUDA(cudaEventRecord(ctx->ev[2], s));
    const bool direct = host_is_pinned(r_host);
    KL_CUDA(cudaMemcpyAsync(direct ? r_host : ctx->h_pinned_r, ctx->d_r, sizeof(double) * n, cudaMemcpyDeviceToHost, s));
    KL_CUDA(cudaEventRecord(ctx->ev[3], s));
    rc = kl_check(ctx, s);
    if (!direct) std::memcpy(r_host, ctx->h_pinned_r, sizeof(double) * n);
    cudaEventElapsedTime(&ctx->ms_h2d, ctx->ev[0], ctx->ev[1]);
    cudaEventElapsedTime(&ctx->ms_kernel, ctx->ev[1], ctx->ev[2]);
    cudaEventElapsedTime(&ctx->ms_d2h, ctx->ev[2], ctx->ev[3]);
    return rc;
}

extern "C" int kl_residual(kl_ctx* ctx, const double* x_host, double* r_host) { return run_residual(ctx, x_host, 1.0, -1.0, r_host); }
extern "C" int kl_al_residual(kl_ctx* ctx, const double* x_host, double lam, double* r_host) { return run_residual(ctx, x_host, lam, 1.0, r_host, true); }

extern "C" int kl_mass(kl_ctx* ctx, double density, double* values_host, double* lumped_host) {
    if (!ctx || (!values_host && !lumped_host)) { kl_set_error("kl_mass: null argument"); return KL_E_ARG; }
    KL_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    // the mass matrix shares the value array of K on the device (it is a set-up quantity; K is re-assembled every call)
    if (values_host) KL_CUDA(cudaMemsetAsync(ctx->d.values, 0, sizeof(double) * (size_t)ctx->nnz, s));
    if (lumped_host) KL_CUDA(cudaMemsetAsync(ctx->d_r, 0, sizeof(double) * ctx->d.nfree, s));
    int rc = ctx->mp ? kl_mp_mass_device(ctx->mp, density, values_host ? ctx->d.values : nullptr, lumped_host ? ctx->d_r : nullptr, s)
                     : kl_launch_mass(ctx, density * ctx->d.mat.t, values_host ? ctx->d.values : nullptr, lumped_host ? ctx->d_r : nullptr, s);
    if (rc) return rc;
    if (values_host) KL_CUDA(cudaMemcpyAsync(values_host, ctx->d.values, sizeof(double) * (size_t)ctx->nnz, cudaMemcpyDeviceToHost, s));
    if (lumped_host) KL_CUDA(cudaMemcpyAsync(lumped_host, ctx->d_r, sizeof(double) * ctx->d.nfree, cudaMemcpyDeviceToHost, s));
    KL_CUDA(cudaStreamSynchronize(s));
    return KL_OK;
}

extern "C" int kl_force(kl_ctx* ctx, double* f_host) {
    if (!ctx || !f_host) return KL_E_ARG;
    KL_CUDA(cudaSetDevice(ctx->device));
    KL_CUDA(cudaMemcpy(f_host, ctx->d_force, sizeof(double) * ctx->d.nfree, cudaMemcpyDeviceToHost));
    return KL_OK;
}

// ---- copy-out of the matrix values -------------------------------------------------------------------
// The library never page-locks memory it does not own (a solver that builds a fresh gsSparseMatrix per call, as
// gsStaticNewton::_computeJacobian does, would leave a stale registration behind when it frees the matrix).  Page-locked
// caller memory (cudaMallocHost, or kl_pin_values by the owner of a long-lived matrix) is written by the DMA engine directly;
// pageable memory is served through a context-owned page-locked staging buffer and copied out by a few host threads while
// the next strip is still in flight.
static int staging_get(kl_ctx* ctx, size_t bytes) {
    if (ctx->h_stage_bytes >= bytes) return 0;
    if (ctx->h_stage) { cudaFreeHost(ctx->h_stage); ctx->h_stage = nullptr; ctx->h_stage_bytes = 0; }
    KL_CUDA(cudaMallocHost((void**)&ctx->h_stage, bytes));
    ctx->h_stage_bytes = bytes;
    return 0;
}

struct CopyRange { size_t dst, src, n; int ev; };   // element offsets; ev = index of the event that completes the range

// device `src_dev` -> host `dst_host` in the given ranges, each waiting for strip event ev on the compute stream
static int copy_out(kl_ctx* ctx, const double* src_dev, double* dst_host, const std::vector<CopyRange>& ranges, size_t total) {
    const bool direct = host_is_pinned(dst_host);
    double* land = dst_host;
    if (!direct) {
        if (int rc = staging_get(ctx, sizeof(double) * total)) return rc;
        land = ctx->h_stage;
    }
    int last_ev = -1;
    std::vector<cudaEvent_t>& cev = ctx->copy_ev;
    for (const auto& r : ranges) {
        if (r.ev != last_ev) {
            if (last_ev >= 0 && !direct) KL_CUDA(cudaEventRecord(cev[last_ev], ctx->copy_stream));
            if (r.ev >= 0) KL_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->strip_ev[r.ev], 0));
            last_ev = r.ev;
        }
        KL_CUDA(cudaMemcpyAsync(land + r.dst, src_dev + r.src, sizeof(double) * r.n, cudaMemcpyDeviceToHost, ctx->copy_stream));
    }
    if (last_ev >= 0 && !direct) KL_CUDA(cudaEventRecord(cev[last_ev], ctx->copy_stream));
    KL_CUDA(cudaEventRecord(ctx->ev[3], ctx->copy_stream));
    if (direct) { KL_CUDA(cudaStreamSynchronize(ctx->copy_stream)); return 0; }
    // pageable destination: helper threads copy each strip out of the staging buffer as soon as its event has fired
    const int NT = 4;
    const int dev = ctx->device;
    auto worker = [&](int t) {
        cudaSetDevice(dev);
        int cur = -2;
        for (const auto& r : ranges) {
            if (r.ev != cur) { if (r.ev >= 0) cudaEventSynchronize(cev[r.ev]); else cudaStreamSynchronize(ctx->copy_stream); cur = r.ev; }
            const size_t a = r.n * t / NT, b = r.n * (t + 1) / NT;
            std::memcpy(dst_host + r.dst + a, land + r.dst + a, sizeof(double) * (b - a));
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < NT; ++t) th.emplace_back(worker, t);
    worker(0);
    for (auto& x : th) x.join();
    KL_CUDA(cudaStreamSynchronize(ctx->copy_stream));
    return 0;
}

static int jacobian_host(kl_ctx* ctx, const double* x_host, double* values_host, bool lower) {
    if (!ctx) { kl_set_error("null context"); return KL_E_ARG; }
    KL_CUDA(cudaSetDevice(ctx->device));
    const int n = ctx->d.nfree;
    cudaStream_t s = ctx->stream;
    KL_CUDA(cudaEventRecord(ctx->ev[0], s));
    const double* xd = nullptr;
    if (x_host) {
        KL_CUDA(cudaMemcpyAsync(ctx->d_x, stage_x(ctx, x_host, n), sizeof(double) * n, cudaMemcpyHostToDevice, s));
        xd = ctx->d_x;
    }
    KL_CUDA(cudaEventRecord(ctx->ev[1], s));
    int rc;
    const bool whole = ctx->e2_begin == 0 && ctx->e2_end == ctx->d.nel2;
    if (lower && (rc = kl_lower_tables(ctx))) return rc;
    if (!values_host || !whole || ctx->d2h_plan.size() < 2) {
        if ((rc = kl_jacobian_device(ctx, xd, s))) return rc;
        if (values_host && lower && (rc = kl_launch_pack_lower(ctx, 0, n, s))) return rc;
        KL_CUDA(cudaEventRecord(ctx->ev[2], s));
        if (values_host) {
            KL_CUDA(cudaEventRecord(ctx->strip_ev[0], s));
            std::vector<CopyRange> one{{0, 0, (size_t)(lower ? ctx->nnz_lower : ctx->nnz), 0}};
            if ((rc = copy_out(ctx, lower ? ctx->d_values_lower : ctx->d.values, values_host, one, one[0].n))) return rc;
        } else {
            KL_CUDA(cudaEventRecord(ctx->ev[3], s));
        }
    } else {
        // strips of element rows on the compute stream; the value ranges a strip completes are copied out on the copy
        // stream while the next strip is assembled (only `double` values ever cross PCIe)
        if ((rc = points_and_jacobian(ctx, xd, s, 0, ctx->d.nel2, false))) return rc;
        std::vector<CopyRange> ranges;
        for (size_t k = 0; k < ctx->d2h_plan.size(); ++k) {
            const auto& st = ctx->d2h_plan[k];
            if ((rc = kl_launch_jacobian(ctx, st.e2_begin, st.e2_end, s))) return rc;
            if (lower) {
                // columns completed by this strip, packed to their lower-triangular part (a contiguous range of the packed array)
                for (const auto& c : st.cols) {
                    if ((rc = kl_launch_pack_lower(ctx, c.first, c.second, s))) return rc;
                    const size_t a = (size_t)ctx->h_outer_lower[c.first], b = (size_t)ctx->h_outer_lower[c.second];
                    ranges.push_back({a, a, b - a, (int)k});
                }
            } else {
                for (const auto& r : st.ranges) ranges.push_back({r.first, r.first, r.second - r.first, (int)k});
            }
            KL_CUDA(cudaEventRecord(ctx->strip_ev[k], s));
        }
        KL_CUDA(cudaEventRecord(ctx->ev[2], s));
        if ((rc = copy_out(ctx, lower ? ctx->d_values_lower : ctx->d.values, values_host, ranges, (size_t)(lower ? ctx->nnz_lower : ctx->nnz)))) return rc;
    }
    rc = kl_check(ctx, s);
    cudaEventElapsedTime(&ctx->ms_h2d, ctx->ev[0], ctx->ev[1]);
    cudaEventElapsedTime(&ctx->ms_kernel, ctx->ev[1], ctx->ev[2]);
    cudaEventElapsedTime(&ctx->ms_d2h, ctx->ev[2], ctx->ev[3]);   // copy-out time NOT hidden behind the assembly
    return rc;
}

extern "C" int kl_jacobian(kl_ctx* ctx, const double* x_host, double* values_host) { return jacobian_host(ctx, x_host, values_host, false); }
extern "C" int kl_jacobian_lower(kl_ctx* ctx, const double* x_host, double* values_lower_host) {
    if (ctx && ctx->d.mat.pressure != 0.0) { kl_set_error("kl_jacobian_lower: the follower-pressure tangent is unsymmetric"); return KL_E_ARG; }
    return jacobian_host(ctx, x_host, values_lower_host, true);
}

extern "C" int kl_pin_values(kl_ctx* ctx, double* values_host, int64_t count) {
    if (!ctx || !values_host || count <= 0) return KL_E_ARG;
    KL_CUDA(cudaSetDevice(ctx->device));
    if (host_is_pinned(values_host)) return KL_OK;
    KL_CUDA(cudaHostRegister(values_host, sizeof(double) * (size_t)count, cudaHostRegisterDefault));
    return KL_OK;
}
extern "C" int kl_unpin_values(kl_ctx* ctx, double* values_host) {
    if (!ctx || !values_host) return KL_E_ARG;
    KL_CUDA(cudaSetDevice(ctx->device));
    KL_CUDA(cudaHostUnregister(values_host));
    return KL_OK;
}

extern "C" int kl_fetch_values(kl_ctx* ctx, double* values_host) {
    if (!ctx || !values_host) return KL_E_ARG;
    KL_CUDA(cudaSetDevice(ctx->device));
    KL_CUDA(cudaStreamSynchronize(ctx->stream));
    KL_CUDA(cudaEventRecord(ctx->strip_ev[0], ctx->stream));
    std::vector<CopyRange> one{{0, 0, (size_t)ctx->nnz, 0}};
    return copy_out(ctx, ctx->d.values, values_host, one, (size_t)ctx->nnz);
}
extern "C" int kl_set_values(kl_ctx* ctx, const double* values_host) {
    if (!ctx || !values_host) return KL_E_ARG;
    KL_CUDA(cudaSetDevice(ctx->device));
    KL_CUDA(cudaMemcpyAsync(ctx->d.values, values_host, sizeof(double) * (size_t)ctx->nnz, cudaMemcpyHostToDevice, ctx->stream));
    KL_CUDA(cudaStreamSynchronize(ctx->stream));
    return KL_OK;
}

extern "C" int kl_pattern_lower_host(kl_ctx* ctx, int32_t* outer_lower, int32_t* inner_lower, int64_t* nnz_lower) {
    if (!ctx) return KL_E_ARG;
    KL_CUDA(cudaSetDevice(ctx->device));
    if (int rc = kl_lower_tables(ctx)) return rc;
    if (nnz_lower) *nnz_lower = ctx->nnz_lower;
    if (outer_lower) KL_CUDA(cudaMemcpy(outer_lower, ctx->d_outer_lower, sizeof(int) * ((size_t)ctx->d.nfree + 1), cudaMemcpyDeviceToHost));
    if (inner_lower) KL_CUDA(cudaMemcpy(inner_lower, ctx->d_inner_lower, sizeof(int) * (size_t)ctx->nnz_lower, cudaMemcpyDeviceToHost));
    return KL_OK;
}

extern "C" int kl_last_timing(const kl_ctx* ctx, float* ms_kernel, float* ms_h2d, float* ms_d2h) {
    if (!ctx) return KL_E_ARG;
    if (ms_kernel) *ms_kernel = ctx->ms_kernel;
    if (ms_h2d) *ms_h2d = ctx->ms_h2d;
    if (ms_d2h) *ms_d2h = ctx->ms_d2h;
    return KL_OK;
}
