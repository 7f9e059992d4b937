/** examples/solid_newton.cpp — the reference's solid tutorial driver shape (tutorials/nonlinear_solid_static.cpp:99-131) on the
    B200 operators: a clamped brick under a dead end traction (benchmarks/benchmark_Elasticity_Beam_APALM.cpp:196-236),
    St.Venant-Kirchhoff, Jacobian_t / Residual_t closures from gsElasticityAssemblerB200, Newton with the CGDiagonal solve
    on the host (as the reference does).  Exit code 0 = converged (or no GPU: prints NO_GPU and exits 0). */
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../include/gsStructuralAnalysisOps_b200.h"

using namespace gismo;

static std::vector<double> open_knots(int p, int nel) {
    std::vector<double> U;
    for (int k = 0; k <= p; ++k) U.push_back(0.0);
    for (int k = 1; k < nel; ++k) U.push_back((double)k / nel);
    for (int k = 0; k <= p; ++k) U.push_back(1.0);
    return U;
}
static std::vector<double> greville(int p, const std::vector<double>& U) {
    std::vector<double> g(U.size() - p - 1);
    for (size_t i = 0; i < g.size(); ++i) { double s = 0; for (int k = 1; k <= p; ++k) s += U[i + k]; g[i] = s / p; }
    return g;
}

static int pcg(const gsSparseMatrix<>& A, const gsVector<>& b, gsVector<>& x, double tol, int maxit) {
    const index_t n = b.size();
    gsVector<> r(n), z(n), p(n), Ap(n), dinv(n);
    x.setZero(n);
    for (index_t i = 0; i < n; ++i) { dinv[i] = 1.0 / A.diagonal(i); r[i] = b[i]; z[i] = dinv[i] * r[i]; p[i] = z[i]; }
    double rz = 0; for (index_t i = 0; i < n; ++i) rz += r[i] * z[i];
    const double bn = b.norm();
    for (int it = 0; it < maxit; ++it) {
        A.apply(p, Ap);
        double pAp = 0; for (index_t i = 0; i < n; ++i) pAp += p[i] * Ap[i];
        const double alpha = rz / pAp;
        for (index_t i = 0; i < n; ++i) { x[i] += alpha * p[i]; r[i] -= alpha * Ap[i]; }
        if (r.norm() <= tol * bn) return it + 1;
        double rz2 = 0; for (index_t i = 0; i < n; ++i) { z[i] = dinv[i] * r[i]; rz2 += r[i] * z[i]; }
        const double beta = rz2 / rz; rz = rz2;
        for (index_t i = 0; i < n; ++i) p[i] = z[i] + beta * p[i];
    }
    return -1;
}

int main() {
    const int p[3] = {2, 2, 2}, nel[3] = {6, 2, 2};
    const double L = 1.0, B = 0.2, H = 0.2;
    std::vector<double> U[3], g[3];
    for (int d = 0; d < 3; ++d) { U[d] = open_knots(p[d], nel[d]); g[d] = greville(p[d], U[d]); }
    const int n1 = (int)g[0].size(), n2 = (int)g[1].size(), n3 = (int)g[2].size();
    std::vector<double> cp((size_t)3 * n1 * n2 * n3);       // a linear map is reproduced by the Greville abscissae
    for (int i3 = 0; i3 < n3; ++i3)
        for (int i2 = 0; i2 < n2; ++i2)
            for (int i1 = 0; i1 < n1; ++i1) {
                double* x = &cp[3 * (size_t)(i1 + n1 * (i2 + n2 * i3))];
                x[0] = L * g[0][i1]; x[1] = B * g[1][i2]; x[2] = H * g[2][i3];
            }
    ks_bc bc = {};
    for (int c = 0; c < 3; ++c) bc.side[KS_WEST][c] = 1;   // BCs.addCondition(0, boundary::west, condition_type::dirichlet, nullptr, c)
    std::vector<int32_t> map(cp.size());
    int32_t nfree = 0, nfixed = 0;
    ks_build_dofmap(n1, n2, n3, &bc, map.data(), &nfree, &nfixed);
    const int32_t side = KS_EAST;
    const double traction[3] = {0.0, 0.0, 2e-3};
    ks_problem P = {};
    for (int d = 0; d < 3; ++d) { P.degree[d] = p[d]; P.n_knots[d] = (int32_t)U[d].size(); P.knots[d] = U[d].data(); }
    P.cp = cp.data(); P.dof_map = map.data(); P.n_free = nfree; P.n_fixed = nfixed;
    P.material_law = KS_LAW_SVK; P.E = 1.0; P.nu = 0.3;
    P.n_tractions = 1; P.traction_side = &side; P.traction_val = traction;

    std::unique_ptr<gsElasticityAssemblerB200> assembler;
    try {
        assembler.reset(new gsElasticityAssemblerB200(P));
    } catch (const std::exception& e) {
        std::printf("NO_GPU %s\n", e.what());
        return 0;
    }
    gsStructuralAnalysisOps<real_t>::Jacobian_t Jacobian = assembler->jacobian();
    gsStructuralAnalysisOps<real_t>::Residual_t Residual = assembler->residual();
    const index_t n = assembler->numDofs();
    std::printf("Solving system with %d DoFs, %lld non-zeros\n", n, (long long)assembler->nonZeros());
    gsVector<> Usol(n), dU(n), R(n);
    gsSparseMatrix<> K;
    Usol.setZero(n);
    if (!Residual(Usol, R)) return 1;
    const double R0 = R.norm();
    gsStatus status = gsStatus::NotConverged;
    for (int it = 0; it < 25; ++it) {
        if (!Jacobian(Usol, K)) { status = gsStatus::AssemblyError; break; }
        const int cg = pcg(K, R, dU, 1e-12, 50 * n);
        if (cg < 0) { status = gsStatus::SolverError; break; }
        Usol += dU;
        if (!Residual(Usol, R)) { status = gsStatus::AssemblyError; break; }
        std::printf("it %2d  |dU|/|U| = %.3e  |R|/|R0| = %.3e  (cg %d)\n", it, dU.norm() / Usol.norm(), R.norm() / R0, cg);
        if (dU.norm() / Usol.norm() < 1e-6 && R.norm() / R0 < 1e-9) { status = gsStatus::Success; break; }
    }
    // one-pass variant must agree with the two closures
    gsVector<> R2(n);
    gsSparseMatrix<> K2;
    if (!assembler->assemble(Usol, K2, R2)) return 1;
    double diff = 0;
    for (index_t i = 0; i < n; ++i) diff = std::max(diff, std::fabs(R2[i] - R[i]));
    std::printf("STATUS %s |U| = %.12e  one-pass rhs diff %.3e\n", status == gsStatus::Success ? "Success" : "NotConverged", Usol.norm(), diff);
    return status == gsStatus::Success && diff <= 1e-12 * R0 ? 0 : 1;
}
