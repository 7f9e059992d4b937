/** examples/newton_shell.cpp — the reference's tutorial driver shape (tutorials/nonlinear_shell_static.cpp:115-160)
    on top of the B200 operators: define Jacobian / Residual closures, solve K du = R with a Newton loop
    (src/gsStaticSolvers/gsStaticNewton.hpp:142-193) and the reference's default "CGDiagonal" linear solver
    (Jacobi-preconditioned CG, gsStaticNewton.hpp:23).  The first pass keeps the linear solve on the host, as in the reference; the
    second pass repeats it with gsThinShellAssemblerB200::newtonSolve (device-resident Jacobian + CGDiagonal + residual).

    usage: newton_shell problem.klp [max_iterations]        (problem.klp written by ShellProblem.save)
    exit code 0 = converged (or no GPU present: prints the reason and exits 0 so that CPU-only CI can build/run it). */
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>

#include "../include/gsStructuralAnalysisOps_b200.h"
#include "problem_file.h"

using namespace gismo;


// Jacobi-preconditioned CG ("CGDiagonal")
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

int main(int argc, char** argv) {
    if (argc < 2) { std::fprintf(stderr, "usage: %s problem.klp [maxit]\n", argv[0]); return 2; }
    ProblemFile pf;
    if (!pf.load(argv[1])) { std::fprintf(stderr, "cannot read %s\n", argv[1]); return 2; }
    const int maxIt = argc > 2 ? std::atoi(argv[2]) : 25;
    std::unique_ptr<gsThinShellAssemblerB200> assembler;
    try {
        assembler.reset(new gsThinShellAssemblerB200(pf.P));
    } catch (const std::exception& e) {
        std::printf("NO_GPU %s\n", e.what());     // libkl_shell has no CPU fallback
        return 0;
    }
    gsStructuralAnalysisOps<real_t>::Jacobian_t Jacobian = assembler->jacobian();
    gsStructuralAnalysisOps<real_t>::Residual_t Residual = assembler->residual();
    const index_t n = assembler->numDofs();
    std::printf("Solving system with %d DoFs, %lld non-zeros\n", n, (long long)assembler->nonZeros());
    gsVector<> U(n), dU(n), R(n);
    gsSparseMatrix<> K;
    U.setZero(n);
    if (!Residual(U, R)) return 1;
    const double R0 = R.norm();
    gsStatus status = gsStatus::NotConverged;
    for (int it = 0; it < maxIt; ++it) {
        if (!Jacobian(U, K)) { status = gsStatus::AssemblyError; break; }
        const int cg = pcg(K, R, dU, 1e-12, 20 * n);
        if (cg < 0) { status = gsStatus::SolverError; break; }
        U += dU;
        if (!Residual(U, R)) { status = gsStatus::AssemblyError; break; }
        std::printf("it %2d  |dU|/|U| = %.3e  |R|/|R0| = %.3e  (cg %d)\n", it, dU.norm() / U.norm(), R.norm() / R0, cg);
        if (dU.norm() / U.norm() < 1e-6 && R.norm() / R0 < 1e-9) { status = gsStatus::Success; break; }   // tolU, tolF
    }
    std::printf("STATUS %s |U| = %.12e\n", status == gsStatus::Success ? "Success" : "NotConverged", U.norm());

    // The same solve with everything resident on the GPU (Jacobian, CGDiagonal, residual, norms): only U comes back.
    kl_newton_options opt = gsThinShellAssemblerB200::defaultNewtonOptions();
    opt.tolU = 1e-6; opt.tolF = 1e-9; opt.max_it = maxIt; opt.cg_tol = 1e-12; opt.cg_max_iter = 20 * n;
    kl_newton_info info;
    gsVector<> Ud(n);
    Ud.setZero(n);
    const gsStatus dstatus = assembler->newtonSolve(Ud, opt, &info);
    double diff = 0;
    for (index_t i = 0; i < n; ++i) diff = std::max(diff, std::fabs(Ud[i] - U[i]));
    std::printf("DEVICE_NEWTON %s iterations %d cg %lld |U| = %.12e  max|U_dev - U_host| = %.3e  assembly %.2f ms  solve %.2f ms\n",
                dstatus == gsStatus::Success ? "Success" : "NotConverged", info.iterations, (long long)info.cg_iterations, Ud.norm(),
                diff, info.ms_assembly, info.ms_solve);
    if (dstatus != gsStatus::Success || diff > 1e-6 * U.norm()) return 1;

    // Third pass: the loop of gsStaticNewton::_solveNonlinear (src/gsStaticSolvers/gsStaticNewton.hpp:142-193) written against the
    // solver interface it uses — jacMat = computeJacobian(U); m_solver->compute(jacMat); deltaU = m_solver->solve(R) (:205-240) — with
    // the assembler in Device copy-out mode and gsSparseSolverB200 as m_solver: K never leaves the GPU.
    {
        assembler->setCopyOut(gsB200CopyOut::Device);
        gsSparseSolver<real_t>::uPtr m_solver(new gsSparseSolverB200<real_t>(*assembler));
        static_cast<gsSparseSolverB200<real_t>*>(m_solver.get())->setTolerance(1e-12);
        static_cast<gsSparseSolverB200<real_t>*>(m_solver.get())->setMaxIterations(20 * n);
        gsVector<> U3(n), R3(n);
        U3.setZero(n);
        if (!Residual(U3, R3)) return 1;
        gsStatus st3 = gsStatus::NotConverged;
        gsSparseMatrix<> jacMat;
        for (int it = 0; it < maxIt; ++it) {
            if (!Jacobian(U3, jacMat)) { st3 = gsStatus::AssemblyError; break; }
            m_solver->compute(jacMat);
            if (m_solver->info() != 0) { st3 = gsStatus::SolverError; break; }
            gsVector<> d3 = m_solver->solve(R3);
            if (!m_solver->succeed()) { st3 = gsStatus::SolverError; break; }
            U3 += d3;
            if (!Residual(U3, R3)) { st3 = gsStatus::AssemblyError; break; }
            if (d3.norm() / U3.norm() < 1e-6 && R3.norm() / R0 < 1e-9) { st3 = gsStatus::Success; break; }
        }
        double d3max = 0;
        for (index_t i = 0; i < n; ++i) d3max = std::max(d3max, std::fabs(U3[i] - U[i]));
        // a host matrix handed to the same solver is uploaded (compute) and gives the same solution; the placeholder can be fetched
        gsSparseMatrix<> Kfetched;
        bool ok3 = assembler->fetch(Kfetched);
        assembler->setCopyOut(gsB200CopyOut::Full);
        gsSparseMatrix<> Kfull;
        ok3 = ok3 && Jacobian(U3, Kfull);
        double dk = 0, kmax = 0;
        for (index_t k = 0; k < Kfull.nonZeros(); ++k) { dk = std::max(dk, std::fabs(Kfull.valuePtr()[k] - Kfetched.valuePtr()[k])); kmax = std::max(kmax, std::fabs(Kfull.valuePtr()[k])); }
        m_solver->compute(Kfull);
        gsVector<> xs = m_solver->solve(R);
        ok3 = ok3 && m_solver->succeed();
        // lower-triangular copy-out: the entries with row >= col of the same matrix
        assembler->setCopyOut(gsB200CopyOut::Lower);
        gsSparseMatrix<> Klow;
        ok3 = ok3 && Jacobian(U3, Klow);
        double dl = 0;
        index_t nlow = 0;
        for (index_t j = 0; j < n && ok3; ++j) {
            index_t kl = Klow.outerIndexPtr()[j];
            for (index_t k = Kfull.outerIndexPtr()[j]; k < Kfull.outerIndexPtr()[j + 1]; ++k) {
                if (Kfull.innerIndexPtr()[k] < j) continue;
                if (kl >= Klow.outerIndexPtr()[j + 1] || Klow.innerIndexPtr()[kl] != Kfull.innerIndexPtr()[k]) { ok3 = false; break; }
                dl = std::max(dl, std::fabs(Klow.valuePtr()[kl] - Kfull.valuePtr()[k]));
                ++kl; ++nlow;
            }
        }
        assembler->setCopyOut(gsB200CopyOut::Full);
        std::printf("SPARSE_SOLVER_B200 %s max|U3 - U| = %.3e  fetched-vs-full %.3e  lower-vs-full %.3e (%d of %d entries)\n",
                    st3 == gsStatus::Success ? "Success" : "NotConverged", d3max, dk / kmax, dl / kmax, (int)nlow, (int)Klow.nonZeros());
        if (st3 != gsStatus::Success || !ok3 || d3max > 1e-6 * U.norm() || dk > 1e-12 * kmax || dl > 1e-12 * kmax || nlow != Klow.nonZeros()) return 1;
    }

    // Post-processing of the converged state as the reference's drivers do it (unittests/gsStaticSolver_test.cpp:313-324,
    // benchmarks/benchmark_Balloon.cpp:381-408): principal stretches, boundary reaction, membrane Cauchy stress.
    const std::vector<real_t> pt = {0.5, 0.5};
    std::vector<real_t> lambdas, sigma;
    real_t fw[3];
    if (!assembler->computePrincipalStretches(pt, U, 0.0, lambdas) || !assembler->boundaryForce(U, KL_WEST, fw) ||
        !assembler->evalStress(U, KL_STRESS_MEMBRANE, pt, sigma)) {
        std::printf("POSTPROCESS failed: %s\n", kl_last_error());
        return 1;
    }
    // the time-dependent operator shapes of the dynamic solvers wrap the same calls
    gsVector<> Rt(n), Rs(n);
    if (!assembler->tResidual()(U, 0.5, Rt) || !assembler->residual()(U, Rs)) return 1;
    double dr = 0;
    for (index_t i = 0; i < n; ++i) dr = std::max(dr, std::fabs(Rt[i] - Rs[i]));
    gsSparseMatrix<> Cd;
    // the assembly accumulates with atomics: two calls agree to rounding, not bit for bit
    if (!assembler->damping()(U, Cd) || Cd.rows() != n || dr > 1e-12 * (Rs.norm() + 1.0)) { std::printf("DYNAMIC_OPS failed\n"); return 1; }
    std::printf("STRETCHES %.12e %.12e %.12e  WEST_FORCE %.6e %.6e %.6e  MEMBRANE %.6e %.6e %.6e\n", lambdas[0], lambdas[1],
                lambdas[2], fw[0], fw[1], fw[2], sigma[0], sigma[1], sigma[2]);
    return status == gsStatus::Success ? 0 : 1;
}
