/** examples/apalm_dispatch.cpp — arc-length intervals dispatched one per GPU (gsAPALM master/worker).

    usage: apalm_dispatch --fake  nWorkers nIntervals                 (no GPU: queue semantics only)
           apalm_dispatch problem.klp nGPUs nIntervals stepsPerJob    (each job = stepsPerJob Jacobian+residual assemblies)

    In the reference every MPI rank builds its own assembler and arc-length solver and receives (start, previous,
    reference) solution vectors per job (benchmarks/benchmark_Frustrum_APALM.cpp:391-458, gsAPALM.hpp:1173-1279); here the
    workers are host threads with one kl_ctx per GPU and the job body is the assembly part of gsALMBase::step
    (one Jacobian + one residual per corrector iteration, src/gsALMSolvers/gsALMBase.hpp:354-416). */
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>

#include "../include/gsAPALMDispatcher_b200.h"
#include "../include/gsStructuralAnalysisOps_b200.h"
#include "problem_file.h"

using namespace gismo;


struct FakeWorker { int device; };

struct GpuWorker {
    std::unique_ptr<gsThinShellAssemblerB200> assembler;
    gsStructuralAnalysisOps<real_t>::Jacobian_t Jacobian;
    gsStructuralAnalysisOps<real_t>::ALResidual_t ALResidual;
    gsVector<> U, R;
    gsSparseMatrix<> K;
};

int main(int argc, char** argv) {
    if (argc < 4) { std::fprintf(stderr, "usage: %s --fake nWorkers nIntervals | problem.klp nGPUs nIntervals steps\n", argv[0]); return 2; }
    const bool fake = std::strcmp(argv[1], "--fake") == 0;
    const int nW = std::atoi(argv[2]), nI = std::atoi(argv[3]);
    gsAPALMQueueB200 queue(/*tolerance*/ 0.1, /*maxLevel*/ 3);
    for (int k = 0; k < nI; ++k) queue.addInterval(k / (double)nI, (k + 1) / (double)nI, 1);
    const auto t0 = std::chrono::steady_clock::now();
    std::vector<int> count;
    if (fake) {
        // error model: a level-1 interval is 25 % off on both halves, deeper levels are exact -> every initial interval is
        // refined exactly once into two children (SubIntervals = 2)
        gsAPALMDispatcherB200<FakeWorker> disp(
            nW, [](int dev) { return std::unique_ptr<FakeWorker>(new FakeWorker{dev}); },
            [](FakeWorker&, const gsAPALMIntervalB200& iv, int) {
                gsAPALMJobResultB200 r;
                r.xi = {0.5 * (iv.xilow + iv.xiupp)};
                const double Dt = iv.xiupp - iv.xilow;
                r.lowerError = r.upperError = (iv.level == 1) ? 0.25 * Dt : 0.0;
                std::this_thread::sleep_for(std::chrono::milliseconds(2));
                return r;
            });
        count = disp.solve(queue);
    } else {
        ProblemFile pf;
        if (!pf.load(argv[1])) { std::fprintf(stderr, "cannot read %s\n", argv[1]); return 2; }
        const int steps = argc > 4 ? std::atoi(argv[4]) : 4;
        try { gsThinShellAssemblerB200 probe(pf.P, 0); } catch (const std::exception& e) { std::printf("NO_GPU %s\n", e.what()); return 0; }
        gsAPALMDispatcherB200<GpuWorker> disp(
            nW,
            [&pf](int dev) {
                std::unique_ptr<GpuWorker> w(new GpuWorker());
                w->assembler.reset(new gsThinShellAssemblerB200(pf.P, dev));
                w->Jacobian = w->assembler->jacobian();
                w->ALResidual = w->assembler->alResidual();
                w->U.setZero(w->assembler->numDofs());
                return w;
            },
            [steps](GpuWorker& w, const gsAPALMIntervalB200& iv, int) {
                gsAPALMJobResultB200 r;
                const double lam = 0.5 * (iv.xilow + iv.xiupp);
                for (int s = 0; s < steps; ++s) {      // corrector iterations of one arc-length step: 1 Jacobian + 1 residual each
                    for (index_t i = 0; i < w.U.size(); ++i) w.U[i] = 1e-7 * lam * ((i * 2654435761u % 1000) / 500.0 - 1.0);
                    r.ok = r.ok && w.Jacobian(w.U, w.K) && w.ALResidual(w.U, lam, w.R);
                }
                r.xi = {lam};
                return r;
            });
        count = disp.solve(queue);
    }
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    int total = 0;
    for (int c : count) total += c;
    int maxLevel = 0;
    for (auto& iv : queue.finished()) maxLevel = std::max(maxLevel, iv.level);
    std::printf("APALM workers=%d jobs=%d points=%zu maxLevel=%d failed=%d wall_s=%.4f per_worker=", nW, total, queue.points().size(),
                maxLevel, queue.nFailed(), secs);
    for (int c : count) std::printf("%d ", c);
    std::printf("\n");
    return queue.nFailed() == 0 ? 0 : 1;
}
