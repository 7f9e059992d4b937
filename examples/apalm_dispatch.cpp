/** examples/apalm_dispatch.cpp — the reference's parallel arc-length benchmark on one multi-GPU box (SURVEY §8e, configs[4]):
    benchmarks/benchmark_Frustrum_APALM.cpp with gsAPALM (src/gsALMSolvers/gsAPALM.hpp) = a serial chain of coarse Crisfield
    arc-length steps (level 0), then every interval is re-traced with SubIntervals finer steps by a worker that owns one GPU and
    one assembler replica; intervals whose length error exceeds the tolerance are refined and queued again.

    usage: apalm_dispatch --fake  nWorkers nIntervals                                   (no GPU: queue / storage semantics only)
           apalm_dispatch problem.klp nGPUs nSteps dL [subIntervals=2] [tolerance=1e-2] [maxLevel=2] [cgTol=1e-12]

    GPU mode prints one line
      APALM gpus=.. steps=.. jobs=.. points=.. maxLevel=.. failed=.. t_chain_s=.. t_parallel_s=.. sum_job_s=.. speedup_parallel_phase=..
            speedup_total=.. lambda_end=.. per_worker=..
    where t_chain_s is the sequential level-0 chain (it cannot be distributed, gsAPALM.hpp:587-598), sum_job_s the time the same
    correction jobs take one after the other, and speedup_total = (t_chain + sum_job) / (t_chain + t_parallel) the honest APALM
    speed-up of the whole traversal. */
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>

#include "../include/gsAPALM_b200.h"
#include "problem_file.h"

using namespace gismo;

static int run_fake(int nW, int nI) {
    // error model: on level 1 the two computed sub-intervals are 25 % of the interval off (lower error), the gap to the reference is
    // exact (upper error 0); deeper levels are exact -> every initial interval is refined exactly once into its first two children
    gsAPALMDataB200 data(/*tolerance*/ 0.1, /*maxLevel*/ 3);
    std::vector<double> times;
    std::vector<gsAPALMSolutionB200> sols;
    for (int k = 0; k <= nI; ++k) { times.push_back((double)k); sols.push_back(std::make_pair(std::vector<double>{(double)k}, 0.1 * k)); }
    data.setData(times, sols);
    std::mutex mtx;
    std::condition_variable cv;
    std::vector<int> count(nW, 0);
    std::vector<std::thread> threads;
    int ready = 0;
    const auto t0 = std::chrono::steady_clock::now();
    for (int w = 0; w < nW; ++w)
        threads.emplace_back([&, w]() {
            { std::lock_guard<std::mutex> lk(mtx); ++ready; }
            cv.notify_all();
            for (;;) {
                gsAPALMDataB200::Job job;
                {
                    std::unique_lock<std::mutex> lk(mtx);
                    cv.wait(lk, [&]() { return !data.empty() || (data.nActive() == 0 && ready == nW); });
                    if (data.empty()) { cv.notify_all(); return; }
                    job = data.pop();
                }
                std::this_thread::sleep_for(std::chrono::milliseconds(2));
                const double Dt = job.tend - job.tstart, u0 = job.start.first[0], u1 = job.reference.first[0];
                std::vector<gsAPALMSolutionB200> s{std::make_pair(std::vector<double>{u0 + 0.4 * (u1 - u0)}, 0.0),
                                                   std::make_pair(std::vector<double>{u0 + 0.8 * (u1 - u0)}, 0.0)};
                std::vector<double> dist{0.4 * Dt, 0.4 * Dt, 0.2 * Dt};
                const double low = job.level == 1 ? 0.75 * Dt : Dt, upp = low;
                {
                    std::lock_guard<std::mutex> lk(mtx);
                    data.submit(job.ID, dist, s, upp, low);
                    ++count[w];
                }
                cv.notify_all();
            }
        });
    for (auto& t : threads) t.join();
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::printf("APALM workers=%d jobs=%d points=%zu maxLevel=%d failed=%d wall_s=%.4f per_worker=", nW, data.nJobs(), data.nPoints() - (size_t)(nI + 1),
                data.maxLevelSeen(), data.nFailed(), secs);
    for (int c : count) std::printf("%d ", c);
    std::printf("\n");
    const std::vector<double>& t = data.times();      // curve times stay sorted
    for (size_t k = 1; k < t.size(); ++k) if (!(t[k] > t[k - 1])) { std::printf("times not ascending\n"); return 1; }
    return data.nFailed() == 0 ? 0 : 1;
}

int main(int argc, char** argv) {
    if (argc < 4) {
        std::fprintf(stderr, "usage: %s --fake nWorkers nIntervals | problem.klp nGPUs nSteps dL [subIntervals tolerance maxLevel cgTol]\n", argv[0]);
        return 2;
    }
    if (std::strcmp(argv[1], "--fake") == 0) return run_fake(std::atoi(argv[2]), std::atoi(argv[3]));
    if (argc < 5) { std::fprintf(stderr, "missing dL\n"); return 2; }
    ProblemFile pf;
    if (!pf.load(argv[1])) { std::fprintf(stderr, "cannot read %s\n", argv[1]); return 2; }
    const int nGPU = std::atoi(argv[2]), nSteps = std::atoi(argv[3]);
    const double dL = std::atof(argv[4]);
    const int sub = argc > 5 ? std::atoi(argv[5]) : 2;
    const double tol = argc > 6 ? std::atof(argv[6]) : 1e-2;
    const int maxLevel = argc > 7 ? std::atoi(argv[7]) : 2;
    const double cgTol = argc > 8 ? std::atof(argv[8]) : 1e-12;
    // benchmark_Frustrum_APALM.cpp:435-452: CGDiagonal, AngleMethod 0, Scaling 0, TolU / TolF, MaxIter
    kl_alm_options opt;
    opt.tolU = 1e-6; opt.tolF = 1e-3; opt.max_it = 20; opt.phi = 0.0; opt.relaxation = 1.0; opt.cg_tol = cgTol; opt.cg_max_iter = 0;
    auto factory = [&](int dev) {
        std::unique_ptr<gsAPALMWorkerB200> w(new gsAPALMWorkerB200());
        if (kl_create(&pf.P, dev, &w->ctx) != KL_OK) throw std::runtime_error(std::string("kl_create: ") + kl_last_error());
        int32_t n = 0;
        kl_sizes(w->ctx, &n, nullptr, nullptr, nullptr);
        w->alm.reset(new gsALMCrisfieldB200(w->ctx, n, opt));
        return w;
    };
    std::unique_ptr<gsAPALMWorkerB200> w0;
    try { w0 = factory(0); } catch (const std::exception& e) { std::printf("NO_GPU %s\n", e.what()); return 0; }
    gsAPALMB200 apalm(factory, dL, sub, tol, maxLevel);
    double t_chain = 0;
    gsAPALMB200::ParallelStats st;
    try {
        t_chain = apalm.serialSolve(*w0, nSteps);
        st = apalm.parallelSolve(nGPU, w0.get());
    } catch (const std::exception& e) {
        std::printf("APALM_ERROR %s\n", e.what());
        return 1;
    }
    gsAPALMDataB200& data = apalm.data();
    std::printf("APALM gpus=%d steps=%d jobs=%d points=%zu maxLevel=%d failed=%d t_chain_s=%.4f t_parallel_s=%.4f sum_job_s=%.4f "
                "speedup_parallel_phase=%.3f speedup_total=%.3f lambda_end=%.8f per_worker=",
                nGPU, nSteps, data.nJobs(), data.nPoints(), data.maxLevelSeen(), data.nFailed(), t_chain, st.wall_s, st.sum_job_s,
                st.sum_job_s / st.wall_s, (t_chain + st.sum_job_s) / (t_chain + st.wall_s), data.loadFactors().back());
    for (int c : st.jobs_per_worker) std::printf("%d ", c);
    std::printf("\n");
    return data.nFailed() == 0 ? 0 : 1;
}
