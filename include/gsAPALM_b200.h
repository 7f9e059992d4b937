/** @file gsAPALM_b200.h

    Single-node, multi-GPU version of the reference's Adaptive Parallel Arc-Length Method:

      gsALMCrisfieldB200   the state and calls of gsALMBase / gsALMCrisfield that gsAPALM uses
                           (src/gsALMSolvers/gsALMBase.h:150-211: setLength, setSolution, setPrevious, step, solutionU/L, distance),
                           every step running device resident through kl_alm_step (include/kl_shell.h)
      gsAPALMDataB200      gsAPALMData<T, solution_t> (src/gsALMSolvers/gsAPALMData.hpp): hierarchical interval queue WITH the
                           solution / previous-solution / curve-time storage (:215-252 pop, :291-433 submit, :435-470 job data)
      gsAPALMB200          gsAPALM<T>: serialSolve = the level-0 initiation chain (_initiation, gsAPALM.hpp:935-1010), parallelSolve =
                           the correction jobs (_correction, :1014-1165) dispatched to worker threads, one GPU and one assembler
                           replica each, exactly as every MPI rank owns its own assembler + arc-length solver in
                           benchmarks/benchmark_Frustrum_APALM.cpp:391-458.  The messages the reference sends per job
                           (start, previous and reference solution, gsAPALM.hpp:1214-1235) are handed over in host memory; no
                           matrix ever leaves a GPU.

    On one 8xB200 box the dispatcher is a host thread, so all N GPUs are workers (the MPI reference keeps rank 0 as a pure
    dispatcher: 8 ranks = 7 workers, gsAPALM.hpp:538-566). */
#pragma once

#include <algorithm>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <deque>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <numeric>
#include <stdexcept>
#include <string>
#include <thread>
#include <tuple>
#include <vector>

#include "kl_shell.h"

namespace gismo {

typedef std::pair<std::vector<double>, double> gsAPALMSolutionB200;      // solution_t = (U, lambda)

/// gsALMCrisfield behind the calls gsAPALM makes; one object per worker (GPU)
class gsALMCrisfieldB200 {
public:
    gsALMCrisfieldB200(kl_ctx* ctx, int numDofs, const kl_alm_options& opt)
        : m_ctx(ctx), m_n(numDofs), m_opt(opt), m_U(numDofs, 0.0), m_DeltaUold(numDofs, 0.0) {}

    void setLength(double dL) { m_arcLength = dL; }
    void setSolution(const std::vector<double>& U, double L) { m_U = U; m_L = L; }
    /// gsALMBase::setPrevious (gsALMBase.h:185-191): the direction of travel
    void setPrevious(const std::vector<double>& Uprev, double Lprev) {
        for (int i = 0; i < m_n; ++i) m_DeltaUold[i] = m_U[i] - Uprev[i];
        m_DeltaLold = m_L - Lprev;
    }
    /// gsStatus: 0 Success, 1 NotConverged, 2 AssemblyError, 3 SolverError, 5 OtherError
    int step() {
        const int rc = kl_alm_step(m_ctx, m_U.data(), &m_L, m_DeltaUold.data(), &m_DeltaLold, m_arcLength, &m_opt, &m_info);
        if (rc != KL_OK) return 5;
        m_phi = m_info.phi;
        m_iterations += m_info.iterations;
        m_cgIterations += m_info.cg_iterations;
        m_msAssembly += m_info.ms_assembly;
        m_msSolve += m_info.ms_solve;
        ++m_steps;
        return m_info.status;
    }
    const std::vector<double>& solutionU() const { return m_U; }
    double solutionL() const { return m_L; }
    /** gsALMBase::computeStability(jacobian = true) with the "Determinant" method (gsALMBase.hpp:546-611): K(m_U) is assembled and
        factorised (L D L^T) on the device, m_indicator = min D, m_negatives = #(D < 0).  Returns false on assembly / factorisation
        failure (the reference throws 2 / 3). */
    bool computeStability() {
        m_stabilityPrev = stability();
        if (kl_jacobian(m_ctx, m_U.data(), nullptr) != KL_OK) return false;
        return kl_stability(m_ctx, &m_indicator, &m_negatives, nullptr) == KL_OK;
    }
    double indicator() const { return m_indicator; }
    int negatives() const { return m_negatives; }
    int stability() const { return m_indicator < 0 ? 1 : -1; }                                   // gsALMBase.hpp:617-621
    bool stabilityChange() const { return stability() * m_stabilityPrev < 0; }                   // gsALMBase.hpp:623-627
    /// gsALMCrisfield::distance: sqrt(DeltaU.DeltaU + phi^2 F.F DeltaL^2); the benchmark sets Scaling = 0 (:440)
    double distance(const std::vector<double>& DU, double DL) const {
        double s = 0;
        for (double v : DU) s += v * v;
        return std::sqrt(s + m_A0 * DL * DL);
    }
    void setA0(double A0) { m_A0 = A0; }
    const kl_alm_info& info() const { return m_info; }
    long steps() const { return m_steps; }
    long iterations() const { return m_iterations; }
    long cgIterations() const { return m_cgIterations; }
    double msAssembly() const { return m_msAssembly; }
    double msSolve() const { return m_msSolve; }

private:
    kl_ctx* m_ctx;
    int m_n;
    kl_alm_options m_opt;
    kl_alm_info m_info{};
    std::vector<double> m_U, m_DeltaUold;
    double m_L = 0, m_DeltaLold = 0, m_arcLength = 1e-2, m_phi = 0, m_A0 = 0;
    double m_indicator = 0;
    int m_stabilityPrev = -1;
    int32_t m_negatives = 0;
    long m_steps = 0, m_iterations = 0, m_cgIterations = 0;
    double m_msAssembly = 0, m_msSolve = 0;
};

/// gsAPALMData<T, solution_t>: not thread safe by itself, gsAPALMB200 serialises access
class gsAPALMDataB200 {
public:
    struct Job {
        int ID = -1, level = 0;
        double dL0 = 0, tstart = 0, tend = 0;
        gsAPALMSolutionB200 start, prev, reference;
    };

    gsAPALMDataB200(double tolerance, int maxLevel) : m_tol(tolerance), m_maxLevel(maxLevel) {}

    /// initialize with the serial solutions at curve times `times` (gsAPALMData::setData + init)
    void setData(const std::vector<double>& times, const std::vector<gsAPALMSolutionB200>& solutions) {
        m_t = times;
        m_xi.resize(times.size());
        for (size_t k = 0; k < times.size(); ++k) m_xi[k] = (times[k] - times.front()) / (times.back() - times.front());
        for (size_t k = 0; k < times.size(); ++k) {
            m_solutions[m_xi[k]] = std::make_shared<gsAPALMSolutionB200>(solutions[k]);
            m_prevs[m_xi[k]] = m_solutions[m_xi[k > 0 ? k - 1 : 0]];
            m_levels[m_xi[k]] = 0;
        }
        for (size_t k = 1; k < times.size(); ++k) m_queue.push_back(std::make_tuple(m_xi[k - 1], m_xi[k], 1));
    }
    bool empty() const { return m_queue.empty(); }
    size_t nActive() const { return m_jobs.size(); }
    size_t nWaiting() const { return m_queue.size(); }

    /// gsAPALMData::pop + jobStartTime / jobLevel / the data gsAPALM::_sendMainToWorker ships (gsAPALM.hpp:1214-1235)
    Job pop() {
        double xilow, xiupp;
        int level;
        std::tie(xilow, xiupp, level) = m_queue.front();
        m_queue.pop_front();
        Job j;
        j.ID = m_ID++;
        j.level = level;
        j.tstart = tmap(xilow);
        j.tend = tmap(xiupp);
        j.dL0 = j.tend - j.tstart;
        j.start = *m_solutions.at(xilow);
        j.prev = *m_prevs.at(xilow);
        j.reference = *m_solutions.at(xiupp);
        m_jobs[j.ID] = std::make_tuple(xilow, xiupp, level);
        return j;
    }

    /// gsAPALMData::submit (:291-433); returns the number of refined intervals queued
    int submit(int ID, const std::vector<double>& distances, const std::vector<gsAPALMSolutionB200>& solutions, double upperDistance,
               double lowerDistance) {
        double xilow, xiupp;
        int level;
        std::tie(xilow, xiupp, level) = m_jobs.at(ID);
        m_jobs.erase(ID);
        const double tlow = tmap(xilow), tupp = tmap(xiupp), dxi = xiupp - xilow, Dt = tupp - tlow;
        std::vector<double> t(distances.size() + 1);
        t[0] = tlow;
        for (size_t k = 0; k < distances.size(); ++k) t[k + 1] = t[k] + distances[k];
        const double dt = t.back() - tlow;
        // the lowerError is the surplus distance over the computed intervals, the upperError the rest of the total error
        const double totalError = dt - upperDistance, lowerError = Dt - lowerDistance, upperError = totalError - lowerError;
        std::vector<double> xi(solutions.size() + 2);
        xi.front() = xilow;
        xi.back() = xiupp;
        for (size_t k = 1; k + 1 < xi.size(); ++k) xi[k] = xilow + dxi * (t[k] - t.front()) / dt;
        size_t kmin = 1, kmax = xi.size();
        if (lowerError / Dt < m_tol) kmin = xi.size() - 1;
        if (upperError / Dt < m_tol) kmax = xi.size() - 1;
        int added = 0;
        for (size_t k = kmin; k < kmax; ++k)
            if (level < m_maxLevel) { m_queue.push_back(std::make_tuple(xi[k - 1], xi[k], level + 1)); ++added; }
        // push the data: later curve times shift by the surplus length, interior points enter the maps
        for (double& tv : m_t) if (tv >= tupp) tv += dt - Dt;
        for (size_t k = 1; k + 1 < t.size(); ++k) {
            const size_t pos = std::lower_bound(m_xi.begin(), m_xi.end(), xi[k]) - m_xi.begin();
            m_xi.insert(m_xi.begin() + pos, xi[k]);
            m_t.insert(m_t.begin() + pos, t[k]);
        }
        for (size_t k = 1; k + 1 < xi.size(); ++k) {
            m_solutions[xi[k]] = std::make_shared<gsAPALMSolutionB200>(solutions[k - 1]);
            m_levels[xi[k]] = level;
        }
        for (size_t k = 1; k + 1 < xi.size(); ++k) m_prevs[xi[k]] = m_solutions.at(xi[k - 1]);
        m_maxLevelSeen = std::max(m_maxLevelSeen, level);
        return added;
    }
    /// a job whose arc-length steps could not be completed: the interval stays as it is
    void abandon(int ID) { m_jobs.erase(ID); ++m_failed; }

    size_t nPoints() const { return m_xi.size(); }
    const std::vector<double>& times() const { return m_t; }
    std::vector<double> loadFactors() const {
        std::vector<double> L;
        for (double x : m_xi) L.push_back(m_solutions.at(x)->second);
        return L;
    }
    int maxLevelSeen() const { return m_maxLevelSeen; }
    int nFailed() const { return m_failed; }
    int nJobs() const { return m_ID; }

private:
    double tmap(double xi) const {
        const size_t k = std::lower_bound(m_xi.begin(), m_xi.end(), xi) - m_xi.begin();
        if (k >= m_xi.size() || m_xi[k] != xi) throw std::runtime_error("gsAPALMDataB200: unknown parametric point");
        return m_t[k];
    }
    double m_tol;
    int m_maxLevel, m_ID = 0, m_failed = 0, m_maxLevelSeen = 0;
    std::vector<double> m_xi, m_t;                                   // sorted parametric points and their curve times
    std::deque<std::tuple<double, double, int>> m_queue;
    std::map<int, std::tuple<double, double, int>> m_jobs;
    std::map<double, std::shared_ptr<gsAPALMSolutionB200>> m_solutions, m_prevs;
    std::map<double, int> m_levels;
};

/// what one worker owns: built inside its thread so that the CUDA context binds there
struct gsAPALMWorkerB200 {
    kl_ctx* ctx = nullptr;
    std::unique_ptr<gsALMCrisfieldB200> alm;
    ~gsAPALMWorkerB200() { alm.reset(); if (ctx) kl_destroy(ctx); }
};

class gsAPALMB200 {
public:
    typedef std::function<std::unique_ptr<gsAPALMWorkerB200>(int device)> Factory;

    gsAPALMB200(Factory factory, double dL, int subIntervals, double tolerance, int maxLevel)
        : m_factory(factory), m_dL(dL), m_subIntervals(subIntervals), m_data(tolerance, maxLevel) {}

    /// gsAPALM::serialSolve: Nsteps arc-length steps from the origin, each from the previous one (the level-0 chain that cannot
    /// be distributed, gsAPALM.hpp:587-598); returns the wall time
    double serialSolve(gsAPALMWorkerB200& w, int Nsteps) {
        const auto t0 = std::chrono::steady_clock::now();
        gsALMCrisfieldB200& alm = *w.alm;
        std::vector<double> Uold(alm.solutionU().size(), 0.0);
        double Lold = 0.0, time = 0.0;
        m_solutions.clear(); m_times.clear();
        m_solutions.push_back(std::make_pair(Uold, Lold));
        m_times.push_back(0.0);
        alm.setSolution(Uold, Lold);
        alm.setPrevious(Uold, Lold);
        for (int k = 0; k < Nsteps; ++k) {
            double dL = m_dL;
            alm.setLength(dL);
            for (int tries = 0;; ++tries) {                 // _initiation: halve the length until the step converges (:968-983)
                const int status = alm.step();
                if (status == 0) break;
                if (tries > 12 || status > 3) throw std::runtime_error("gsAPALMB200::serialSolve: arc-length step failed");
                dL *= 0.5;
                alm.setLength(dL);
                alm.setSolution(Uold, Lold);
            }
            std::vector<double> DU(Uold.size());
            for (size_t i = 0; i < DU.size(); ++i) DU[i] = alm.solutionU()[i] - Uold[i];
            time += alm.distance(DU, alm.solutionL() - Lold);
            Uold = alm.solutionU(); Lold = alm.solutionL();
            m_solutions.push_back(std::make_pair(Uold, Lold));
            m_times.push_back(time);
        }
        m_data.setData(m_times, m_solutions);
        return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    }

    /// gsAPALM::_correction (:1014-1165) on one worker
    bool correction(gsAPALMWorkerB200& w, const gsAPALMDataB200::Job& job, std::vector<double>& distances,
                    std::vector<gsAPALMSolutionB200>& stepSolutions, double& upperDistance, double& lowerDistance) {
        gsALMCrisfieldB200& alm = *w.alm;
        int Nintervals = m_subIntervals;
        std::vector<double> Uold = job.start.first, Uori = job.start.first;
        double Lold = job.start.second, Lori = Lold;
        const double dL0 = job.dL0 / Nintervals;
        double dL = dL0, dL_rem = 0;
        bool bisected = false;
        stepSolutions.assign(Nintervals, gsAPALMSolutionB200());
        distances.assign(Nintervals + 1, 0.0);
        alm.setLength(dL);
        alm.setSolution(Uold, Lold);
        alm.setPrevious(job.prev.first, job.prev.second);
        int failures = 0;
        for (int k = 0; k != Nintervals; ++k) {
            const int status = alm.step();
            if (status == 1 || status == 2) {
                if (++failures > 12) return false;
                dL *= 0.5;
                dL_rem += dL;                       // the remainder of the interval
                alm.setLength(dL);
                alm.setSolution(Uold, Lold);
                bisected = true;
                --k;
                continue;
            }
            if (status != 0) return false;
            stepSolutions.at(k) = std::make_pair(alm.solutionU(), alm.solutionL());
            std::vector<double> DU(Uold.size());
            for (size_t i = 0; i < DU.size(); ++i) DU[i] = alm.solutionU()[i] - Uold[i];
            distances.at(k) = alm.distance(DU, alm.solutionL() - Lold);
            Uold = alm.solutionU(); Lold = alm.solutionL();
            if (!bisected) dL = dL0;
            else {
                dL = dL_rem;
                ++Nintervals;
                stepSolutions.resize(Nintervals);
                distances.resize(Nintervals + 1);
            }
            alm.setLength(dL);
            dL_rem = 0;
            bisected = false;
        }
        std::vector<double> D(Uold.size());
        for (size_t i = 0; i < D.size(); ++i) D[i] = job.reference.first[i] - alm.solutionU()[i];
        distances.back() = alm.distance(D, job.reference.second - alm.solutionL());
        for (size_t i = 0; i < D.size(); ++i) D[i] = job.reference.first[i] - Uori[i];
        upperDistance = alm.distance(D, job.reference.second - Lori);
        for (size_t i = 0; i < D.size(); ++i) D[i] = stepSolutions.back().first[i] - Uori[i];
        lowerDistance = alm.distance(D, stepSolutions.back().second - Lori);
        return true;
    }

    struct ParallelStats {
        double wall_s = 0, sum_job_s = 0;
        std::vector<int> jobs_per_worker;
        std::vector<double> busy_s_per_worker;
    };

    /// gsAPALM::parallelSolve: the queue is drained by `nWorkers` threads; worker 0 may reuse an existing replica
    ParallelStats parallelSolve(int nWorkers, gsAPALMWorkerB200* worker0 = nullptr) {
        ParallelStats st;
        st.jobs_per_worker.assign(nWorkers, 0);
        st.busy_s_per_worker.assign(nWorkers, 0.0);
        std::vector<std::thread> threads;
        std::vector<std::unique_ptr<gsAPALMWorkerB200>> owned(nWorkers);
        std::vector<gsAPALMWorkerB200*> workers(nWorkers, nullptr);
        // every rank builds its own assembler + solver before the clock starts (as in the reference's main())
        for (int w = 0; w < nWorkers; ++w) {
            if (w == 0 && worker0) { workers[0] = worker0; continue; }
            owned[w] = m_factory(w);
            workers[w] = owned[w].get();
        }
        int ready = 0;
        const auto t0 = std::chrono::steady_clock::now();
        for (int w = 0; w < nWorkers; ++w)
            threads.emplace_back([&, w]() {
                { std::lock_guard<std::mutex> lk(m_mutex); ++ready; }
                m_cv.notify_all();
                for (;;) {
                    gsAPALMDataB200::Job job;
                    {
                        std::unique_lock<std::mutex> lk(m_mutex);
                        m_cv.wait(lk, [&]() { return !m_data.empty() || (m_data.nActive() == 0 && ready == nWorkers); });
                        if (m_data.empty()) { m_cv.notify_all(); return; }
                        job = m_data.pop();
                    }
                    const auto j0 = std::chrono::steady_clock::now();
                    std::vector<double> distances;
                    std::vector<gsAPALMSolutionB200> sols;
                    double upp = 0, low = 0;
                    const bool ok = correction(*workers[w], job, distances, sols, upp, low);
                    const double js = std::chrono::duration<double>(std::chrono::steady_clock::now() - j0).count();
                    {
                        std::lock_guard<std::mutex> lk(m_mutex);
                        if (ok) m_data.submit(job.ID, distances, sols, upp, low);
                        else m_data.abandon(job.ID);
                        ++st.jobs_per_worker[w];
                        st.busy_s_per_worker[w] += js;
                        st.sum_job_s += js;
                    }
                    m_cv.notify_all();
                }
            });
        for (auto& t : threads) t.join();
        st.wall_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        return st;
    }

    gsAPALMDataB200& data() { return m_data; }
    const std::vector<gsAPALMSolutionB200>& serialSolutions() const { return m_solutions; }
    const std::vector<double>& serialTimes() const { return m_times; }

private:
    Factory m_factory;
    double m_dL;
    int m_subIntervals;
    gsAPALMDataB200 m_data;
    std::vector<gsAPALMSolutionB200> m_solutions;
    std::vector<double> m_times;
    std::mutex m_mutex;
    std::condition_variable m_cv;
};

}  // namespace gismo
