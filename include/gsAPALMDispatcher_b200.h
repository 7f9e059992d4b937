/** @file gsAPALMDispatcher_b200.h

    Single-node replacement of gsAPALM's MPI master/worker plumbing (src/gsALMSolvers/gsAPALM.hpp:509-820, message
    shapes :1173-1672): on one 8xB200 box the "ranks" are host threads, one per GPU, each owning its own assembler /
    arc-length solver replica exactly as every MPI rank does in the reference
    (benchmarks/benchmark_Frustrum_APALM.cpp:391-458), and the hierarchical interval queue of gsAPALMData
    (src/gsALMSolvers/gsAPALMData.hpp:215-252 pop, :291-433 submit) lives in shared host memory.  Only solution vectors
    (ndof doubles) ever move between threads — no matrix crosses a worker, as in the reference.

    The queue semantics mirrored from gsAPALMData:
      - an interval (xilow, xiupp, level) is popped FIFO and becomes an active job with an ID           (:215-252)
      - a finished job submits its interior points with the lower / upper error of the interval;
        the sub-intervals [xi_{k-1}, xi_k] are queued on level+1 unless the respective error relative to
        the interval length is below the tolerance or the maximum level is reached                     (:391-409)
      - the traversal ends when the queue is empty and no job is active                                (gsAPALM.hpp:617-745)
*/
#pragma once

#include <atomic>
#include <condition_variable>
#include <deque>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <thread>
#include <tuple>
#include <vector>

namespace gismo {

struct gsAPALMIntervalB200 {
    double xilow, xiupp;
    int level;
};

/// what a worker reports for one interval (gsAPALM::_correction, gsAPALM.hpp:1014-1165)
struct gsAPALMJobResultB200 {
    std::vector<double> xi;            ///< interior points of the interval, ascending (may be empty)
    double lowerError = 0, upperError = 0;   ///< absolute errors of the first / last sub-interval
    bool ok = true;                    ///< false = bisection exhausted (gsStatus != Success)
};

/// Thread-safe hierarchical interval queue (gsAPALMData without the solution storage, which stays with the caller)
class gsAPALMQueueB200 {
public:
    gsAPALMQueueB200(double tolerance, int maxLevel) : m_tol(tolerance), m_maxLevel(maxLevel) {}

    void addInterval(double xilow, double xiupp, int level = 1) { m_queue.push_back({xilow, xiupp, level}); }

    /// gsAPALMData::pop
    std::pair<int, gsAPALMIntervalB200> pop() {
        gsAPALMIntervalB200 iv = m_queue.front();
        m_queue.pop_front();
        m_jobs[m_ID] = iv;
        return {m_ID++, iv};
    }
    /// gsAPALMData::submit + finishJob: returns the number of refined intervals queued
    int submit(int ID, const gsAPALMJobResultB200& r) {
        const gsAPALMIntervalB200 iv = m_jobs.at(ID);
        m_jobs.erase(ID);
        m_done.push_back(iv);
        if (!r.ok) { ++m_failed; return 0; }
        std::vector<double> xi;
        xi.push_back(iv.xilow);
        for (double v : r.xi) xi.push_back(v);
        xi.push_back(iv.xiupp);
        const double Dt = iv.xiupp - iv.xilow;
        size_t kmin = 1, kmax = xi.size();
        if (r.lowerError / Dt < m_tol) kmin = xi.size() - 1;
        if (r.upperError / Dt < m_tol) kmax = xi.size() - 1;
        int added = 0;
        for (size_t k = kmin; k < kmax; ++k)
            if (iv.level < m_maxLevel) { m_queue.push_back({xi[k - 1], xi[k], iv.level + 1}); ++added; }
        for (double v : r.xi) m_points.push_back(v);
        return added;
    }
    bool empty() const { return m_queue.empty(); }
    size_t nActive() const { return m_jobs.size(); }
    size_t nWaiting() const { return m_queue.size(); }
    const std::vector<gsAPALMIntervalB200>& finished() const { return m_done; }
    const std::vector<double>& points() const { return m_points; }
    int nFailed() const { return m_failed; }

private:
    double m_tol;
    int m_maxLevel, m_ID = 0, m_failed = 0;
    std::deque<gsAPALMIntervalB200> m_queue;
    std::map<int, gsAPALMIntervalB200> m_jobs;
    std::vector<gsAPALMIntervalB200> m_done;
    std::vector<double> m_points;
};

/** One worker thread per device.  `Worker` is whatever a rank builds in the reference's main(): assembler + operators +
    arc-length solver; it is constructed INSIDE its thread by `factory(device)` (so that the CUDA context binds to that
    thread) and used by `job(worker, interval, jobID)`. */
template <class Worker>
class gsAPALMDispatcherB200 {
public:
    typedef std::function<std::unique_ptr<Worker>(int device)> Factory;
    typedef std::function<gsAPALMJobResultB200(Worker&, const gsAPALMIntervalB200&, int jobID)> Job;

    gsAPALMDispatcherB200(int nWorkers, Factory factory, Job job) : m_n(nWorkers), m_factory(factory), m_job(job) {}

    /// runs until the queue is drained; returns the number of jobs executed per worker
    std::vector<int> solve(gsAPALMQueueB200& queue) {
        std::vector<int> count(m_n, 0);
        std::vector<std::thread> threads;
        std::atomic<int> ready(0);
        for (int w = 0; w < m_n; ++w)
            threads.emplace_back([&, w]() {
                std::unique_ptr<Worker> worker = m_factory(w);
                ++ready;
                for (;;) {
                    int id;
                    gsAPALMIntervalB200 iv;
                    {
                        std::unique_lock<std::mutex> lk(m_mutex);
                        m_cv.wait(lk, [&]() { return !queue.empty() || (queue.nActive() == 0 && ready.load() == m_n); });
                        if (queue.empty()) {               // nothing waiting and nothing running: done
                            m_cv.notify_all();
                            return;
                        }
                        std::tie(id, iv) = queue.pop();
                    }
                    gsAPALMJobResultB200 res = m_job(*worker, iv, id);
                    {
                        std::lock_guard<std::mutex> lk(m_mutex);
                        queue.submit(id, res);
                        ++count[w];
                    }
                    m_cv.notify_all();
                }
            });
        for (auto& t : threads) t.join();
        return count;
    }

private:
    int m_n;
    Factory m_factory;
    Job m_job;
    std::mutex m_mutex;
    std::condition_variable m_cv;
};

}  // namespace gismo
