// workerpool.h — host worker threads of the ingest path (plain C++, no CUDA).
#pragma once
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>
#include <unistd.h>

namespace boss {

// Host worker threads of the ingest path, kept between batches: spawning and joining ~15 threads costs 0.3-0.5 ms per batch,
// as much as the copy + packing work they share. One pool per process; a caller that finds it busy (another handle's ingest
// on another thread) falls back to its own short-lived threads.
class WorkerPool {
public:
    static WorkerPool& instance() { static WorkerPool p; return p; }
    std::mutex owner;                       // held by the ingest that uses the pool
    // workers 0..n-1 run fn(t); returns at once
    void start(int n, const std::function<void(int)>& fn) {
        std::unique_lock<std::mutex> lk(m_);
        if (pid_ != getpid()) {             // forked child: the parent's workers do not exist here (their handles are leaked)
            threads_ = new std::vector<std::thread>();
            pid_ = getpid();
        }
        while ((int)threads_->size() < n) { const int id = (int)threads_->size(); threads_->emplace_back([this, id] { loop(id); }); }
        job_ = &fn; n_active_ = n; remaining_ = n; ++gen_;
        lk.unlock();
        cv_start_.notify_all();
    }
    void wait() {
        std::unique_lock<std::mutex> lk(m_);
        cv_done_.wait(lk, [this] { return remaining_ == 0; });
        job_ = nullptr;
    }
    ~WorkerPool() {
        { std::lock_guard<std::mutex> lk(m_); stop_ = true; }
        cv_start_.notify_all();
        if (pid_ == getpid()) for (auto& t : *threads_) t.join();
    }
private:
    void loop(int id) {
        unsigned long seen = 0;
        for (;;) {
            const std::function<void(int)>* job = nullptr;
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_start_.wait(lk, [&] { return stop_ || (gen_ != seen && id < n_active_); });
                if (stop_) return;
                seen = gen_;
                job = job_;
            }
            (*job)(id);
            {
                std::lock_guard<std::mutex> lk(m_);
                if (--remaining_ == 0) cv_done_.notify_all();
            }
        }
    }
    std::vector<std::thread>* threads_ = new std::vector<std::thread>();
    pid_t pid_ = getpid();
    std::mutex m_;
    std::condition_variable cv_start_, cv_done_;
    const std::function<void(int)>* job_ = nullptr;
    int n_active_ = 0, remaining_ = 0;
    unsigned long gen_ = 0;
    bool stop_ = false;
};


}  // namespace boss
