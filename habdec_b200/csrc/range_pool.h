// A few persistent host threads for the per-channel-range work of the result path (sentence-layer replay in
// hbd_collect*, record packing in hbd_pack_results).  One small drain is ~100 us of work; creating std::threads for it
// costs about as much again, and a worker that keeps "its" channel range from drain to drain finds the channels' text
// state in its own cache.  A worker that is slow to wake up (a busy core) never holds the drain up: the calling thread
// takes every part nobody has started by the time it gets there, so the worst case is the single-threaded drain.
// Host side only (the reference's counterpart is the tail of Decoder::process, Decoder.h:572-632), no device code.
#pragma once
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace hbd {

class RangePool {
public:
    static constexpr int kMaxParts = 64;

    RangePool() { for (auto& c : claimed_) c.store(0, std::memory_order_relaxed); }
    RangePool(const RangePool&) = delete;
    RangePool& operator=(const RangePool&) = delete;
    ~RangePool() { stop(); }

    // fn(t) for every t in [0, parts), each exactly once: part 0 on the calling thread, part t > 0 on worker t unless the
    // caller gets to it first.  Returns when all parts are done.  One run at a time (the callers hold the decoder's mutex).
    void run(int parts, const std::function<void(int)>& fn)
    {
        if (parts <= 1 || parts > kMaxParts) { for (int t = 0; t < std::max(parts, 1); ++t) fn(t); return; }
        unsigned gen;
        {
            std::unique_lock<std::mutex> l(m_);
            while (int(th_.size()) < parts - 1) {
                const int id = int(th_.size()) + 1;
                const unsigned seen = gen_;          // a new worker must not take a job that was posted before it existed
                th_.emplace_back([this, id, seen] { loop(id, seen); });
            }
            fn_ = &fn; parts_ = parts; gen = ++gen_;
            finished_.store(0, std::memory_order_relaxed);
        }
        go_.notify_all();
        for (int t = 0; t < parts; ++t)
            if (claim(t, gen)) { fn(t); finished_.fetch_add(1, std::memory_order_acq_rel); }
        std::unique_lock<std::mutex> l(m_);          // parts a worker is running right now
        done_.wait(l, [&] { return finished_.load(std::memory_order_acquire) == parts; });
        fn_ = nullptr;
    }

    // joins the workers; run() starts new ones when it is called again
    void stop()
    {
        {
            std::unique_lock<std::mutex> l(m_);
            quit_ = true;
        }
        go_.notify_all();
        for (std::thread& t : th_) if (t.joinable()) t.join();
        th_.clear();
        quit_ = false;
    }

    int workers() const { return int(th_.size()); }

private:
    // Part t of job `gen` goes to whoever asks first.  Generations only move forward, so a worker that wakes up for a job
    // which is already over (its part was taken by the caller) cannot claim anything.
    bool claim(int t, unsigned gen)
    {
        unsigned prev = claimed_[size_t(t)].load(std::memory_order_relaxed);
        while (int(gen - prev) > 0)
            if (claimed_[size_t(t)].compare_exchange_weak(prev, gen, std::memory_order_acq_rel)) return true;
        return false;
    }

    void loop(int id, unsigned seen)
    {
        for (;;) {
            const std::function<void(int)>* fn = nullptr;
            int parts = 0;
            {
                std::unique_lock<std::mutex> l(m_);
                go_.wait(l, [&] { return quit_ || gen_ != seen; });
                if (quit_) return;
                seen = gen_;
                if (id < parts_) { fn = fn_; parts = parts_; }
            }
            // a successful claim means run() of this job is still waiting for the part: *fn is alive
            if (!fn || !claim(id, seen)) continue;
            (*fn)(id);
            if (finished_.fetch_add(1, std::memory_order_acq_rel) + 1 == parts) {
                std::unique_lock<std::mutex> l(m_);
                done_.notify_one();
            }
        }
    }

    std::vector<std::thread> th_;
    std::mutex m_;
    std::condition_variable go_, done_;
    const std::function<void(int)>* fn_ = nullptr;
    int parts_ = 0;
    unsigned gen_ = 0;
    bool quit_ = false;
    std::atomic<int> finished_{0};
    std::atomic<unsigned> claimed_[kMaxParts];
};

} // namespace hbd
