// A few persistent host threads for the per-channel-range work of the result drain (sentence-layer replay in
// hbd_collect*, record packing in hbd_pack_results).  One small drain is ~100 us of work; creating std::threads for it
// costs about as much again, and a worker that keeps "its" channel range from drain to drain finds the channels' text
// state in its own cache.  Host side of the reference's Decoder::process tail (Decoder.h:572-632), no device code.
#pragma once
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace hbd {

class RangePool {
public:
    RangePool() = default;
    RangePool(const RangePool&) = delete;
    RangePool& operator=(const RangePool&) = delete;
    ~RangePool() { stop(); }

    // fn(t) for every t in [0, parts): part 0 on the calling thread, part t > 0 on worker t.  Returns when all parts are
    // done.  One run at a time (the callers hold the decoder's mutex).
    void run(int parts, const std::function<void(int)>& fn)
    {
        if (parts <= 1) { fn(0); return; }
        {
            std::unique_lock<std::mutex> l(m_);
            while (int(th_.size()) < parts - 1) {
                const int id = int(th_.size()) + 1;
                const unsigned seen = gen_;          // a new worker must not take a job that was posted before it existed
                th_.emplace_back([this, id, seen] { loop(id, seen); });
            }
            fn_ = &fn; parts_ = parts; left_ = parts - 1; ++gen_;
        }
        go_.notify_all();
        fn(0);
        std::unique_lock<std::mutex> l(m_);
        done_.wait(l, [this] { return left_ == 0; });
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
    void loop(int id, unsigned seen)
    {
        for (;;) {
            const std::function<void(int)>* fn = nullptr;
            {
                std::unique_lock<std::mutex> l(m_);
                go_.wait(l, [&] { return quit_ || gen_ != seen; });
                if (quit_) return;
                seen = gen_;
                if (id < parts_) fn = fn_;
            }
            if (!fn) continue;
            (*fn)(id);
            std::unique_lock<std::mutex> l(m_);
            if (--left_ == 0) done_.notify_one();
        }
    }

    std::vector<std::thread> th_;
    std::mutex m_;
    std::condition_variable go_, done_;
    const std::function<void(int)>* fn_ = nullptr;
    int parts_ = 0, left_ = 0;
    unsigned gen_ = 0;
    bool quit_ = false;
};

} // namespace hbd
