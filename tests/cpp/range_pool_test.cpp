// CPU test of habdec_b200/csrc/range_pool.h (the drain's persistent worker threads): every part of every run executes
// exactly once, on a stable thread per part index, across changing part counts and a stop()/restart.
#include <atomic>
#include <cstdio>
#include <thread>
#include <vector>
#include "../../habdec_b200/csrc/range_pool.h"

int main()
{
    hbd::RangePool pool;
    std::vector<long long> sum(8, 0);
    std::vector<std::thread::id> owner(8);
    long long want[8] = {0};
    std::atomic<int> calls{0};
    for (int round = 0; round < 2; ++round) {
        for (int it = 0; it < 20000; ++it) {
            const int parts = 1 + (it * 7 + round) % 6;
            pool.run(parts, [&](int t) {
                sum[size_t(t)] += it + t;            // part t only ever touches slot t
                if (it > 100 && t > 0 && owner[size_t(t)] != std::this_thread::get_id()) std::abort();   // same worker every time
                owner[size_t(t)] = std::this_thread::get_id();
                calls.fetch_add(1, std::memory_order_relaxed);
            });
            for (int t = 0; t < parts; ++t) want[t] += it + t;
        }
        if (pool.workers() != 5) { std::printf("workers %d\n", pool.workers()); return 1; }
        pool.stop();
        if (pool.workers() != 0) return 1;
        for (auto& o : owner) o = std::thread::id();
    }
    for (int t = 0; t < 8; ++t) if (sum[size_t(t)] != want[t]) { std::printf("slot %d: %lld != %lld\n", t, sum[size_t(t)], want[t]); return 1; }
    std::printf("OK %d\n", calls.load());
    return 0;
}
