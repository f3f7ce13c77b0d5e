// CPU test of habdec_b200/csrc/range_pool.h (the drain's persistent worker threads): every part of every run executes
// exactly once -- on its worker or, when that one is late, on the caller -- across changing part counts, parts of very
// different length and a stop()/restart.
#include <atomic>
#include <chrono>
#include <cstdio>
#include <thread>
#include <vector>
#include "../../habdec_b200/csrc/range_pool.h"

int main()
{
    hbd::RangePool pool;
    std::vector<long long> sum(8, 0);
    std::vector<int> ran(8, 0);
    long long want[8] = {0};
    long long on_caller = 0, on_worker = 0;
    for (int round = 0; round < 2; ++round) {
        for (int it = 0; it < 20000; ++it) {
            const int parts = 1 + (it * 7 + round) % 6;
            const std::thread::id me = std::this_thread::get_id();
            std::atomic<int> stolen{0};
            for (int t = 0; t < parts; ++t) ran[size_t(t)] = 0;
            pool.run(parts, [&](int t) {
                sum[size_t(t)] += it + t;            // part t only ever touches slot t
                ++ran[size_t(t)];
                if (t > 0 && std::this_thread::get_id() == me) stolen.fetch_add(1, std::memory_order_relaxed);
                if (it % 997 == 0 && t == 0) std::this_thread::sleep_for(std::chrono::microseconds(300));   // a long part 0: workers do theirs
            });
            for (int t = 0; t < parts; ++t) {
                if (ran[size_t(t)] != 1) { std::printf("iteration %d: part %d ran %d times\n", it, t, ran[size_t(t)]); return 1; }
                want[t] += it + t;
            }
            on_caller += stolen.load();
            on_worker += parts - 1 - stolen.load();
        }
        if (pool.workers() != 5) { std::printf("workers %d\n", pool.workers()); return 1; }
        pool.stop();
        if (pool.workers() != 0) return 1;
    }
    for (int t = 0; t < 8; ++t) if (sum[size_t(t)] != want[t]) { std::printf("slot %d: %lld != %lld\n", t, sum[size_t(t)], want[t]); return 1; }
    std::printf("OK parts on workers %lld, taken over by the caller %lld\n", on_worker, on_caller);
    return 0;
}
