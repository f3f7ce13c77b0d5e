// Host-side drain cost model: the per-(call, channel) work of hbd_decoder::collect_locked (segment replay ->
// TextChannel::feed) on a synthetic character log of the bench workload.  No GPU.
//   g++ -O2 -std=c++17 -I habdec_b200/csrc tools/micro/drain_bench.cpp habdec_b200/csrc/build/host_tail.o -o /tmp/drain_bench
#include <chrono>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include "host_tail.h"
using namespace hbd;
struct U2 { unsigned x, y; };
int main(int argc, char** argv)
{
    const int n_ch = argc > 1 ? atoi(argv[1]) : 4096, n_calls = argc > 2 ? atoi(argv[2]) : 100;
    std::vector<std::string> msg(n_ch);
    for (int c = 0; c < n_ch; ++c) { char b[64]; snprintf(b, sizeof b, "$$C%04d,%03d,%03d*ABCD\n", c, (c * 7) % 1000, (c * 13) % 1000); msg[c] = b; }
    // 0.87 characters per call and channel: a character in 7 of 8 calls
    std::vector<TextChannel> text(n_ch);
    std::vector<hbd_result_record> pend(n_ch);
    memset(pend.data(), 0, pend.size() * sizeof(pend[0]));
    std::vector<size_t> pos(n_ch, 0);
    double total = 0; size_t chars = 0;
    std::vector<U2> log;
    for (int call = 0; call < n_calls; ++call) {
        log.clear();
        for (int k = 0; k < n_ch; ++k) {
            const int c = (k * 2654435761u) % n_ch;   // CTAs finish in scattered order
            if ((c + call) % 8 == 7) continue;
            log.push_back({unsigned(c), (unsigned(call) << 8) | (unsigned char)msg[c][pos[c]++ % msg[c].size()]});
        }
        auto t0 = std::chrono::steady_clock::now();
        SentenceSink sink;
        unsigned char seg[64];
        for (size_t i = 0; i < log.size(); ++i) {
            if (i + 8 < log.size()) { __builtin_prefetch(&text[log[i + 8].x]); __builtin_prefetch(&pend[log[i + 8].x]); }
            if (i + 4 < log.size()) text[log[i + 4].x].prefetch_tails();
            seg[0] = (unsigned char)(log[i].y & 0xff);
            text[log[i].x].feed(seg, 1, int(log[i].x), sink, false, pend[log[i].x]);
        }
        total += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        chars += log.size();
        if (call % 16 == 15) for (auto& p : pend) { p.n_chars = 0; p.sentence_bytes = 0; }
    }
    printf("channels %d: %.1f us per call, %.1f ns per character (%zu chars)\n", n_ch, total / n_calls * 1e6, total / chars * 1e9, chars);
    return 0;
}
