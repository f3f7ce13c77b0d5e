// Multi-GPU result gather: channels are block-partitioned over ranks and never exchange signal data; what
// travels -- once per batch of calls, to rank 0 -- is one fixed-size hbd_result_record per channel: the characters and
// CRC-valid sentences decoded since the previous gather plus the AFC scalars.  This file holds the transport-independent
// half (record packing, the rank-0 sink, the sharding-invariant hash); dist.cu moves the records over NCCL.
//
// The reference has no counterpart (one Decoder, one process); the records carry exactly what its callbacks and getters
// deliver: character_callback_ / sentence_callback_ payloads (Decoder.h:135-138,604-606,625-626) and
// getFrequencyCorrection / getShift / getNoiseFloor / getPeaks (Decoder.h:108-113).
#include "../../include/habdec_b200.h"
#include "api_internal.h"

#include <algorithm>
#include <cstddef>
#include <cstring>
#include <mutex>
#include <sched.h>
#include <string>
#include <thread>
#include <vector>

static_assert(sizeof(hbd_result_record) == 256, "hbd_result_record is a wire format: 256 bytes");

extern "C" {

// Fill one record from the heads of a character stream and of a sentence stream ('\n'-terminated sentences; the
// streams are cut wherever the record is full).  *chars_used / *sentence_bytes_used = bytes taken.
void hbd_record_set(hbd_result_record* r, uint32_t channel, const char* chars, size_t n_chars, const char* sentences, size_t sentence_bytes,
                    const double stats[6], size_t* chars_used, size_t* sentence_bytes_used)
{
    memset(r, 0, sizeof(*r));
    r->channel = channel;
    const size_t nc = std::min(n_chars, sizeof(r->chars));
    if (nc) memcpy(r->chars, chars, nc);
    r->n_chars = uint16_t(nc);
    if (nc < n_chars) r->flags |= 1u;
    const size_t nsb = std::min(sentence_bytes, sizeof(r->sentences));
    if (nsb) memcpy(r->sentences, sentences, nsb);
    r->sentence_bytes = uint16_t(nsb);
    r->n_sentences = uint16_t(std::count(r->sentences, r->sentences + nsb, '\n'));
    if (nsb < sentence_bytes) r->flags |= 2u;
    if (stats) {
        r->frequency_correction = float(stats[0]); r->shift = float(stats[1]); r->noise_floor = float(stats[2]); r->noise_variance = float(stats[3]);
        r->peak_left = int32_t(stats[4]); r->peak_right = int32_t(stats[5]);
    }
    if (chars_used) *chars_used = nc;
    if (sentence_bytes_used) *sentence_bytes_used = nsb;
}

} // extern "C"

// ---- rank-0 sink ---------------------------------------------------------------------------------------------------
namespace {
// Streaming CRC-32C of the per-channel character / sentence streams: the state carries over record boundaries, so the
// value does not depend on how a stream was cut into records.  Hardware instruction where the CPU has it (same value).
struct Crc32cTable {
    uint32_t t[256];
    Crc32cTable()
    {
        for (uint32_t i = 0; i < 256; ++i) {
            uint32_t c = i;
            for (int k = 0; k < 8; ++k) c = (c & 1u) ? (c >> 1) ^ 0x82F63B78u : c >> 1;
            t[i] = c;
        }
    }
};
const Crc32cTable kCrcTab;
inline uint32_t crc32c_sw(uint32_t crc, const unsigned char* p, size_t n)
{
    for (size_t i = 0; i < n; ++i) crc = kCrcTab.t[(crc ^ p[i]) & 0xffu] ^ (crc >> 8);
    return crc;
}
#if defined(__x86_64__)
__attribute__((target("sse4.2"))) inline uint32_t crc32c_hw(uint32_t crc, const unsigned char* p, size_t n)
{
    uint64_t c = crc;
    for (; n >= 8; n -= 8, p += 8) { uint64_t v; memcpy(&v, p, 8); c = __builtin_ia32_crc32di(c, v); }
    uint32_t c32 = uint32_t(c);
    for (; n; --n, ++p) c32 = __builtin_ia32_crc32qi(c32, *p);
    return c32;
}
const bool kHaveCrcHw = __builtin_cpu_supports("sse4.2");
#else
inline uint32_t crc32c_hw(uint32_t crc, const unsigned char* p, size_t n) { return crc32c_sw(crc, p, n); }
const bool kHaveCrcHw = false;
#endif
inline uint32_t crc32c(uint32_t crc, const void* p, size_t n)
{
    return kHaveCrcHw ? crc32c_hw(crc, static_cast<const unsigned char*>(p), n) : crc32c_sw(crc, static_cast<const unsigned char*>(p), n);
}
inline uint64_t fnv1a(uint64_t h, const void* p, size_t n)
{
    const unsigned char* b = static_cast<const unsigned char*>(p);
    for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 0x100000001b3ull; }
    return h;
}
constexpr uint64_t kFnvInit = 0xcbf29ce484222325ull;
struct SinkChan {
    std::string chars, sentences;            // since the previous poll
    uint32_t h_chars = ~0u, h_sent = ~0u;    // running CRC-32C of everything ever fed
    uint64_t n_chars = 0, n_sent = 0;
    float stats[4] = {0, 0, 0, 0}; int32_t peaks[2] = {0, 0};   // scalars of the newest record
    bool seen = false;
};
} // namespace

struct hbd_result_sink {
    std::mutex mtx;                          // a server thread may poll while the gather thread feeds
    std::vector<SinkChan> ch;
    // Feeding is lazy: a fed block is only validated and copied in compact form (header + used bytes); the per-channel
    // strings are built when somebody looks (poll / stats / totals / hash).  On rank 0 of an 8-GPU box a gather brings
    // 28 672 records; touching that many scattered per-channel strings costs milliseconds that the caller -- the thread
    // that also feeds the GPU -- does not have, while most channels are never polled between two gathers.
    struct Block { std::vector<unsigned char> own; const hbd_result_record* borrowed = nullptr; size_t n = 0; };
    std::vector<Block> pending;              // own: compact copy of a fed block; borrowed: records that stay valid until hbd::sink_flush
    unsigned long long records = 0;
    int threads = 1;
    void materialize();
};

extern "C" {

hbd_result_sink* hbd_sink_create(int total_channels)
{
    if (total_channels < 1) return nullptr;
    hbd_result_sink* s = new (std::nothrow) hbd_result_sink;
    if (s) {
        s->ch.resize(size_t(total_channels));
        cpu_set_t set;
        int cores = int(std::thread::hardware_concurrency());
        if (sched_getaffinity(0, sizeof(set), &set) == 0) cores = CPU_COUNT(&set);
        s->threads = std::max(1, std::min(4, cores / 2));
    }
    return s;
}
void hbd_sink_destroy(hbd_result_sink* s) { delete s; }
int hbd_sink_set_threads(hbd_result_sink* s, int n) { if (!s) return HBD_ERR_ARG; std::lock_guard<std::mutex> l(s->mtx); s->threads = std::max(1, std::min(n, 64)); return HBD_OK; }

} // extern "C"

namespace {
constexpr size_t kHdr = offsetof(hbd_result_record, chars);   // 40 bytes: everything in front of the text fields

// compact entries [header | chars | sentence bytes] of channels in [c_lo, c_hi) into the per-channel state
void apply_range(hbd_result_sink* s, const unsigned char* blob, size_t bytes, size_t c_lo, size_t c_hi)
{
    for (size_t o = 0; o + kHdr <= bytes;) {
        hbd_result_record r;                                  // header only
        memcpy(&r, blob + o, kHdr);
        const unsigned char* chars = blob + o + kHdr;
        const unsigned char* sent = chars + r.n_chars;
        o += kHdr + r.n_chars + r.sentence_bytes;
        if (r.channel < c_lo || r.channel >= c_hi) continue;
        SinkChan& c = s->ch[r.channel];
        if (r.n_chars) { c.chars.append(reinterpret_cast<const char*>(chars), r.n_chars); c.h_chars = crc32c(c.h_chars, chars, r.n_chars); c.n_chars += r.n_chars; }
        if (r.sentence_bytes) { c.sentences.append(reinterpret_cast<const char*>(sent), r.sentence_bytes); c.h_sent = crc32c(c.h_sent, sent, r.sentence_bytes); c.n_sent += r.n_sentences; }
        c.stats[0] = r.frequency_correction; c.stats[1] = r.shift; c.stats[2] = r.noise_floor; c.stats[3] = r.noise_variance;
        c.peaks[0] = r.peak_left; c.peaks[1] = r.peak_right; c.seen = true;
    }
}
// whole records (a borrowed block, e.g. the pinned receive buffer of the NCCL gather) of channels in [c_lo, c_hi)
void apply_records(hbd_result_sink* s, const hbd_result_record* recs, size_t n, size_t c_lo, size_t c_hi)
{
    for (size_t i = 0; i < n; ++i) {
        const hbd_result_record& r = recs[i];
        if (r.channel < c_lo || r.channel >= c_hi || r.n_chars > sizeof(r.chars) || r.sentence_bytes > sizeof(r.sentences)) continue;
        SinkChan& c = s->ch[r.channel];
        if (r.n_chars) { c.chars.append(r.chars, r.n_chars); c.h_chars = crc32c(c.h_chars, r.chars, r.n_chars); c.n_chars += r.n_chars; }
        if (r.sentence_bytes) { c.sentences.append(r.sentences, r.sentence_bytes); c.h_sent = crc32c(c.h_sent, r.sentences, r.sentence_bytes); c.n_sent += r.n_sentences; }
        c.stats[0] = r.frequency_correction; c.stats[1] = r.shift; c.stats[2] = r.noise_floor; c.stats[3] = r.noise_variance;
        c.peaks[0] = r.peak_left; c.peaks[1] = r.peak_right; c.seen = true;
    }
}
} // namespace

namespace hbd {
// Library-internal (dist.cu): feed without copying.  `recs` must stay valid and unchanged until sink_flush(s) returns.
int sink_feed_borrowed(hbd_result_sink* s, const hbd_result_record* recs, size_t n)
{
    if (!s || (!recs && n)) return HBD_ERR_ARG;
    std::lock_guard<std::mutex> l(s->mtx);
    hbd_result_sink::Block b;
    b.borrowed = recs; b.n = n;
    s->pending.push_back(std::move(b));
    s->records += n;
    return HBD_OK;
}
void sink_flush(hbd_result_sink* s)
{
    if (!s) return;
    std::lock_guard<std::mutex> l(s->mtx);
    s->materialize();
}
} // namespace hbd

void hbd_result_sink::materialize()
{
    if (pending.empty()) return;
    size_t total = 0;
    for (const auto& b : pending) total += b.own.size() + b.n * 64;
    const int n_thr = std::max(1, std::min(threads, int(total / (256 * 1024))));
    auto run = [&](int t) {
        const size_t c_lo = ch.size() * size_t(t) / size_t(n_thr), c_hi = ch.size() * size_t(t + 1) / size_t(n_thr);
        for (const auto& b : pending) {      // blocks in feed order: streams stay in order
            if (b.borrowed) apply_records(this, b.borrowed, b.n, c_lo, c_hi);
            else apply_range(this, b.own.data(), b.own.size(), c_lo, c_hi);
        }
    };
    if (n_thr == 1) run(0);
    else {
        std::vector<std::thread> th;
        for (int t = 1; t < n_thr; ++t) th.emplace_back(run, t);
        run(0);
        for (auto& x : th) x.join();
    }
    pending.clear();
}

extern "C" {

int hbd_sink_feed(hbd_result_sink* s, const hbd_result_record* recs, size_t n)
{
    if (!s || (!recs && n)) return HBD_ERR_ARG;
    if (!n) return HBD_OK;
    int rc = HBD_OK;
    std::vector<unsigned char> blob;
    blob.resize(n * (kHdr + 48));                 // grows below if the records are fuller than that on average
    size_t o = 0, kept = 0;
    for (size_t i = 0; i < n; ++i) {
        const hbd_result_record& r = recs[i];
        if (r.n_chars > sizeof(r.chars) || r.sentence_bytes > sizeof(r.sentences)) { rc = HBD_ERR_ARG; continue; }
        const size_t need = kHdr + r.n_chars + r.sentence_bytes;
        if (o + need > blob.size()) blob.resize(std::max(blob.size() * 2, o + need));
        memcpy(blob.data() + o, &r, kHdr);
        memcpy(blob.data() + o + kHdr, r.chars, r.n_chars);
        memcpy(blob.data() + o + kHdr + r.n_chars, r.sentences, r.sentence_bytes);
        o += need; ++kept;
    }
    blob.resize(o);
    std::lock_guard<std::mutex> l(s->mtx);
    // the channel numbers are checked against THIS sink here, so that a bad block is reported by the feed that brought it
    for (size_t q = 0; q + kHdr <= blob.size();) {
        hbd_result_record r; memcpy(&r, blob.data() + q, kHdr);
        if (r.channel >= s->ch.size()) { rc = HBD_ERR_ARG; uint32_t bad = ~0u; memcpy(blob.data() + q, &bad, 4); --kept; }
        q += kHdr + r.n_chars + r.sentence_bytes;
    }
    s->records += kept;
    hbd_result_sink::Block b;
    b.own = std::move(blob);
    s->pending.push_back(std::move(b));
    if (s->pending.size() > 64) s->materialize();   // bound the backlog of a sink nobody reads
    return rc;
}

static size_t sink_take(std::string& src, char* out, size_t cap)
{
    const size_t n = src.size();
    if (out && cap) memcpy(out, src.data(), std::min(cap, n));
    if (out && cap >= n) src.clear();
    return n;
}
size_t hbd_sink_poll_chars(hbd_result_sink* s, int ch, char* out, size_t cap)
{
    if (!s || ch < 0 || size_t(ch) >= s->ch.size()) return 0;
    std::lock_guard<std::mutex> l(s->mtx);
    s->materialize();
    return sink_take(s->ch[size_t(ch)].chars, out, cap);
}
size_t hbd_sink_poll_sentences(hbd_result_sink* s, int ch, char* out, size_t cap)
{
    if (!s || ch < 0 || size_t(ch) >= s->ch.size()) return 0;
    std::lock_guard<std::mutex> l(s->mtx);
    s->materialize();
    return sink_take(s->ch[size_t(ch)].sentences, out, cap);
}
// the newest record's scalars: out[0..5] = frequency correction, shift, noise floor, noise variance, peak left, peak right
int hbd_sink_stats(hbd_result_sink* s, int ch, double out[6])
{
    if (!s || !out || ch < 0 || size_t(ch) >= s->ch.size()) return HBD_ERR_ARG;
    std::lock_guard<std::mutex> l(s->mtx);
    s->materialize();
    const SinkChan& c = s->ch[size_t(ch)];
    if (!c.seen) return HBD_ERR_STATE;
    for (int i = 0; i < 4; ++i) out[i] = c.stats[i];
    out[4] = c.peaks[0]; out[5] = c.peaks[1];
    return HBD_OK;
}
// totals over all channels (what was ever fed, polled or not)
void hbd_sink_totals(hbd_result_sink* s, unsigned long long* chars, unsigned long long* sentences, unsigned long long* min_sentences, unsigned long long* records)
{
    unsigned long long nc = 0, ns = 0, mn = ~0ull;
    std::unique_lock<std::mutex> l;
    if (s) { l = std::unique_lock<std::mutex>(s->mtx); s->materialize(); }
    if (s) for (const SinkChan& c : s->ch) { nc += c.n_chars; ns += c.n_sent; mn = std::min<unsigned long long>(mn, c.n_sent); }
    if (chars) *chars = nc;
    if (sentences) *sentences = ns;
    if (min_sentences) *min_sentences = s && !s->ch.empty() ? mn : 0;
    if (records) *records = s ? s->records : 0;
}
// One number for "the same characters and sentences arrived for every channel": the per-channel streams are hashed as
// streams, so the value does not depend on how the channels were sharded over ranks or on the gather cadence.
uint64_t hbd_sink_hash(hbd_result_sink* s)
{
    uint64_t h = kFnvInit;
    if (!s) return h;
    std::lock_guard<std::mutex> l(s->mtx);
    s->materialize();
    for (size_t c = 0; c < s->ch.size(); ++c) {
        const SinkChan& x = s->ch[c];
        const uint64_t rec[5] = {uint64_t(c), uint64_t(x.h_chars), uint64_t(x.h_sent), x.n_chars, x.n_sent};
        h = fnv1a(h, rec, sizeof(rec));
    }
    return h;
}

} // extern "C"
