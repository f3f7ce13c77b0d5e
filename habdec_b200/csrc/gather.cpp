// Multi-GPU result gather: channels are block-partitioned over ranks and never exchange signal data; what
// travels -- once per batch of calls, to rank 0 -- is one fixed-size hbd_result_record per channel: the characters and
// CRC-valid sentences decoded since the previous gather plus the AFC scalars.  This file holds the transport-independent
// half (record packing, the rank-0 sink, the sharding-invariant hash); gather_nccl.cu moves the records over NCCL.
//
// The reference has no counterpart (one Decoder, one process); the records carry exactly what its callbacks and getters
// deliver: character_callback_ / sentence_callback_ payloads (Decoder.h:135-138,604-606,625-626) and
// getFrequencyCorrection / getShift / getNoiseFloor / getPeaks (Decoder.h:108-113).
#include "../../include/habdec_b200.h"

#include <algorithm>
#include <cstring>
#include <string>
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
inline uint64_t fnv1a(uint64_t h, const void* p, size_t n)
{
    const unsigned char* b = static_cast<const unsigned char*>(p);
    for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 0x100000001b3ull; }
    return h;
}
constexpr uint64_t kFnvInit = 0xcbf29ce484222325ull;
struct SinkChan {
    std::string chars, sentences;            // since the previous poll
    uint64_t h_chars = kFnvInit, h_sent = kFnvInit, n_chars = 0, n_sent = 0;
    float stats[4] = {0, 0, 0, 0}; int32_t peaks[2] = {0, 0};   // scalars of the newest record
    bool seen = false;
};
} // namespace

struct hbd_result_sink {
    std::vector<SinkChan> ch;
    unsigned long long records = 0;
};

extern "C" {

hbd_result_sink* hbd_sink_create(int total_channels)
{
    if (total_channels < 1) return nullptr;
    hbd_result_sink* s = new (std::nothrow) hbd_result_sink;
    if (s) s->ch.resize(size_t(total_channels));
    return s;
}
void hbd_sink_destroy(hbd_result_sink* s) { delete s; }

int hbd_sink_feed(hbd_result_sink* s, const hbd_result_record* recs, size_t n)
{
    if (!s || (!recs && n)) return HBD_ERR_ARG;
    int rc = HBD_OK;
    for (size_t i = 0; i < n; ++i) {
        if (i + 8 < n && recs[i + 8].channel < s->ch.size()) {   // the per-channel strings are scattered over the heap
            const SinkChan& nx = s->ch[recs[i + 8].channel];
            __builtin_prefetch(nx.chars.data() + nx.chars.size());
            __builtin_prefetch(nx.sentences.data() + nx.sentences.size());
        }
        const hbd_result_record& r = recs[i];
        if (r.channel >= s->ch.size() || r.n_chars > sizeof(r.chars) || r.sentence_bytes > sizeof(r.sentences)) { rc = HBD_ERR_ARG; continue; }
        SinkChan& c = s->ch[r.channel];
        c.chars.append(r.chars, r.n_chars);
        c.sentences.append(r.sentences, r.sentence_bytes);
        c.h_chars = fnv1a(c.h_chars, r.chars, r.n_chars); c.n_chars += r.n_chars;
        c.h_sent = fnv1a(c.h_sent, r.sentences, r.sentence_bytes); c.n_sent += r.n_sentences;
        c.stats[0] = r.frequency_correction; c.stats[1] = r.shift; c.stats[2] = r.noise_floor; c.stats[3] = r.noise_variance;
        c.peaks[0] = r.peak_left; c.peaks[1] = r.peak_right; c.seen = true;
        ++s->records;
    }
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
    return sink_take(s->ch[size_t(ch)].chars, out, cap);
}
size_t hbd_sink_poll_sentences(hbd_result_sink* s, int ch, char* out, size_t cap)
{
    if (!s || ch < 0 || size_t(ch) >= s->ch.size()) return 0;
    return sink_take(s->ch[size_t(ch)].sentences, out, cap);
}
// the newest record's scalars: out[0..5] = frequency correction, shift, noise floor, noise variance, peak left, peak right
int hbd_sink_stats(hbd_result_sink* s, int ch, double out[6])
{
    if (!s || !out || ch < 0 || size_t(ch) >= s->ch.size()) return HBD_ERR_ARG;
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
    for (size_t c = 0; c < s->ch.size(); ++c) {
        const SinkChan& x = s->ch[c];
        const uint64_t rec[5] = {uint64_t(c), x.h_chars, x.h_sent, x.n_chars, x.n_sent};
        h = fnv1a(h, rec, sizeof(rec));
    }
    return h;
}

} // extern "C"
