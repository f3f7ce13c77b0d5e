// Host-side pieces of the path that are not data parallel: low-pass tap design and the
// character -> sentence layer.  Plain C++17, no CUDA.
//
//   lp design        code/Decoder/FirFilter.h:173-209, code/Decoder/habdec_windows.h:27-53
//   printable filter code/Decoder/Decoder.h:572-580
//   sentence scan    code/Decoder/Decoder.h:591-613,635-636, code/Decoder/sentence_extract.cpp:58-98
//   CRC16            code/Decoder/CRC.cpp:21-47
#include "host_tail.h"

#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstring>

namespace hbd {

// ---- low-pass design ----------------------------------------------------------------------------------
// The reference evaluates sin/cos in double on float arguments, multiplies in float, sums in double
// and normalises float/double; any other mix changes tap bits.
static float bh4_window(size_t x, size_t N)
{
    const float a0 = 0.35874, a1 = 0.48829, a2 = 0.14128, a3 = 0.01168;
    const float pi2 = 2.0 * M_PI, pi4 = 4.0 * M_PI, pi6 = 6.0 * M_PI;
    const float n1 = N - 1;
    const double c1 = ::cos(double(pi2 * x / n1)), c2 = ::cos(double(pi4 * x / n1)), c3 = ::cos(double(pi6 * x / n1));
    const float w = a0 - a1 * c1 + a2 * c2 - a3 * c3;
    return w;
}

static float sinc_without_pi(float x)
{
    return x ? float(::sin(double(x)) / x) : 1.0f;
}

size_t design_lowpass(float rel_width, float trans, size_t input_size, size_t current_taps, std::vector<float>& taps)
{
    if (!input_size) return current_taps;                       // "No Input set"
    const float tbw = trans ? trans : rel_width * rel_width;
    size_t T = size_t(4.0f / tbw);
    if (T > input_size) T = input_size;
    T |= 1;
    if (T <= 4) return current_taps;
    if (T == current_taps) return current_taps;                 // same count: the old design stays
    taps.assign(T, 0.f);
    double sum = 0;
    const int mid = int(T / 2);
    for (int i = 0; i < int(T); ++i) {
        taps[i] = sinc_without_pi(2.0f * rel_width * (i - mid)) * bh4_window(size_t(i), T);
        sum += taps[i];
    }
    for (size_t i = 0; i < T; ++i) taps[i] /= sum;
    return T;
}

// ---- CRC16-CCITT-FALSE, 4 upper-case hex digits --------------------------------------------------------
namespace {
struct CrcTable {
    uint16_t t[256];
    CrcTable()
    {
        for (unsigned b = 0; b < 256; ++b) {
            unsigned crc = b << 8;
            for (int j = 0; j < 8; ++j) crc = (crc & 0x8000) ? ((crc << 1) ^ 0x1021) : (crc << 1);
            t[b] = uint16_t(crc);
        }
    }
};
const CrcTable kCrc;
// character classes of the sentence pattern in the "C" locale (what std::regex's \w and \s and isprint() give the reference)
struct CharClass {
    unsigned char t[256];   // bit 0: \w, bit 1: callsign class [\w,\-,\s], bit 2: printable or '\n' (Decoder.h:575-577)
    CharClass()
    {
        for (int c = 0; c < 256; ++c) {
            const bool w = (c >= '0' && c <= '9') || (c >= 'A' && c <= 'Z') || (c >= 'a' && c <= 'z') || c == '_';
            const bool sp = c == ' ' || (c >= 9 && c <= 13);
            t[c] = (unsigned char)((w ? 1 : 0) | ((w || c == ',' || c == '-' || sp) ? 2 : 0) | (((c >= 0x20 && c < 0x7f) || c == '\n') ? 4 : 0));
        }
    }
};
const CharClass kCls;
} // namespace

// the bitwise loop of CRC.cpp:21-47 (init 0xFFFF, poly 0x1021, MSB first), one table step per byte
static unsigned crc16_raw(const char* p, size_t n, unsigned crc = 0xffff)
{
    for (size_t i = 0; i < n; ++i) crc = ((crc << 8) ^ kCrc.t[((crc >> 8) ^ (unsigned char)p[i]) & 0xff]) & 0xffff;
    return crc;
}
std::string crc16_hex(const std::string& s)
{
    const unsigned crc = crc16_raw(s.data(), s.size());
    static const char hex[] = "0123456789ABCDEF";
    std::string r(4, '0');
    r[0] = hex[(crc >> 12) & 15]; r[1] = hex[(crc >> 8) & 15]; r[2] = hex[(crc >> 4) & 15]; r[3] = hex[crc & 15];
    return r;
}

// ---- sentence extraction -------------------------------------------------------------------------------
// The reference runs std::regex_match (ECMAScript, backtracking, whole string) with
//     .*?(\$+)([\w,\-,\s]+?),(.+?)(\*|\$)(\w\w\w\w).*
// on the stream with '\n' replaced by ' '.  std::regex costs tens of microseconds per call, which
// would make the host the bottleneck at GPU decode rates, so the same match is found directly.
// Equivalence (also fuzz-tested against std::regex in tests/test_host_side.py):
//  * `.*?` lazy  => the earliest start position with a '$' from which the rest can match; a start
//    inside a run of '$' behaves like the start of that run (`\$+` must swallow the rest of the
//    run, because '$' is not in the callsign class), so runs are tried left to right.
//  * callsign `[\w,\-,\s]+?` lazy => the first ',' preceded by >= 1 class characters (',' itself is
//    in the class); a later ',' can only shrink the set of possible data ends, so if the first one
//    fails the run fails.
//  * data `(.+?)` lazy, >= 1 char => the first '*' or '$' at >= comma+2 followed by four \w.
static inline bool is_word(unsigned char c) { return kCls.t[c] & 1; }
static inline bool is_callsign_char(unsigned char c) { return kCls.t[c] & 2; }   // '\n' (== ' ' after the substitution) is in \s either way

bool extract_sentence(const std::string& stream_in, SentenceMatch& m)
{
    m.ok = false;
    const size_t n = stream_in.size();
    if (!memchr(stream_in.data(), '*', n)) return false;  // sentence_extract.cpp:74
    // the reference matches on a copy with '\n' replaced by ' ' (:70-71); both are in \s and in neither of the other classes the
    // pattern uses, so the match positions are the same on the stream itself and only the extracted fields need the replacement
    const std::string& s = stream_in;
    size_t p = 0;
    while (p < n) {
        if (s[p] != '$') { ++p; continue; }
        size_t q = p;
        while (q < n && s[q] == '$') ++q;                       // q: first char after the run
        // callsign: s[q..c), all class chars, c > q, s[c] == ','  -- first such c
        size_t c = q;
        bool found_c = false;
        while (c < n && is_callsign_char((unsigned char)s[c])) {
            if (s[c] == ',' && c > q) { found_c = true; break; }
            ++c;
        }
        if (found_c) {
            for (size_t e = c + 2; e + 4 < n; ++e) { // s[e+1..e+4] must exist
                if ((s[e] == '*' || s[e] == '$') && is_word((unsigned char)s[e + 1]) && is_word((unsigned char)s[e + 2]) &&
                    is_word((unsigned char)s[e + 3]) && is_word((unsigned char)s[e + 4])) {
                    m.ok = true;
                    m.callsign.assign(s, q, c - q);
                    m.data.assign(s, c + 1, e - (c + 1));
                    m.crc.assign(s, e + 1, 4);
                    std::replace(m.callsign.begin(), m.callsign.end(), '\n', ' ');
                    std::replace(m.data.begin(), m.data.end(), '\n', ' ');
                    m.rest_offset = std::min(n, e + 4);           // sentence_extract.cpp:86-87 (keeps the last CRC char)
                    return true;
                }
            }
        }
        p = q; // next run
    }
    return false;
}

// ---- per-channel text state: Decoder.h:572-613,635-636 ------------------------------------------------------
void TextChannel::feed(const unsigned char* raw, size_t n, int ch, const SentenceSink& sink, bool keep_raw, hbd_result_record& pend)
{
    if (!n) return; // the reference returns before touching the streams when no char was decoded (:568-569)
    if (keep_raw) raw_pending.insert(raw_pending.end(), raw, raw + n);
    const size_t old_len = text_stream.size();
    for (size_t i = 0; i < n; ++i) {
        const char c = char(raw[i]);
        if (kCls.t[(unsigned char)c] & 4) text_stream.push_back(c);   // isprint(char) in the "C" locale, or '\n' (Decoder.h:575-577)
    }
    if (text_stream.size() > old_len)   // chr_callback_stream_ (:581) == what waits for hbd_poll_chars / the next gather
        append(pend.chars, pend.n_chars, sizeof(pend.chars), chars_spill, text_stream.data() + old_len, text_stream.size() - old_len);
    if (text_stream.size() > 20) {
        // The scan loop below leaves a stream without a match (scan_clean).  New characters can only complete a match
        // whose CRC group ends among them -- the pattern's trailing `.*` swallows everything behind the CRC, so a match that
        // ends earlier would have been found before -- i.e. some new position i with [*$] at i-4 and \w at i-3..i.
        // One more way: extractSentence refuses any stream without a '*' (sentence_extract.cpp:74), so a match that ends in
        // "$" + CRC stays latent until a '*' shows up anywhere.  `latent` remembers that a scan was cut short by that rule;
        // only then does a new '*' by itself call for a scan.  Without either, the whole scan is skipped.
        bool candidate = !scan_clean;
        for (size_t i = old_len; !candidate && i < text_stream.size(); ++i) {
            const char* p = text_stream.data() + i;
            candidate = (p[0] == '*' && latent) || (i >= 4 && (p[-4] == '*' || p[-4] == '$') && is_word((unsigned char)p[-3]) &&
                                                   is_word((unsigned char)p[-2]) && is_word((unsigned char)p[-1]) && is_word((unsigned char)p[0]));
        }
        if (candidate) {
            latent = !memchr(text_stream.data(), '*', text_stream.size());   // the scan below gives up at once: the groups seen so far stay latent
            SentenceMatch m;
            while (extract_sentence(text_stream, m)) {
                text_stream.erase(0, m.rest_offset);                   // in place: the buffer keeps its capacity
                std::replace(text_stream.begin(), text_stream.end(), '\n', ' ');    // the reference keeps the space-substituted copy (:599)
                last_sentence.clear();
                last_sentence.reserve(m.callsign.size() + m.data.size() + 7);
                last_sentence += m.callsign; last_sentence += ','; last_sentence += m.data;
                unsigned crc = crc16_raw(last_sentence.data(), last_sentence.size());   // CRC of "callsign,data" (Decoder.h:603-604)
                last_sentence += '*'; last_sentence += m.crc;
                static const char hex[] = "0123456789ABCDEF";
                const char want[4] = {hex[(crc >> 12) & 15], hex[(crc >> 8) & 15], hex[(crc >> 4) & 15], hex[crc & 15]};
                if (m.crc.size() == 4 && memcmp(m.crc.data(), want, 4) == 0) {
                    append(pend.sentences, pend.sentence_bytes, sizeof(pend.sentences), sent_spill, last_sentence.data(), last_sentence.size());
                    append(pend.sentences, pend.sentence_bytes, sizeof(pend.sentences), sent_spill, "\n", 1);
                    if (sink) sink(ch, m.callsign, m.data, m.crc);
                }
            }
        }
        scan_clean = true;
    } else if (text_stream.size() != old_len) {
        scan_clean = false; latent = true;   // a short stream is not scanned (Decoder.h:591): it may hold a complete sentence already
    }
    if (text_stream.size() > 1000) text_stream.erase(0, text_stream.rfind('$')); // npos => erase everything
}

// ---- SSDV: header fields the bookkeeping needs (published layout of fsphil/ssdv: [2..5] base-40 callsign, [6] image id,
// [7..8] packet id, [9] width / 16, [10] height / 16) -----------------------------------------------------------
SsdvHeader ssdv_decode_header(const unsigned char* pkt)
{
    SsdvHeader h;
    uint32_t code = (uint32_t(pkt[2]) << 24) | (uint32_t(pkt[3]) << 16) | (uint32_t(pkt[4]) << 8) | uint32_t(pkt[5]);
    if (code <= 0xF423FFFFu) {          // larger codes are not a callsign: the string stays empty
        int k = 0;
        for (; code && k < 7; code /= 40) {
            const unsigned s = code % 40;
            h.callsign[k++] = s == 0 ? '-' : s < 11 ? char('0' + s - 1) : s < 14 ? '-' : char('A' + s - 14);
        }
    }
    h.image_id = pkt[6];
    h.packet_id = uint16_t((pkt[7] << 8) | pkt[8]);
    h.width = uint16_t(pkt[9] << 4);
    h.height = uint16_t(pkt[10] << 4);
    return h;
}

bool SsdvChannel::push(const unsigned char* chars, size_t n, SsdvEvent& ev)
{
    buff.insert(buff.end(), chars, chars + n);                               // ssdv_wrapper.cpp:42-45
    stream_end += uint32_t(n);
    if (buff.size() < 256) return false;                                     // :47-48
    auto scan_sync = [this] { while (++packet_begin < long(buff.size()) && buff[size_t(packet_begin)] != 0x55) {} };
    if (packet_begin == -1) {                                                // :51-61
        scan_sync();
        if (packet_begin == long(buff.size())) { buff.clear(); packet_begin = -1; return false; }
    }
    if (buff.size() - size_t(packet_begin) < 256) return false;              // :63-64
    // ssdv_dec_is_packet(buff + packet_begin), :66 -- looked up by stream position
    const uint32_t pos = stream_end - uint32_t(buff.size() - size_t(packet_begin));
    while (!verdicts.empty() && int32_t(verdicts.front().pos - pos) < 0) verdicts.pop_front();   // positions only move forward
    if (verdicts.empty() || verdicts.front().pos != pos) {                   // not a packet, :67-85
        scan_sync();
        if (packet_begin == long(buff.size())) { buff.clear(); packet_begin = -1; }
        else { buff.erase(buff.begin(), buff.begin() + packet_begin); packet_begin = 0; }
        return false;
    }
    const SsdvVerdict v = verdicts.front();
    verdicts.pop_front();
    buff.erase(buff.begin() + packet_begin, buff.begin() + packet_begin + 256);   // :89-91
    packet_begin = -1;
    Filed f;
    f.data = v.data;
    f.header = ssdv_decode_header(v.data.data());                            // :92

    const ImageKey key(f.header.callsign, f.header.image_id);                // :105-141
    auto it = packets.find(key);
    if (it == packets.end()) {
        packets[key][f.header.packet_id] = f;
    } else {
        auto& set = it->second;
        if (!set.empty()) {
            const Filed& last = set.rbegin()->second;                        // the highest packet id filed so far (:125-126)
            if (set.count(f.header.packet_id) || last.header.height != f.header.height || last.header.width != f.header.width)
                set.clear();                                                 // retransmission or a new image under the same id
        }
        set[f.header.packet_id] = f;
    }
    last_key = key;                                                          // make_jpeg, :171
    ev.header = f.header; ev.errors = v.errors; ev.set_size = int(packets[key].size()); ev.data = f.data;
    return true;
}

size_t SsdvChannel::image(const std::string& callsign, int image_id, unsigned char* out, size_t cap) const
{
    auto it = packets.find(ImageKey(callsign, uint16_t(image_id)));
    if (it == packets.end()) return 0;
    size_t n = 0;
    for (const auto& kv : it->second) { if (out && n + 256 <= cap) std::copy(kv.second.data.begin(), kv.second.data.end(), out + n); n += 256; }
    return n;
}

} // namespace hbd
