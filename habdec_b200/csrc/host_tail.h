// Host-side, non data-parallel pieces of the path (see host_tail.cpp)
#pragma once
#include <functional>
#include <string>
#include <vector>

namespace hbd {

// FirFilter::LP_BlackmanHarris; returns the (possibly unchanged) tap count, fills `taps` only on a redesign
size_t design_lowpass(float rel_width, float trans, size_t input_size, size_t current_taps, std::vector<float>& taps);

std::string crc16_hex(const std::string& s);

struct SentenceMatch {
    bool ok = false;
    std::string callsign, data, crc;
    size_t rest_offset = 0;
};
bool extract_sentence(const std::string& stream, SentenceMatch& m);

using SentenceSink = std::function<void(int ch, const std::string& callsign, const std::string& data, const std::string& crc)>;

// text side of one channel (Decoder.h:188-200: rtty_char_stream_, last_sentence_, chr_callback_stream_)
struct TextChannel {
    std::string text_stream;        // getRTTY()
    std::string last_sentence;      // getLastSentence()
    std::string chars_pending;      // printable chars since the last poll / callback
    std::string sentences_pending;  // CRC-valid sentences since the last poll
    std::vector<unsigned char> raw_pending; // raw chars since the last poll (SSDV consumers)
    void feed(const unsigned char* raw, size_t n, int ch, const SentenceSink& sink);
};

} // namespace hbd
