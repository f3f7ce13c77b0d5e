// A small server-less client of the C++ facade (include/habdec_b200/Decoder.hpp), written the way the reference's
// websocket server drives its habdec::Decoder -- the SAME member calls in the same order, in this repo's own code:
//   feed loop            IQSource::get() -> IQVector -> pushSamples() -> operator()      (what DECODER_THREAD does,
//                                                                                          code/websocketServer/main.cpp:233-245)
//   three callbacks      sentence_callback_, character_callback_, ssdv_callback_ assigned as std::function members
//                        (main.cpp:573-604); the messages the server would send are only measured and printed here
//   spectrum for a GUI   getSpectrumInfo() -> centre zoom -> peak re-indexing -> nearest-bin decimation -> PWR_ header + floats
//                        (the contract of habdec_ws_protocol.cpp:338-405 with NetTransport.h:29-85; the byte layout is the
//                        wire format, tests/test_gpu_wire.py pins the library's own frames to it)
// tests/test_gpu_cpp_facade.py compares what is printed with the oracle and with the library's hbd_get_spectrum_frame().
//   decoder_thread file.cf32 fs baud bits stops factor
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <memory>
#include <sstream>
#include <string>
#include <vector>
#include "habdec_b200/Decoder.hpp"
#include "habdec_b200/IQSource.hpp"

using Decoder = habdec_b200::Decoder;
static std::unique_ptr<Decoder> decoder;      // the server keeps one decoder in its globals; so does this client

// ---- the PWR_ payload: 13 little-endian 4-byte fields, then the bins as float ------------------------------------
#pragma pack(push, 1)
struct PwrHeader {
    int32_t bytes;
    float noise_floor, noise_variance, sampling_rate, shift;
    int32_t peak_left, peak_right, peak_left_valid, peak_right_valid;
    float lowest, highest;
    int32_t bytes_per_bin, bins;
};
#pragma pack(pop)
static_assert(sizeof(PwrHeader) == 52, "PWR_ header is 52 bytes");

// One "cmd::power:res=R,zoom=Z" request: what a client asks for and what goes back to it.  Returns the bins sent.
// Every step goes through the container interface of SpectrumInfo (a vector of bins with the AFC scalars as members),
// i.e. the operations the reference's server performs on its own SpectrumInfo.
static size_t power_frame(std::string& frame, float zoom, int resolution)
{
    auto info = decoder->getSpectrumInfo();
    frame.clear();
    if (info.size() == 0) return 0;

    // centre zoom: drop zoom/2 of the bins at either end (the zoom factor is kept inside [0.01, 0.99])
    const float z = std::max(0.01f, std::min(zoom, 0.99f));
    const size_t all = info.size();
    const size_t first = size_t(z / 2 * all), last = size_t((1.0f - z / 2) * all);
    info.erase(info.begin() + last, info.end());
    info.erase(info.begin(), info.begin() + first);

    // the two FSK peaks follow the slice; one that falls outside is reported as absent
    auto reindex = [&](int& peak, bool& valid) {
        peak -= int(first);
        if (peak < 0 || size_t(peak) > info.size()) { peak = 0; valid = false; }
    };
    reindex(info.peak_left_, info.peak_left_valid_);
    reindex(info.peak_right_, info.peak_right_valid_);

    // fewer bins than the slice holds: bin i of the reply is bin floor(i/resolution * size) of the slice (float ratio)
    if (resolution >= 0 && size_t(resolution) < info.size()) {
        const size_t before = info.size();
        info.peak_left_ = int(double(info.peak_left_) * resolution / before);
        info.peak_right_ = int(double(info.peak_right_) * resolution / before);
        for (size_t i = 0; i < size_t(resolution); ++i) {
            const float ratio = float(i) / resolution;
            info[i] = info[size_t(ratio * before)];
        }
        info.resize(size_t(resolution));
    }

    PwrHeader h{};
    h.bytes = int32_t(sizeof(h));
    h.noise_floor = float(info.noise_floor_); h.noise_variance = float(info.noise_variance_);
    h.sampling_rate = float(info.sampling_rate_); h.shift = float(info.shift_);
    h.peak_left = info.peak_left_; h.peak_right = info.peak_right_;
    h.peak_left_valid = info.peak_left_valid_; h.peak_right_valid = info.peak_right_valid_;
    const auto range = std::minmax_element(info.begin(), info.end());
    h.lowest = *range.first; h.highest = *range.second;
    h.bytes_per_bin = 4; h.bins = int32_t(info.size());
    frame.assign(reinterpret_cast<const char*>(&h), sizeof(h));
    frame.append(reinterpret_cast<const char*>(info.data()), info.size() * sizeof(float));
    return info.size();
}

static std::string to_hex(const std::string& bytes)
{
    std::string out;
    char two[3];
    for (unsigned char b : bytes) { std::snprintf(two, sizeof(two), "%02x", b); out += two; }
    return out;
}

// zlib's CRC-32: the checksum the oracle's SSDV transcript carries for an image's packet set
static uint32_t crc32_of(const std::vector<uint8_t>& bytes)
{
    uint32_t crc = ~0u;
    for (uint8_t b : bytes) {
        crc ^= b;
        for (int bit = 0; bit < 8; ++bit) crc = (crc >> 1) ^ (0xEDB88320u & (0u - (crc & 1u)));
    }
    return ~crc;
}

int main(int argc, char** argv)
{
    if (argc < 7) { std::fprintf(stderr, "usage: decoder_thread file.cf32 fs baud bits stops factor\n"); return 2; }
    const std::string file = argv[1];
    double rate = std::atof(argv[2]);
    bool off = false;

    habdec_b200::IQSourceFile source;                 // option names: IQSource_File.h:205-232
    source.quiet = true;
    source.setOption("file_string", &file);
    source.setOption("sampling_rate_double", &rate);
    source.setOption("realtime_bool", &off);
    source.setOption("loop_bool", &off);
    if (!source.init() || !source.start()) { std::fprintf(stderr, "cannot open %s\n", file.c_str()); return 1; }

    decoder.reset(new Decoder());
    Decoder& dec = *decoder;
    dec.livePrint(false);
    dec.baud(std::atof(argv[3]));                     // the setters a server applies from its command line (main.cpp:544-553)
    dec.rtty_bits(std::atoi(argv[4]));
    dec.rtty_stops(float(std::atof(argv[5])));
    dec.dc_remove(false);
    dec.lowpass_bw(1500);
    dec.lowpass_trans(0.025f);
    dec.setupDecimationStagesFactor(std::atoi(argv[6]));
    dec.ssdvBaseFile("ssdv_");

    std::string live_print;       // what the "cmd::info:liveprint=" messages would carry, concatenated
    size_t sentences = 0, ssdv_packets = 0;
    const std::string live_tag = "cmd::info:liveprint=";

    dec.sentence_callback_ = [&](std::string callsign, std::string data, std::string crc) {
        ++sentences;
        const std::string whole = callsign + "," + data + "*" + crc;
        // a callback may call back into the decoder (the server's SentenceCallback does, under its own locks)
        std::cout << "SENT " << whole << " last=" << (dec.getLastSentence() == whole) << "\n";
    };
    dec.character_callback_ = [&](std::string rtty_characters) {
        std::ostringstream message;
        message << live_tag << rtty_characters;
        live_print += message.str().substr(live_tag.size());
    };
    dec.ssdv_callback_ = [&](std::string callsign, int image_id, std::vector<uint8_t> jpeg) {
        // binary "SDV_" message: tag, two ints (callsign length, image id), callsign; the JPEG would follow base64 coded
        std::ostringstream message;
        const int32_t lengths[2] = {int32_t(callsign.size()), int32_t(image_id)};
        message << "SDV_";
        message.write(reinterpret_cast<const char*>(lengths), sizeof(lengths));
        message << callsign;
        ++ssdv_packets;
        std::cout << "SSDV " << callsign << " " << image_id << " " << jpeg.size() << " " << crc32_of(jpeg) << " hdr=" << message.str().size() << "\n";
    };

    // the feed loop: blocks of 65 536 samples from the source, one process() per block
    const size_t block = 256 * 256;
    habdec_b200::IQVector samples;
    samples.samplingRate(source.samplingRate());
    for (bool more = true; more;) {
        samples.resize(block);
        const size_t got = source.get(samples.data(), samples.size());
        if (got == 0) break;
        samples.resize(got);
        dec.pushSamples(samples);
        dec();
        more = got == block;
    }

    std::cout << "NSENT " << sentences << "\n" << "NSSDV " << ssdv_packets << "\n" << "LAST " << dec.getLastSentence() << "\n";
    std::cout << "RATE " << dec.getDecimatedSamplingRate() << " " << dec.getDecimationFactor() << " " << dec.getBinsCount() << "\n";
    int left = 0, right = 0;
    dec.getPeaks(left, right);
    std::cout << "PEAKS " << left << " " << right << "\n";

    // three spectrum requests of a GUI client; each frame is held against the library's own (GPU-built) PWR_ payload
    const struct { float zoom; int resolution; } requests[] = {{0.0f, 100000}, {0.5f, 1024}, {0.9f, 300}};
    int k = 0;
    for (const auto& rq : requests) {
        std::string mine;
        const size_t bins = power_frame(mine, rq.zoom, rq.resolution);
        std::string theirs(mine.size() + 64, '\0');
        const size_t n = hbd_get_spectrum_frame(dec.batch().handle(), 0, rq.zoom, rq.resolution, 4, reinterpret_cast<unsigned char*>(&theirs[0]), theirs.size());
        theirs.resize(std::min(n, theirs.size()));
        std::cout << "PWR " << k++ << " " << bins << " " << mine.size() << " " << (theirs == mine ? "same" : "DIFFERENT") << " " << to_hex(mine.substr(0, sizeof(PwrHeader))) << "\n";
    }
    std::cout << "CHARS " << live_print.size() << "\n" << live_print << "\nEND\n";
    return 0;
}
