// The reference's caller code with habdec_b200's types swapped in -- nothing else changed:
//   DECODER_THREAD                     code/websocketServer/main.cpp:203-283
//       IQSource::get() -> IQVector -> Decoder::pushSamples -> Decoder::operator() -> callbacks / getters
//   the three callback installs        code/websocketServer/main.cpp:573-604
//       sentence_callback_, character_callback_, ssdv_callback_ (messages are printed instead of sent to websocket sessions)
//   SpectrumToStream                   code/websocketServer/habdec_ws_protocol.cpp:355-405 (+ ShrinkVector :338-351)
//       Decoder::getSpectrumInfo() -> zoom / peak shift / ShrinkVector -> SerializeSpectrum (NetTransport.h:61-85)
// It prints what callbacks, getters and the spectrum stream deliver so that tests/test_gpu_cpp_facade.py can hold it
// against the oracle and against the library's own PWR_ frames (which are pinned to the reference's serialiser).
//   decoder_thread file.cf32 fs baud bits stops factor
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <memory>
#include <sstream>
#include <utility>
#include "habdec_b200/Decoder.hpp"
#include "habdec_b200/IQSource.hpp"

typedef habdec_b200::Decoder TDecoder;                     // typedef habdec::Decoder<TReal> TDecoder (GLOBALS.h:36-37)
static std::unique_ptr<TDecoder> g_decoder;                // GLOBALS::get().decoder_
#define DECODER (*g_decoder)

// ---- NetTransport.h:29-47,61-85 for TTransport = float (CompressedVector<float> keeps the values, min/max of the data)
struct SpectrumInfoHeader {
    int32_t header_size_ = (int32_t)sizeof(SpectrumInfoHeader);
    float noise_floor_ = 0, noise_variance_ = 0, sampling_rate_ = 0, shift_ = 0;
    int32_t peak_left_ = 0, peak_right_ = 0, peak_left_valid_ = 0, peak_right_valid_ = 0;
    float min_ = 0, max_ = 0;
    int32_t type_size_ = 0, size_ = 0;
};
template <typename TSpectrumInfo>
void SerializeSpectrum(const TSpectrumInfo& spectrum_info, std::stringstream& ostr, float*)
{
    SpectrumInfoHeader header;
    header.noise_floor_ = spectrum_info.noise_floor_;
    header.noise_variance_ = spectrum_info.noise_variance_;
    header.sampling_rate_ = spectrum_info.sampling_rate_;
    header.shift_ = spectrum_info.shift_;
    header.peak_left_ = spectrum_info.peak_left_;
    header.peak_right_ = spectrum_info.peak_right_;
    header.peak_left_valid_ = spectrum_info.peak_left_valid_;
    header.peak_right_valid_ = spectrum_info.peak_right_valid_;
    header.size_ = spectrum_info.size();
    header.min_ = *std::min_element(spectrum_info.begin(), spectrum_info.end());
    header.max_ = *std::max_element(spectrum_info.begin(), spectrum_info.end());
    header.type_size_ = sizeof(float);
    ostr.write(reinterpret_cast<char*>(&header), sizeof(header));
    ostr.write(reinterpret_cast<const char*>(spectrum_info.data()), spectrum_info.size() * sizeof(float));
}

// ---- habdec_ws_protocol.cpp:338-351
template <typename T>
void ShrinkVector(T& vec, size_t new_size)
{
    if (new_size >= vec.size()) return;
    for (size_t i = 0; i < new_size; ++i) {
        float i_0_1 = float(i) / new_size;
        size_t I = i_0_1 * vec.size();
        vec[i] = vec[I];
    }
    vec.resize(new_size);
}

// ---- habdec_ws_protocol.cpp:355-405
size_t SpectrumToStream(std::stringstream& res_stream, float zoom, int resolution)
{
    using namespace std;
    auto spectrum_info = DECODER.getSpectrumInfo();
    if (!spectrum_info.size()) return 0;
    zoom = min(max(zoom, 0.01f), 0.99f);
    const size_t zoom_slice_begin = zoom / 2 * spectrum_info.size();
    const size_t zoom_slice_end = (1.0f - zoom / 2) * spectrum_info.size();
    spectrum_info.erase(spectrum_info.begin() + zoom_slice_end, spectrum_info.end());
    spectrum_info.erase(spectrum_info.begin(), spectrum_info.begin() + zoom_slice_begin);
    spectrum_info.peak_left_ -= zoom_slice_begin;
    if (spectrum_info.peak_left_ < 0 || spectrum_info.peak_left_ > spectrum_info.size()) {
        spectrum_info.peak_left_ = 0;
        spectrum_info.peak_left_valid_ = false;
    }
    spectrum_info.peak_right_ -= zoom_slice_begin;
    if (spectrum_info.peak_right_ < 0 || spectrum_info.peak_right_ > spectrum_info.size()) {
        spectrum_info.peak_right_ = 0;
        spectrum_info.peak_right_valid_ = false;
    }
    if (resolution < spectrum_info.size()) {
        spectrum_info.peak_left_ = double(spectrum_info.peak_left_) * resolution / spectrum_info.size();
        spectrum_info.peak_right_ = double(spectrum_info.peak_right_) * resolution / spectrum_info.size();
        ShrinkVector(spectrum_info, resolution);
    }
    SerializeSpectrum(spectrum_info, res_stream, (float*)0);      // TransportDataType::kFloat
    return spectrum_info.size();
}

static std::string hex(const std::string& s)
{
    static const char d[] = "0123456789abcdef";
    std::string r;
    for (unsigned char c : s) { r.push_back(d[c >> 4]); r.push_back(d[c & 15]); }
    return r;
}
static uint32_t crc32_zlib(const uint8_t* p, size_t n)   // the checksum the oracle's SSDV transcript carries for an image's packet set
{
    uint32_t c = 0xffffffffu;
    for (size_t i = 0; i < n; ++i) { c ^= p[i]; for (int k = 0; k < 8; ++k) c = (c & 1u) ? (c >> 1) ^ 0xEDB88320u : c >> 1; }
    return ~c;
}

int main(int argc, char** argv)
{
    using namespace std;
    if (argc < 7) return 2;
    const std::string path = argv[1];
    double fs = atof(argv[2]);
    bool no = false;
    habdec_b200::IQSourceFile src;
    src.quiet = true;
    src.setOption("file_string", &path);                 // IQSource_File.h:205-232 option names
    src.setOption("sampling_rate_double", &fs);
    src.setOption("realtime_bool", &no);
    src.setOption("loop_bool", &no);
    if (!src.init() || !src.start()) { std::cerr << "source\n"; return 1; }

    g_decoder.reset(new TDecoder());
    DECODER.livePrint(false);
    DECODER.baud(atof(argv[3])); DECODER.rtty_bits(atoi(argv[4])); DECODER.rtty_stops(float(atof(argv[5])));   // main.cpp:544-553
    DECODER.dc_remove(false);
    DECODER.lowpass_bw(1500); DECODER.lowpass_trans(0.025f);
    DECODER.setupDecimationStagesFactor(atoi(argv[6]));
    DECODER.ssdvBaseFile("ssdv_");

    std::string chars;
    size_t n_sent = 0, n_ssdv = 0;
    // ---- main.cpp:573-604 ------------------------------------------------------------------------------------
    DECODER.sentence_callback_ =
        [&n_sent](string callsign, string data, string crc)
        {
            ++n_sent;
            // a callback may use the decoder (the reference's SentenceCallback reads GLOBALS state under its own locks)
            cout << "SENT " << callsign << "," << data << "*" << crc << " last=" << (DECODER.getLastSentence() == callsign + "," + data + "*" + crc) << "\n";
        };
    DECODER.character_callback_ =
        [&chars](string rtty_characters)
        {
            stringstream data_stream_;
            data_stream_ << "cmd::info:liveprint=" << rtty_characters;
            chars += data_stream_.str().substr(20);
        };
    DECODER.ssdv_callback_ =
        [&n_ssdv](string callsign, int image_id, std::vector<uint8_t> jpeg)
        {
            stringstream data_stream_;
            pair<int, int> ssdv_header{(int)callsign.size(), (int)image_id};
            data_stream_ << "SDV_";
            data_stream_.write(reinterpret_cast<char*>(&ssdv_header), sizeof(ssdv_header));
            data_stream_ << callsign;
            ++n_ssdv;
            cout << "SSDV " << callsign << " " << image_id << " " << jpeg.size() << " " << crc32_zlib(jpeg.data(), jpeg.size()) << " hdr=" << data_stream_.str().size() << "\n";
        };

    // ---- main.cpp:233-245 --------------------------------------------------------------------------------------
    habdec_b200::IQVector samples;                        // TIQVector
    samples.resize(256 * 256);
    samples.samplingRate(src.samplingRate());
    for (;;) {
        const size_t count = src.get(samples.data(), samples.size());   // main.cpp:238
        if (!count) break;
        samples.resize(count);
        DECODER.pushSamples(samples);                     // main.cpp:243
        DECODER();                                        // main.cpp:245
        samples.resize(256 * 256);
        if (count < samples.size()) break;
    }
    cout << "NSENT " << n_sent << "\n";
    cout << "NSSDV " << n_ssdv << "\n";
    cout << "LAST " << DECODER.getLastSentence() << "\n";
    cout << "RATE " << DECODER.getDecimatedSamplingRate() << " " << DECODER.getDecimationFactor() << " " << DECODER.getBinsCount() << "\n";
    int pl = 0, pr = 0; DECODER.getPeaks(pl, pr);
    cout << "PEAKS " << pl << " " << pr << "\n";
    // ---- the "cmd::power:res=R,zoom=Z" requests of a client (habdec_ws_protocol.cpp:92-124 -> SpectrumToStream) --------
    const float zooms[] = {0.0f, 0.5f, 0.9f};
    const int ress[] = {100000, 1024, 300};
    for (int k = 0; k < 3; ++k) {
        stringstream s;
        const size_t n = SpectrumToStream(s, zooms[k], ress[k]);
        // the same frame from the library (PWR_ payload produced on the GPU, type_size 4)
        std::string lib(s.str().size() + 64, '\0');
        const size_t ln = hbd_get_spectrum_frame(DECODER.batch().handle(), 0, zooms[k], ress[k], 4, reinterpret_cast<unsigned char*>(&lib[0]), lib.size());
        lib.resize(std::min(ln, lib.size()));
        cout << "PWR " << k << " " << n << " " << s.str().size() << " " << (lib == s.str() ? "same" : "DIFFERENT") << " " << hex(s.str().substr(0, 52)) << "\n";
    }
    cout << "CHARS " << chars.size() << "\n" << chars << "\nEND\n";
    return 0;
}
