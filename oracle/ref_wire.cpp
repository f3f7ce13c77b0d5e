// TEST INFRASTRUCTURE ONLY -- the reference's websocket wire formats behind the oracle ABI (oracle/oracle_abi.h).
// Header layout + min/max quantisation come from the reference's OWN code (NetTransport.h SerializeSpectrum /
// SerializeDemodulation, CompressedVector.cpp, compiled from /root/reference by oracle/Makefile).  The part of
// SpectrumToStream / DemodToStream that lives in habdec_ws_protocol.cpp (needs boost::beast, not compilable here) is
// restated below, each step citing its line.
#include <algorithm>
#include <cstring>
#include <sstream>
#include <string>
#include <vector>

#include "websocketServer/NetTransport.h"
#include "oracle_abi.h"

namespace {

// habdec_ws_protocol.cpp:338-351
template <typename T>
void ShrinkVector_(T& vec, size_t new_size)
{
    if (new_size >= vec.size()) return;
    for (size_t i = 0; i < new_size; ++i) {
        float i_0_1 = float(i) / new_size;
        size_t I = i_0_1 * vec.size();
        vec[i] = vec[I];
    }
    vec.resize(new_size);
}

size_t emit(std::stringstream& ss, unsigned char* out, size_t cap)
{
    const std::string s = ss.str();
    if (out && cap) memcpy(out, s.data(), std::min(cap, s.size()));
    return s.size();
}

} // namespace

extern "C" {

size_t ref_spectrum_frame(const float* power, size_t n, const hbo_spectrum_meta* meta, float zoom, int resolution, int type_size,
                          unsigned char* out, size_t cap)
{
    using namespace std;
    // Decoder::getSpectrumInfo, Decoder.h:814-836
    habdec::SpectrumInfo<float> spectrum_info;
    spectrum_info = std::vector<float>(power, power + n);
    if (!spectrum_info.size()) return 0;
    spectrum_info.min_ = *std::min_element(spectrum_info.cbegin(), spectrum_info.cend());
    spectrum_info.max_ = *std::max_element(spectrum_info.cbegin(), spectrum_info.cend());
    spectrum_info.peak_left_ = std::abs(meta->peak_left);
    spectrum_info.peak_left_valid_ = meta->peak_left > 0;
    spectrum_info.peak_right_ = std::abs(meta->peak_right);
    spectrum_info.peak_right_valid_ = meta->peak_right > 0;
    spectrum_info.noise_floor_ = meta->noise_floor;
    spectrum_info.noise_variance_ = meta->noise_variance;
    spectrum_info.sampling_rate_ = meta->sampling_rate;
    spectrum_info.shift_ = meta->shift;

    // SpectrumToStream, habdec_ws_protocol.cpp:364-392
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
        ShrinkVector_(spectrum_info, resolution);
    }
    if (!spectrum_info.size()) return 0; // the reference would dereference end() in CompressedVector: no frame
    std::stringstream ss;
    if (type_size == 1) SerializeSpectrum(spectrum_info, ss, (unsigned char*)0);        // :396-401
    else if (type_size == 2) SerializeSpectrum(spectrum_info, ss, (unsigned short int*)0);
    else SerializeSpectrum(spectrum_info, ss, (float*)0);
    return emit(ss, out, cap);
}

size_t ref_demod_frame(const float* demod, size_t n, int resolution, int type_size, unsigned char* out, size_t cap)
{
    // DemodToStream, habdec_ws_protocol.cpp:408-429
    std::vector<float> demod_acc(demod, demod + n);
    if (!demod_acc.size()) return 0;
    ShrinkVector_(demod_acc, resolution);
    if (!demod_acc.size()) return 0;
    std::stringstream ss;
    if (type_size == 1) SerializeDemodulation(demod_acc, ss, (unsigned char*)0);
    else if (type_size == 2) SerializeDemodulation(demod_acc, ss, (unsigned short int*)0);
    else SerializeDemodulation(demod_acc, ss, (float*)0);
    return emit(ss, out, cap);
}

} // extern "C"
