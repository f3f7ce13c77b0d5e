// TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's websocket wire formats (oracle/oracle_abi.h, orc_*).
//   SpectrumToStream / DemodToStream / ShrinkVector   code/websocketServer/habdec_ws_protocol.cpp:338-429
//   SpectrumInfoHeader / DemodHeader / Serialize*      code/websocketServer/NetTransport.h:29-102
//   CompressedVector (min/max, float -> u8 / u16)      code/websocketServer/CompressedVector.h:48-55, CompressedVector.cpp:72-116
//   Decoder::getSpectrumInfo                           code/Decoder/Decoder.h:814-836
// Pinned against oracle/_ref (ref_spectrum_frame / ref_demod_frame) and tests/golden/wire_frames.npz.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "oracle_abi.h"

namespace {

#pragma pack(push, 1)
struct SpectrumHeader { // NetTransport.h:29-47 (14 x 4 bytes, no padding)
    int32_t header_size; float noise_floor, noise_variance, sampling_rate, shift;
    int32_t peak_left, peak_right, peak_left_valid, peak_right_valid; float min_, max_; int32_t type_size, size;
};
struct DemodHdr { int32_t header_size; float min_, max_; int32_t type_size, size; }; // NetTransport.h:50-57
#pragma pack(pop)

void shrink(std::vector<float>& v, size_t new_size) // habdec_ws_protocol.cpp:338-351
{
    if (new_size >= v.size()) return;
    for (size_t i = 0; i < new_size; ++i) {
        const float i_0_1 = float(i) / float(new_size);
        const size_t I = size_t(i_0_1 * float(v.size()));
        v[i] = v[I];
    }
    v.resize(new_size);
}

// CompressedVector<T>(const std::vector<float>&): min/max in double, then copyValues (CompressedVector.cpp:72-116)
void quantise(const std::vector<float>& v, int type_size, double& mn, double& mx, std::vector<unsigned char>& bytes)
{
    mn = *std::min_element(v.begin(), v.end());
    mx = *std::max_element(v.begin(), v.end());
    bytes.clear();
    for (float x : v) {
        if (type_size == 4) { unsigned char b[4]; memcpy(b, &x, 4); bytes.insert(bytes.end(), b, b + 4); continue; }
        const float r = float(float(double(x) - mn) / (mx - mn));           // rhs_v = float(rhs_v - i_min) / (i_max - i_min)
        const float scaled = r * float(type_size == 1 ? 255 : 65535);        // rhs_v * numeric_limits<T>::max()
        // float -> unsigned conversion of the x86-64 build: cvttss2si, NaN / out of range -> 0x80000000 -> low bits 0
        const int32_t iv = std::isnan(scaled) || scaled >= 2147483648.0f || scaled < -2147483648.0f ? INT32_MIN : int32_t(scaled);
        if (type_size == 1) bytes.push_back((unsigned char)(iv & 0xff));
        else { const uint16_t u = uint16_t(iv & 0xffff); bytes.push_back(u & 0xff); bytes.push_back(u >> 8); }
    }
}

size_t emit(const void* hdr, size_t hs, const std::vector<unsigned char>& body, unsigned char* out, size_t cap)
{
    const size_t total = hs + body.size();
    if (out && cap) {
        std::vector<unsigned char> all(total);
        memcpy(all.data(), hdr, hs);
        if (!body.empty()) memcpy(all.data() + hs, body.data(), body.size());
        memcpy(out, all.data(), std::min(cap, total));
    }
    return total;
}

} // namespace

extern "C" {

size_t orc_spectrum_frame(const float* power, size_t n, const hbo_spectrum_meta* meta, float zoom, int resolution, int type_size,
                          unsigned char* out, size_t cap)
{
    if (!n) return 0;
    std::vector<float> v(power, power + n);
    int pl = std::abs(meta->peak_left), pr = std::abs(meta->peak_right);
    bool plv = meta->peak_left > 0, prv = meta->peak_right > 0;
    zoom = std::min(std::max(zoom, 0.01f), 0.99f);
    const size_t zb = size_t(zoom / 2 * float(v.size()));
    const size_t ze = size_t((1.0f - zoom / 2) * float(v.size()));
    v.erase(v.begin() + ze, v.end());
    v.erase(v.begin(), v.begin() + zb);
    pl -= int(zb);
    if (pl < 0 || size_t(pl) > v.size()) { pl = 0; plv = false; }
    pr -= int(zb);
    if (pr < 0 || size_t(pr) > v.size()) { pr = 0; prv = false; }
    if (size_t(resolution) < v.size()) { // int promoted to size_t like the reference's comparison
        pl = int(double(pl) * resolution / double(v.size()));
        pr = int(double(pr) * resolution / double(v.size()));
        shrink(v, size_t(resolution));
    }
    if (v.empty()) return 0;
    double mn, mx; std::vector<unsigned char> body;
    quantise(v, type_size, mn, mx, body);
    SpectrumHeader h;
    h.header_size = int32_t(sizeof(SpectrumHeader));
    h.noise_floor = float(meta->noise_floor); h.noise_variance = float(meta->noise_variance);
    h.sampling_rate = float(meta->sampling_rate); h.shift = float(meta->shift);
    h.peak_left = pl; h.peak_right = pr; h.peak_left_valid = plv; h.peak_right_valid = prv;
    h.min_ = float(mn); h.max_ = float(mx); h.type_size = type_size; h.size = int32_t(v.size());
    return emit(&h, sizeof(h), body, out, cap);
}

size_t orc_demod_frame(const float* demod, size_t n, int resolution, int type_size, unsigned char* out, size_t cap)
{
    if (!n) return 0;
    std::vector<float> v(demod, demod + n);
    shrink(v, size_t(resolution));
    if (v.empty()) return 0;
    double mn, mx; std::vector<unsigned char> body;
    quantise(v, type_size, mn, mx, body);
    DemodHdr h;
    h.header_size = int32_t(sizeof(DemodHdr)); h.min_ = float(mn); h.max_ = float(mx); h.type_size = type_size; h.size = int32_t(v.size());
    return emit(&h, sizeof(h), body, out, cap);
}

} // extern "C"
