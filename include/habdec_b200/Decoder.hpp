// C++17 facade over the C ABI (include/habdec_b200.h) with the reference's member names, so that code written
// against habdec::Decoder<float> (code/Decoder/Decoder.h:65-141) -- e.g. DECODER_THREAD in
// code/websocketServer/main.cpp:203-283 -- compiles against the B200 implementation by changing one type.
//
//   habdec_b200::Decoder        one channel  == one reference Decoder<float>
//   habdec_b200::BatchDecoder   N channels processed together (the reason this library exists)
//
// Header only; link with -lhabdec_b200.
#pragma once
#include <complex>
#include <functional>
#include <stdexcept>
#include <string>
#include <vector>

#include "../habdec_b200.h"

namespace habdec_b200 {

// same shape as habdec::IQVector<float> (code/Decoder/IQVector.h:33-64): a complex vector that knows its rate
class IQVector : public std::vector<std::complex<float>> {
public:
    double samplingRate() const { return sampling_rate_; }
    void samplingRate(double r) { sampling_rate_ = r; }
private:
    double sampling_rate_ = 0;
};

class BatchDecoder {
public:
    explicit BatchDecoder(int n_channels, int cuda_device = 0)
    {
        if (hbd_create(n_channels, cuda_device, &h_) != HBD_OK)
            throw std::runtime_error("habdec_b200: no usable CUDA device (there is no CPU fallback)");
    }
    ~BatchDecoder() { hbd_destroy(h_); }
    BatchDecoder(const BatchDecoder&) = delete;
    BatchDecoder& operator=(const BatchDecoder&) = delete;

    int channels() const { return hbd_n_channels(h_); }
    hbd_decoder* handle() { return h_; }

    // feed: one channel, all channels from a host matrix, or a device matrix (zero copy)
    bool pushSamples(int ch, const IQVector& v) { return hbd_push_samples(h_, ch, reinterpret_cast<const float*>(v.data()), v.size(), v.samplingRate()) == HBD_OK; }
    bool pushSamples(const std::complex<float>* host_matrix, size_t n, size_t pitch, double fs) { return hbd_push_samples_batch(h_, reinterpret_cast<const float*>(host_matrix), n, pitch, fs) == HBD_OK; }
    bool pushSamplesDevice(const void* device_matrix, size_t n, size_t pitch, double fs) { return hbd_push_samples_device(h_, static_cast<const float*>(device_matrix), n, pitch, fs) == HBD_OK; }

    // one wideband capture for all channels, each through its own NCO (frequency-offset channels)
    bool pushWideband(const IQVector& v) { return hbd_push_wideband(h_, reinterpret_cast<const float*>(v.data()), v.size(), v.samplingRate()) == HBD_OK; }
    bool pushWidebandDevice(const void* device_row, size_t n, double fs) { return hbd_push_wideband_device(h_, static_cast<const float*>(device_row), n, fs) == HBD_OK; }
    // NCO pre-mixer: replaces the SDR retune of code/websocketServer/main.cpp:247-265
    void nco(double freq_hz, int ch = -1) { hbd_set_nco(h_, ch, freq_hz); }
    double nco(int ch) const { return hbd_get_nco(h_, ch); }
    int afcRetune(double min_abs_hz = 100.0, double* applied = nullptr) { return hbd_afc_retune(h_, min_abs_hz, applied); }

    // configure (ch = -1: every channel)
    void baud(double v, int ch = -1) { hbd_set_baud(h_, ch, v); }
    double baud(int ch) const { return hbd_get_baud(h_, ch); }
    void rtty_bits(size_t v, int ch = -1) { hbd_set_rtty_bits(h_, ch, v); }
    size_t rtty_bits(int ch) const { return hbd_get_rtty_bits(h_, ch); }
    void rtty_stops(float v, int ch = -1) { hbd_set_rtty_stops(h_, ch, v); }
    float rtty_stops(int ch) const { return hbd_get_rtty_stops(h_, ch); }
    void lowpass_bw(float v, int ch = -1) { hbd_set_lowpass_bw(h_, ch, v); }
    float lowpass_bw(int ch) const { return hbd_get_lowpass_bw(h_, ch); }
    void lowpass_trans(float v, int ch = -1) { hbd_set_lowpass_trans(h_, ch, v); }
    float lowpass_trans(int ch) const { return hbd_get_lowpass_trans(h_, ch); }
    void dc_remove(bool v, int ch = -1) { hbd_set_dc_remove(h_, ch, v); }
    bool dc_remove(int ch) const { return hbd_get_dc_remove(h_, ch) != 0; }
    size_t setupDecimationStagesFactor(size_t f) { return hbd_setup_decimation_factor(h_, f); }
    size_t setupDecimationStagesBW(double max_rate) { return hbd_setup_decimation_bw(h_, max_rate); }

    // run
    void process() { hbd_process(h_); }
    void operator()() { process(); }
    void processAsync() { hbd_process_async(h_); }
    void collect() { hbd_collect(h_); }

    // results
    std::string getRTTY(int ch) { return str(hbd_get_rtty, ch); }
    std::string getLastSentence(int ch) { return str(hbd_get_last_sentence, ch); }
    std::string pollChars(int ch) { return str(hbd_poll_chars, ch); }
    std::string pollSentences(int ch) { return str(hbd_poll_sentences, ch); }

    // info
    int getDecimationFactor() const { return hbd_get_decimation_factor(h_); }
    double getInputSamplingRate() const { return hbd_get_input_sampling_rate(h_); }
    double getDecimatedSamplingRate() const { return hbd_get_decimated_sampling_rate(h_); }
    double getSymbolRate(int ch) const { return hbd_get_symbol_rate(h_, ch); }

    // GUI data
    size_t getBinsCount() const { return hbd_get_bins_count(h_); }
    IQVector getFFT(int ch)
    {
        IQVector v; v.resize(hbd_get_fft(h_, ch, nullptr, 0) / 2);
        if (!v.empty()) hbd_get_fft(h_, ch, reinterpret_cast<float*>(v.data()), 2 * v.size());
        v.samplingRate(getDecimatedSamplingRate());
        return v;
    }
    std::vector<float> getDemodulated(int ch) { return floats(hbd_get_demodulated, ch); }
    std::vector<float> getPowerSpectrum(int ch) { return floats(hbd_get_power_spectrum, ch); }
    void getPeaks(int ch, int& pl, int& pr) { hbd_get_peaks(h_, ch, &pl, &pr); }
    void getNoiseFloor(int ch, double& nf, double& nv) { hbd_get_noise_floor(h_, ch, &nf, &nv); }
    double getShift(int ch) { return hbd_get_shift(h_, ch); }
    double getFrequencyCorrection(int ch) { return hbd_get_frequency_correction(h_, ch); }
    void resetFrequencyCorrection(int ch, double c) { hbd_reset_frequency_correction(h_, ch, c); }

    // callbacks: (channel, callsign, data, crc) / (channel, chars); fired inside process()/collect()
    std::function<void(int, std::string, std::string, std::string)> sentence_callback_;
    std::function<void(int, std::string)> character_callback_;
    void installCallbacks()
    {
        hbd_set_sentence_callback(h_, sentence_callback_ ? &BatchDecoder::on_sentence : nullptr, this);
        hbd_set_chars_callback(h_, character_callback_ ? &BatchDecoder::on_chars : nullptr, this);
    }

private:
    static void on_sentence(void* u, int ch, const char* cs, const char* d, const char* crc) { static_cast<BatchDecoder*>(u)->sentence_callback_(ch, cs, d, crc); }
    static void on_chars(void* u, int ch, const char* p, size_t n) { static_cast<BatchDecoder*>(u)->character_callback_(ch, std::string(p, n)); }
    template <typename F> std::string str(F fn, int ch)
    {
        std::string s(fn(h_, ch, nullptr, 0), '\0');
        if (!s.empty()) fn(h_, ch, &s[0], s.size());
        return s;
    }
    template <typename F> std::vector<float> floats(F fn, int ch)
    {
        std::vector<float> v(fn(h_, ch, nullptr, 0));
        if (!v.empty()) fn(h_, ch, v.data(), v.size());
        return v;
    }
    hbd_decoder* h_ = nullptr;
};

// One channel with exactly the reference's public surface (Decoder.h:73-141); livePrint/ssdvBaseFile are kept as
// inert properties because the text console and the SSDV JPEG writer are outside the accelerated path.
class Decoder {
public:
    explicit Decoder(int cuda_device = 0) : b_(1, cuda_device) {}
    bool pushSamples(const IQVector& v) { return b_.pushSamples(0, v); }
    void lowpass_bw(float v) { b_.lowpass_bw(v, 0); }
    float lowpass_bw() const { return b_.lowpass_bw(0); }
    void lowpass_trans(float v) { b_.lowpass_trans(v, 0); }
    float lowpass_trans() const { return b_.lowpass_trans(0); }
    void baud(double v) { b_.baud(v, 0); }
    double baud() const { return b_.baud(0); }
    void rtty_bits(size_t v) { b_.rtty_bits(v, 0); }
    size_t rtty_bits() const { return b_.rtty_bits(0); }
    void rtty_stops(float v) { b_.rtty_stops(v, 0); }
    float rtty_stops() const { return b_.rtty_stops(0); }
    void dc_remove(bool v) { b_.dc_remove(v, 0); }
    bool dc_remove() const { return b_.dc_remove(0); }
    size_t setupDecimationStagesFactor(size_t f) { return b_.setupDecimationStagesFactor(f); }
    size_t setupDecimationStagesBW(double r) { return b_.setupDecimationStagesBW(r); }
    std::string getRTTY() { return b_.getRTTY(0); }
    std::string getLastSentence() { return b_.getLastSentence(0); }
    int getDecimationFactor() const { return b_.getDecimationFactor(); }
    double getInputSamplingRate() const { return b_.getInputSamplingRate(); }
    double getDecimatedSamplingRate() const { return b_.getDecimatedSamplingRate(); }
    double getSymbolRate() const { return b_.getSymbolRate(0); }
    size_t getBinsCount() const { return b_.getBinsCount(); }
    IQVector getFFT() { return b_.getFFT(0); }
    std::vector<float> getDemodulated() { return b_.getDemodulated(0); }
    std::vector<float> getPowerSpectrum() { return b_.getPowerSpectrum(0); }
    void getPeaks(int& pl, int& pr) { b_.getPeaks(0, pl, pr); }
    void getNoiseFloor(double& nf, double& nv) { b_.getNoiseFloor(0, nf, nv); }
    double getShift() { return b_.getShift(0); }
    double getFrequencyCorrection() { return b_.getFrequencyCorrection(0); }
    void resetFrequencyCorrection(double c) { b_.resetFrequencyCorrection(0, c); }
    void process()
    {
        if (!wired_) {
            b_.sentence_callback_ = [this](int, std::string a, std::string b, std::string c) { if (sentence_callback_) sentence_callback_(a, b, c); };
            b_.character_callback_ = [this](int, std::string s) { if (character_callback_) character_callback_(s); };
            b_.installCallbacks();
            wired_ = true;
        }
        b_.process();
    }
    void operator()() { process(); }
    bool livePrint() const { return live_print_; }
    void livePrint(bool v) { live_print_ = v; }
    std::string ssdvBaseFile() const { return ssdv_base_; }
    void ssdvBaseFile(const std::string& f) { ssdv_base_ = f; }
    std::function<void(std::string, std::string, std::string)> sentence_callback_;
    std::function<void(std::string)> character_callback_;
private:
    BatchDecoder b_;
    bool wired_ = false, live_print_ = false;
    std::string ssdv_base_;
};

} // namespace habdec_b200
