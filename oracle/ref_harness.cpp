// TEST INFRASTRUCTURE ONLY -- wraps the UNMODIFIED reference habdec::Decoder<float>
// (compiled from /root/reference/code by oracle/Makefile into
// oracle/_ref/libhabdec_ref.so) behind the oracle C ABI (oracle/oracle_abi.h).
// No reference source is copied or edited: private members are reached with the
// usual "#define private public" test trick so per-stage arrays can be compared.
//
// Reference usage rules honoured here:
//  * ONE OS THREAD PER DECODER for its whole life -- FSK2_Demod keeps its carry in
//    a `thread_local static` (code/Decoder/FSK2_Demod.h:35) and so does the
//    character-callback timer (code/Decoder/Decoder.h:617).
//  * livePrint(false); stdout of the reference is discarded.
//  * characters = concat(character_callback_) + the still unflushed
//    chr_callback_stream_ (the callback is paced by wall-clock, Decoder.h:617-629).
//  * configuration order mirrors code/websocketServer/main.cpp:544-553.

#include <atomic>
#include <chrono>
#include <complex>
#include <condition_variable>
#include <cstring>
#include <functional>
#include <iostream>
#include <map>
#include <mutex>
#include <regex>
#include <set>
#include <sstream>
#include <string>
#include <thread>
#include <vector>
#include <array>
#include <memory>
#include <numeric>
#include <algorithm>
#include <future>
#include <cmath>

#define private public
#include "Decoder/Decoder.h"
#undef private

#include "oracle_abi.h"
#include <stdlib.h>
#include <unistd.h>

namespace {

struct NullBuf : std::streambuf { int overflow(int c) override { return c; } };
struct Silencer {
    NullBuf nb; std::streambuf* old = nullptr;
    Silencer() { if (!getenv("HBD_REF_VERBOSE")) old = std::cout.rdbuf(&nb); }
    ~Silencer() { if (old) std::cout.rdbuf(old); }
};
Silencer g_silencer;

// a dedicated thread that executes closures in order; run() blocks until done
class Worker {
public:
    Worker() : th_([this] { loop(); }) {}
    ~Worker() {
        { std::lock_guard<std::mutex> l(m_); stop_ = true; }
        cv_.notify_all();
        th_.join();
    }
    void run(std::function<void()> f) {
        std::unique_lock<std::mutex> l(m_);
        job_ = std::move(f); has_job_ = true; done_ = false;
        cv_.notify_all();
        cv_.wait(l, [this] { return done_; });
    }
private:
    void loop() {
        std::unique_lock<std::mutex> l(m_);
        for (;;) {
            cv_.wait(l, [this] { return has_job_ || stop_; });
            if (stop_) return;
            auto f = std::move(job_); has_job_ = false;
            l.unlock(); f(); l.lock();
            done_ = true; cv_.notify_all();
        }
    }
    std::mutex m_; std::condition_variable cv_;
    std::function<void()> job_; bool has_job_ = false, done_ = true, stop_ = false;
    std::thread th_;
};

void configure(habdec::Decoder<float>& D, const hbo_config& c)
{
    D.baud(c.baud);
    D.rtty_bits(size_t(c.rtty_bits));
    D.rtty_stops(c.rtty_stops);
    D.livePrint(false);
    D.dc_remove(c.dc_remove != 0);
    D.lowpass_bw(c.lowpass_bw);
    D.lowpass_trans(c.lowpass_trans);
    D.setupDecimationStagesFactor(size_t(c.dec_factor));
    if (c.fft_bins && size_t(c.fft_bins) != D.fft_bins_cnt_) {
        // 16k-bin extension: FFT and AFC classes are size agnostic, only this const member pins 4096.
        // It is a per-object runtime value (read in process(), Decoder.h:466-506), patched here without touching sources.
        *const_cast<size_t*>(&D.fft_bins_cnt_) = size_t(c.fft_bins);
    }
}

struct Ref {
    hbo_config cfg;
    std::unique_ptr<Worker> w;
    std::unique_ptr<habdec::Decoder<float>> D;
    std::string chars_cb, sentences;
    std::vector<float> decimated, filtered, demod;
    size_t q_in = 0, dec_pending = 0;
    uint32_t n_calls = 0;
    std::vector<hbo_ssdv_event> ssdv_events;
    std::string ssdv_dir;
    std::map<std::pair<std::string, uint16_t>, std::set<const void*>> ssdv_prev;
};

// one transcript record per accepted packet, read from the wrapper's own bookkeeping (ssdv_wrapper.h:62,82)
void record_ssdv_event(Ref* r)
{
    auto& W = r->D->ssdv_;
    const auto key = W.last_img_k_;
    auto it = W.packets_.find(key);
    if (it == W.packets_.end() || it->second.empty()) return;
    hbo_ssdv_event e;
    memset(&e, 0, sizeof(e));
    e.call = r->n_calls - 1;
    // the wrapper does not say which packet it just filed: it is the one object in the set that was not there at the
    // previous event of this key (the harness keeps that snapshot; the new packet_t is allocated while the old ones
    // are still alive, ssdv_wrapper.cpp:89 vs :120, so its address is distinct from every one in the snapshot)
    std::set<const void*>& prev = r->ssdv_prev[key];
    std::set<const void*> now;
    std::vector<uint8_t> cat;
    e.packet_id = 0xFFFF;
    for (auto& p : it->second) {
        cat.insert(cat.end(), p->data_.begin(), p->data_.end());
        now.insert(p.get());
        if (!prev.count(p.get())) { e.packet_id = p->header_.packet_id; e.width = p->header_.width; e.height = p->header_.height; }
    }
    prev.swap(now);
    e.set_size = uint16_t(it->second.size());
    e.set_crc32 = ssdvp_crc32(cat.data(), cat.size());
    e.image_id = uint16_t(key.second);
    strncpy(e.callsign, key.first.c_str(), sizeof(e.callsign) - 1);
    r->ssdv_events.push_back(e);
}

size_t copy_str(const std::string& s, char* out, size_t cap)
{
    if (out && cap) memcpy(out, s.data(), std::min(cap, s.size()));
    return s.size();
}

size_t copy_floats(const float* src, size_t n, float* out, size_t cap)
{
    if (out && cap && n) memcpy(out, src, std::min(cap, n) * sizeof(float));
    return n;
}

} // namespace

extern "C" {

void* ref_create(const hbo_config* cfg)
{
    Ref* r = new Ref;
    r->cfg = *cfg;
    r->w.reset(new Worker);
    r->w->run([r] {
        r->D.reset(new habdec::Decoder<float>);
        configure(*r->D, r->cfg);
        r->D->character_callback_ = [r](std::string s) { r->chars_cb += s; };
        r->D->sentence_callback_ = [r](std::string cs, std::string data, std::string crc) {
            r->sentences += cs + "," + data + "*" + crc + "\n";
        };
        // SSDV_wraper_t::save_jpeg (ssdv_wrapper.cpp:188-215) writes a file per packet: keep them in a scratch directory
        char tmpl[] = "/tmp/hbd_ref_ssdv_XXXXXX";
        if (const char* d = mkdtemp(tmpl)) { r->ssdv_dir = d; r->D->ssdvBaseFile(r->ssdv_dir + "/ssdv_"); }
        r->D->ssdv_callback_ = [r](std::string, int, std::vector<uint8_t>) { record_ssdv_event(r); };
    });
    return r;
}

void ref_destroy(void* h)
{
    Ref* r = static_cast<Ref*>(h);
    if (!r) return;
    r->w->run([r] { r->D.reset(); });
    if (!r->ssdv_dir.empty()) { std::string cmd = "rm -rf '" + r->ssdv_dir + "'"; if (system(cmd.c_str())) {} }
    delete r;
}

void ref_set_param(void* h, int which, double value)
{
    Ref* r = static_cast<Ref*>(h);
    r->w->run([=] {
        auto& D = *r->D;
        switch (which) {
        case 0: D.baud(value); break;
        case 1: D.rtty_bits(size_t(value)); break;
        case 2: D.rtty_stops(float(value)); break;
        case 3: D.dc_remove(value != 0); break;
        case 4: D.lowpass_bw(float(value)); break;
        case 5: D.lowpass_trans(float(value)); break;
        case 6: D.setupDecimationStagesBW(value); break;     // Decoder.h:336-412 (needs a latched sampling rate: after the first push)
        default: break;
        }
    });
}

void ref_push_process(void* h, const float* iq, size_t n, double fs)
{
    Ref* r = static_cast<Ref*>(h);
    r->w->run([=] {
        habdec::IQVector<float> v;
        v.resize(n);
        v.samplingRate(fs);
        if (n) memcpy(v.data(), iq, n * sizeof(std::complex<float>));
        auto& D = *r->D;
        ++r->n_calls;
        D.pushSamples(v);
        D();
        if (!r->cfg.record) return;
        // mirror the buffer bookkeeping of Decoder.h:429-435,461,492-495,522-542
        const size_t dec = size_t(D.getDecimationFactor());
        r->q_in += n;
        if (r->q_in < dec) return;
        const size_t consumed = r->q_in - r->q_in % dec;
        r->q_in -= consumed;
        const size_t ndec = consumed / dec;
        const bool too_fast = D.getDecimatedSamplingRate() > 4 * D.max_decimated_sampling_rate_;
        if (D.iq_samples_temp_.size() == ndec && ndec)
            r->decimated.insert(r->decimated.end(), (const float*)D.iq_samples_temp_.data(),
                                (const float*)D.iq_samples_temp_.data() + 2 * ndec);
        r->dec_pending += ndec;
        if (r->dec_pending < 256) return;
        if (too_fast) { r->dec_pending = 0; return; }
        const size_t nf = r->dec_pending - r->dec_pending % 256;
        r->dec_pending -= nf;
        if (D.iq_samples_filtered_.size() == nf) {
            r->filtered.insert(r->filtered.end(), (const float*)D.iq_samples_filtered_.data(),
                               (const float*)D.iq_samples_filtered_.data() + 2 * nf);
            r->demod.insert(r->demod.end(), D.demodulated_.begin(), D.demodulated_.end());
        }
    });
}

size_t ref_ssdv_events(void* h, hbo_ssdv_event* out, size_t cap)
{
    Ref* r = static_cast<Ref*>(h);
    size_t n = 0;
    r->w->run([&] {
        n = r->ssdv_events.size();
        if (out && cap) memcpy(out, r->ssdv_events.data(), std::min(cap, n) * sizeof(hbo_ssdv_event));
    });
    return n;
}

void ref_ssdv_push(void* h, const uint8_t* chars, size_t n)
{
    Ref* r = static_cast<Ref*>(h);
    r->w->run([&] {
        ++r->n_calls;
        std::vector<char> v(reinterpret_cast<const char*>(chars), reinterpret_cast<const char*>(chars) + n);
        if (r->D->ssdv_.push(v)) record_ssdv_event(r);   // what Decoder.h:573,631-632 do
    });
}

size_t ref_ssdv_image(void* h, const char* callsign, int image_id, uint8_t* out, size_t cap)
{
    Ref* r = static_cast<Ref*>(h);
    size_t n = 0;
    r->w->run([&] {
        auto& W = r->D->ssdv_;
        auto it = W.packets_.find(std::make_pair(std::string(callsign), uint16_t(image_id)));
        if (it == W.packets_.end()) return;
        for (auto& p : it->second) { if (out && n + 256 <= cap) memcpy(out + n, p->data_.data(), 256); n += 256; }
    });
    return n;
}

size_t ref_chars(void* h, char* out, size_t cap)
{
    Ref* r = static_cast<Ref*>(h);
    std::string s;
    r->w->run([&] { s = r->chars_cb + r->D->chr_callback_stream_; });
    return copy_str(s, out, cap);
}

size_t ref_rtty(void* h, char* out, size_t cap)
{
    Ref* r = static_cast<Ref*>(h);
    std::string s;
    r->w->run([&] { s = r->D->getRTTY(); });
    return copy_str(s, out, cap);
}

size_t ref_last_sentence(void* h, char* out, size_t cap)
{
    Ref* r = static_cast<Ref*>(h);
    std::string s;
    r->w->run([&] { s = r->D->getLastSentence(); });
    return copy_str(s, out, cap);
}

size_t ref_sentences(void* h, char* out, size_t cap)
{
    Ref* r = static_cast<Ref*>(h);
    std::string s;
    r->w->run([&] { s = r->sentences; });
    return copy_str(s, out, cap);
}

size_t ref_stage(void* h, int stage, float* out, size_t cap)
{
    Ref* r = static_cast<Ref*>(h);
    size_t n = 0;
    r->w->run([&] {
        auto& D = *r->D;
        switch (stage) {
        case HBO_STAGE_DECIMATED: n = copy_floats(r->decimated.data(), r->decimated.size(), out, cap); break;
        case HBO_STAGE_FILTERED:  n = copy_floats(r->filtered.data(), r->filtered.size(), out, cap); break;
        case HBO_STAGE_DEMOD:     n = copy_floats(r->demod.data(), r->demod.size(), out, cap); break;
        case HBO_STAGE_FFT: {
            auto f = D.getFFT();
            n = copy_floats((const float*)f.data(), 2 * f.size(), out, cap); break;
        }
        case HBO_STAGE_POWER: {
            auto p = D.getPowerSpectrum();
            n = copy_floats(p.data(), p.size(), out, cap); break;
        }
        case HBO_STAGE_LPTAPS:
            n = copy_floats(D.lowpass_fir_.taps_.data(), D.lowpass_fir_.taps_.size(), out, cap); break;
        case HBO_STAGE_PENDING:
            n = copy_floats(D.symbol_extractor_.samples_.data(), D.symbol_extractor_.samples_.size(), out, cap); break;
        default: n = 0;
        }
    });
    return n;
}

void ref_afc(void* h, hbo_afc_info* o)
{
    Ref* r = static_cast<Ref*>(h);
    r->w->run([&] {
        auto& D = *r->D;
        o->frequency_correction = D.getFrequencyCorrection();
        o->shift_hz = D.getShift();
        D.getNoiseFloor(o->noise_floor, o->noise_variance);
        D.getPeaks(o->peak_left, o->peak_right);
    });
}

void ref_reset_frequency_correction(void* h, double corr)
{
    Ref* r = static_cast<Ref*>(h);
    r->w->run([&] { r->D->resetFrequencyCorrection(corr); });
}

double ref_bench(const hbo_config* cfg, int n_threads, const float* iq, size_t n, size_t stride,
                 size_t chunk, double fs, int reps, uint64_t* o_chars)
{
    std::atomic<int> ready{0};
    std::atomic<bool> go{false};
    std::atomic<uint64_t> chars{0};
    std::vector<double> secs(n_threads, 0.0);
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; ++t)
        th.emplace_back([&, t] {
            habdec::Decoder<float> D;
            configure(D, *cfg);
            uint64_t my_chars = 0;
            D.character_callback_ = [&](std::string s) { my_chars += s.size(); };
            habdec::IQVector<float> v;
            v.samplingRate(fs);
            const std::complex<float>* src = reinterpret_cast<const std::complex<float>*>(iq) + size_t(t) * stride;
            ready++;
            while (!go.load()) std::this_thread::yield();
            auto t0 = std::chrono::steady_clock::now();
            for (int rep = 0; rep < reps; ++rep)
                for (size_t o = 0; o < n; o += chunk) {
                    const size_t c = std::min(chunk, n - o);
                    v.resize(c);
                    memcpy(v.data(), src + o, c * sizeof(std::complex<float>)); // what IQSource::get does
                    D.pushSamples(v);
                    D();
                }
            secs[t] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            my_chars += D.chr_callback_stream_.size();
            chars += my_chars;
        });
    while (ready.load() < n_threads) std::this_thread::yield();
    go = true;
    for (auto& x : th) x.join();
    if (o_chars) *o_chars = chars.load();
    return *std::max_element(secs.begin(), secs.end());
}

void ref_run_ring(const hbo_config* cfg, int n_threads, const float* iq, size_t n_channels, size_t stride, size_t ring_n, size_t chunk,
                  size_t first_chunk, size_t n_chunks, double fs, char* chars_out, size_t chars_pitch, uint32_t* chars_len,
                  char* sent_out, size_t sent_pitch, uint32_t* sent_len)
{
    std::atomic<size_t> next{0};
    const size_t slices = ring_n / chunk;
    auto one_channel = [&](size_t c) {
        habdec::Decoder<float> D;
        configure(D, *cfg);
        std::string chars, sentences;
        D.character_callback_ = [&](std::string s) { chars += s; };
        D.sentence_callback_ = [&](std::string cs, std::string data, std::string crc) { sentences += cs + "," + data + "*" + crc + "\n"; };
        habdec::IQVector<float> v;
        v.samplingRate(fs);
        const std::complex<float>* src = reinterpret_cast<const std::complex<float>*>(iq) + c * stride;
        for (size_t k = 0; k < n_chunks; ++k) {
            const size_t o = ((first_chunk + k) % slices) * chunk;
            v.resize(chunk);
            memcpy(v.data(), src + o, chunk * sizeof(std::complex<float>));
            D.pushSamples(v);
            D();
        }
        chars += D.chr_callback_stream_;
        chars_len[c] = uint32_t(chars.size()); sent_len[c] = uint32_t(sentences.size());
        memcpy(chars_out + c * chars_pitch, chars.data(), std::min(chars.size(), chars_pitch));
        memcpy(sent_out + c * sent_pitch, sentences.data(), std::min(sentences.size(), sent_pitch));
    };
    std::vector<std::thread> pool;
    for (int t = 0; t < std::max(1, n_threads); ++t)
        pool.emplace_back([&] {
            // a fresh OS thread per Decoder: FSK2_Demod's carry and the callback timer are thread_local statics
            for (size_t c; (c = next.fetch_add(1)) < n_channels;) std::thread(one_channel, c).join();
        });
    for (auto& x : pool) x.join();
}

} // extern "C"
