// TEST INFRASTRUCTURE ONLY -- CPU restatement ("port") of habdec's IQ -> characters
// path, used solely as the checker in tests/, __graft_entry__.smoke() and the
// cpu_baseline leg of bench.py.  The product (habdec_b200/) never links this.
//
// PARITY STATUS: pinned.  tests/test_oracle_port_vs_ref.py compares every stage of
// this file against the UNMODIFIED reference compiled into oracle/_ref
// (bit-exact floats for decimator / low-pass / slicer, identical characters and
// sentences).  Exception, stated once: the FFT itself is UNPINNED -- FFTW is not
// in the image, both oracles evaluate the DFT in float64 (oracle/shim/inc/fftw3.h).
//
// Every function cites the reference lines it restates (paths under
// /root/reference/code/Decoder).  It is written as a plain streaming state
// machine with one struct per stage; it keeps the reference's call-based quirks
//  because they change characters or AFC numbers.
//
// Build: see oracle/Makefile (-O3, no -ffast-math, no -march => no FMA
// contraction, same float evaluation as the reference build).

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstring>
#include <regex>
#include <string>
#include <thread>
#include <vector>

#include "ssdv_oracle.h"

#include "decim_taps.inc"
#include "fftw3.h" // hbd_dft_f64 only
#include "oracle_abi.h"

namespace {

typedef std::complex<float> cf32;

// ----------------------------------------------------------------------------------------------
// Decimation plan: factor -> list of (M, taps).  Decoder.h:286-320.
// ----------------------------------------------------------------------------------------------
struct TapTable { int M; const uint32_t* bits; int len; };

static bool decimation_plan(int factor, std::vector<TapTable>& plan)
{
#define TT(M, name) TapTable{M, hbd_taps_bits_##name, int(sizeof(hbd_taps_bits_##name) / 4)}
    plan.clear();
    switch (factor) {
    case 256: plan = {TT(64, d256_m64), TT(4, d4_m4)}; return true;
    case 128: plan = {TT(32, d128_m32), TT(4, d4_m4)}; return true;
    case 64:  plan = {TT(32, d64_m32), TT(2, d2_m2)}; return true;
    case 32:  plan = {TT(16, d32_m16), TT(2, d2_m2)}; return true;
    case 16:  plan = {TT(8, d16_m8), TT(2, d2_m2)}; return true;
    case 8:   plan = {TT(8, d8_m8)}; return true;
    case 4:   plan = {TT(4, d4_m4)}; return true;
    case 2:   plan = {TT(2, d2_m2)}; return true;
    case 1:   return true; // no stages (reference: default-constructed Decoder)
    default:  return false;
    }
#undef TT
}

// ----------------------------------------------------------------------------------------------
// One FIR decimator stage.  Decimator.h:69-80 (buffer growth re-zeroes the history),
// :122-138 (y[k] = sum_t buf[kM+t] h[t], accumulated left to right from 0),
// :141-143 (history = last T-1 inputs, read from the caller's buffer AFTER the outputs
// were written over its head, since Decoder.h:443-444 runs the stage in place).
// ----------------------------------------------------------------------------------------------
struct DecimStage {
    int M = 1;
    std::vector<float> h;
    std::vector<cf32> hist;      // T-1 samples preceding the next input
    size_t grown_to = 0;         // size of the reference's work buffer so far

    void setup(const TapTable& t)
    {
        M = t.M;
        h.resize(t.len);
        memcpy(h.data(), t.bits, 4 * size_t(t.len));
        hist.assign(h.size() - 1, cf32(0, 0));
        grown_to = 0;
    }

    // in place on `io`: consumes n samples, leaves n/M outputs at the head; returns n/M
    size_t run(cf32* io, size_t n, std::vector<cf32>& scratch)
    {
        const size_t T = h.size();
        const size_t need = n + T + size_t(M);
        if (grown_to < need) { // Decimator.h:74-79
            grown_to = need;
            std::fill(hist.begin(), hist.end(), cf32(0, 0));
        }
        scratch.resize(T - 1 + n);
        std::copy(hist.begin(), hist.end(), scratch.begin());
        std::copy(io, io + n, scratch.begin() + (T - 1));
        const size_t n_out = n / size_t(M);
        for (size_t k = 0; k < n_out; ++k) {
            const cf32* w = scratch.data() + k * size_t(M);
            float re = 0.0f, im = 0.0f;
            for (size_t t = 0; t < T; ++t) { // complex * real, then complex +=
                re += w[t].real() * h[t];
                im += w[t].imag() * h[t];
            }
            io[k] = cf32(re, im);
        }
        if (n >= T - 1) {
            std::copy(io + n - (T - 1), io + n, hist.begin()); // after the in-place overwrite
        } else { // the reference reads out of bounds here (undefined); keep streaming semantics
            std::vector<cf32> joined(hist);
            joined.insert(joined.end(), scratch.begin() + (T - 1), scratch.end());
            std::copy(joined.end() - (T - 1), joined.end(), hist.begin());
        }
        return n_out;
    }
};

// ----------------------------------------------------------------------------------------------
// Low-pass: tap design FirFilter.h:173-209 + habdec_windows.h:27-53, filtering FirFilter.h:117-169.
// ----------------------------------------------------------------------------------------------
static float window_bh4(size_t x, size_t N) // 4-term Blackman-Harris, habdec_windows.h:37-53
{
    const float a0 = 0.35874, a1 = 0.48829, a2 = 0.14128, a3 = 0.01168;
    const float pi2 = 2.0 * M_PI, pi4 = 4.0 * M_PI, pi6 = 6.0 * M_PI;
    const float n1 = N - 1;
    // float arguments, double cosines, double combination, one rounding to float
    const float w = a0 - a1 * ::cos(double(pi2 * x / n1)) + a2 * ::cos(double(pi4 * x / n1)) - a3 * ::cos(double(pi6 * x / n1));
    return w;
}

static float sinc_nopi(float x) // habdec_windows.h:27-34: sin(x)/x, no pi
{
    if (x) return float(::sin(double(x)) / x);
    return 1;
}

struct LowPass {
    std::vector<float> taps;
    std::vector<cf32> buff;   // the work buffer itself, FirFilter.h:139-160: [T-1 history | last inputs | older leftovers]
    size_t input_size = 0;    // last setInput size (0: "No Input set")

    void design(float rel_width, float trans)
    {
        if (!input_size) return; // FirFilter.h:175-179
        const float tbw = trans ? trans : rel_width * rel_width;
        size_t T = size_t(4.0f / tbw);
        if (T > input_size) T = input_size;
        T |= 1;
        if (T <= 4) return;
        if (T == taps.size()) return; // FirFilter.h:193-194: same count => keep the old design
        taps.resize(T);
        double sum = 0;
        const int mid = int(T / 2);
        for (int i = 0; i < int(T); ++i) {
            taps[i] = sinc_nopi(2.0f * rel_width * (i - mid)) * window_bh4(size_t(i), T);
            sum += taps[i];
        }
        for (size_t i = 0; i < T; ++i) taps[i] /= sum;
    }

    // returns false if the reference would have bailed out (FirFilter.h:121-131)
    bool run(const cf32* in, size_t n, cf32* out, std::vector<cf32>& scratch)
    {
        const size_t T = taps.size();
        if (!T) return false;
        if (T > n + 1) return false;
        const size_t need = n + T;
        if (buff.size() < need) { // FirFilter.h:141-147: vector grows (old content kept, new tail zero), first T entries zeroed
            buff.resize(need, cf32(0, 0));
            std::fill(buff.begin(), buff.begin() + T, cf32(0, 0));
        }
        // a tap count that changed without growth finds whatever the buffer holds: the head of the old history when the
        // filter got shorter, the old history plus stale inputs of the previous call when it got longer (:149-160)
        std::copy(in, in + n, buff.begin() + (T - 1));
        (void)scratch;
        for (size_t i = 0; i < n; ++i) {
            float re = 0.0f, im = 0.0f;
            for (size_t t = 0; t < T; ++t) {
                re += buff[i + t].real() * taps[t];
                im += buff[i + t].imag() * taps[t];
            }
            out[i] = cf32(re, im);
        }
        std::copy(in + n - (T - 1), in + n, buff.begin());
        return true;
    }
};

// ----------------------------------------------------------------------------------------------
// Average<T>: Average.h:34-70.  The constructor add()s 0, the full-window update assigns a double
// expression back into T (truncation for T=int), add() returns the signed pre-update difference.
// ----------------------------------------------------------------------------------------------
template <typename T>
struct Avg {
    T sum = 0; size_t count = 0; size_t cap;
    explicit Avg(size_t n) : cap(std::max<size_t>(1, n)) { add(0); }
    double get() const { return count ? double(sum) / count : double(sum); }
    double add(const T& v)
    {
        const double diff = get() - v;
        if (count == cap) sum = get() * (cap - 1) + v;
        else { ++count; sum += v; }
        return diff;
    }
    void reset(const T& v) { sum = v; count = 1; }
};

// ----------------------------------------------------------------------------------------------
// AFC: AFC.h:92-194 (state machine), :236-286 (power in dB), :225-232 (std-dev), :290-329 (peaks)
// ----------------------------------------------------------------------------------------------
struct Afc {
    std::vector<cf32> spectrum; double spectrum_fs = 0;
    std::vector<float> power;
    double correction = 0, noise_floor = 0, noise_var = 0, shift_hz = 0;
    Avg<double> nf_avg{100}, nv_avg{100};
    Avg<int> left_avg{4}, right_avg{4};
    int gui_left = 0, gui_right = 0;

    bool power_db()
    {
        const size_t N = spectrum.size();
        if (!N || !spectrum_fs) return false;
        const float* raw = reinterpret_cast<const float*>(spectrum.data());
        for (size_t i = 0; i < 2 * N; ++i) if (raw[i] != raw[i] || std::isinf(raw[i])) return false;
        power.resize(N);
        for (size_t i = 0; i < N; ++i) {
            float p = std::norm(spectrum[i]) / N;
            p = p * p;
            p /= spectrum_fs;
            p = 10.0f * std::log10(p); // `using namespace std` in the reference selects the float overload
            power[i] = p;
        }
        for (size_t i = 0; i < N; ++i) if (power[i] != power[i] || std::isinf(power[i])) return false;
        return true;
    }

    static void two_peaks(const std::vector<float>& v, double rel_sep, int* o1, int* o2)
    {
        int sep = round(rel_sep * v.size());
        sep = std::max(8, sep);
        const auto it = std::max_element(v.begin(), v.end());
        int p1 = int(it - v.begin());
        const int lo = std::max(p1 - 2 * sep, 0), hi = std::min(p1 + 2 * sep, int(v.size()));
        int p2 = 0; float best = v[0];
        for (int i = lo; i < hi; ++i)
            if (v[i] > best && std::abs(i - p1) > sep / 2) { p2 = i; best = v[i]; }
        if (p2 < p1) std::swap(p1, p2);
        *o1 = p1; *o2 = p2;
    }

    void step()
    {
        if (!power_db()) { correction = 0; return; }
        double s = 0; for (float p : power) s += p;
        noise_floor = s / power.size();
        double var = 0; for (float p : power) var += (p - noise_floor) * (p - noise_floor);
        noise_var = sqrt(var / power.size());
        nf_avg.add(noise_floor);
        nv_avg.add(noise_var);

        const float fsk_shift = 500;
        const float rel_sep = fsk_shift / spectrum_fs;
        int p1, p2;
        two_peaks(power, rel_sep, &p1, &p2);
        const float thr = nf_avg.get() + 3 * fabs(nv_avg.get());
        const bool d1 = power[p1] > thr, d2 = power[p2] > thr;
        bool stable_l = false, stable_r = false;
        if (d1 && d2) {
            if (p2 < p1) std::swap(p1, p2);
            if (left_avg.add(p1) <= 2) stable_l = true;
            if (right_avg.add(p2) <= 2) stable_r = true;
        }
        gui_left = 0;
        if (d1) gui_left = stable_l ? int(left_avg.get()) : int(-left_avg.get());
        gui_right = 0;
        if (d2) gui_right = stable_r ? int(right_avg.get()) : int(-right_avg.get());
        if (stable_l && stable_r) {
            const int pl = round(left_avg.get()), pr = round(right_avg.get());
            const int dist = pr - pl;
            const double hz_per_bin = spectrum_fs / spectrum.size();
            shift_hz = hz_per_bin * dist;
            const double mid = pl + dist / 2;
            const double err_bins = mid - double(spectrum.size()) / 2;
            if (4 < std::abs(err_bins)) correction = hz_per_bin * err_bins;
        }
    }

    void reset_correction(double c) // AFC.h:188-194
    {
        const double bins_per_hz = double(spectrum.size()) / spectrum_fs;
        left_avg.reset(std::max(0.0, left_avg.get() - c * bins_per_hz));
        right_avg.reset(std::max(0.0, right_avg.get() - c * bins_per_hz));
        correction = 0;
    }
};

// ----------------------------------------------------------------------------------------------
// Bit slicer: SymbolExtractor.h:108-255
// ----------------------------------------------------------------------------------------------
static inline int sign3(float v) { return (0.0f < v) - (v < 0.0f); }

struct Slicer {
    double fs = 0, baud = 1;
    std::vector<float> pend;

    size_t spb() const { return size_t(round(fs / baud)); }

    void window_means(size_t i, size_t R, float& l, float& r) const // SymbolExtractor.h:51-63
    {
        const int lo = std::max(int(i - R), 0);
        const int hi = int(std::min(i + R, pend.size()));
        float sl = 0, sr = 0;
        for (int k = lo; k < int(i); ++k) sl += pend[k];
        for (int k = int(i); k < hi; ++k) sr += pend[k];
        l = sl / (i - lo);
        r = sr / (hi - i);
    }

    size_t next_flip(size_t start) const // SymbolExtractor.h:162-224; 0 == none
    {
        const size_t S = spb();
        if (pend.size() - start < S) return 0;
        const size_t R = std::max(4, int(S / 4));
        size_t p = start + R;
        float l, r;
        window_means(p, R, l, r);
        while (sign3(l) == sign3(r)) {
            ++p;
            if (p >= pend.size() - S) return 0;
            window_means(p, R, l, r);
        }
        const size_t first = p;
        while (sign3(l) != sign3(r)) {
            ++p;
            if (p >= pend.size() - S) return 0;
            window_means(p, R, l, r);
        }
        size_t best = first; float best_w = -1;
        for (size_t i = first; i < p; ++i) {
            window_means(i, R, l, r);
            // SymbolExtractor.h:212: unqualified abs() on a float resolves to ::abs(int) in the
            // reference build, i.e. the weight is |trunc(r - l)| (verified against oracle/_ref)
            const float w = float(::abs(int(r - l)));
            if (i == first || w > best_w) { best = i; best_w = w; } // first maximum
        }
        return best;
    }

    void push(const float* v, size_t n) // :108-125
    {
        if (!n) return;
        if (pend.size() > 3e4) pend.clear();
        pend.insert(pend.end(), v, v + n);
    }

    void run(std::vector<uint8_t>& bits_out) // :129-158
    {
        if (!fs || !baud) return;
        if (pend.size() < fs / baud * 3) return;
        size_t last = 0, off = 0, flip;
        bool any = false;
        const size_t S = spb();
        while ((flip = next_flip(off)) != 0) {
            any = true;
            float sum = 0;
            for (size_t k = last; k < flip; ++k) sum += pend[k];
            const float avg = sum / (flip - last);
            const bool bit = avg > 0;
            size_t cnt = size_t(round(float(flip - last) / float(S)));
            last = flip; off = flip;
            while (cnt--) bits_out.push_back(bit);
        }
        if (!any) return;
        pend.erase(pend.begin(), pend.begin() + std::min(last, pend.size()));
    }
};

// ----------------------------------------------------------------------------------------------
// UART deframer: RTTY.h:77-137
// ----------------------------------------------------------------------------------------------
struct Uart {
    size_t nbits = 0; float nstops = 0;
    std::vector<uint8_t> bits;

    void run(std::vector<char>& out)
    {
        if (!nbits && !nstops) return;
        if (bits.size() < (1 + nbits + nstops)) return;
        size_t last_stop = 0;
        for (size_t i = 0; i < bits.size();) {
            bool ok = bits[i] == 0;
            ok &= (i + 1 + nbits + nstops) <= bits.size();
            for (size_t s = 0; s < nstops; ++s) {
                const size_t idx = i + 1 + nbits + s;
                ok &= (idx < bits.size()) && bits[idx] == 1; // reference reads past the end here; result is masked by `ok`
            }
            if (!ok) { ++i; continue; }
            ++i;
            char c = 0;
            for (size_t k = 0; k < nbits; ++k) { c += bits[i] << k; ++i; }
            out.push_back(c);
            i += nstops; // size_t += float
            last_stop = i - 1;
        }
        if (last_stop) bits.erase(bits.begin(), bits.begin() + last_stop + 1);
    }
};

// ----------------------------------------------------------------------------------------------
// Sentence layer: sentence_extract.cpp:30,58-98 and CRC.cpp:21-47
// ----------------------------------------------------------------------------------------------
static std::string crc16_hex(const std::string& s)
{
    unsigned crc = 0xffff;
    for (unsigned char ch : s) {
        crc ^= (unsigned(int(char(ch))) << 8);
        for (int j = 0; j < 8; ++j) crc = (crc & 0x8000) ? ((crc << 1) ^ 0x1021) : (crc << 1);
    }
    static const char* hex = "0123456789ABCDEF";
    std::string r;
    r += hex[(crc >> 12) & 15]; r += hex[(crc >> 8) & 15]; r += hex[(crc >> 4) & 15]; r += hex[crc & 15];
    return r;
}

static const std::regex& sentence_regex()
{
    static const std::regex re(R"_(.*?(\$+)([\w,\-,\s]+?),(.+?)(\*|\$)(\w\w\w\w).*)_");
    return re;
}

struct Sentence { bool ok = false; std::string callsign, data, crc, rest; };

static Sentence extract_sentence(std::string stream)
{
    Sentence r;
    std::replace(stream.begin(), stream.end(), '\n', ' ');
    if (stream.find("*") < std::string::npos - 4) {
        std::smatch m;
        std::regex_match(stream, m, sentence_regex());
        if (m.size() >= 5) {
            r.callsign = m[2]; r.data = m[3]; r.crc = m[5];
            const size_t cut = std::min(stream.size(), size_t(m.position(4) + 4));
            r.rest = stream.substr(cut);
            r.ok = true;
        }
    }
    return r;
}

// ----------------------------------------------------------------------------------------------
// The decoder: Decoder.h:206-219 (push), :416-638 (process)
// ----------------------------------------------------------------------------------------------
struct Port {
    hbo_config cfg{};
    double fs_in = 0;
    int factor = 1;
    std::vector<DecimStage> stages;
    std::vector<cf32> in_q, work, scratch, dec_q, filtered, fft_in;
    std::vector<float> demod;
    LowPass lp;
    float lp_bw = 1500, lp_trans = 0.025f;
    Afc afc;
    Slicer slicer;
    Uart uart;
    bool demod_primed = false; cf32 demod_last;
    std::string text_stream, last_sentence, chars_all, sentences;
    // recordings
    std::vector<float> rec_dec, rec_filt, rec_demod, rec_bits, rec_raw;

    double fs_dec() const { return fs_in / factor; }

    void configure(const hbo_config& c) // call order of websocketServer/main.cpp:544-553
    {
        cfg = c;
        slicer.baud = c.baud;
        uart.nbits = size_t(c.rtty_bits);
        uart.nstops = c.rtty_stops;
        lp_bw = c.lowpass_bw;   // setter would call design(): no input yet -> no-op (Decoder.h:238-243)
        lp_trans = c.lowpass_trans;
        if (c.dec_factor >= 1 && c.dec_factor <= 256) {
            std::vector<TapTable> plan;
            if (decimation_plan(c.dec_factor, plan)) {
                stages.clear();
                for (auto& t : plan) { stages.emplace_back(); stages.back().setup(t); }
                factor = c.dec_factor;
            } else { stages.clear(); factor = 1; } // Decoder.h:283-284,317-319
        }
    }

    void push(const float* iq, size_t n, double fs)
    {
        const cf32* p = reinterpret_cast<const cf32*>(iq);
        in_q.insert(in_q.end(), p, p + n);
        if (!fs_in) fs_in = float(fs); // init(const float), Decoder.h:215-216,223-225
    }

    void process()
    {
        if (!fs_in) return;
        if (int(in_q.size()) < factor) return;
        const size_t take = in_q.size() - in_q.size() % size_t(factor);
        work.assign(in_q.begin(), in_q.begin() + take);
        in_q.erase(in_q.begin(), in_q.begin() + take);

        size_t n = work.size();
        for (auto& st : stages) n = st.run(work.data(), n, scratch);
        work.resize(n);

        if (cfg.dc_remove) { // Decoder.h:450-459
            cf32 w_prev = .97f * cf32(work[0]);
            for (size_t i = 0; i < work.size(); ++i) {
                const cf32 w = work[i] + .97f * w_prev;
                work[i] = w - w_prev;
                w_prev = w;
            }
        }
        if (cfg.record) rec_dec.insert(rec_dec.end(), (float*)work.data(), (float*)work.data() + 2 * work.size());
        dec_q.insert(dec_q.end(), work.begin(), work.end());

        // FFT frame assembly, Decoder.h:467-489
        const size_t NFFT = cfg.fft_bins ? size_t(cfg.fft_bins) : 4096; // Decoder.h:163 (16384: configs[1] extension)
        if (fft_in.size() < NFFT && work.size()) {
            const size_t k = std::min(NFFT - fft_in.size(), work.size());
            fft_in.insert(fft_in.end(), work.begin(), work.begin() + k);
        }
        if (fft_in.size() >= NFFT) {
            afc.spectrum.resize(NFFT);
            std::vector<cf32> tmp(NFFT);
            hbd_dft_f64((const float*)fft_in.data(), (float*)tmp.data(), int(NFFT));
            for (size_t i = 0; i < NFFT / 2; ++i) { // swap halves, FFT.cpp:77-87
                afc.spectrum[i] = tmp[i + NFFT / 2];
                afc.spectrum[i + NFFT / 2] = tmp[i];
            }
            afc.spectrum_fs = fs_dec();
            have_spectrum = true;
            fft_in.clear();
        }

        if (dec_q.size() < 256) return;
        if (have_spectrum) afc.spectrum_fs = fs_dec();
        afc.step(); // once per call that gets here, Decoder.h:501-509

        if (fs_dec() > 4 * 40e3) { work.clear(); dec_q.clear(); return; } // Decoder.h:522-527

        const size_t nf = dec_q.size() - dec_q.size() % 256;
        filtered.resize(nf);
        lp.input_size = nf;
        lp.design(float(lp_bw / fs_dec()), lp_trans);
        const bool lp_ran = lp.run(dec_q.data(), nf, filtered.data(), scratch);
        (void)lp_ran;
        dec_q.erase(dec_q.begin(), dec_q.begin() + nf);

        demod.resize(nf); // FSK2_Demod.h:30-42 with a per-decoder carry
        if (!demod_primed) { demod_last = filtered[0]; demod_primed = true; }
        for (size_t i = 1; i < nf; ++i) demod[i] = std::arg(filtered[i] * std::conj(filtered[i - 1]));
        demod[0] = std::arg(filtered[0] * std::conj(demod_last));
        demod_last = filtered[nf - 1];
        if (cfg.record) {
            rec_filt.insert(rec_filt.end(), (float*)filtered.data(), (float*)filtered.data() + 2 * nf);
            rec_demod.insert(rec_demod.end(), demod.begin(), demod.end());
        }

        slicer.fs = fs_dec();
        slicer.push(demod.data(), nf);
        std::vector<uint8_t> bits;
        slicer.run(bits);
        std::vector<char> raw;
        if (!bits.empty()) {
            if (cfg.record) for (uint8_t b : bits) rec_bits.push_back(b);
            uart.bits.insert(uart.bits.end(), bits.begin(), bits.end());
            uart.run(raw);
        }
        if (raw.empty()) return;
        ssdv.push(reinterpret_cast<const uint8_t*>(raw.data()), raw.size(), n_calls - 1); // Decoder.h:573
        if (cfg.record) for (char c : raw) rec_raw.push_back(float((unsigned char)c));

        std::string printable; // Decoder.h:575-580
        for (char c : raw) if (isprint(c) || c == '\n') printable += c;
        text_stream += printable;
        chars_all += printable;

        if (text_stream.size() > 20) { // Decoder.h:591-613
            for (;;) {
                Sentence s = extract_sentence(text_stream);
                if (!s.ok) break;
                text_stream = s.rest;
                last_sentence = s.callsign + "," + s.data + "*" + s.crc;
                if (s.crc == crc16_hex(s.callsign + "," + s.data))
                    sentences += s.callsign + "," + s.data + "*" + s.crc + "\n";
            }
        }
        if (text_stream.size() > 1000) text_stream.erase(0, text_stream.rfind('$')); // Decoder.h:635-636
    }

    bool have_spectrum = false;
    hbo_ssdv::Port ssdv;     // Decoder.h:193
    uint32_t n_calls = 0;    // push_process / ssdv_push calls so far
};

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

void* orc_create(const hbo_config* cfg) { Port* p = new Port; p->configure(*cfg); return p; }
void orc_destroy(void* h) { delete static_cast<Port*>(h); }

void orc_set_param(void* h, int which, double value) // Decoder.h:654-706: plain member updates, nothing pending is touched
{
    Port* p = static_cast<Port*>(h);
    switch (which) {
    case 0: p->slicer.baud = value; p->cfg.baud = value; break;
    case 1: p->uart.nbits = size_t(value); p->cfg.rtty_bits = int(value); break;
    case 2: p->uart.nstops = float(value); p->cfg.rtty_stops = float(value); break;
    case 3: p->cfg.dc_remove = value != 0; break;
    case 4: p->lp_bw = float(value); p->lp.design(float(p->lp_bw / p->fs_dec()), p->lp_trans); break;     // Decoder.h:238-243
    case 5: p->lp_trans = float(value); p->lp.design(float(p->lp_bw / p->fs_dec()), p->lp_trans); break;  // Decoder.h:252-257
    case 6: {   // setupDecimationStagesBW, Decoder.h:336-412: plans of <= 256 each are appended until the rate is under the limit
        if (!p->fs_in) break;
        double rate = p->fs_in;
        p->stages.clear(); p->factor = 1;
        while (rate > value) {
            int div;
            for (div = 2; div < 256; div *= 2) if (rate / div <= value) break;
            rate /= div; p->factor *= div;
            std::vector<TapTable> plan;
            decimation_plan(div, plan);
            for (auto& t : plan) { p->stages.emplace_back(); p->stages.back().setup(t); }
        }
        break;
    }
    default: break;
    }
}

void orc_push_process(void* h, const float* iq, size_t n, double fs)
{
    Port* p = static_cast<Port*>(h);
    ++p->n_calls;
    p->push(iq, n, fs);
    p->process();
}

size_t orc_ssdv_events(void* h, hbo_ssdv_event* out, size_t cap)
{
    const auto& ev = static_cast<Port*>(h)->ssdv.events;
    if (out && cap) memcpy(out, ev.data(), std::min(cap, ev.size()) * sizeof(hbo_ssdv_event));
    return ev.size();
}
void orc_ssdv_push(void* h, const uint8_t* chars, size_t n)
{
    Port* p = static_cast<Port*>(h);
    ++p->n_calls;
    p->ssdv.push(chars, n, p->n_calls - 1);
}
size_t orc_ssdv_image(void* h, const char* callsign, int image_id, uint8_t* out, size_t cap)
{
    return static_cast<Port*>(h)->ssdv.image(callsign, image_id, out, cap);
}

int hbo_ssdv_is_packet(uint8_t pkt[256], int* errors) { return ssdv_dec_is_packet(pkt, errors); }

void hbo_ssdv_make_packet(uint8_t out[256], int type, const char* callsign, int image_id, int packet_id, int width16,
                          int height16, int flags, int mcu_offset, int mcu_id, const uint8_t* payload)
{
    // layout of the published format, see oracle/ssdv_published.h
    memset(out, 0, 256);
    out[0] = 0x55; out[1] = uint8_t(0x66 + type);
    uint32_t code = 0;   // base-40, last character most significant
    for (int i = int(strlen(callsign)) - 1; i >= 0; --i) {
        const char c = callsign[i];
        code *= 40;
        if (c >= 'A' && c <= 'Z') code += uint32_t(c - 'A' + 14);
        else if (c >= 'a' && c <= 'z') code += uint32_t(c - 'a' + 14);
        else if (c >= '0' && c <= '9') code += uint32_t(c - '0' + 1);
    }
    out[2] = uint8_t(code >> 24); out[3] = uint8_t(code >> 16); out[4] = uint8_t(code >> 8); out[5] = uint8_t(code);
    out[6] = uint8_t(image_id); out[7] = uint8_t(packet_id >> 8); out[8] = uint8_t(packet_id);
    out[9] = uint8_t(width16); out[10] = uint8_t(height16); out[11] = uint8_t(flags); out[12] = uint8_t(mcu_offset);
    out[13] = uint8_t(mcu_id >> 8); out[14] = uint8_t(mcu_id);
    const int n_payload = type == 1 ? 237 : 205;
    memcpy(out + 15, payload, size_t(n_payload));
    const uint32_t crc = ssdvp_crc32(out + 1, size_t(14 + n_payload));
    uint8_t* q = out + 15 + n_payload;
    q[0] = uint8_t(crc >> 24); q[1] = uint8_t(crc >> 16); q[2] = uint8_t(crc >> 8); q[3] = uint8_t(crc);
    if (type == 0) ssdvp_rs_encode(out + 1, out + 224);
}

size_t orc_chars(void* h, char* out, size_t cap) { return copy_str(static_cast<Port*>(h)->chars_all, out, cap); }
size_t orc_rtty(void* h, char* out, size_t cap) { return copy_str(static_cast<Port*>(h)->text_stream, out, cap); }
size_t orc_last_sentence(void* h, char* out, size_t cap) { return copy_str(static_cast<Port*>(h)->last_sentence, out, cap); }
size_t orc_sentences(void* h, char* out, size_t cap) { return copy_str(static_cast<Port*>(h)->sentences, out, cap); }

size_t orc_stage(void* h, int stage, float* out, size_t cap)
{
    Port* p = static_cast<Port*>(h);
    switch (stage) {
    case HBO_STAGE_DECIMATED: return copy_floats(p->rec_dec.data(), p->rec_dec.size(), out, cap);
    case HBO_STAGE_FILTERED:  return copy_floats(p->rec_filt.data(), p->rec_filt.size(), out, cap);
    case HBO_STAGE_DEMOD:     return copy_floats(p->rec_demod.data(), p->rec_demod.size(), out, cap);
    case HBO_STAGE_FFT:       return copy_floats((const float*)p->afc.spectrum.data(), 2 * p->afc.spectrum.size(), out, cap);
    case HBO_STAGE_POWER:     return copy_floats(p->afc.power.data(), p->afc.power.size(), out, cap);
    case HBO_STAGE_LPTAPS:    return copy_floats(p->lp.taps.data(), p->lp.taps.size(), out, cap);
    case HBO_STAGE_PENDING:   return copy_floats(p->slicer.pend.data(), p->slicer.pend.size(), out, cap);
    case HBO_STAGE_BITS:      return copy_floats(p->rec_bits.data(), p->rec_bits.size(), out, cap);
    case HBO_STAGE_RAWCHARS:  return copy_floats(p->rec_raw.data(), p->rec_raw.size(), out, cap);
    default: return 0;
    }
}

void orc_afc(void* h, hbo_afc_info* o)
{
    Port* p = static_cast<Port*>(h);
    o->frequency_correction = p->afc.correction;
    o->shift_hz = p->afc.shift_hz;
    o->noise_floor = p->afc.noise_floor;
    o->noise_variance = p->afc.noise_var;
    o->peak_left = p->afc.gui_left;
    o->peak_right = p->afc.gui_right;
}

void orc_reset_frequency_correction(void* h, double corr) { static_cast<Port*>(h)->afc.reset_correction(corr); }

// std::regex sentence extraction alone (the reference's extractSentence), for fuzzing the product's matcher
int orc_extract_sentence(const char* stream, size_t n, char* callsign, char* data, char* crc, size_t cap, size_t* rest)
{
    Sentence s = extract_sentence(std::string(stream, n));
    if (!s.ok) return 0;
    auto put = [cap](char* dst, const std::string& v) { if (dst && cap) { const size_t k = std::min(cap - 1, v.size()); memcpy(dst, v.data(), k); dst[k] = 0; } };
    put(callsign, s.callsign); put(data, s.data); put(crc, s.crc);
    if (rest) *rest = n - s.rest.size();
    return 1;
}

void orc_crc16(const char* s, size_t n, char out[5])
{
    const std::string r = crc16_hex(std::string(s, n));
    memcpy(out, r.data(), 4); out[4] = 0;
}

double orc_bench(const hbo_config* cfg, int n_threads, const float* iq, size_t n, size_t stride,
                 size_t chunk, double fs, int reps, uint64_t* o_chars)
{
    std::atomic<int> ready{0};
    std::atomic<bool> go{false};
    std::atomic<uint64_t> chars{0};
    std::vector<double> secs(n_threads, 0.0);
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; ++t)
        th.emplace_back([&, t] {
            Port P;
            hbo_config c = *cfg; c.record = 0;
            P.configure(c);
            const float* src = iq + 2 * size_t(t) * stride;
            ready++;
            while (!go.load()) std::this_thread::yield();
            auto t0 = std::chrono::steady_clock::now();
            for (int rep = 0; rep < reps; ++rep)
                for (size_t o = 0; o < n; o += chunk) {
                    P.push(src + 2 * o, std::min(chunk, n - o), fs);
                    P.process();
                }
            secs[t] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            chars += P.chars_all.size();
        });
    while (ready.load() < n_threads) std::this_thread::yield();
    go = true;
    for (auto& x : th) x.join();
    if (o_chars) *o_chars = chars.load();
    return *std::max_element(secs.begin(), secs.end());
}

void orc_run_ring(const hbo_config* cfg, int n_threads, const float* iq, size_t n_channels, size_t stride, size_t ring_n, size_t chunk,
                  size_t first_chunk, size_t n_chunks, double fs, char* chars_out, size_t chars_pitch, uint32_t* chars_len,
                  char* sent_out, size_t sent_pitch, uint32_t* sent_len)
{
    std::atomic<size_t> next{0};
    const size_t slices = ring_n / chunk;
    std::vector<std::thread> pool;
    for (int t = 0; t < std::max(1, n_threads); ++t)
        pool.emplace_back([&] {
            for (size_t c; (c = next.fetch_add(1)) < n_channels;) {
                Port P;
                hbo_config cc = *cfg; cc.record = 0;
                P.configure(cc);
                const float* src = iq + 2 * c * stride;
                for (size_t k = 0; k < n_chunks; ++k) {
                    P.push(src + 2 * (((first_chunk + k) % slices) * chunk), chunk, fs);
                    P.process();
                }
                chars_len[c] = uint32_t(P.chars_all.size()); sent_len[c] = uint32_t(P.sentences.size());
                memcpy(chars_out + c * chars_pitch, P.chars_all.data(), std::min(P.chars_all.size(), chars_pitch));
                memcpy(sent_out + c * sent_pitch, P.sentences.data(), std::min(P.sentences.size(), sent_pitch));
            }
        });
    for (auto& x : pool) x.join();
}

} // extern "C"
