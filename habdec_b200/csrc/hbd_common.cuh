// habdec_b200 -- shared device/host definitions for the batched IQ -> characters path.
//
// One `hbd_decoder` handles N independent channels (each one equivalent to one
// reference habdec::Decoder<float>, code/Decoder/Decoder.h:51-203).  All per-channel
// stream state lives in HBM so that every hbd_process() is a fixed sequence of
// kernels with no host round trip:
//
//   K1 decim1   stage-1 FIR decimator          (Decimator.h:99-146, first stage)
//   K2 tail     stage-2 FIR -> DC-remove -> FFT frame -> low-pass -> discriminator -> bit slicer + UART
//               (Decoder.h:440-555, FirFilter.h:117-169, FSK2_Demod.h:30-42, SymbolExtractor.h:108-255, RTTY.h:77-137)
//   K4 fft_afc  4096-pt FFT + power + peaks + AFC state machine (FFT.cpp:77-99, AFC.h:92-329)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace hbd {

// ---- capacities (per channel) ------------------------------------------------------------------
constexpr int kCarryCapMin  = 1024;   // stage-1 carry row: (T1-1) history + (< total factor) unconsumed + margin; grows with cascaded plans
constexpr int kMidHist      = 352;    // history slots of the buffers between cascaded decimator stages (>= 348 - 1)
constexpr int kS1Hist       = 320;    // stage-2 history slots (>= T2-1 = 138; doubles as the single-stage carry)
constexpr int kLpMaxTaps    = 1025;   // low-pass taps upper bound ((4/trans)|1, trans >= 0.0039)
constexpr int kLpHist       = 1024;   // low-pass history slots (>= kLpMaxTaps-1)
constexpr int kLpBatch      = 256;    // Decoder.h:492 batch_size
constexpr int kFftN         = 4096;   // Decoder.h:163 fft_bins_cnt_ (default; hbd_set_fft_size selects 16384)
constexpr int kFftNMax      = 16384;
constexpr int kSlicerVent   = 30000;  // SymbolExtractor.h:116 safety vent (3e4)
constexpr int kBitsCap      = 16384;  // bits the slicer may emit in one call (+ pending UART bits)
constexpr unsigned kMaskWords = 192;    // cached slicer mask words per channel and mask (6144 positions, the most the tail kernel stages)
constexpr unsigned kUartRunsCap = 2048; // UART backlog entries per channel (slicer runs since the last decoded character)
constexpr unsigned kLogCapMin = 1u << 20; // decoded-character log entries (ring) between host drains: max(this, 512 per channel)

// control words of the result logs (one device array, copied to a pinned per-call slot after the kernels of every call)
enum LogCtl {
    kCtlCharHead = 0,     // monotonic append counter of the character log
    kCtlSsdvHead = 1,     // ... of the SSDV packet log
    kCtlSsdvRingOvf = 2,  // a call appended more raw characters than a channel's SSDV ring holds
    kCtlCharTail = 3,     // entries below this index have been consumed by the host (uploaded after every drain)
    kCtlCharOvf = 4,      // characters a writer had to DROP because the ring was full (never overwrites unread entries)
    kCtlSsdvTail = 5,
    kCtlSsdvOvf = 6,      // SSDV packets dropped for the same reason
    kCtlWords = 8
};

// ---- per-channel persistent state (device resident, one struct per channel) -----------------------
struct ChanState {
    // configuration (written by the host when a setter is called)
    double baud;            // SymbolExtractor::symbolRate
    float  rtty_stops;      // RTTY::ascii_stops (float on purpose, RTTY.h:85)
    int    rtty_bits;       // RTTY::ascii_bits
    int    dc_remove;       // Decoder::dc_remove
    int    lp_ntaps;        // current low-pass tap count (0: not designed yet)

    // stream bookkeeping written by the kernels
    unsigned dec_pending;   // decimated samples queued in front of the low-pass (< 256 after a call)
    unsigned fft_have;      // samples collected for the next FFT frame
    unsigned fft_ready;     // 1: frame complete, K4 must transform it
    unsigned afc_tick;      // 1: this call reaches the AFC step (Decoder.h:494-509)
    unsigned have_spectrum; // 1 after the first FFT
    unsigned demod_primed;  // discriminator carry valid
    float    demod_last_re, demod_last_im;
    unsigned n_filtered;    // low-pass outputs produced by this call (multiple of 256)
    unsigned demod_n;       // size of the reference's demodulated_ vector: n_filtered of the last call that produced any
    unsigned slicer_n;      // pending slicer samples
    unsigned uart_n;        // pending UART bits (< one frame)
    unsigned long long uart_win; // those bits, LSB = oldest
    unsigned uart_runs_n;   // entries of the channel's UART backlog (slicer_dev.cuh: bits since the last decoded character)
    unsigned uart_ovf;      // the backlog outgrew its buffer
    unsigned uart_rescan;   // rtty_bits / rtty_stops changed: the backlog is re-evaluated when the next bit arrives
    unsigned mask_valid;    // slicer positions [0, mask_valid) whose flip-search mask bits are cached in HBM (tail.cu), for radius mask_R
    int      mask_R;
    unsigned pad0_;

    // AFC (AFC.h:72-90): two Average<double>(100), two Average<int>(4)
    double   afc_correction, afc_noise_floor, afc_noise_var, afc_shift_hz;
    double   nf_sum;  unsigned nf_cnt;
    double   nv_sum;  unsigned nv_cnt;
    int      pl_sum;  unsigned pl_cnt;
    int      pr_sum;  unsigned pr_cnt;
    int      gui_left, gui_right;
    // cached results of the last spectrum (recomputed by the reference on every call, identical values)
    int      spec_ok;       // FftPower() verdict
    int      spec_p1, spec_p2;
    float    spec_p1_val, spec_p2_val;
    double   spec_nf, spec_nv;
};

// ---- per-call plan (host mirror -> device; identical from call to call in steady state) ------------
struct ChanPlan {
    unsigned r;         // unconsumed input samples carried from the previous call
    unsigned n;         // samples pushed for this call
    unsigned consumed;  // (r+n) - (r+n) % factor      (Decoder.h:432)
    unsigned n1;        // stage-1 outputs  = consumed / M1
    unsigned n2;        // decimated outputs = consumed / factor
    unsigned flags;     // bit0: nothing to do this call
    unsigned dec_pending; // decimated samples queued in front of the low-pass BEFORE this call (host mirror of ChanState)
    unsigned lp_ntaps;  // low-pass tap count in force for this call
};

struct DecimGeometry {
    int M1, T1;         // first stage
    int M2, T2;         // second stage (M2 == 1: none)
    int factor;
};

__host__ __device__ inline unsigned hbd_min_u(unsigned a, unsigned b) { return a < b ? a : b; }

#ifdef __CUDACC__
// complex sample x real tap, accumulated: ONE packed FFMA2 (fma.rn.f32x2, sm_100+) instead of two FFMA.  Each half is an
// IEEE fused multiply-add, so the value equals fmaf() per component; the FMA pipe does the same work, the issue
// slot count halves (tools/micro/ffma2_bench.cu: same 127 FMA/clk/SM at half the instructions).
__device__ __forceinline__ float2 cfma(float2 x, float h, float2 acc)
{
    float2 hh = make_float2(h, h);
    unsigned long long rx = *reinterpret_cast<unsigned long long*>(&x), rh = *reinterpret_cast<unsigned long long*>(&hh),
                       ra = *reinterpret_cast<unsigned long long*>(&acc), rd;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(rx), "l"(rh), "l"(ra));
    return *reinterpret_cast<float2*>(&rd);
}
__device__ __forceinline__ float2 cadd2(float2 a, float2 b) // packed FADD2
{
    unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b), rd;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
    return *reinterpret_cast<float2*>(&rd);
}
#endif

} // namespace hbd

#define HBD_CUDA_CHECK(call)                                                                    \
    do {                                                                                        \
        cudaError_t e__ = (call);                                                               \
        if (e__ != cudaSuccess) {                                                               \
            set_error(std::string(#call) + ": " + cudaGetErrorString(e__));                     \
            return HBD_ERR_CUDA;                                                                \
        }                                                                                       \
    } while (0)
