// C ABI of libhabdec_b200.so (include/habdec_b200.h): handle, HBM state, per-call planning
// (the host mirror of the reference's buffer bookkeeping, Decoder.h:426-436,492-542), kernel
// sequencing and the host sentence layer.  There is no CPU compute path in here: every
// process call launches K1..K4 or fails.
#include "../../include/habdec_b200.h"

#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <sched.h>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "decim1.cuh"
#include "mid.cuh"
#include "decim_taps.inc"
#include "fft_afc.cuh"
#include "hbd_common.cuh"
#include "host_tail.h"
#include "range_pool.h"
#include "telemetry_abi.h"
#include "api_internal.h"
#include "nco.cuh"
#include "ssdv.cuh"
#include "tail.cuh"
#include "wire.cuh"

using namespace hbd;

namespace {

struct TapTable { int M; const uint32_t* bits; int len; };
#define HBD_TT(M, name) TapTable{M, hbd_taps_bits_##name, int(sizeof(hbd_taps_bits_##name) / 4)}

// factor -> stage list, Decoder.h:286-320
bool plan_for_factor(size_t factor, std::vector<TapTable>& st)
{
    st.clear();
    switch (factor) {
    case 256: st = {HBD_TT(64, d256_m64), HBD_TT(4, d4_m4)}; return true;
    case 128: st = {HBD_TT(32, d128_m32), HBD_TT(4, d4_m4)}; return true;
    case 64:  st = {HBD_TT(32, d64_m32), HBD_TT(2, d2_m2)}; return true;
    case 32:  st = {HBD_TT(16, d32_m16), HBD_TT(2, d2_m2)}; return true;
    case 16:  st = {HBD_TT(8, d16_m8), HBD_TT(2, d2_m2)}; return true;
    case 8:   st = {HBD_TT(8, d8_m8)}; return true;
    case 4:   st = {HBD_TT(4, d4_m4)}; return true;
    case 2:   st = {HBD_TT(2, d2_m2)}; return true;
    case 1:   return true;
    default:  return false;
    }
}

constexpr int kMaxMidStages = 6;   // a cascaded plan has at most 8 stages: K1's, up to six in the middle, the tail kernel's

// host mirror of one channel, part 1: stream counters that are pure functions of the push sizes (and of the
// low-pass configuration).  While every channel is fed the same way they are identical for all channels, and only
// hc[0]'s copy is kept current (hbd_decoder::ctr_uniform / sync_ctrs): a steady-state call costs O(1) host work.
struct ChanCtr {
    unsigned in_r = 0, dec_pending = 0;
    size_t grown1 = 0, grown2 = 0, grown_lp = 0;
    size_t grown_mid[kMaxMidStages] = {0, 0, 0, 0, 0, 0};   // work-buffer sizes of the middle stages of a cascaded plan
    unsigned fft_have = 0;   // mirror of ChanState::fft_have: tells the host which calls complete an FFT frame (K4 launch)
    size_t lp_input_size = 0, lp_ntaps = 0;
    unsigned pushed = 0;     // samples waiting in the staging row
    unsigned last_nf = 0, last_n2 = 0;
    unsigned demod_n = 0;    // size of the reference's demodulated_ (last call that produced any; Decoder.h:546-555)
    bool lp_dirty = true;    // bw / trans / input size changed since the last design attempt
    bool same(const ChanCtr& o) const
    {
        for (int k = 0; k < kMaxMidStages; ++k) if (grown_mid[k] != o.grown_mid[k]) return false;
        return in_r == o.in_r && dec_pending == o.dec_pending && grown1 == o.grown1 && grown2 == o.grown2 && grown_lp == o.grown_lp &&
               fft_have == o.fft_have && lp_input_size == o.lp_input_size && lp_ntaps == o.lp_ntaps && pushed == o.pushed &&
               last_nf == o.last_nf && last_n2 == o.last_n2 && demod_n == o.demod_n && lp_dirty == o.lp_dirty;
    }
};
// part 2: configuration (always per channel)
struct HostChan : ChanCtr {
    double baud = 1;         // SymbolExtractor default symbol_rate_ = 1
    size_t rtty_bits = 0;
    float rtty_stops = 0;
    float lp_bw = 1500, lp_trans = 0.025f;
    bool dc_remove = false;
    double nco_freq = 0, nco_phase = 0; // pre-mixer: frequency in Hz, phase (cycles) of the next pushed sample
    bool cfg_dirty = true;
    ChanCtr& ctr() { return *this; }
};

constexpr unsigned kMaxCallsBetweenCollects = 256; // process_async forces a drain beyond this (log capacity)

__global__ void init_cfg_kernel(ChanState* st, const double* baud, const float* stops, const int* bits, const int* dc, const int* ntaps,
                                const unsigned char* dirty, int n_ch)
{
    const int ch = blockIdx.x * blockDim.x + threadIdx.x;
    if (ch >= n_ch || !dirty[ch]) return;
    st[ch].baud = baud[ch];
    if (st[ch].rtty_stops != stops[ch] || st[ch].rtty_bits != bits[ch]) st[ch].uart_rescan = 1;   // RTTY.h:90-134 rescans bits_ under the new framing
    st[ch].rtty_stops = stops[ch];
    st[ch].rtty_bits = bits[ch];
    st[ch].dc_remove = dc[ch];
    st[ch].lp_ntaps = ntaps[ch];
}

// the per-channel scalars that travel with the gathered results (hbd_result_record), packed after every call
__global__ void stats_snap_kernel(const ChanState* __restrict__ st, float* __restrict__ out, int n_ch)
{
    const int ch = blockIdx.x * blockDim.x + threadIdx.x;
    if (ch >= n_ch) return;
    const ChanState& s = st[ch];
    float* o = out + size_t(ch) * 6;
    o[0] = float(s.afc_correction); o[1] = float(s.afc_shift_hz); o[2] = float(s.afc_noise_floor); o[3] = float(s.afc_noise_var);
    o[4] = float(s.gui_left); o[5] = float(s.gui_right);
}

} // namespace

// a user callback recorded under the handle's mutex and fired after it is released (a callback may call hbd_* again)
struct Deferred {
    enum Kind { kSentence, kChars, kSsdv } kind;
    int ch;
    std::string a, b, c;                                   // sentence: callsign, data, crc; chars: a
    hbd_ssdv_packet_info info; std::array<unsigned char, 256> pkt;
};

// what one thread of the drain records for its channel range (collect_locked)
struct ReplayPart { std::vector<std::pair<unsigned, Deferred>> events; bool spilled = false; };

struct hbd_decoder {
    std::mutex mtx;
    std::string err;
    int n_ch = 0, device = 0, n_sms = 148;
    cudaStream_t stream = nullptr;   // the caller's stream: inputs are ordered on it, it waits until inputs are consumed
    bool own_stream = false;
    // Two streams: the HBM-bound K1 of call s+1 overlaps the FP32/latency-bound tail kernels of call s (the stage-1
    // stream and the carry are double buffered over calls, so K1 only ever waits for the tail of call s-1).
    //   hi (high priority): K1 (+ carry) of every call
    //   lo (low priority):  K2 tail / K4 fft_afc; they fill the SM resources K1 leaves free
    cudaStream_t hi = nullptr, lo = nullptr;
    cudaEvent_t ev_consumed = nullptr;      // K1 + carry done (input consumed, stage-1 output ready)
    cudaEvent_t ev_tail[2] = {nullptr, nullptr};   // per stage-1 buffer: tail done, that buffer may be overwritten by K1
    bool tail_pending[2] = {false, false};
    bool tail_ev_late = false;              // HBD_TAIL_EV_LATE=1 (measurement hook): record ev_tail after the whole lo-stream sequence
    cudaEvent_t ev_in = nullptr;
    int sync_groups()
    {
        if (hi && cudaStreamSynchronize(hi) != cudaSuccess) return HBD_ERR_CUDA;
        if (lo && cudaStreamSynchronize(lo) != cudaSuccess) return HBD_ERR_CUDA;
        return HBD_OK;
    }
    bool record = false;
    unsigned long long launches = 0;

    double fs_in = 0;
    int factor = 1;
    int M1 = 1, T1 = 1, M2 = 1, T2 = 1;   // first stage (K1) and last stage (tail kernel; M2 == 1: the plan has one stage)
    std::vector<TapTable> stages;         // the whole plan (Decoder.h:286-320, cascaded by setupDecimationStagesBW :350-399)
    struct MidStage { int M = 1, T = 1; unsigned div_in = 1; float* d_taps = nullptr; float2* d_out = nullptr; size_t pitch = 0; };
    std::vector<MidStage> mids;           // stages[1 .. size-2]
    float2* d_midlast[2] = {nullptr, nullptr}; size_t midlast_pitch = 0, midlast_pitch_b = 0;   // output of the last middle stage = the tail kernel's
                                          // input, ping-pong over calls like d_s1x
    unsigned div_last_in = 1;             // samples entering the last stage per call = consumed / div_last_in
    int carry_cap = kCarryCapMin;
    int set_plan(const std::vector<TapTable>& st, size_t total_factor);
    int fft_n = kFftN;       // spectrum size: 4096 like the reference, or 16384 (hbd_set_fft_size)
    int alloc_fft();
    std::vector<HostChan> hc;        // per-channel configuration + counters (see ChanCtr)
    // Uniform mode: every channel has been fed the same way since creation (whole-batch pushes only), so all ChanCtr are
    // equal and only hc[0]'s is stepped per call; hc[1..] lag behind (ctr_stale) until sync_ctrs() copies them.
    bool ctr_uniform = true, ctr_stale = false;
    bool lp_cfg_uniform = true;      // lp_bw / lp_trans are the same for all channels (setters with ch == -1 only)
    bool cfg_dirty_any = true;       // some channel has cfg_dirty set
    bool derived_dirty = true;       // max_lp_taps_c / max_spb_c need recomputing
    size_t max_lp_taps_c = 1; double max_spb_c = 1;
    void sync_ctrs() { if (ctr_stale) { for (size_t c = 1; c < hc.size(); ++c) hc[c].ctr() = hc[0].ctr(); ctr_stale = false; } }
    void leave_uniform() { sync_ctrs(); ctr_uniform = false; }
    // staged (host-pushed, not yet processed) samples: the same count in every channel?
    bool queue_depth_equal(unsigned& base)
    {
        base = hc[0].pushed;
        if (ctr_uniform) return true;
        for (const auto& x : hc) if (x.pushed != base) return false;
        return true;
    }
    void add_pushed(unsigned n)
    {
        if (ctr_uniform) { hc[0].pushed += n; ctr_stale = true; }
        else for (auto& x : hc) x.pushed += n;
    }
    std::vector<TextChannel> text;   // sentence layer state, touched only when a channel produced characters
    bool spill_free = true;                // no channel holds spilled (beyond-the-slot) pending bytes: hbd_pack_results stays sequential
    std::vector<hbd_result_record> pend;   // per channel: characters / sentence bytes waiting for a poll or the next gather (host_tail.h)

    // device state
    ChanState* d_state = nullptr;
    ChanPlan* d_plan = nullptr;
    std::vector<ChanPlan> h_plan, h_plan_uploaded;
    float2* d_carry2[2] = {nullptr, nullptr}; // stage-1 carry, ping-pong: K1 reads [carry_cur], writes [carry_cur ^ 1]
    int carry_cur = 0;
    float2* d_s1x[2] = {nullptr, nullptr}; size_t s1_pitch = 0, s1_pitch_b = 0; // stage-1 stream, ping-pong over calls:
    int s1_cur = 0;              // K1 + tail of a call use [s1_cur]; the tail leaves the stage-2 history in [s1_cur ^ 1]
    float2* d_decq = nullptr;    size_t dq_pitch = 0;
    float2* d_fftbuf = nullptr;
    float2* d_spectrum = nullptr;
    float* d_power = nullptr;
    float* d_lptaps = nullptr;
    float* d_slicer = nullptr;   size_t slicer_pitch = 0;
    float* d_demod = nullptr;    size_t demod_pitch = 0;
    unsigned* d_maskc = nullptr;             // cached slicer position masks [n][2][kMaskWords] (tail.cu)
    unsigned short* d_uart_runs = nullptr;   // UART backlog [n][kUartRunsCap] (slicer_dev.cuh)
    // Decoded-character log: ring of log_cap entries appended by the tail kernels (monotonic head in d_log_ctl[0]).
    // After the kernels of a call the control words are copied to that call's slot of h_heads (pinned), in stream order:
    // the host only ever reads a COMMITTED head, never the live one (entries below it are completely written).
    uint2* d_log = nullptr;
    unsigned log_cap = 0;                // power of two, sized from the channel count
    unsigned* d_log_ctl = nullptr;       // kCtlWords control words (see LogCtl in hbd_common.cuh)
    unsigned* h_heads = nullptr;         // pinned, [ev_call.size()][kCtlWords]
    unsigned* h_tail_pin = nullptr;      // pinned, [2]: char log tail, SSDV log tail (source of the async upload)
    unsigned log_tail = 0;               // host: entries below this index are already processed
    unsigned call_seq = 0;               // calls enqueued so far (24 bits travel in the log entries)
    unsigned calls_collected = 0;        // calls whose characters went through the sentence layer
    unsigned done_seq = 0;               // calls known to have completed (event queries, see log_pressure)
    unsigned done_head = 0;              // committed character-log head of call done_seq - 1
    unsigned per_call_max = 0;           // largest number of characters one completed call appended so far
    unsigned long long chars_lost = 0;   // characters dropped by log overflows (reported by hbd_collect*)
    unsigned ovf_seen[3] = {0, 0, 0};    // last seen kCtlSsdvRingOvf / kCtlSsdvOvf / kCtlCharOvf (monotonic counters)
    std::vector<cudaEvent_t> ev_call;    // ring: completion of call (seq % size)
    cudaStream_t copy_stream = nullptr;  // result read-back, independent of the compute streams
    std::vector<uint2> h_log;            // host staging
    std::vector<std::string> call_chars; // per channel: raw chars of the call being replayed
    std::vector<unsigned> seg_start;     // collect_locked: first log entry of every (call, channel) segment
    static constexpr int kCharBufHost = 64;   // == kCharBuf (slicer_dev.cuh): characters per device-side flush
    double drain_host_ms = 0; unsigned drain_calls = 0;   // host time of the replay part of the drains (hbd_get_kernel_timing which = 5)
    hbd::RangePool pool;                 // the drain's worker threads (part t of a drain / a pack runs on the same thread every time)
    int host_threads = 1;                // threads the drain may use (hbd_set_host_threads; default: min(4, half the cores this process may run on))
    bool keep_raw = true;                // hbd_set_raw_chars: retain the raw (unfiltered) characters for hbd_poll_raw_chars
    float* d_taps1 = nullptr; float* d_taps2 = nullptr;
    float2* d_twiddle = nullptr;
    // config upload scratch
    double* d_cfg_baud = nullptr; float* d_cfg_stops = nullptr; int* d_cfg_bits = nullptr; int* d_cfg_dc = nullptr;
    int* d_cfg_ntaps = nullptr; unsigned char* d_cfg_dirty = nullptr;
    // recordings
    float2* d_rec_dec = nullptr; float2* d_rec_filt = nullptr; size_t rec_pitch = 0;
    unsigned char* d_rec_bits = nullptr; unsigned* d_rec_bits_n = nullptr; unsigned rec_bits_pitch = 0;

    // input
    float2* d_stage = nullptr; size_t stage_pitch = 0;   // host pushes land here
    const float2* ext = nullptr; size_t ext_pitch = 0; size_t ext_n = 0; // zero-copy device push
    // fused-NCO parameter ring: pinned host slots + device slots, so a push never waits for the GPU
    static constexpr unsigned kNcoSlots = 16;
    NcoChan* h_nco_pin = nullptr; NcoChan* d_nco_ring = nullptr; cudaEvent_t ev_nco[kNcoSlots] = {}; unsigned nco_seq = 0;
    const NcoChan* ext_nco_ptr = nullptr;
    bool ext_nco = false;    // ... of RAW samples that K1 mixes through the channels' NCOs itself (fused K0; ext_pitch 0 = wideband row)
    bool nco_fused = true;   // HBD_NCO_FUSED=0 (measurement / test hook): always pre-mix with K0 into the staging matrix
    int push_fused_nco(const float2* src, size_t pitch, size_t n);
    float* h_pinned = nullptr; size_t pinned_bytes = 0;  // staging for pageable host memory

    // websocket wire formats (wire.cu)
    bool demod_acc_on = false;
    float* d_dacc = nullptr; size_t dacc_pitch = 0; unsigned* d_dacc_n = nullptr;
    unsigned char* d_frames = nullptr; size_t frames_pitch = 0; unsigned* d_frame_sizes = nullptr;
    std::vector<unsigned> h_frame_sizes;
    int ensure_frames(size_t pitch);
    int ensure_dacc();
    // NCO pre-mixer (nco.cu)
    NcoChan* d_nco = nullptr; std::vector<NcoChan> h_nco;
    float2* d_wide = nullptr; size_t wide_cap = 0;       // wideband capture row shared by all channels
    bool nco_any = false;    // some channel has (had) a non-zero NCO frequency or phase (set by hbd_set_nco, never cleared)
    bool nco_active() const { return nco_any; }
    int mix_into_stage(const float2* src, size_t src_pitch, int c0, int nc, size_t dst_off, size_t n);
    int pending_marks = 0;   // async calls since the last collect
    int sv_override = -1;
    bool mask_cache = true;      // HBD_MASK_CACHE=0 (test hook): the slicer's position masks are rebuilt from scratch in every call
    // optional per-kernel CUDA-event timing (bench roofline): one event pair per K1 launch / per rest-of-step
    int timing = 0;          // 0 off, 1: events around K1 only (what the roofline needs), 2: also around the rest of the step (pipeline diagnostics)
    std::vector<cudaEvent_t> ev_k1, ev_rest; // pairs: [2i] start, [2i+1] stop
    size_t ev_used_k1 = 0, ev_used_rest = 0;
    cudaEvent_t next_event(std::vector<cudaEvent_t>& pool, size_t& used)
    {
        if (used == pool.size()) { cudaEvent_t e; cudaEventCreate(&e); pool.push_back(e); }
        return pool[used++];
    }
    size_t cap_n_in = 0;     // largest (r + n) the per-call buffers are sized for

    // SSDV packet sync (ssdv.cu + host_tail.cpp SsdvChannel); off until hbd_set_ssdv / hbd_set_ssdv_callback
    bool ssdv_on = false;
    unsigned char* d_ssdv_ring = nullptr; unsigned* d_ssdv_total = nullptr; unsigned* d_ssdv_scanned = nullptr;
    SsdvLogEntry* d_ssdv_log = nullptr; unsigned ssdv_log_tail = 0;
    std::vector<SsdvLogEntry> h_ssdv_log;
    std::vector<SsdvChannel> ssdv;
    hbd_ssdv_cb ssdv_cb = nullptr; void* ssdv_user = nullptr;
    int enable_ssdv(bool on);

    hbd_sentence_cb sentence_cb = nullptr; void* sentence_user = nullptr;
    hbd_tracker* tracker = nullptr; int tracker_off = 0;   // telemetry layer fed after the sentence callback
    hbd_chars_cb chars_cb = nullptr; void* chars_user = nullptr;
    std::vector<Deferred> deferred;  // callbacks recorded by collect_locked, fired by the entry point after unlocking
    // result gather (gather.cpp / dist.cu): AFC scalars of every channel snapshotted after each call into a small device
    // ring, so that a drain can read the values that belong to the calls it drains without stopping the pipeline
    static constexpr unsigned kSnapSlots = 32;
    bool snap_on = false;
    float* d_snap = nullptr;         // [kSnapSlots][n][6]
    float* h_snap = nullptr;         // pinned, [n][6]: snapshot of the newest drained call
    bool h_snap_valid = false;
    void* dist_ctx = nullptr;        // dist.cu
    int enable_snap();

    void set_error(const std::string& e) { err = e; }
    int ensure_call_capacity(size_t n_in_max);
    int alloc_fixed();
    int process_async_locked();
    int collect_locked(unsigned lag);
    int log_pressure();
    void free_all();
};

// Runs `body` (a lambda returning int) under the handle's mutex, then fires the callbacks it recorded with the mutex
// released -- a callback may call any hbd_* function of the same handle (the reference's getRTTY()/getLastSentence()
// are callable from sentence_callback_ as well).
namespace {
struct CallbackSet { hbd_sentence_cb s; void* su; hbd_chars_cb c; void* cu; hbd_ssdv_cb v; void* vu; hbd_tracker* trk; int off; };
void fire_deferred(const std::vector<Deferred>& ev, const CallbackSet& cb)
{
    for (const Deferred& d : ev) {
        switch (d.kind) {
        case Deferred::kSentence:
            if (cb.s) cb.s(cb.su, d.ch, d.a.c_str(), d.b.c_str(), d.c.c_str());
            if (cb.trk) tracker_feed(cb.trk, d.ch + cb.off, d.a, d.b, d.c);      // SentenceCallback, websocketServer/main.cpp:292-366
            break;
        case Deferred::kChars: if (cb.c) cb.c(cb.cu, d.ch, d.a.data(), d.a.size()); break;
        case Deferred::kSsdv: if (cb.v) cb.v(cb.vu, d.ch, &d.info, d.pkt.data()); break;
        }
    }
}
template <typename F>
int locked_then_fire(hbd_decoder* h, F body)
{
    std::vector<Deferred> ev;
    CallbackSet cb{};
    int rc;
    {
        std::lock_guard<std::mutex> l(h->mtx);
        rc = body();
        ev.swap(h->deferred);
        cb = CallbackSet{h->sentence_cb, h->sentence_user, h->chars_cb, h->chars_user, h->ssdv_cb, h->ssdv_user, h->tracker, h->tracker_off};
    }
    if (!ev.empty()) fire_deferred(ev, cb);
    return rc;
}
} // namespace

static void fill_ssdv_info(hbd_ssdv_packet_info& info, const SsdvEvent& ev)
{
    memset(&info, 0, sizeof(info));
    memcpy(info.callsign, ev.header.callsign, sizeof(info.callsign));
    info.image_id = ev.header.image_id; info.packet_id = ev.header.packet_id;
    info.width = ev.header.width; info.height = ev.header.height;
    info.errors = ev.errors; info.set_size = ev.set_size;
}

#define HBD_CHECK_H(h)  if (!(h)) return HBD_ERR_ARG
#define HBD_CHECK_CH(h, ch) if (!(h) || (ch) < 0 || (ch) >= (h)->n_ch) return HBD_ERR_ARG

template <typename T>
static cudaError_t dalloc(T** p, size_t count) { return cudaMalloc((void**)p, std::max<size_t>(count, 1) * sizeof(T)); }

// grow a [rows][pitch] device matrix, keeping the first `keep` elements of every row
template <typename T>
static cudaError_t grow_rows(T** p, size_t* pitch, size_t new_pitch, size_t rows, size_t keep, cudaStream_t s)
{
    if (new_pitch <= *pitch && *p) return cudaSuccess;
    T* np = nullptr;
    cudaError_t e = cudaMalloc((void**)&np, rows * new_pitch * sizeof(T));
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(np, 0, rows * new_pitch * sizeof(T), s);
    if (e != cudaSuccess) return e;
    if (*p && keep) {
        e = cudaMemcpy2DAsync(np, new_pitch * sizeof(T), *p, *pitch * sizeof(T), std::min(keep, *pitch) * sizeof(T), rows,
                              cudaMemcpyDeviceToDevice, s);
        if (e != cudaSuccess) return e;
    }
    e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) return e;
    if (*p) cudaFree(*p);
    *p = np;
    *pitch = new_pitch;
    return cudaSuccess;
}

int hbd_decoder::alloc_fixed()
{
    const size_t n = size_t(n_ch);
    HBD_CUDA_CHECK(dalloc(&d_state, n));
    {
        // Average<T>'s constructor add()s one 0: count starts at 1 (Average.h:34-37)
        ChanState init;
        memset(&init, 0, sizeof(init));
        init.baud = 1; // SymbolExtractor.h:96
        init.nf_cnt = init.nv_cnt = init.pl_cnt = init.pr_cnt = 1;
        std::vector<ChanState> all(n, init);
        HBD_CUDA_CHECK(cudaMemcpy(d_state, all.data(), n * sizeof(ChanState), cudaMemcpyHostToDevice));
    }
    HBD_CUDA_CHECK(dalloc(&d_plan, n));
    for (int i = 0; i < 2; ++i) {
        HBD_CUDA_CHECK(dalloc(&d_carry2[i], n * size_t(carry_cap)));
        HBD_CUDA_CHECK(cudaMemset(d_carry2[i], 0, n * size_t(carry_cap) * sizeof(float2)));
    }
    { const int rc = alloc_fft(); if (rc) return rc; }
    HBD_CUDA_CHECK(dalloc(&d_uart_runs, n * size_t(kUartRunsCap)));
    HBD_CUDA_CHECK(dalloc(&d_maskc, n * size_t(2 * kMaskWords)));
    HBD_CUDA_CHECK(dalloc(&d_lptaps, n * kLpMaxTaps));
    HBD_CUDA_CHECK(cudaMemset(d_lptaps, 0, n * kLpMaxTaps * sizeof(float)));
    // character log: at least 512 entries per channel (a 600 baud channel decodes ~2 characters per 65 536-sample call,
    // so that is > 250 calls between drains even when every channel is noise); log_pressure() drains before it fills
    log_cap = kLogCapMin;
    while (size_t(log_cap) < 512 * n && log_cap < (1u << 30)) log_cap <<= 1;
    HBD_CUDA_CHECK(dalloc(&d_log, size_t(log_cap)));
    HBD_CUDA_CHECK(dalloc(&d_log_ctl, size_t(kCtlWords)));
    HBD_CUDA_CHECK(cudaMemset(d_log_ctl, 0, kCtlWords * sizeof(unsigned)));
    HBD_CUDA_CHECK(cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking));
    ev_call.resize(kMaxCallsBetweenCollects * 2);
    for (auto& e : ev_call) HBD_CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    HBD_CUDA_CHECK(cudaHostAlloc((void**)&h_heads, ev_call.size() * kCtlWords * sizeof(unsigned), cudaHostAllocDefault));
    memset(h_heads, 0, ev_call.size() * kCtlWords * sizeof(unsigned));
    HBD_CUDA_CHECK(cudaHostAlloc((void**)&h_tail_pin, 2 * sizeof(unsigned), cudaHostAllocDefault));
    call_chars.resize(n);
    HBD_CUDA_CHECK(dalloc(&d_taps1, 512));
    HBD_CUDA_CHECK(dalloc(&d_taps2, 512));
    HBD_CUDA_CHECK(dalloc(&d_cfg_baud, n));
    HBD_CUDA_CHECK(dalloc(&d_cfg_stops, n));
    HBD_CUDA_CHECK(dalloc(&d_cfg_bits, n));
    HBD_CUDA_CHECK(dalloc(&d_cfg_dc, n));
    HBD_CUDA_CHECK(dalloc(&d_cfg_ntaps, n));
    HBD_CUDA_CHECK(dalloc(&d_cfg_dirty, n));
    HBD_CUDA_CHECK(dalloc(&d_nco, n));
    h_nco.resize(n);
    return HBD_OK;
}

// spectrum buffers + twiddle table for the current fft_n (also called by hbd_set_fft_size)
int hbd_decoder::alloc_fft()
{
    const size_t n = size_t(n_ch), N = size_t(fft_n);
    for (void* p : {(void*)d_fftbuf, (void*)d_spectrum, (void*)d_power, (void*)d_twiddle}) if (p) cudaFree(p);
    d_fftbuf = d_spectrum = nullptr; d_power = nullptr; d_twiddle = nullptr;
    HBD_CUDA_CHECK(dalloc(&d_fftbuf, n * N));
    HBD_CUDA_CHECK(cudaMemset(d_fftbuf, 0, n * N * sizeof(float2)));
    HBD_CUDA_CHECK(dalloc(&d_spectrum, n * N));
    HBD_CUDA_CHECK(cudaMemset(d_spectrum, 0, n * N * sizeof(float2)));
    HBD_CUDA_CHECK(dalloc(&d_power, n * N));
    HBD_CUDA_CHECK(cudaMemset(d_power, 0, n * N * sizeof(float)));
    // [0, N): exp(-2 pi i e / N); behind it the twiddles of the 4096-point (sub-)transform in the order its threads read
    // them: pass 1 [k1][t] = W_4096^(t k1) (16 x 256), pass 2 [j1][m2] = W_256^(m2 j1) (16 x 16); all evaluated in float64
    HBD_CUDA_CHECK(dalloc(&d_twiddle, N + 4096 + 256));
    std::vector<float2> tw(N + 4096 + 256);
    auto w = [](double num, double den) { const double ang = -2.0 * M_PI * num / den; return make_float2(float(std::cos(ang)), float(std::sin(ang))); };
    for (size_t e = 0; e < N; ++e) tw[e] = w(double(e), double(N));
    for (size_t k1 = 0; k1 < 16; ++k1) for (size_t t = 0; t < 256; ++t) tw[N + k1 * 256 + t] = w(double(t * k1), 4096.0);
    for (size_t j1 = 0; j1 < 16; ++j1) for (size_t m2 = 0; m2 < 16; ++m2) tw[N + 4096 + j1 * 16 + m2] = w(double(16 * m2 * j1), 4096.0);
    HBD_CUDA_CHECK(cudaMemcpy(d_twiddle, tw.data(), sizeof(float2) * tw.size(), cudaMemcpyHostToDevice));
    return HBD_OK;
}

// Install a decimation plan (a list of (M, taps) stages with the given total factor): K1 runs stages[0], the tail kernel the
// last stage, mid.cu the ones in between.  A new plan starts from fresh decimators (Decoder.h:283-284,347-348:
// decimation_stages_.clear()); unconsumed input stays queued (iq_in_buffer_ is untouched) -- it sits at the end of the carry.
int hbd_decoder::set_plan(const std::vector<TapTable>& st, size_t total_factor)
{
    HBD_CUDA_CHECK(cudaSetDevice(device));
    if (sync_groups()) return HBD_ERR_CUDA;
    HBD_CUDA_CHECK(cudaStreamSynchronize(stream));
    const size_t n = size_t(n_ch);
    sync_ctrs();
    const int old_cap = carry_cap;
    stages = st;
    factor = int(total_factor);
    M1 = T1 = M2 = T2 = 1;
    if (!st.empty()) {
        M1 = st[0].M; T1 = st[0].len;
        HBD_CUDA_CHECK(cudaMemcpy(d_taps1, st[0].bits, 4 * size_t(T1), cudaMemcpyHostToDevice));
    }
    if (st.size() >= 2) {
        M2 = st.back().M; T2 = st.back().len;
        HBD_CUDA_CHECK(cudaMemcpy(d_taps2, st.back().bits, 4 * size_t(T2), cudaMemcpyHostToDevice));
    }
    for (MidStage& m : mids) { if (m.d_taps) cudaFree(m.d_taps); if (m.d_out) cudaFree(m.d_out); }
    mids.clear();
    for (int i = 0; i < 2; ++i) { if (d_midlast[i]) cudaFree(d_midlast[i]); d_midlast[i] = nullptr; }
    midlast_pitch = midlast_pitch_b = 0;
    unsigned div = unsigned(M1);
    for (size_t k = 1; k + 1 < st.size(); ++k) {
        MidStage m;
        m.M = st[k].M; m.T = st[k].len; m.div_in = div;
        HBD_CUDA_CHECK(dalloc(&m.d_taps, size_t(m.T)));
        HBD_CUDA_CHECK(cudaMemcpy(m.d_taps, st[k].bits, 4 * size_t(m.T), cudaMemcpyHostToDevice));
        mids.push_back(m);
        div *= unsigned(m.M);
    }
    div_last_in = div;
    // the carry holds T1-1 samples of history plus the unconsumed remainder (< total factor).  Fresh, zeroed rows (new
    // decimators have an empty history); the remainder queued under the old plan stays queued, as in the reference
    unsigned max_r = 0;
    for (const auto& x : hc) max_r = std::max(max_r, x.in_r);
    carry_cap = std::max(kCarryCapMin, int((size_t(T1) + std::max<size_t>(total_factor, max_r) + 64 + 63) & ~size_t(63)));
    float2* fresh[2] = {nullptr, nullptr};
    for (int i = 0; i < 2; ++i) {
        HBD_CUDA_CHECK(cudaMalloc((void**)&fresh[i], n * size_t(carry_cap) * sizeof(float2)));
        HBD_CUDA_CHECK(cudaMemset(fresh[i], 0, n * size_t(carry_cap) * sizeof(float2)));
    }
    for (size_t c = 0; c < n; ++c)
        if (hc[c].in_r)
            HBD_CUDA_CHECK(cudaMemcpy(fresh[carry_cur] + c * size_t(carry_cap) + size_t(carry_cap) - hc[c].in_r,
                                      d_carry2[carry_cur] + c * size_t(old_cap) + size_t(old_cap) - hc[c].in_r, sizeof(float2) * hc[c].in_r, cudaMemcpyDeviceToDevice));
    for (int i = 0; i < 2; ++i) { cudaFree(d_carry2[i]); d_carry2[i] = fresh[i]; }
    for (int i = 0; i < 2; ++i) if (d_s1x[i]) HBD_CUDA_CHECK(cudaMemset(d_s1x[i], 0, n * s1_pitch * sizeof(float2)));
    for (auto& x : hc) { x.grown1 = x.grown2 = 0; for (size_t& g : x.grown_mid) g = 0; }
    cap_n_in = 0;
    h_plan_uploaded.clear();
    return HBD_OK;
}

void hbd_decoder::free_all()
{
    cudaSetDevice(device);
    sync_groups();
    if (stream) cudaStreamSynchronize(stream);
    if (hi) cudaStreamDestroy(hi);
    if (lo) cudaStreamDestroy(lo);
    if (ev_consumed) cudaEventDestroy(ev_consumed);
    for (cudaEvent_t e : ev_tail) if (e) cudaEventDestroy(e);
    if (ev_in) cudaEventDestroy(ev_in);
    void* ptrs[] = {d_state, d_plan, d_carry2[0], d_carry2[1], d_s1x[0], d_s1x[1], d_decq, d_fftbuf, d_spectrum, d_power, d_lptaps, d_slicer, d_demod, d_uart_runs, d_maskc, d_log, d_log_ctl,
                    d_taps1, d_taps2, d_twiddle, d_cfg_baud, d_cfg_stops, d_cfg_bits, d_cfg_dc, d_cfg_ntaps,
                    d_cfg_dirty, d_rec_dec, d_rec_filt, d_rec_bits, d_rec_bits_n, d_stage, d_nco, d_wide, d_dacc, d_dacc_n, d_frames, d_frame_sizes,
                    d_ssdv_ring, d_ssdv_total, d_ssdv_scanned, d_ssdv_log};
    for (void* p : ptrs) if (p) cudaFree(p);
    if (h_pinned) cudaFreeHost(h_pinned);
    for (MidStage& m : mids) { if (m.d_taps) cudaFree(m.d_taps); if (m.d_out) cudaFree(m.d_out); }
    for (int i = 0; i < 2; ++i) if (d_midlast[i]) cudaFree(d_midlast[i]);
    if (dist_ctx) { internal_free_dist(dist_ctx); dist_ctx = nullptr; }
    if (d_snap) cudaFree(d_snap);
    if (h_snap) cudaFreeHost(h_snap);
    if (h_heads) cudaFreeHost(h_heads);
    if (h_tail_pin) cudaFreeHost(h_tail_pin);
    for (cudaEvent_t e : ev_call) cudaEventDestroy(e);
    if (copy_stream) cudaStreamDestroy(copy_stream);
    for (cudaEvent_t e : ev_k1) cudaEventDestroy(e);
    for (cudaEvent_t e : ev_rest) cudaEventDestroy(e);
    if (h_nco_pin) { cudaFreeHost(h_nco_pin); h_nco_pin = nullptr; }
    if (d_nco_ring) { cudaFree(d_nco_ring); d_nco_ring = nullptr; }
    for (cudaEvent_t& e : ev_nco) if (e) { cudaEventDestroy(e); e = nullptr; }
    if (own_stream && stream) cudaStreamDestroy(stream);
}

// size the per-call buffers for pushes of up to n_in_max (carried + new) samples per channel
int hbd_decoder::ensure_call_capacity(size_t n_in_max)
{
    if (n_in_max <= cap_n_in && d_s1x[0]) return HBD_OK;
    const size_t n = size_t(n_ch);
    const size_t grow_to = std::max(n_in_max, cap_n_in);
    const size_t n1_max = grow_to / size_t(M1) + 2, n2_max = grow_to / size_t(factor) + 2;
    HBD_CUDA_CHECK(grow_rows(&d_s1x[0], &s1_pitch, (kS1Hist + n1_max + 15) & ~size_t(15), n, kS1Hist, stream));
    HBD_CUDA_CHECK(grow_rows(&d_s1x[1], &s1_pitch_b, (kS1Hist + n1_max + 15) & ~size_t(15), n, kS1Hist, stream));
    for (size_t k = 0; k < mids.size(); ++k) {   // cascaded plan: the buffers between the stages
        const size_t n_out_max = grow_to / (size_t(mids[k].div_in) * size_t(mids[k].M)) + 2;
        if (k + 1 < mids.size()) HBD_CUDA_CHECK(grow_rows(&mids[k].d_out, &mids[k].pitch, (kMidHist + n_out_max + 15) & ~size_t(15), n, kMidHist, stream));
        else {
            HBD_CUDA_CHECK(grow_rows(&d_midlast[0], &midlast_pitch, (kS1Hist + n_out_max + 15) & ~size_t(15), n, kS1Hist, stream));
            HBD_CUDA_CHECK(grow_rows(&d_midlast[1], &midlast_pitch_b, (kS1Hist + n_out_max + 15) & ~size_t(15), n, kS1Hist, stream));
        }
    }
    HBD_CUDA_CHECK(grow_rows(&d_decq, &dq_pitch, (kLpHist + kLpBatch + n2_max + 15) & ~size_t(15), n, kLpHist + kLpBatch, stream));
    HBD_CUDA_CHECK(grow_rows(&d_slicer, &slicer_pitch, (size_t(kSlicerVent) + 1 + kLpBatch + n2_max + 15) & ~size_t(15), n,
                             slicer_pitch, stream));
    HBD_CUDA_CHECK(grow_rows(&d_demod, &demod_pitch, (kLpBatch + n2_max + 15) & ~size_t(15), n, 0, stream));
    if (record) {
        HBD_CUDA_CHECK(grow_rows(&d_rec_dec, &rec_pitch, (kLpBatch + n2_max + 15) & ~size_t(15), n, 0, stream));
        size_t p2 = 0;
        if (d_rec_filt) { cudaFree(d_rec_filt); d_rec_filt = nullptr; }
        HBD_CUDA_CHECK(grow_rows(&d_rec_filt, &p2, rec_pitch, n, 0, stream));
        if (!d_rec_bits) {
            rec_bits_pitch = 1u << 16;
            HBD_CUDA_CHECK(dalloc(&d_rec_bits, n * rec_bits_pitch));
            HBD_CUDA_CHECK(dalloc(&d_rec_bits_n, n));
            HBD_CUDA_CHECK(cudaMemset(d_rec_bits_n, 0, n * sizeof(unsigned)));
        }
    }
    cap_n_in = grow_to;
    return HBD_OK;
}

namespace {
// What one channel's share of a process() call amounts to on the host: the reference's buffer arithmetic
// (Decoder.h:426-436,492-542) stepped on the channel's counters.  No CUDA calls in here; the caller applies the
// side effects to that channel's rows -- or, in uniform mode, to the rows of every channel at once.
struct ChanStep {
    ChanPlan plan;
    bool zero_carry = false;       // Decimator.h:74-79: the stage's work buffer grew -> its history is zeroed
    bool zero_s1hist = false;      // ... of the LAST stage (the tail kernel's input buffer)
    unsigned zero_mid = 0;         // bit k: ... of middle stage k (cascaded plans)
    bool zero_lphist = false;      // FirFilter.h:141-147
    bool taps_changed = false;     // low-pass redesigned: upload `taps`
    size_t taps_old = 0, taps_new = 0;
    bool frame_done = false;       // this call completes an FFT frame (K4 runs)
    unsigned nf = 0;
};
} // namespace

// returns HBD_OK or HBD_ERR_ARG (nothing has been changed in that case)
static int step_channel(hbd_decoder* h, HostChan& x, unsigned pushed, ChanStep& o, std::vector<float>& taps)
{
    const int factor = h->factor, M1 = h->M1, T1 = h->T1, M2 = h->M2, T2 = h->T2, fft_n = h->fft_n;
    const double fs_dec = h->fs_in / factor;
    ChanPlan& p = o.plan;
    p.r = x.in_r; p.n = pushed;
    p.dec_pending = x.dec_pending; p.lp_ntaps = unsigned(x.lp_ntaps);
    const unsigned total = x.in_r + pushed;
    if (int(total) < factor) {            // Decoder.h:429-430
        p.consumed = 0; p.n1 = p.n2 = 0; p.flags = 1;
        x.in_r = total; x.pushed = 0; x.last_n2 = 0; x.last_nf = 0;
        return HBD_OK;
    }
    p.consumed = total - total % unsigned(factor);
    p.n1 = p.consumed / unsigned(M1);
    p.n2 = p.consumed / unsigned(factor);
    p.flags = 0;
    // ---- everything that can fail comes first: the low-pass design (pure) --------------------------------
    const unsigned total_dec = x.dec_pending + p.n2;
    const bool gated = total_dec < unsigned(kLpBatch);                 // Decoder.h:494-495
    const bool cut_off = !gated && fs_dec > 4 * 40e3;                  // Decoder.h:522-527
    unsigned nf = 0;
    size_t T = x.lp_ntaps;
    bool redesign = false;
    if (!gated && !cut_off) {
        nf = total_dec - total_dec % unsigned(kLpBatch);
        const bool dirty = x.lp_dirty || x.lp_input_size != nf;
        if (dirty) {                                                   // Decoder.h:536-538
            T = design_lowpass(float(x.lp_bw / fs_dec), x.lp_trans, nf, x.lp_ntaps, taps);
            if (T > size_t(kLpMaxTaps)) { h->set_error("low-pass needs more than kLpMaxTaps (1025) taps: raise lowpass_trans"); return HBD_ERR_ARG; }
            redesign = T != x.lp_ntaps;
        }
    }
    // ---- commit ----------------------------------------------------------------------------------------------
    x.in_r = total - p.consumed;
    x.pushed = 0;
    x.last_n2 = p.n2; x.last_nf = 0;
    if (p.n2 && x.fft_have < unsigned(fft_n)) {   // FFT frame assembly, Decoder.h:467-473 (mirror of K2's arithmetic)
        x.fft_have += std::min(unsigned(fft_n) - x.fft_have, p.n2);
        if (x.fft_have >= unsigned(fft_n)) { o.frame_done = true; x.fft_have = 0; }   // transformed and cleared in this call
    }
    if (M1 > 1) {   // history re-zeroing when the reference's work buffers grow (Decimator.h:74-79)
        const size_t need = size_t(p.consumed) + size_t(T1) + size_t(M1);
        if (x.grown1 < need) { x.grown1 = need; o.zero_carry = true; }
    }
    size_t n_in = p.n1;            // samples entering the next stage
    for (size_t k = 0; k < h->mids.size(); ++k) {
        const size_t need = n_in + size_t(h->mids[k].T) + size_t(h->mids[k].M);
        if (x.grown_mid[k] < need) { x.grown_mid[k] = need; o.zero_mid |= 1u << k; }
        n_in /= size_t(h->mids[k].M);
    }
    if (M2 > 1) {
        const size_t need = n_in + size_t(T2) + size_t(M2);
        if (x.grown2 < need) { x.grown2 = need; o.zero_s1hist = true; }
    }
    if (gated) { x.dec_pending = total_dec; return HBD_OK; }
    if (cut_off) { x.dec_pending = 0; return HBD_OK; }
    x.dec_pending = total_dec - nf;
    x.last_nf = nf; x.demod_n = nf; o.nf = nf;
    x.lp_input_size = nf; x.lp_dirty = false;
    if (redesign) {
        o.taps_changed = true; o.taps_old = x.lp_ntaps; o.taps_new = T;
        x.lp_ntaps = T; x.cfg_dirty = true; h->cfg_dirty_any = true; h->derived_dirty = true;
    }
    p.lp_ntaps = unsigned(x.lp_ntaps);
    const size_t need = size_t(nf) + x.lp_ntaps;
    if (x.grown_lp < need) { x.grown_lp = need; o.zero_lphist = true; }   // FirFilter.h:141-147
    return HBD_OK;
}

// Keep the character / SSDV logs from filling up: look at what the completed calls have appended (their committed
// heads sit in pinned host memory, so this costs an event query, not a copy) and drain early when the calls still in
// flight could exhaust the space that is left.
int hbd_decoder::log_pressure()
{
    while (done_seq != call_seq && cudaEventQuery(ev_call[done_seq % ev_call.size()]) == cudaSuccess) {
        const unsigned head = h_heads[size_t(done_seq % ev_call.size()) * kCtlWords + kCtlCharHead];
        per_call_max = std::max(per_call_max, head - done_head);
        done_head = head;
        ++done_seq;
    }
    const unsigned in_flight = call_seq - done_seq + 1u;     // the call about to be enqueued included
    const unsigned long long worst = (unsigned long long)(done_head - log_tail) + 2ull * in_flight * std::max(per_call_max, unsigned(n_ch));
    bool drain = worst > log_cap;
    if (ssdv_on && done_seq) {
        const unsigned sh = h_heads[size_t((done_seq - 1u) % ev_call.size()) * kCtlWords + kCtlSsdvHead];
        if (sh - ssdv_log_tail > kSsdvLogCap / 2u) drain = true;
    }
    if (call_seq - calls_collected >= ev_call.size() / 2u) drain = true;   // event / head slots are about to be reused
    if (!drain || call_seq == calls_collected) return HBD_OK;
    return collect_locked(call_seq - calls_collected > 2u ? 1u : 0u);
}

int hbd_decoder::process_async_locked()
{
    if (!fs_in) return HBD_OK; // Decoder.h:418-419: uninitialised -> silently nothing
    HBD_CUDA_CHECK(cudaSetDevice(device));
    { const int rc = log_pressure(); if (rc) return rc; }
    const size_t n = size_t(n_ch);
    const double fs_dec = fs_in / factor;
    bool groups_idle = false; // set once we had to wait for the compute streams (rare host-side state changes)
    auto quiesce = [&]() -> int { if (!groups_idle) { if (sync_groups()) return HBD_ERR_CUDA; groups_idle = true; } return HBD_OK; };

    // ---- plan ------------------------------------------------------------------------------------------------
    // uniform mode (whole-batch pushes only, one low-pass setting): channel 0 stands for all of them, O(1) per call
    const bool uni = ctr_uniform && lp_cfg_uniform;
    if (!uni) sync_ctrs();
    size_t max_total = 0; unsigned max_n1 = 0, max_nf = 0;
    bool any_work = false, need_k4 = false;
    std::vector<float> new_taps;
    ChanPlan uplan{};
    if (uni) {
        ChanStep st;
        const unsigned pushed = ext ? unsigned(ext_n) : hc[0].pushed;
        max_total = size_t(hc[0].in_r) + pushed;
        if (max_total > cap_n_in || !d_s1x[0]) { if (quiesce()) return HBD_ERR_CUDA; }   // buffers are about to be reallocated
        { const int rc = ensure_call_capacity(max_total); if (rc) return rc; }
        const int rc = step_channel(this, hc[0], pushed, st, new_taps);
        if (rc) return rc;   // nothing has been stepped
        ctr_stale = true;
        uplan = st.plan;
        any_work = !(st.plan.flags & 1u); need_k4 = st.frame_done; max_n1 = st.plan.n1; max_nf = st.nf;
        if (st.zero_carry || st.zero_s1hist || st.zero_mid || st.zero_lphist || st.taps_changed) {
            if (quiesce()) return HBD_ERR_CUDA;
            if (st.zero_carry)
                HBD_CUDA_CHECK(cudaMemset2DAsync(d_carry2[carry_cur] + (carry_cap - (T1 - 1) - int(st.plan.r)), size_t(carry_cap) * sizeof(float2), 0,
                                                 sizeof(float2) * size_t(T1 - 1), n, stream));
            for (size_t k = 0; k < mids.size(); ++k)
                if (st.zero_mid & (1u << k)) {
                    if (k == 0) HBD_CUDA_CHECK(cudaMemset2DAsync(d_s1x[s1_cur], s1_pitch * sizeof(float2), 0, sizeof(float2) * kS1Hist, n, stream));
                    else HBD_CUDA_CHECK(cudaMemset2DAsync(mids[k - 1].d_out, mids[k - 1].pitch * sizeof(float2), 0, sizeof(float2) * kMidHist, n, stream));
                }
            if (st.zero_s1hist) {
                if (mids.empty()) HBD_CUDA_CHECK(cudaMemset2DAsync(d_s1x[s1_cur], s1_pitch * sizeof(float2), 0, sizeof(float2) * kS1Hist, n, stream));
                else HBD_CUDA_CHECK(cudaMemset2DAsync(d_midlast[s1_cur], midlast_pitch * sizeof(float2), 0, sizeof(float2) * kS1Hist, n, stream));
            }
            if (st.taps_changed) {
                std::vector<float> all(n * st.taps_new);
                for (size_t c = 0; c < n; ++c) memcpy(all.data() + c * st.taps_new, new_taps.data(), 4 * st.taps_new);
                HBD_CUDA_CHECK(cudaMemcpy2DAsync(d_lptaps, kLpMaxTaps * sizeof(float), all.data(), st.taps_new * sizeof(float), st.taps_new * sizeof(float), n,
                                                 cudaMemcpyHostToDevice, stream));
                HBD_CUDA_CHECK(cudaStreamSynchronize(stream)); // `all` goes out of scope
                for (size_t c = 1; c < n; ++c) hc[c].cfg_dirty = true;   // lp_ntaps travels with the configuration upload
            }
            if (st.zero_lphist) HBD_CUDA_CHECK(cudaMemset2DAsync(d_decq, dq_pitch * sizeof(float2), 0, sizeof(float2) * std::min<size_t>(hc[0].lp_ntaps, kLpHist), n, stream));
        }
    } else {
        h_plan.resize(n);
        for (size_t c = 0; c < n; ++c) max_total = std::max<size_t>(max_total, size_t(hc[c].in_r) + (ext ? unsigned(ext_n) : hc[c].pushed));
        if (max_total > cap_n_in || !d_s1x[0]) { if (quiesce()) return HBD_ERR_CUDA; }
        { const int rc = ensure_call_capacity(max_total); if (rc) return rc; }
        for (size_t c = 0; c < n; ++c) {
            HostChan& x = hc[c];
            ChanStep st;
            const int rc = step_channel(this, x, ext ? unsigned(ext_n) : x.pushed, st, new_taps);
            if (rc) return rc;   // channels below c have been stepped: their samples are gone (documented in the header)
            h_plan[c] = st.plan;
            any_work |= !(st.plan.flags & 1u); need_k4 |= st.frame_done;
            max_n1 = std::max(max_n1, st.plan.n1); max_nf = std::max(max_nf, st.nf);
            if (!(st.zero_carry || st.zero_s1hist || st.zero_mid || st.zero_lphist || st.taps_changed)) continue;
            if (quiesce()) return HBD_ERR_CUDA;
            if (st.zero_carry) HBD_CUDA_CHECK(cudaMemsetAsync(d_carry2[carry_cur] + c * size_t(carry_cap) + (carry_cap - (T1 - 1) - int(st.plan.r)), 0, sizeof(float2) * size_t(T1 - 1), stream));
            for (size_t k = 0; k < mids.size(); ++k)
                if (st.zero_mid & (1u << k)) {
                    if (k == 0) HBD_CUDA_CHECK(cudaMemsetAsync(d_s1x[s1_cur] + c * s1_pitch, 0, sizeof(float2) * kS1Hist, stream));
                    else HBD_CUDA_CHECK(cudaMemsetAsync(mids[k - 1].d_out + c * mids[k - 1].pitch, 0, sizeof(float2) * kMidHist, stream));
                }
            if (st.zero_s1hist) {
                if (mids.empty()) HBD_CUDA_CHECK(cudaMemsetAsync(d_s1x[s1_cur] + c * s1_pitch, 0, sizeof(float2) * kS1Hist, stream));
                else HBD_CUDA_CHECK(cudaMemsetAsync(d_midlast[s1_cur] + c * midlast_pitch, 0, sizeof(float2) * kS1Hist, stream));
            }
            if (st.taps_changed) {
                HBD_CUDA_CHECK(cudaMemcpyAsync(d_lptaps + c * kLpMaxTaps, new_taps.data(), 4 * st.taps_new, cudaMemcpyHostToDevice, stream));
                HBD_CUDA_CHECK(cudaStreamSynchronize(stream)); // new_taps is reused
            }
            if (st.zero_lphist) HBD_CUDA_CHECK(cudaMemsetAsync(d_decq + c * dq_pitch, 0, sizeof(float2) * std::min<size_t>(x.lp_ntaps, kLpHist), stream));
        }
        // back to uniform mode as soon as the channels agree again (e.g. after a round of per-channel pushes of one size)
        bool same = true;
        for (size_t c = 1; c < n && same; ++c) same = hc[c].same(hc[0]);
        ctr_uniform = same;
    }
    if (demod_acc_on) { const int rc = ensure_dacc(); if (rc) return rc; }

    if (cfg_dirty_any) {
        if (quiesce()) return HBD_ERR_CUDA;
        sync_ctrs();
        std::vector<double> b(n); std::vector<float> s(n); std::vector<int> bi(n), dc(n), nt(n); std::vector<unsigned char> d(n);
        for (size_t c = 0; c < n; ++c) {
            b[c] = hc[c].baud; s[c] = hc[c].rtty_stops; bi[c] = int(hc[c].rtty_bits); dc[c] = hc[c].dc_remove; nt[c] = int(hc[c].lp_ntaps);
            d[c] = hc[c].cfg_dirty; hc[c].cfg_dirty = false;
        }
        HBD_CUDA_CHECK(cudaMemcpyAsync(d_cfg_baud, b.data(), 8 * n, cudaMemcpyHostToDevice, stream));
        HBD_CUDA_CHECK(cudaMemcpyAsync(d_cfg_stops, s.data(), 4 * n, cudaMemcpyHostToDevice, stream));
        HBD_CUDA_CHECK(cudaMemcpyAsync(d_cfg_bits, bi.data(), 4 * n, cudaMemcpyHostToDevice, stream));
        HBD_CUDA_CHECK(cudaMemcpyAsync(d_cfg_dc, dc.data(), 4 * n, cudaMemcpyHostToDevice, stream));
        HBD_CUDA_CHECK(cudaMemcpyAsync(d_cfg_ntaps, nt.data(), 4 * n, cudaMemcpyHostToDevice, stream));
        HBD_CUDA_CHECK(cudaMemcpyAsync(d_cfg_dirty, d.data(), n, cudaMemcpyHostToDevice, stream));
        init_cfg_kernel<<<(n_ch + 127) / 128, 128, 0, stream>>>(d_state, d_cfg_baud, d_cfg_stops, d_cfg_bits, d_cfg_dc, d_cfg_ntaps, d_cfg_dirty, n_ch);
        ++launches;
        HBD_CUDA_CHECK(cudaStreamSynchronize(stream)); // host vectors go out of scope
        cfg_dirty_any = false;
    }
    bool plan_uniform = uni;
    if (!uni) {
        plan_uniform = true;
        for (size_t c = 1; c < n && plan_uniform; ++c) plan_uniform = memcmp(&h_plan[c], &h_plan[0], sizeof(ChanPlan)) == 0;
        uplan = h_plan[0];
        if (!plan_uniform && (h_plan_uploaded.size() != n || memcmp(h_plan.data(), h_plan_uploaded.data(), n * sizeof(ChanPlan)) != 0)) {
            if (quiesce()) return HBD_ERR_CUDA;
            HBD_CUDA_CHECK(cudaMemcpyAsync(d_plan, h_plan.data(), n * sizeof(ChanPlan), cudaMemcpyHostToDevice, stream));
            HBD_CUDA_CHECK(cudaStreamSynchronize(stream));
            h_plan_uploaded = h_plan;
        }
    }

    // what the tail kernel stages in shared memory: sized from host-side knowledge of all channels
    if (derived_dirty) {
        sync_ctrs();
        max_lp_taps_c = 1; max_spb_c = 1;
        for (size_t c = 0; c < n; ++c) {
            max_lp_taps_c = std::max(max_lp_taps_c, hc[c].lp_ntaps);
            if (hc[c].baud > 0) max_spb_c = std::max(max_spb_c, fs_dec / hc[c].baud);
        }
        derived_dirty = false;
    }
    // a channel normally holds < ~12 symbols of pending samples after a slicer pass (plus this call's batch)
    int sv_want = int(std::min<double>(6144.0, 16.0 * max_spb_c + double(max_nf) + 64.0));
    if (sv_override >= 0) sv_want = sv_override;   // test hook (HBD_SV_WANT): 0 forces the slicer's HBM path

    const float2* chunk = ext ? ext : d_stage;
    const size_t chunk_pitch = ext ? ext_pitch : stage_pitch;
    int nl = 0;
    // everything the caller (and the host-side updates above) put on `stream` happens before the kernels start
    HBD_CUDA_CHECK(cudaEventRecord(ev_in, stream));
    HBD_CUDA_CHECK(cudaStreamWaitEvent(hi, ev_in, 0));
    bool tail_recorded = false;
    // K1 overwrites the stage-1 buffer: the tail of the call before the previous one must be done with it
    if (tail_pending[s1_cur]) HBD_CUDA_CHECK(cudaStreamWaitEvent(hi, ev_tail[s1_cur], 0));
    {   // K1 also writes the next call's carry (even when no channel has a full decimation block yet)
        DecimArgs da{};
        da.chunk = chunk; da.chunk_pitch = chunk_pitch; da.carry = d_carry2[carry_cur]; da.carry_next = d_carry2[carry_cur ^ 1]; da.carry_cap = carry_cap;
        da.s1 = d_s1x[s1_cur]; da.s1_pitch = s1_pitch; da.s1_hist = kS1Hist;
        da.plan = d_plan; da.uplan = uplan; da.uniform = plan_uniform ? 1 : 0; da.taps = d_taps1; da.ch0 = 0; da.n_channels = n_ch;
        da.nco = ext_nco ? ext_nco_ptr : nullptr;
        if (timing) HBD_CUDA_CHECK(cudaEventRecord(next_event(ev_k1, ev_used_k1), hi));
        HBD_CUDA_CHECK(launch_decim1(da, M1, T1, max_n1, n_sms, hi, &nl));
        if (timing) HBD_CUDA_CHECK(cudaEventRecord(next_event(ev_k1, ev_used_k1), hi));
    }
    HBD_CUDA_CHECK(cudaEventRecord(ev_consumed, hi)); // input no longer needed; stage-1 output ready
    HBD_CUDA_CHECK(cudaStreamWaitEvent(lo, ev_consumed, 0));
    if (timing > 1) HBD_CUDA_CHECK(cudaEventRecord(next_event(ev_rest, ev_used_rest), lo));
    if (any_work) {
        // cascaded plan: the stages between K1's and the tail kernel's (1/64 of the input rate or less)
        for (size_t k = 0; k < mids.size(); ++k) {
            MidArgs ma{};
            ma.plan = d_plan; ma.uplan = uplan; ma.uniform = plan_uniform ? 1 : 0;
            if (k == 0) { ma.in = d_s1x[s1_cur]; ma.in_next = d_s1x[s1_cur ^ 1]; ma.in_pitch = s1_pitch; ma.in_hist = kS1Hist; }
            else { ma.in = mids[k - 1].d_out; ma.in_next = mids[k - 1].d_out; ma.in_pitch = mids[k - 1].pitch; ma.in_hist = kMidHist; }
            if (k + 1 < mids.size()) { ma.out = mids[k].d_out; ma.out_pitch = mids[k].pitch; ma.out_hist = kMidHist; }
            else { ma.out = d_midlast[s1_cur]; ma.out_pitch = midlast_pitch; ma.out_hist = kS1Hist; }
            ma.taps = mids[k].d_taps; ma.M = mids[k].M; ma.T = mids[k].T; ma.div_in = mids[k].div_in;
            HBD_CUDA_CHECK(launch_mid_stage(ma, n_ch, lo, &nl));
        }
        TailArgs ta{};
        ta.plan = d_plan; ta.uplan = uplan; ta.uniform = plan_uniform ? 1 : 0; ta.state = d_state; ta.ch0 = 0;
        if (mids.empty()) { ta.s1 = d_s1x[s1_cur]; ta.s1_next = d_s1x[s1_cur ^ 1]; ta.s1_pitch = s1_pitch; }
        else { ta.s1 = d_midlast[s1_cur]; ta.s1_next = d_midlast[s1_cur ^ 1]; ta.s1_pitch = midlast_pitch; }
        ta.taps2 = d_taps2; ta.M2 = M2; ta.T2 = T2;
        ta.decq = d_decq; ta.dq_pitch = dq_pitch; ta.fs_dec = fs_dec; ta.fftbuf = d_fftbuf; ta.fft_n = fft_n; ta.lptaps = d_lptaps;
        ta.max_lp_taps = int(max_lp_taps_c);
        ta.slicer = d_slicer; ta.slicer_pitch = slicer_pitch; ta.sv_want = sv_want;
        ta.log = d_log; ta.log_ctl = d_log_ctl; ta.log_mask = log_cap - 1u; ta.call_seq = call_seq & 0xffffffu;
        ta.uart_runs = d_uart_runs;
        ta.maskc = mask_cache ? d_maskc : nullptr;
        ta.ssdv_ring = ssdv_on ? d_ssdv_ring : nullptr; ta.ssdv_total = d_ssdv_total;
        ta.demod_last = d_demod; ta.demod_pitch = demod_pitch;
        ta.rec_decimated = record ? d_rec_dec : nullptr; ta.rec_filtered = record ? d_rec_filt : nullptr; ta.rec_pitch = rec_pitch;
        ta.rec_bits = record ? d_rec_bits : nullptr; ta.rec_bits_n = d_rec_bits_n; ta.rec_bits_pitch = rec_bits_pitch;
        HBD_CUDA_CHECK(launch_tail(ta, n_ch, lo, &nl));
        // only the tail kernel reads the stage-1 buffer: K1 of the call after next may overwrite it as soon as the
        // tail is done, without waiting for the FFT / SSDV / demod kernels behind it on this stream
        if (!tail_ev_late) { HBD_CUDA_CHECK(cudaEventRecord(ev_tail[s1_cur], lo)); tail_recorded = true; }
        FftArgs fa{};
        fa.state = d_state; fa.fftbuf = d_fftbuf; fa.spectrum = d_spectrum; fa.power = d_power; fa.twiddle = d_twiddle; fa.fs_dec = fs_dec;
        fa.ch0 = 0; fa.fft_n = fft_n;
        if (need_k4) HBD_CUDA_CHECK(launch_fft_afc(fa, n_ch, lo, &nl));   // other calls: K2 has stepped the AFC itself
        if (ssdv_on) {   // test every 0x55-started window the new characters completed
            SsdvScanArgs sa{};
            sa.ring = d_ssdv_ring; sa.total = d_ssdv_total; sa.scanned = d_ssdv_scanned; sa.log = d_ssdv_log;
            sa.ctl = d_log_ctl; sa.call_seq = call_seq & 0xffffffu; sa.ch0 = 0;
            HBD_CUDA_CHECK(launch_ssdv_scan(sa, n_ch, lo, &nl));
        }
    }
    if (demod_acc_on && d_demod) { // main.cpp:267-282 runs after EVERY process(), also when nothing new was demodulated
        DemodAccArgs aa{};
        aa.state = d_state; aa.demod = d_demod; aa.demod_pitch = demod_pitch; aa.acc = d_dacc; aa.acc_pitch = dacc_pitch; aa.acc_n = d_dacc_n;
        aa.fs_dec = fs_dec; aa.ch0 = 0;
        HBD_CUDA_CHECK(launch_demod_accumulate(aa, n_ch, lo, &nl));
    }
    if (timing > 1) HBD_CUDA_CHECK(cudaEventRecord(next_event(ev_rest, ev_used_rest), lo));
    if (!tail_recorded) HBD_CUDA_CHECK(cudaEventRecord(ev_tail[s1_cur], lo));
    tail_pending[s1_cur] = true;
    // the caller's stream resumes once the input has been consumed (it does not wait for the tail kernels)
    HBD_CUDA_CHECK(cudaStreamWaitEvent(stream, ev_consumed, 0));
    carry_cur ^= 1;
    if (any_work) s1_cur ^= 1; // the tail (which moves the stage-2 history to the other buffer) ran
    launches += unsigned(nl);
    if (snap_on) {
        stats_snap_kernel<<<(n_ch + 127) / 128, 128, 0, lo>>>(d_state, d_snap + size_t(call_seq % kSnapSlots) * n * 6, n_ch);
        ++launches;
    }
    // commit: the log control words as they stand after this call, into the call's pinned slot (stream order: every
    // entry below these heads is completely written when ev_call fires)
    const size_t slot = call_seq % ev_call.size();
    HBD_CUDA_CHECK(cudaMemcpyAsync(h_heads + slot * kCtlWords, d_log_ctl, kCtlWords * sizeof(unsigned), cudaMemcpyDeviceToHost, lo));
    HBD_CUDA_CHECK(cudaEventRecord(ev_call[slot], lo)); // everything of this call is done
    ++call_seq;
    ++pending_marks;
    ext = nullptr; ext_n = 0; ext_nco = false;
    return HBD_OK;
}

// Drain decoded characters of all calls up to (last enqueued - lag) and run the sentence layer on them, call by
// call like Decoder::process().  lag == 0 waits for everything; lag > 0 leaves the newest calls in flight so the GPU
// keeps working while the host is busy here.  User callbacks are only RECORDED here (h->deferred).
int hbd_decoder::collect_locked(unsigned lag)
{
    HBD_CUDA_CHECK(cudaSetDevice(device));
    if (call_seq == calls_collected) {
        if (lag == 0) { if (sync_groups()) return HBD_ERR_CUDA; HBD_CUDA_CHECK(cudaStreamSynchronize(stream)); }
        return HBD_OK;
    }
    if (call_seq - calls_collected <= lag) return HBD_OK;
    if (lag == 0 && call_seq - calls_collected > 2) {
        // a full drain behind a queue of calls: take the finished calls first, in halving steps, so the host's sentence
        // layer works while the GPU is still busy with the newest calls instead of after it has gone idle
        int first_rc = HBD_OK;
        for (unsigned l = (call_seq - calls_collected) / 2; l >= 1; l /= 2) { const int rc = collect_locked(l); if (rc == HBD_ERR_CUDA) return rc; if (rc && !first_rc) first_rc = rc; }
        const int rc = collect_locked(0);
        return rc ? rc : first_rc;
    }
    const unsigned upto = call_seq - lag;            // calls [calls_collected, upto) get drained
    if (lag == 0) {
        if (sync_groups()) { set_error("stream sync failed"); return HBD_ERR_CUDA; }
        HBD_CUDA_CHECK(cudaStreamSynchronize(stream));
    } else {
        HBD_CUDA_CHECK(cudaEventSynchronize(ev_call[(upto - 1) % ev_call.size()]));
    }
    // the COMMITTED control words of call upto-1 (never the live head: newer calls may be appending right now)
    unsigned heads[kCtlWords];
    memcpy(heads, h_heads + size_t((upto - 1) % ev_call.size()) * kCtlWords, sizeof(heads));
    const unsigned head = heads[kCtlCharHead];
    const unsigned upto24 = upto & 0xffffffu;
    int result = HBD_OK;
    if (ssdv_on) {
        // accepted SSDV packets of the calls being drained go to their channels first: the buffer automaton below
        // looks a window's verdict up when it reaches it, which is never before the call that completed the window
        // the overflow words are monotonic counters: a change since the previous drain is reported once, decoding goes on
        if (heads[kCtlSsdvRingOvf] != ovf_seen[0]) {
            ovf_seen[0] = heads[kCtlSsdvRingOvf];
            set_error("SSDV character ring overflow: more than 3840 characters in one call (windows lost); push smaller chunks");
            result = HBD_ERR_STATE;
        }
        if (heads[kCtlSsdvOvf] != ovf_seen[1]) {
            set_error("SSDV packet log overflow: " + std::to_string(heads[kCtlSsdvOvf] - ovf_seen[1]) + " packet(s) lost; collect more often");
            ovf_seen[1] = heads[kCtlSsdvOvf];
            result = HBD_ERR_STATE;
        }
        const unsigned avail_p = std::min(heads[kCtlSsdvHead] - ssdv_log_tail, kSsdvLogCap);
        h_ssdv_log.resize(avail_p);
        if (avail_p) {
            const unsigned i0 = ssdv_log_tail & (kSsdvLogCap - 1u);
            const unsigned first = std::min(avail_p, kSsdvLogCap - i0);
            HBD_CUDA_CHECK(cudaMemcpyAsync(h_ssdv_log.data(), d_ssdv_log + i0, size_t(first) * sizeof(SsdvLogEntry), cudaMemcpyDeviceToHost, copy_stream));
            if (avail_p > first)
                HBD_CUDA_CHECK(cudaMemcpyAsync(h_ssdv_log.data() + first, d_ssdv_log, size_t(avail_p - first) * sizeof(SsdvLogEntry), cudaMemcpyDeviceToHost, copy_stream));
            HBD_CUDA_CHECK(cudaStreamSynchronize(copy_stream));
        }
        for (unsigned k = 0; k < avail_p; ++k) {
            const SsdvLogEntry& e = h_ssdv_log[k];
            if (e.ch >= unsigned(n_ch)) continue;
            SsdvVerdict v; v.pos = e.pos; v.errors = e.errors; memcpy(v.data.data(), e.data, 256);
            ssdv[e.ch].verdicts.push_back(v);
        }
        ssdv_log_tail = heads[kCtlSsdvHead];
    }
    if (heads[kCtlCharOvf] != ovf_seen[2]) {
        // a writer found the ring full and dropped its characters (it never overwrites unread entries).  Which slots of
        // the range were skipped is not recorded, so the whole range is given up; decoding continues behind it.
        ovf_seen[2] = heads[kCtlCharOvf];
        chars_lost += head - log_tail;
        set_error("decoded-character log overflow: " + std::to_string(head - log_tail) + " characters of calls [" + std::to_string(calls_collected) + ", " +
                  std::to_string(upto) + ") lost; collect more often");
        log_tail = head;
        result = HBD_ERR_STATE;
    }
    if (snap_on && call_seq - (upto - 1u) < kSnapSlots - 1u) {   // the slot of call upto-1 has not been reused yet
        HBD_CUDA_CHECK(cudaMemcpyAsync(h_snap, d_snap + size_t((upto - 1u) % kSnapSlots) * size_t(n_ch) * 6, sizeof(float) * 6 * size_t(n_ch),
                                       cudaMemcpyDeviceToHost, copy_stream));
        h_snap_valid = true;
    }
    const unsigned avail = head - log_tail;
    h_log.resize(avail);
    if (avail) {
        const unsigned i0 = log_tail & (log_cap - 1u);
        const unsigned first = std::min(avail, log_cap - i0);
        HBD_CUDA_CHECK(cudaMemcpyAsync(h_log.data(), d_log + i0, size_t(first) * sizeof(uint2), cudaMemcpyDeviceToHost, copy_stream));
        if (avail > first)
            HBD_CUDA_CHECK(cudaMemcpyAsync(h_log.data() + first, d_log, size_t(avail - first) * sizeof(uint2), cudaMemcpyDeviceToHost, copy_stream));
        HBD_CUDA_CHECK(cudaStreamSynchronize(copy_stream));
    }
    // the writers may reuse everything below the new tail
    log_tail = head;
    h_tail_pin[0] = log_tail; h_tail_pin[1] = ssdv_log_tail;
    HBD_CUDA_CHECK(cudaMemcpyAsync(d_log_ctl + kCtlCharTail, h_tail_pin, sizeof(unsigned), cudaMemcpyHostToDevice, copy_stream));
    HBD_CUDA_CHECK(cudaMemcpyAsync(d_log_ctl + kCtlSsdvTail, h_tail_pin + 1, sizeof(unsigned), cudaMemcpyHostToDevice, copy_stream));
    HBD_CUDA_CHECK(cudaStreamSynchronize(copy_stream));   // h_tail_pin is rewritten by the next drain

    // The log is sorted by call, and within a call a channel's characters form ONE contiguous segment (one reservation per
    // flush; only a channel that decodes more than kCharBuf = 64 characters in a call flushes twice).  Replay segment by
    // segment: one TextChannel::feed per channel and call, like one Decoder::process().
    const auto t_replay0 = std::chrono::steady_clock::now();
    const size_t n_log = h_log.size();
    seg_start.clear();
    for (size_t i = 0; i < n_log;) {           // segment boundaries: one sequential pass over the copied log
        size_t j = i + 1;
        while (j < n_log && h_log[j].x == h_log[i].x && (h_log[j].y >> 8) == (h_log[i].y >> 8)) ++j;
        seg_start.push_back(unsigned(i));
        i = j;
    }
    seg_start.push_back(unsigned(n_log));
    const size_t n_seg = seg_start.size() - 1;
    // Channels are independent, so the replay is cut by CHANNEL RANGE over a few host threads (the caller's included):
    // every thread walks all segments in log order and takes the ones of its channels -- per channel the order of the
    // calls is kept, and the callbacks the threads record are merged back into log order afterwards.
    const bool want_sent = sentence_cb || tracker;
    int n_thr = std::max(1, std::min(host_threads, int(n_seg / 1024)));
    n_thr = std::min(n_thr, n_ch);
    std::vector<ReplayPart> parts;
    parts.resize(size_t(n_thr));
    auto replay = [&](int t) {
        ReplayPart& part = parts[size_t(t)];
        const int c_lo = int((long long)n_ch * t / n_thr), c_hi = int((long long)n_ch * (t + 1) / n_thr);
        unsigned cur_seg = 0;
        SentenceSink sink;
        if (want_sent)
            sink = [&part, &cur_seg](int ch, const std::string& cs, const std::string& d, const std::string& crc) {
                Deferred ev; ev.kind = Deferred::kSentence; ev.ch = ch; ev.a = cs; ev.b = d; ev.c = crc;
                part.events.emplace_back(cur_seg, std::move(ev));
            };
        auto feed_channel = [&](int ch, const unsigned char* chars, size_t n_chars) {
            TextChannel& tc = text[size_t(ch)];
            hbd_result_record& pr = pend[size_t(ch)];
            const size_t before = chars_cb ? tc.chars_size(pr) : 0;
            tc.feed(chars, n_chars, ch, sink, keep_raw, pr);
            if (!tc.chars_spill.empty() || !tc.sent_spill.empty()) part.spilled = true;
            if (chars_cb && tc.chars_size(pr) > before) {   // character_callback_, Decoder.h:617-629 (not paced by wall clock)
                Deferred ev; ev.kind = Deferred::kChars; ev.ch = ch; ev.a = tc.chars_from(pr, before);
                part.events.emplace_back(cur_seg, std::move(ev));
            }
            if (ssdv_on) {   // Decoder.h:573: one SSDV_wraper_t::push per call that decoded characters
                SsdvEvent sev;
                if (ssdv[size_t(ch)].push(chars, n_chars, sev)) {
                    ssdv[size_t(ch)].events_pending.push_back(sev);
                    if (ssdv_cb) {   // ssdv_callback_, Decoder.h:631-632
                        Deferred ev; ev.kind = Deferred::kSsdv; ev.ch = ch; fill_ssdv_info(ev.info, sev); ev.pkt = sev.data;
                        part.events.emplace_back(cur_seg, std::move(ev));
                    }
                }
            }
        };
        std::vector<int> split;                    // channels of this range whose characters of the current call span several segments
        unsigned char seg_chars[kCharBufHost];
        // the per-channel state is scattered over the heap: a short look-ahead queue of this range's next segments feeds prefetches
        size_t look = 0;
        unsigned ahead[8]; int n_ahead = 0, head = 0;
        auto refill = [&] {
            while (n_ahead < 8 && look < n_seg) {
                const unsigned c = h_log[seg_start[look]].x;
                if (c >= unsigned(c_lo) && c < unsigned(c_hi)) {
                    ahead[(head + n_ahead) & 7] = unsigned(look); ++n_ahead;
                    __builtin_prefetch(&text[c]); __builtin_prefetch(&pend[c]);
                }
                ++look;
            }
        };
        refill();
        while (n_ahead) {
            const size_t k = ahead[head]; head = (head + 1) & 7; --n_ahead;
            refill();
            if (n_ahead >= 4) text[h_log[seg_start[ahead[(head + 3) & 7]]].x].prefetch_tails();
            cur_seg = unsigned(k);
            const size_t i0 = seg_start[k], i1 = seg_start[k + 1];
            const int ch = int(h_log[i0].x);
            const unsigned seq = h_log[i0].y >> 8;
            const size_t len = i1 - i0;
            // a full buffer means the channel flushed more than once in this call: join the pieces (rare: > 64 characters of
            // one channel in one call); they need not be adjacent in the log
            bool joined = len >= size_t(kCharBufHost);
            if (!joined && !split.empty()) joined = std::find(split.begin(), split.end(), ch) != split.end();
            if (joined) {
                if (call_chars[size_t(ch)].empty()) split.push_back(ch);
                for (size_t i = i0; i < i1; ++i) call_chars[size_t(ch)].push_back(char(h_log[i].y & 0xffu));
            } else {
                for (size_t i = 0; i < len; ++i) seg_chars[i] = (unsigned char)(h_log[i0 + i].y & 0xffu);
                feed_channel(ch, seg_chars, len);
            }
            // end of this call's entries (for this range): feed the joined channels
            if (!split.empty() && (!n_ahead || (h_log[seg_start[ahead[head]]].y >> 8) != seq)) {
                for (int sc : split) {
                    std::string& cc = call_chars[size_t(sc)];
                    feed_channel(sc, reinterpret_cast<const unsigned char*>(cc.data()), cc.size());
                    cc.clear();
                }
                split.clear();
            }
        }
    };
    pool.run(n_thr, replay);     // part t on worker t (it finds its channels' text state in its cache), late workers' parts on this thread
    {   // callbacks back into log order (stable: a channel's events keep the order they were recorded in)
        size_t total = 0;
        for (const ReplayPart& pt : parts) { total += pt.events.size(); if (pt.spilled) spill_free = false; }
        if (total) {
            std::vector<std::pair<unsigned, Deferred>> all;
            all.reserve(total);
            for (ReplayPart& pt : parts) for (auto& e : pt.events) all.push_back(std::move(e));
            if (n_thr > 1) std::stable_sort(all.begin(), all.end(), [](const auto& x, const auto& y) { return x.first < y.first; });
            for (auto& e : all) deferred.push_back(std::move(e.second));
        }
    }
    (void)upto24;
    drain_host_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_replay0).count();
    drain_calls += upto - calls_collected;
    calls_collected = upto;
    pending_marks = int(call_seq - calls_collected);
    return result;
}

// SSDV packet sync on / off.  Switching it on starts from an empty character stream (like a freshly constructed
// SSDV_wraper_t): every call in flight is drained first, so device ring, scan position and host automaton agree.
int hbd_decoder::enable_ssdv(bool on)
{
    if (on == ssdv_on) return HBD_OK;
    HBD_CUDA_CHECK(cudaSetDevice(device));
    { const int rc = collect_locked(0); if (rc) return rc; }
    const size_t n = size_t(n_ch);
    if (on) {
        if (!d_ssdv_ring) {
            HBD_CUDA_CHECK(dalloc(&d_ssdv_ring, n * kSsdvRing));
            HBD_CUDA_CHECK(dalloc(&d_ssdv_total, n));
            HBD_CUDA_CHECK(dalloc(&d_ssdv_scanned, n));
            HBD_CUDA_CHECK(dalloc(&d_ssdv_log, size_t(kSsdvLogCap)));
        }
        HBD_CUDA_CHECK(cudaMemset(d_ssdv_ring, 0, n * kSsdvRing));
        HBD_CUDA_CHECK(cudaMemset(d_ssdv_total, 0, n * sizeof(unsigned)));
        HBD_CUDA_CHECK(cudaMemset(d_ssdv_scanned, 0, n * sizeof(unsigned)));
        // the SSDV log keeps its (monotonic) head: nothing below it is of interest any more
        ssdv_log_tail = h_heads[size_t((call_seq ? call_seq - 1u : 0u) % ev_call.size()) * kCtlWords + kCtlSsdvHead];
        h_tail_pin[1] = ssdv_log_tail;
        HBD_CUDA_CHECK(cudaMemcpy(d_log_ctl + kCtlSsdvTail, h_tail_pin + 1, sizeof(unsigned), cudaMemcpyHostToDevice));
        ssdv.assign(n, SsdvChannel());
    }
    ssdv_on = on;
    return HBD_OK;
}

int hbd_decoder::enable_snap()
{
    if (snap_on) return HBD_OK;
    HBD_CUDA_CHECK(cudaSetDevice(device));
    if (!d_snap) {
        HBD_CUDA_CHECK(dalloc(&d_snap, size_t(kSnapSlots) * size_t(n_ch) * 6));
        HBD_CUDA_CHECK(cudaMemset(d_snap, 0, sizeof(float) * size_t(kSnapSlots) * size_t(n_ch) * 6));
        HBD_CUDA_CHECK(cudaHostAlloc((void**)&h_snap, sizeof(float) * 6 * size_t(n_ch), cudaHostAllocDefault));
        memset(h_snap, 0, sizeof(float) * 6 * size_t(n_ch));
    }
    snap_on = true;
    return HBD_OK;
}

namespace hbd {
int internal_device(hbd_decoder* h) { return h->device; }
void internal_set_error(hbd_decoder* h, const std::string& what) { std::lock_guard<std::mutex> l(h->mtx); h->set_error(what); }
void** internal_dist_slot(hbd_decoder* h) { return &h->dist_ctx; }
}

// =====================================================================================================
extern "C" {

// One record per channel with what the sentence layer has produced since the previous pack (printable characters, CRC-valid
// sentences; what does not fit a record stays for the next one) and the AFC scalars as of the newest drained call.
size_t hbd_pack_results(hbd_decoder* h, int ch_offset, hbd_result_record* out, size_t cap_records)
{
    if (!h) return 0;
    std::lock_guard<std::mutex> l(h->mtx);
    const size_t n = size_t(h->n_ch);
    if (!out || cap_records < n) return n;
    std::vector<float> live;     // AFC scalars when no snapshot has been taken yet
    const float* st = h->h_snap;
    if (!(h->snap_on && h->h_snap_valid)) {   // read the live state (waits for the calls in flight)
        if (h->enable_snap() != HBD_OK) return 0;
        cudaSetDevice(h->device);
        h->sync_groups();
        cudaStreamSynchronize(h->stream);
        std::vector<ChanState> cs(n);
        if (cudaMemcpy(cs.data(), h->d_state, n * sizeof(ChanState), cudaMemcpyDeviceToHost) != cudaSuccess) return 0;
        live.resize(6 * n);
        for (size_t c = 0; c < n; ++c) {
            live[6 * c + 0] = float(cs[c].afc_correction); live[6 * c + 1] = float(cs[c].afc_shift_hz); live[6 * c + 2] = float(cs[c].afc_noise_floor);
            live[6 * c + 3] = float(cs[c].afc_noise_var); live[6 * c + 4] = float(cs[c].gui_left); live[6 * c + 5] = float(cs[c].gui_right);
        }
        st = live.data();
    }
    // the pending slots ARE the records: copy, stamp, reset (and move spilled bytes up).  Cut by channel range like the
    // replay of the drain, so that a range is packed by the thread that filled it.
    const bool spill_free = h->spill_free;
    const int n_thr = int(std::max<size_t>(1, std::min<size_t>(size_t(h->host_threads), n / 1024)));
    std::vector<unsigned char> left(size_t(n_thr), 0);
    h->pool.run(n_thr, [&](int t) {
        const size_t c_lo = n * size_t(t) / size_t(n_thr), c_hi = n * size_t(t + 1) / size_t(n_thr);
        bool any_left = false;
        for (size_t c = c_lo; c < c_hi; ++c) {
            hbd_result_record& pr = h->pend[c];
            hbd_result_record& o = out[c];
            // header + the bytes in use only (the rest of a record is not read by anybody)
            o.n_chars = pr.n_chars; o.sentence_bytes = pr.sentence_bytes; o.reserved = 0;
            if (pr.n_chars) memcpy(o.chars, pr.chars, pr.n_chars);
            if (pr.sentence_bytes) memcpy(o.sentences, pr.sentences, pr.sentence_bytes);
            o.channel = uint32_t(ch_offset + int(c));
            const float* sc = st + 6 * c;
            o.frequency_correction = sc[0]; o.shift = sc[1]; o.noise_floor = sc[2]; o.noise_variance = sc[3];
            o.peak_left = int32_t(sc[4]); o.peak_right = int32_t(sc[5]);
            o.n_sentences = pr.sentence_bytes ? uint16_t(std::count(pr.sentences, pr.sentences + pr.sentence_bytes, '\n')) : uint16_t(0);
            pr.n_chars = 0; pr.sentence_bytes = 0;
            uint16_t fl = 0;
            if (__builtin_expect(!spill_free, 0)) {
                TextChannel& tc = h->text[c];
                if (!tc.chars_spill.empty()) { fl |= 1u; TextChannel::refill(pr.chars, pr.n_chars, sizeof(pr.chars), tc.chars_spill); }
                if (!tc.sent_spill.empty()) { fl |= 2u; TextChannel::refill(pr.sentences, pr.sentence_bytes, sizeof(pr.sentences), tc.sent_spill); }
                any_left = any_left || !tc.chars_spill.empty() || !tc.sent_spill.empty();
            }
            o.flags = fl;
        }
        left[size_t(t)] = any_left;
    });
    if (!h->spill_free) h->spill_free = std::find(left.begin(), left.end(), 1) == left.end();
    return n;
}
int hbd_set_stats_snapshot(hbd_decoder* h, int on)
{
    HBD_CHECK_H(h);
    std::lock_guard<std::mutex> l(h->mtx);
    if (on) return h->enable_snap();
    h->snap_on = false; h->h_snap_valid = false;
    return HBD_OK;
}

int hbd_create(int n_channels, int cuda_device, hbd_decoder** out)
{
    if (!out || n_channels < 1) return HBD_ERR_ARG;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || cuda_device < 0 || cuda_device >= count) return HBD_ERR_CUDA;
    if (cudaSetDevice(cuda_device) != cudaSuccess) return HBD_ERR_CUDA;
    hbd_decoder* h = new (std::nothrow) hbd_decoder;
    if (!h) return HBD_ERR_NOMEM;
    h->n_ch = n_channels; h->device = cuda_device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, cuda_device) == cudaSuccess) h->n_sms = prop.multiProcessorCount;
    h->hc.resize(size_t(n_channels));
    h->text.resize(size_t(n_channels));
    for (auto& tc : h->text) { tc.text_stream.reserve(120); tc.last_sentence.reserve(96); }   // no allocation on the drain path in steady state
    { hbd_result_record z; memset(&z, 0, sizeof(z)); h->pend.assign(size_t(n_channels), z); }
    if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) { delete h; return HBD_ERR_CUDA; }
    h->own_stream = true;
    {
        if (const char* sv = getenv("HBD_SV_WANT")) h->sv_override = atoi(sv);
        if (const char* mc = getenv("HBD_MASK_CACHE")) h->mask_cache = atoi(mc) != 0;
        if (const char* tl = getenv("HBD_TAIL_EV_LATE")) h->tail_ev_late = atoi(tl) != 0;
        if (const char* nf = getenv("HBD_NCO_FUSED")) h->nco_fused = atoi(nf) != 0;
        {
            cpu_set_t set;
            int cores = int(std::thread::hardware_concurrency());
            if (sched_getaffinity(0, sizeof(set), &set) == 0) cores = CPU_COUNT(&set);
            h->host_threads = std::max(1, std::min(4, cores / 2));   // measured on 8 ranks x 4 cores: 2 threads 0.44 ms/step, 4 threads 0.51 (stalls)
            if (const char* ht = getenv("HBD_HOST_THREADS")) h->host_threads = std::max(1, atoi(ht));
        }
        int prio_lo = 0, prio_hi = 0;
        cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi); // numerically lower = higher priority
        if (cudaStreamCreateWithPriority(&h->hi, cudaStreamNonBlocking, prio_hi) != cudaSuccess ||
            cudaStreamCreateWithPriority(&h->lo, cudaStreamNonBlocking, prio_lo) != cudaSuccess) { h->free_all(); delete h; return HBD_ERR_CUDA; }
        if (cudaEventCreateWithFlags(&h->ev_consumed, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&h->ev_tail[0], cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&h->ev_tail[1], cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&h->ev_in, cudaEventDisableTiming) != cudaSuccess) { h->free_all(); delete h; return HBD_ERR_CUDA; }
    }
    const int rc = h->alloc_fixed();
    if (rc) { h->free_all(); delete h; return rc; }
    *out = h;
    return HBD_OK;
}

void hbd_destroy(hbd_decoder* h)
{
    if (!h) return;
    h->free_all();
    delete h;
}

const char* hbd_last_error(hbd_decoder* h) { return h ? h->err.c_str() : "null handle"; }

int hbd_set_stream(hbd_decoder* h, void* s)
{
    HBD_CHECK_H(h);
    std::lock_guard<std::mutex> l(h->mtx);
    cudaSetDevice(h->device);
    h->sync_groups();
    cudaStreamSynchronize(h->stream);
    if (h->own_stream) cudaStreamDestroy(h->stream);
    h->stream = (cudaStream_t)s; h->own_stream = false;
    return HBD_OK;
}

// 4096 (the reference's fft_bins_cnt_) or 16384; restarts spectrum collection and the AFC of every channel
int hbd_set_fft_size(hbd_decoder* h, size_t n_bins)
{
    HBD_CHECK_H(h);
    if (n_bins != size_t(kFftN) && n_bins != size_t(kFftNMax)) return HBD_ERR_ARG;
    std::lock_guard<std::mutex> l(h->mtx);
    cudaSetDevice(h->device);
    if (h->sync_groups() || cudaStreamSynchronize(h->stream) != cudaSuccess) return HBD_ERR_CUDA;
    if (int(n_bins) == h->fft_n) return HBD_OK;
    h->fft_n = int(n_bins);
    const int rc = h->alloc_fft();
    if (rc) return rc;
    // spectrum / AFC state back to a freshly constructed decoder's (Average<T> starts with one 0 sample)
    std::vector<ChanState> st(size_t(h->n_ch));
    if (cudaMemcpy(st.data(), h->d_state, st.size() * sizeof(ChanState), cudaMemcpyDeviceToHost) != cudaSuccess) return HBD_ERR_CUDA;
    h->sync_ctrs();
    for (auto& x : h->hc) x.fft_have = 0;
    for (auto& s : st) {
        s.fft_have = s.fft_ready = s.have_spectrum = 0;
        s.afc_correction = s.afc_noise_floor = s.afc_noise_var = s.afc_shift_hz = 0;
        s.nf_sum = s.nv_sum = 0; s.pl_sum = s.pr_sum = 0; s.nf_cnt = s.nv_cnt = s.pl_cnt = s.pr_cnt = 1;
        s.gui_left = s.gui_right = 0; s.spec_ok = 0;
    }
    if (cudaMemcpy(h->d_state, st.data(), st.size() * sizeof(ChanState), cudaMemcpyHostToDevice) != cudaSuccess) return HBD_ERR_CUDA;
    return HBD_OK;
}

int hbd_set_record(hbd_decoder* h, int on)
{
    HBD_CHECK_H(h);
    std::lock_guard<std::mutex> l(h->mtx);
    h->record = on != 0;
    h->cap_n_in = 0; // force (re)allocation of the recording buffers on the next call
    return HBD_OK;
}

#define HBD_SETTER(NAME, TYPE, FIELD)                                                \
    int hbd_set_##NAME(hbd_decoder* h, int ch, TYPE v)                               \
    {                                                                                \
        if (!h || ch < -1 || ch >= h->n_ch) return HBD_ERR_ARG;                      \
        std::lock_guard<std::mutex> l(h->mtx);                                       \
        for (int c = (ch < 0 ? 0 : ch); c < (ch < 0 ? h->n_ch : ch + 1); ++c) {      \
            h->hc[size_t(c)].FIELD = v;                                              \
            h->hc[size_t(c)].cfg_dirty = true;                                       \
        }                                                                            \
        h->cfg_dirty_any = h->derived_dirty = true;                                  \
        return HBD_OK;                                                               \
    }
HBD_SETTER(baud, double, baud)
HBD_SETTER(rtty_bits, size_t, rtty_bits)
HBD_SETTER(rtty_stops, float, rtty_stops)
HBD_SETTER(dc_remove, int, dc_remove)

// Decoder::lowpass_bw / lowpass_trans re-run the design immediately (Decoder.h:238-257)
static int set_lp(hbd_decoder* h, int ch, float bw, float trans, bool set_bw)
{
    if (!h || ch < -1 || ch >= h->n_ch) return HBD_ERR_ARG;
    std::lock_guard<std::mutex> l(h->mtx);
    // FirFilter::LP_BlackmanHarris takes (size_t)(4 / trans) taps (clipped to the block length); the tap rows hold kLpMaxTaps
    if (!set_bw && trans != 0.f && !(4.0f / trans <= float(kLpMaxTaps - 1))) {
        h->set_error("lowpass_trans below 4/1024 needs more than kLpMaxTaps (1025) low-pass taps");
        return HBD_ERR_ARG;
    }
    cudaSetDevice(h->device);
    h->sync_groups();
    h->sync_ctrs();
    std::vector<float> taps;
    for (int c = (ch < 0 ? 0 : ch); c < (ch < 0 ? h->n_ch : ch + 1); ++c) {
        HostChan& x = h->hc[size_t(c)];
        if (set_bw) x.lp_bw = bw; else x.lp_trans = trans;
        x.lp_dirty = true;
        if (!h->fs_in) continue;
        const double fs_dec = h->fs_in / h->factor;
        const size_t T = design_lowpass(float(x.lp_bw / fs_dec), x.lp_trans, x.lp_input_size, x.lp_ntaps, taps);
        if (T != x.lp_ntaps && T <= size_t(kLpMaxTaps)) {
            x.lp_ntaps = T; x.cfg_dirty = true; h->cfg_dirty_any = h->derived_dirty = true;
            if (cudaMemcpy(h->d_lptaps + size_t(c) * kLpMaxTaps, taps.data(), 4 * T, cudaMemcpyHostToDevice) != cudaSuccess) return HBD_ERR_CUDA;
        }
    }
    bool cfg_same = true, ctr_same = true;
    for (size_t c = 1; c < h->hc.size(); ++c) {
        cfg_same = cfg_same && h->hc[c].lp_bw == h->hc[0].lp_bw && h->hc[c].lp_trans == h->hc[0].lp_trans;
        ctr_same = ctr_same && h->hc[c].same(h->hc[0]);
    }
    h->lp_cfg_uniform = cfg_same; h->ctr_uniform = ctr_same;
    return HBD_OK;
}
int hbd_set_lowpass_bw(hbd_decoder* h, int ch, float bw) { return set_lp(h, ch, bw, 0, true); }
int hbd_set_lowpass_trans(hbd_decoder* h, int ch, float tr) { return set_lp(h, ch, 0, tr, false); }

#define HBD_GETTER(NAME, TYPE, FIELD)                                                \
    TYPE hbd_get_##NAME(hbd_decoder* h, int ch)                                      \
    {                                                                                \
        if (!h || ch < 0 || ch >= h->n_ch) return 0;                                 \
        std::lock_guard<std::mutex> l(h->mtx);                                       \
        return h->hc[size_t(ch)].FIELD;                                              \
    }
HBD_GETTER(baud, double, baud)
HBD_GETTER(rtty_bits, size_t, rtty_bits)
HBD_GETTER(rtty_stops, float, rtty_stops)
HBD_GETTER(lowpass_bw, float, lp_bw)
HBD_GETTER(lowpass_trans, float, lp_trans)
HBD_GETTER(dc_remove, int, dc_remove)

size_t hbd_setup_decimation_factor(hbd_decoder* h, size_t factor)
{
    if (!h) return 0;
    std::lock_guard<std::mutex> l(h->mtx);
    if (factor < 1 || factor > 256) return size_t(h->factor);
    std::vector<TapTable> st;
    if (!plan_for_factor(factor, st)) { h->set_plan(st, 1); return 0; }   // Decoder.h:283-284,317-319: stages cleared, factor 1, returns 0
    return h->set_plan(st, factor) == HBD_OK ? factor : 0;
}

// Decoder::setupDecimationStagesBW (Decoder.h:336-412): while the rate is above the limit, divide by the smallest power of
// two below 256 that gets under it -- or by 256 when none does -- and append that factor's stages.  Limits far below the
// input rate therefore cascade several plans (20 MS/s -> 5 kHz: /256 then /16 = (64,348t)(4,139t)(8,54t)(2,69t)).
size_t hbd_setup_decimation_bw(hbd_decoder* h, double max_rate)
{
    if (!h) return 0;
    std::lock_guard<std::mutex> l(h->mtx);
    if (!h->fs_in) return 0;
    double rate = h->fs_in;
    size_t total = 1;
    std::vector<TapTable> plan, st;
    while (rate > max_rate) {
        int div;
        for (div = 2; div < 256; div *= 2) if (rate / div <= max_rate) break;
        rate /= div; total *= size_t(div);
        plan_for_factor(size_t(div), st);
        plan.insert(plan.end(), st.begin(), st.end());
        if (total > 65536 || plan.size() > size_t(kMaxMidStages + 2)) {
            h->set_error("setup_decimation_bw: more than 8 stages / a total factor above 65536 is not supported");
            return 0;
        }
    }
    return h->set_plan(plan, total) == HBD_OK ? total : 0;
}

static int latch_rate(hbd_decoder* h, double fs)
{
    if (!h->fs_in) h->fs_in = double(float(fs)); // Decoder::init(const float)
    return HBD_OK;
}

static int ensure_stage(hbd_decoder* h, size_t need)
{
    if (need <= h->stage_pitch && h->d_stage) return HBD_OK;
    const size_t np = (std::max(need, h->stage_pitch) + 15) & ~size_t(15);
    cudaError_t e = grow_rows(&h->d_stage, &h->stage_pitch, np, size_t(h->n_ch), h->stage_pitch, h->stream);
    if (e != cudaSuccess) { h->set_error(std::string("staging alloc: ") + cudaGetErrorString(e)); return HBD_ERR_CUDA; }
    return HBD_OK;
}

// Mix `n` samples per channel of channels [c0, c0+nc) from src into the staging rows at dst_off and advance the
// channels' NCO phases.  src_pitch 0: one shared (wideband) row.  Runs on the handle's stream.
int hbd_decoder::mix_into_stage(const float2* src, size_t src_pitch, int c0, int nc, size_t dst_off, size_t n)
{
    if (!n) return HBD_OK;
    for (int c = c0; c < c0 + nc; ++c) {
        HostChan& x = hc[size_t(c)];
        NcoChan& k = h_nco[size_t(c)];
        k.inc = x.nco_freq / fs_in;
        k.ph0 = x.nco_phase;
        const double a = -2.0 * M_PI * k.inc * double(kNcoThreads);
        k.step_re = std::cos(a); k.step_im = std::sin(a);
        double ph = x.nco_phase + double(n) * k.inc;
        x.nco_phase = ph - std::floor(ph);
    }
    HBD_CUDA_CHECK(cudaMemcpyAsync(d_nco + c0, h_nco.data() + c0, sizeof(NcoChan) * size_t(nc), cudaMemcpyHostToDevice, stream));
    int nl = 0;
    HBD_CUDA_CHECK(launch_nco_mix(src, src_pitch, d_stage, stage_pitch, dst_off, n, d_nco, c0, nc, stream, &nl));
    launches += unsigned(nl);
    HBD_CUDA_CHECK(cudaStreamSynchronize(stream)); // h_nco is rewritten by the next push
    return HBD_OK;
}

// Fused form of mix_into_stage for a whole-batch device push with nothing staged: the raw samples stay where they are
// (src_pitch 0: one wideband row), the per-channel NCO parameters go to the device and K1 mixes while it decimates
// (decim1.cu, DecimArgs::nco) -- no K0 launch, no 8-byte write + 8-byte read per channel-sample through HBM.
int hbd_decoder::push_fused_nco(const float2* src, size_t pitch, size_t n)
{
    if (!h_nco_pin) {
        HBD_CUDA_CHECK(cudaHostAlloc((void**)&h_nco_pin, sizeof(NcoChan) * size_t(n_ch) * kNcoSlots, cudaHostAllocDefault));
        HBD_CUDA_CHECK(dalloc(&d_nco_ring, size_t(n_ch) * kNcoSlots));
        for (unsigned i = 0; i < kNcoSlots; ++i) HBD_CUDA_CHECK(cudaEventCreateWithFlags(&ev_nco[i], cudaEventDisableTiming));
    }
    const unsigned slot = nco_seq % kNcoSlots;
    if (nco_seq >= kNcoSlots) HBD_CUDA_CHECK(cudaEventSynchronize(ev_nco[slot]));   // the upload that last used this pinned slot (16 pushes ago)
    ++nco_seq;
    NcoChan* hp = h_nco_pin + size_t(slot) * size_t(n_ch);
    for (int c = 0; c < n_ch; ++c) {
        HostChan& x = hc[size_t(c)];
        NcoChan& k = hp[c];
        k.inc = x.nco_freq / fs_in;
        k.ph0 = x.nco_phase;
        k.step_re = 1.0; k.step_im = 0.0;   // K0's rotation, unused by K1
        double ph = x.nco_phase + double(n) * k.inc;
        x.nco_phase = ph - std::floor(ph);
    }
    NcoChan* dp = d_nco_ring + size_t(slot) * size_t(n_ch);
    HBD_CUDA_CHECK(cudaMemcpyAsync(dp, hp, sizeof(NcoChan) * size_t(n_ch), cudaMemcpyHostToDevice, stream));
    HBD_CUDA_CHECK(cudaEventRecord(ev_nco[slot], stream));
    ext = src; ext_pitch = pitch; ext_n = n; ext_nco = true; ext_nco_ptr = dp;
    return HBD_OK;
}

int hbd_push_samples(hbd_decoder* h, int ch, const float* iq, size_t n, double fs)
{
    HBD_CHECK_CH(h, ch);
    if (!iq && n) return HBD_ERR_ARG;
    std::lock_guard<std::mutex> l(h->mtx);
    if (h->ext) { h->set_error("device push pending"); return HBD_ERR_STATE; }
    if (cudaSetDevice(h->device) != cudaSuccess) return HBD_ERR_CUDA;
    if (h->n_ch > 1) h->leave_uniform();   // per-channel feeding: the counters are stepped channel by channel from here on
    HostChan& x = h->hc[size_t(ch)];
    const int rc = ensure_stage(h, size_t(x.pushed) + n);
    if (rc) return rc;
    if (n) {
        cudaError_t e = cudaMemcpyAsync(h->d_stage + size_t(ch) * h->stage_pitch + x.pushed, iq, n * sizeof(float2), cudaMemcpyHostToDevice, h->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream); // the caller owns `iq` again on return (Decoder.h:209-213 copies)
        if (e != cudaSuccess) { h->set_error(cudaGetErrorString(e)); return HBD_ERR_CUDA; }
    }
    latch_rate(h, fs);
    if (x.nco_freq != 0 || x.nco_phase != 0) {
        const float2* row = h->d_stage + size_t(ch) * h->stage_pitch + x.pushed;
        const int rc2 = h->mix_into_stage(row - size_t(ch) * h->stage_pitch, h->stage_pitch, ch, 1, x.pushed, n); // in place
        if (rc2) return rc2;
    }
    x.pushed += unsigned(n);
    return HBD_OK;
}

int hbd_push_samples_batch(hbd_decoder* h, const float* iq, size_t n, size_t pitch, double fs)
{
    HBD_CHECK_H(h);
    if ((!iq && n) || pitch < n) return HBD_ERR_ARG;
    std::lock_guard<std::mutex> l(h->mtx);
    if (h->ext) { h->set_error("device push pending"); return HBD_ERR_STATE; }
    if (cudaSetDevice(h->device) != cudaSuccess) return HBD_ERR_CUDA;
    unsigned base = 0;
    if (!h->queue_depth_equal(base)) { h->set_error("batch push needs equal queue depth in all channels"); return HBD_ERR_STATE; }
    const int rc = ensure_stage(h, size_t(base) + n);
    if (rc) return rc;
    if (n) {
        cudaError_t e = cudaMemcpy2DAsync(h->d_stage + base, h->stage_pitch * sizeof(float2), iq, pitch * sizeof(float2), n * sizeof(float2),
                                          size_t(h->n_ch), cudaMemcpyHostToDevice, h->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
        if (e != cudaSuccess) { h->set_error(cudaGetErrorString(e)); return HBD_ERR_CUDA; }
    }
    latch_rate(h, fs);
    if (h->nco_active()) {
        const int rc2 = h->mix_into_stage(h->d_stage + base, h->stage_pitch, 0, h->n_ch, base, n); // in place
        if (rc2) return rc2;
    }
    h->add_pushed(unsigned(n));
    return HBD_OK;
}

// One capture shared by every channel (frequency-offset channels of a wideband capture): each channel receives
// iq * exp(-2 pi i f_nco t).  `device` != 0: iq is a device pointer.
static int push_wideband(hbd_decoder* h, const float* iq, size_t n, double fs, bool device)
{
    HBD_CHECK_H(h);
    if (!iq && n) return HBD_ERR_ARG;
    std::lock_guard<std::mutex> l(h->mtx);
    if (h->ext) { h->set_error("device push pending"); return HBD_ERR_STATE; }
    if (cudaSetDevice(h->device) != cudaSuccess) return HBD_ERR_CUDA;
    unsigned base = 0;
    if (!h->queue_depth_equal(base)) { h->set_error("wideband push needs equal queue depth in all channels"); return HBD_ERR_STATE; }
    latch_rate(h, fs);
    const float2* src = reinterpret_cast<const float2*>(iq);
    if (!device) {
        if (n > h->wide_cap) {
            if (h->d_wide) { cudaStreamSynchronize(h->stream); cudaFree(h->d_wide); h->d_wide = nullptr; }
            if (cudaMalloc((void**)&h->d_wide, n * sizeof(float2)) != cudaSuccess) { h->set_error("wideband buffer alloc"); return HBD_ERR_NOMEM; }
            h->wide_cap = n;
        }
        cudaError_t e = cudaMemcpyAsync(h->d_wide, iq, n * sizeof(float2), cudaMemcpyHostToDevice, h->stream);
        if (e != cudaSuccess) { h->set_error(cudaGetErrorString(e)); return HBD_ERR_CUDA; }
        src = h->d_wide;
    }
    if (h->nco_fused && base == 0 && n && decim1_supports_fused_nco(h->M1, h->T1) && !(reinterpret_cast<uintptr_t>(src) & 15)) {
        // host capture: the caller owns `iq` again on return (Decoder.h:209-213 copies), so the upload has to be complete
        if (!device && cudaStreamSynchronize(h->stream) != cudaSuccess) return HBD_ERR_CUDA;
        return h->push_fused_nco(src, 0, n);
    }
    const int rc = ensure_stage(h, size_t(base) + n);
    if (rc) return rc;
    const int rc2 = h->mix_into_stage(src, 0, 0, h->n_ch, base, n);
    if (rc2) return rc2;
    h->add_pushed(unsigned(n));
    return HBD_OK;
}
int hbd_push_wideband(hbd_decoder* h, const float* iq, size_t n, double fs) { return push_wideband(h, iq, n, fs, false); }
int hbd_push_wideband_device(hbd_decoder* h, const float* d_iq, size_t n, double fs) { return push_wideband(h, d_iq, n, fs, true); }

int hbd_set_nco(hbd_decoder* h, int ch, double freq_hz)
{
    if (!h || ch < -1 || ch >= h->n_ch) return HBD_ERR_ARG;
    std::lock_guard<std::mutex> l(h->mtx);
    for (int c = (ch < 0 ? 0 : ch); c < (ch < 0 ? h->n_ch : ch + 1); ++c) h->hc[size_t(c)].nco_freq = freq_hz;
    if (freq_hz != 0) h->nco_any = true;
    return HBD_OK;
}
double hbd_get_nco(hbd_decoder* h, int ch)
{
    if (!h || ch < 0 || ch >= h->n_ch) return 0;
    std::lock_guard<std::mutex> l(h->mtx);
    return h->hc[size_t(ch)].nco_freq;
}

int hbd_push_samples_device(hbd_decoder* h, const float* d_iq, size_t n, size_t pitch, double fs)
{
    HBD_CHECK_H(h);
    if (!d_iq || pitch < n || (pitch & 1) || (reinterpret_cast<uintptr_t>(d_iq) & 15)) return HBD_ERR_ARG;
    std::lock_guard<std::mutex> l(h->mtx);
    if (h->ext) { h->set_error("device push pending"); return HBD_ERR_STATE; }
    latch_rate(h, fs);
    if (h->nco_active()) { // zero copy is not possible: the mixed samples go through the staging matrix
        if (cudaSetDevice(h->device) != cudaSuccess) return HBD_ERR_CUDA;
        unsigned base = 0;
        if (!h->queue_depth_equal(base)) { h->set_error("batch push needs equal queue depth in all channels"); return HBD_ERR_STATE; }
        if (h->nco_fused && base == 0 && n && decim1_supports_fused_nco(h->M1, h->T1))
            return h->push_fused_nco(reinterpret_cast<const float2*>(d_iq), pitch, n);   // zero copy after all: K1 mixes
        const int rc = ensure_stage(h, size_t(base) + n);
        if (rc) return rc;
        const int rc2 = h->mix_into_stage(reinterpret_cast<const float2*>(d_iq), pitch, 0, h->n_ch, base, n);
        if (rc2) return rc2;
        h->add_pushed(unsigned(n));
        return HBD_OK;
    }
    { unsigned base = 0; if (!h->queue_depth_equal(base) || base) { h->set_error("host push pending"); return HBD_ERR_STATE; } }
    h->ext = reinterpret_cast<const float2*>(d_iq); h->ext_pitch = pitch; h->ext_n = n;
    return HBD_OK;
}

int hbd_process_async(hbd_decoder* h) { HBD_CHECK_H(h); return locked_then_fire(h, [h] { return h->process_async_locked(); }); }
int hbd_collect(hbd_decoder* h) { HBD_CHECK_H(h); return locked_then_fire(h, [h] { return h->collect_locked(0); }); }
int hbd_collect_ready(hbd_decoder* h, unsigned lag) { HBD_CHECK_H(h); return locked_then_fire(h, [h, lag] { return h->collect_locked(lag); }); }
int hbd_process(hbd_decoder* h)
{
    HBD_CHECK_H(h);
    return locked_then_fire(h, [h] {
        const int rc = h->process_async_locked();
        if (rc) return rc;
        return h->collect_locked(0);
    });
}
int hbd_synchronize(hbd_decoder* h)
{
    HBD_CHECK_H(h);
    std::lock_guard<std::mutex> l(h->mtx);
    cudaSetDevice(h->device);
    if (h->sync_groups()) return HBD_ERR_CUDA;
    return cudaStreamSynchronize(h->stream) == cudaSuccess ? HBD_OK : HBD_ERR_CUDA;
}
unsigned long long hbd_kernel_launches(hbd_decoder* h) { return h ? h->launches : 0; }

int hbd_set_kernel_timing(hbd_decoder* h, int on)
{
    HBD_CHECK_H(h);
    std::lock_guard<std::mutex> l(h->mtx);
    h->timing = on < 0 ? 0 : on;
    h->ev_used_k1 = h->ev_used_rest = 0;
    h->drain_host_ms = 0; h->drain_calls = 0;
    if (on) {   // events are created here, not on the issue path of the calls being measured
        cudaSetDevice(h->device);
        for (auto* pool : {&h->ev_k1, &h->ev_rest})
            while (pool->size() < 1024) { cudaEvent_t e; if (cudaEventCreate(&e) != cudaSuccess) return HBD_ERR_CUDA; pool->push_back(e); }
    }
    return HBD_OK;
}

int hbd_get_kernel_timing(hbd_decoder* h, int which, double* total_ms, unsigned* count)
{
    HBD_CHECK_H(h);
    std::lock_guard<std::mutex> l(h->mtx);
    cudaSetDevice(h->device);
    if (h->sync_groups() || cudaStreamSynchronize(h->stream) != cudaSuccess) return HBD_ERR_CUDA;
    double tot = 0; unsigned cnt = 0;
    if (which == 5) {   // host time spent replaying drained calls through the sentence layer (wall clock of the caller's thread)
        if (total_ms) *total_ms = h->drain_host_ms;
        if (count) *count = h->drain_calls;
        return HBD_OK;
    }
    if (which >= 2) {   // pipeline diagnostics (one channel group): signed gaps between events of different pairs
        const size_t calls = std::min(h->ev_used_k1, h->ev_used_rest) / 2;
        for (size_t i = 0; i < calls; ++i) {
            cudaEvent_t a = nullptr, b = nullptr;
            if (which == 2 && i + 1 < calls) { a = h->ev_k1[2 * i + 1]; b = h->ev_k1[2 * i + 2]; }        // K1 end -> next K1 start
            else if (which == 3) { a = h->ev_k1[2 * i + 1]; b = h->ev_rest[2 * i]; }                      // K1 end -> own tail start
            else if (which == 4 && i + 2 < calls) { a = h->ev_rest[2 * i + 1]; b = h->ev_k1[2 * i + 4]; } // tail end -> K1 start two calls on
            if (!a) continue;
            float ms = 0;
            if (cudaEventElapsedTime(&ms, a, b) != cudaSuccess) return HBD_ERR_CUDA;
            tot += ms; ++cnt;
        }
        if (total_ms) *total_ms = tot;
        if (count) *count = cnt;
        return HBD_OK;
    }
    const std::vector<cudaEvent_t>& pool = which == 0 ? h->ev_k1 : h->ev_rest;
    const size_t used = which == 0 ? h->ev_used_k1 : h->ev_used_rest;
    for (size_t i = 0; i + 1 < used; i += 2) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, pool[i], pool[i + 1]) != cudaSuccess) return HBD_ERR_CUDA;
        tot += ms; ++cnt;
    }
    if (total_ms) *total_ms = tot;
    if (count) *count = cnt;
    return HBD_OK;
}

static size_t copy_out(const std::string& s, char* out, size_t cap)
{
    if (out && cap) memcpy(out, s.data(), std::min(cap, s.size()));
    return s.size();
}

size_t hbd_get_rtty(hbd_decoder* h, int ch, char* out, size_t cap)
{
    if (!h || ch < 0 || ch >= h->n_ch) return 0;
    std::lock_guard<std::mutex> l(h->mtx);
    return copy_out(h->text[size_t(ch)].text_stream, out, cap);
}
size_t hbd_get_last_sentence(hbd_decoder* h, int ch, char* out, size_t cap)
{
    if (!h || ch < 0 || ch >= h->n_ch) return 0;
    std::lock_guard<std::mutex> l(h->mtx);
    return copy_out(h->text[size_t(ch)].last_sentence, out, cap);
}
size_t hbd_poll_chars(hbd_decoder* h, int ch, char* out, size_t cap)
{
    if (!h || ch < 0 || ch >= h->n_ch) return 0;
    std::lock_guard<std::mutex> l(h->mtx);
    TextChannel& tc = h->text[size_t(ch)];
    hbd_result_record& pr = h->pend[size_t(ch)];
    const size_t n = tc.chars_size(pr);
    if (out && cap && n) { const std::string s = tc.chars_from(pr, 0); memcpy(out, s.data(), std::min(cap, n)); }
    if (out && cap >= n) tc.clear_chars(pr);
    return n;
}
size_t hbd_poll_sentences(hbd_decoder* h, int ch, char* out, size_t cap)
{
    if (!h || ch < 0 || ch >= h->n_ch) return 0;
    std::lock_guard<std::mutex> l(h->mtx);
    TextChannel& tc = h->text[size_t(ch)];
    hbd_result_record& pr = h->pend[size_t(ch)];
    const size_t n = tc.sent_size(pr);
    if (out && cap && n) { const std::string s = tc.sentences_all(pr); memcpy(out, s.data(), std::min(cap, n)); }
    if (out && cap >= n) tc.clear_sentences(pr);
    return n;
}
size_t hbd_poll_raw_chars(hbd_decoder* h, int ch, unsigned char* out, size_t cap)
{
    if (!h || ch < 0 || ch >= h->n_ch) return 0;
    std::lock_guard<std::mutex> l(h->mtx);
    auto& v = h->text[size_t(ch)].raw_pending;
    const size_t n = v.size();
    if (out && cap) memcpy(out, v.data(), std::min(cap, n));
    if (out && cap >= n) v.clear();
    return n;
}
int hbd_set_host_threads(hbd_decoder* h, int n)
{
    HBD_CHECK_H(h); std::lock_guard<std::mutex> l(h->mtx); h->host_threads = std::max(1, std::min(n, 64)); return HBD_OK;
}
int hbd_set_raw_chars(hbd_decoder* h, int on)
{
    HBD_CHECK_H(h); std::lock_guard<std::mutex> l(h->mtx); h->keep_raw = on != 0; return HBD_OK;
}
int hbd_attach_tracker(hbd_decoder* h, hbd_tracker* t, int ch_offset)
{
    HBD_CHECK_H(h); std::lock_guard<std::mutex> l(h->mtx); h->tracker = t; h->tracker_off = ch_offset; return HBD_OK;
}

int hbd_set_sentence_callback(hbd_decoder* h, hbd_sentence_cb cb, void* user)
{
    HBD_CHECK_H(h); std::lock_guard<std::mutex> l(h->mtx); h->sentence_cb = cb; h->sentence_user = user; return HBD_OK;
}
int hbd_set_chars_callback(hbd_decoder* h, hbd_chars_cb cb, void* user)
{
    HBD_CHECK_H(h); std::lock_guard<std::mutex> l(h->mtx); h->chars_cb = cb; h->chars_user = user; return HBD_OK;
}

// ---- SSDV packet sync ---------------------------------------------------------------------------------------
int hbd_set_ssdv(hbd_decoder* h, int on)
{
    HBD_CHECK_H(h);
    return locked_then_fire(h, [h, on] { return h->enable_ssdv(on != 0); });
}
int hbd_set_ssdv_callback(hbd_decoder* h, hbd_ssdv_cb cb, void* user)
{
    HBD_CHECK_H(h);
    return locked_then_fire(h, [h, cb, user] {
        const int rc = cb ? h->enable_ssdv(true) : HBD_OK;   // drains the calls in flight (under the previous callback)
        h->ssdv_cb = cb; h->ssdv_user = user;
        return rc;
    });
}
size_t hbd_poll_ssdv_packets(hbd_decoder* h, int ch, hbd_ssdv_packet_info* infos, unsigned char* packets, size_t cap_packets)
{
    if (!h || ch < 0 || ch >= h->n_ch) return 0;
    std::lock_guard<std::mutex> l(h->mtx);
    if (h->ssdv.empty()) return 0;
    auto& ev = h->ssdv[size_t(ch)].events_pending;
    const size_t n = ev.size();
    for (size_t i = 0; i < std::min(n, cap_packets); ++i) {
        if (infos) fill_ssdv_info(infos[i], ev[i]);
        if (packets) memcpy(packets + 256 * i, ev[i].data.data(), 256);
    }
    if ((infos || packets) && cap_packets >= n) ev.clear();
    return n;
}
size_t hbd_get_ssdv_image(hbd_decoder* h, int ch, const char* callsign, int image_id, unsigned char* out, size_t cap)
{
    if (!h || ch < 0 || ch >= h->n_ch || !callsign) return 0;
    std::lock_guard<std::mutex> l(h->mtx);
    if (h->ssdv.empty()) return 0;
    return h->ssdv[size_t(ch)].image(callsign, image_id, out, cap);
}
int hbd_get_ssdv_last_image(hbd_decoder* h, int ch, char callsign[8], int* image_id)
{
    HBD_CHECK_CH(h, ch); std::lock_guard<std::mutex> l(h->mtx);
    if (h->ssdv.empty()) return HBD_ERR_STATE;
    const auto& k = h->ssdv[size_t(ch)].last_key;
    if (callsign) { memset(callsign, 0, 8); strncpy(callsign, k.first.c_str(), 7); }
    if (image_id) *image_id = k.second;
    return HBD_OK;
}
int hbd_ssdv_check_packets(hbd_decoder* h, unsigned char* windows, size_t n, int* verdict, int* errors)
{
    HBD_CHECK_H(h);
    if (!windows || !verdict || !errors) return HBD_ERR_ARG;
    if (!n) return HBD_OK;
    std::lock_guard<std::mutex> l(h->mtx);
    if (cudaSetDevice(h->device) != cudaSuccess) return HBD_ERR_CUDA;
    unsigned char* d_w = nullptr; int* d_v = nullptr;
    auto fail = [&](const char* what) { h->set_error(what); if (d_w) cudaFree(d_w); if (d_v) cudaFree(d_v); return HBD_ERR_CUDA; };
    if (cudaMalloc((void**)&d_w, n * 256) != cudaSuccess || cudaMalloc((void**)&d_v, n * 2 * sizeof(int)) != cudaSuccess) return fail("cudaMalloc (ssdv check)");
    if (cudaMemcpyAsync(d_w, windows, n * 256, cudaMemcpyHostToDevice, h->stream) != cudaSuccess) return fail("cudaMemcpyAsync (ssdv check)");
    int nl = 0;
    if (launch_ssdv_check(d_w, int(n), d_v, d_v + n, h->stream, &nl) != cudaSuccess) return fail("ssdv_check_kernel launch");
    h->launches += unsigned(nl);
    if (cudaMemcpyAsync(windows, d_w, n * 256, cudaMemcpyDeviceToHost, h->stream) != cudaSuccess ||
        cudaMemcpyAsync(verdict, d_v, n * sizeof(int), cudaMemcpyDeviceToHost, h->stream) != cudaSuccess ||
        cudaMemcpyAsync(errors, d_v + n, n * sizeof(int), cudaMemcpyDeviceToHost, h->stream) != cudaSuccess ||
        cudaStreamSynchronize(h->stream) != cudaSuccess) return fail("ssdv check read-back");
    cudaFree(d_w); cudaFree(d_v);
    return HBD_OK;
}

// host half of the SSDV path alone (test hook, no GPU): replay SsdvChannel over `n_chunks` pushes of the character
// stream `chars`, with the accepted windows given by position
size_t hbd_ssdv_host_replay(const unsigned char* chars, const size_t* chunk_sizes, size_t n_chunks, const unsigned* accepted_pos,
                            const unsigned char* accepted_packets, const int* accepted_errors, size_t n_accepted,
                            hbd_ssdv_packet_info* out_infos, unsigned* out_chunk, unsigned char* out_packets, size_t cap)
{
    if (!chars || !chunk_sizes) return 0;
    SsdvChannel sc;
    for (size_t i = 0; i < n_accepted; ++i) {
        SsdvVerdict v; v.pos = accepted_pos[i]; v.errors = accepted_errors ? accepted_errors[i] : 0;
        memcpy(v.data.data(), accepted_packets + 256 * i, 256);
        sc.verdicts.push_back(v);
    }
    size_t off = 0, n_ev = 0;
    for (size_t c = 0; c < n_chunks; ++c) {
        SsdvEvent ev;
        if (chunk_sizes[c] && sc.push(chars + off, chunk_sizes[c], ev)) {
            if (n_ev < cap) {
                if (out_infos) fill_ssdv_info(out_infos[n_ev], ev);
                if (out_chunk) out_chunk[n_ev] = unsigned(c);
                if (out_packets) memcpy(out_packets + 256 * n_ev, ev.data.data(), 256);
            }
            ++n_ev;
        }
        off += chunk_sizes[c];
    }
    return n_ev;
}

int hbd_get_decimation_factor(hbd_decoder* h) { if (!h) return 0; std::lock_guard<std::mutex> l(h->mtx); return h->factor; }
double hbd_get_input_sampling_rate(hbd_decoder* h) { if (!h) return 0; std::lock_guard<std::mutex> l(h->mtx); return h->fs_in; }
double hbd_get_decimated_sampling_rate(hbd_decoder* h) { if (!h) return 0; std::lock_guard<std::mutex> l(h->mtx); return h->fs_in / h->factor; }
double hbd_get_symbol_rate(hbd_decoder* h, int ch) { return hbd_get_baud(h, ch); }
int hbd_n_channels(hbd_decoder* h) { return h ? h->n_ch : 0; }
size_t hbd_get_bins_count(hbd_decoder* h) { if (!h) return size_t(kFftN); std::lock_guard<std::mutex> l(h->mtx); return size_t(h->fft_n); }

static int fetch_state(hbd_decoder* h, int ch, ChanState* st)
{
    cudaSetDevice(h->device);
    if (h->sync_groups() || cudaStreamSynchronize(h->stream) != cudaSuccess) return HBD_ERR_CUDA;
    return cudaMemcpy(st, h->d_state + ch, sizeof(ChanState), cudaMemcpyDeviceToHost) == cudaSuccess ? HBD_OK : HBD_ERR_CUDA;
}
static size_t fetch_floats(hbd_decoder* h, const void* dsrc, size_t n, float* out, size_t cap)
{
    if (out && cap && n) {
        cudaSetDevice(h->device);
        h->sync_groups();
        cudaStreamSynchronize(h->stream);
        if (cudaMemcpy(out, dsrc, std::min(n, cap) * sizeof(float), cudaMemcpyDeviceToHost) != cudaSuccess) return 0;
    }
    return n;
}

size_t hbd_get_fft(hbd_decoder* h, int ch, float* out, size_t cap)
{
    if (!h || ch < 0 || ch >= h->n_ch) return 0;
    std::lock_guard<std::mutex> l(h->mtx);
    ChanState st; if (fetch_state(h, ch, &st)) return 0;
    if (!st.have_spectrum) return 0; // freq_out_ is empty before the first FFT
    return fetch_floats(h, h->d_spectrum + size_t(ch) * size_t(h->fft_n), 2 * size_t(h->fft_n), out, cap);
}
size_t hbd_get_power_spectrum(hbd_decoder* h, int ch, float* out, size_t cap)
{
    if (!h || ch < 0 || ch >= h->n_ch) return 0;
    std::lock_guard<std::mutex> l(h->mtx);
    ChanState st; if (fetch_state(h, ch, &st)) return 0;
    if (!st.have_spectrum) return 0;
    return fetch_floats(h, h->d_power + size_t(ch) * size_t(h->fft_n), size_t(h->fft_n), out, cap);
}
size_t hbd_get_demodulated(hbd_decoder* h, int ch, float* out, size_t cap)
{
    if (!h || ch < 0 || ch >= h->n_ch) return 0;
    std::lock_guard<std::mutex> l(h->mtx);
    h->sync_ctrs();
    const size_t n = h->hc[size_t(ch)].demod_n;   // a call that demodulates nothing leaves the previous block in place
    if (!h->d_demod) return 0;
    return fetch_floats(h, h->d_demod + size_t(ch) * h->demod_pitch, n, out, cap);
}
int hbd_get_peaks(hbd_decoder* h, int ch, int* pl, int* pr)
{
    HBD_CHECK_CH(h, ch); std::lock_guard<std::mutex> l(h->mtx);
    ChanState st; const int rc = fetch_state(h, ch, &st); if (rc) return rc;
    if (pl) *pl = st.gui_left; if (pr) *pr = st.gui_right; return HBD_OK;
}
int hbd_get_noise_floor(hbd_decoder* h, int ch, double* nf, double* nv)
{
    HBD_CHECK_CH(h, ch); std::lock_guard<std::mutex> l(h->mtx);
    ChanState st; const int rc = fetch_state(h, ch, &st); if (rc) return rc;
    if (nf) *nf = st.afc_noise_floor; if (nv) *nv = st.afc_noise_var; return HBD_OK;
}
double hbd_get_shift(hbd_decoder* h, int ch)
{
    if (!h || ch < 0 || ch >= h->n_ch) return 0; std::lock_guard<std::mutex> l(h->mtx);
    ChanState st; if (fetch_state(h, ch, &st)) return 0; return st.afc_shift_hz;
}
double hbd_get_frequency_correction(hbd_decoder* h, int ch)
{
    if (!h || ch < 0 || ch >= h->n_ch) return 0; std::lock_guard<std::mutex> l(h->mtx);
    ChanState st; if (fetch_state(h, ch, &st)) return 0; return st.afc_correction;
}
int hbd_reset_frequency_correction(hbd_decoder* h, int ch, double corr)
{
    HBD_CHECK_CH(h, ch); std::lock_guard<std::mutex> l(h->mtx);
    cudaSetDevice(h->device);
    h->sync_groups(); // ordered after everything in flight, like a call between two process() calls
    if (launch_afc_reset(h->d_state, ch, corr, h->fs_in / h->factor, h->fft_n, h->stream) != cudaSuccess) return HBD_ERR_CUDA;
    cudaStreamSynchronize(h->stream);
    ++h->launches;
    return HBD_OK;
}
// ---- websocket wire formats ---------------------------------------------------------------------------------
int hbd_decoder::ensure_frames(size_t pitch)
{
    if (pitch <= frames_pitch && d_frames) return HBD_OK;
    if (d_frames) cudaFree(d_frames);
    d_frames = nullptr;
    pitch = (pitch + 255) & ~size_t(255);
    HBD_CUDA_CHECK(cudaMalloc((void**)&d_frames, size_t(n_ch) * pitch));
    frames_pitch = pitch;
    if (!d_frame_sizes) HBD_CUDA_CHECK(dalloc(&d_frame_sizes, size_t(n_ch)));
    h_frame_sizes.resize(size_t(n_ch));
    return HBD_OK;
}

// accumulation buffers: 50 symbols of the slowest channel plus one call's worth of new samples
int hbd_decoder::ensure_dacc()
{
    const double fs_dec = fs_in / factor;
    size_t need = 0;
    for (const auto& x : hc) if (x.baud > 0) need = std::max(need, size_t(fs_dec / x.baud * 50));
    need += demod_pitch + 64;
    if (!d_dacc_n) { HBD_CUDA_CHECK(dalloc(&d_dacc_n, size_t(n_ch))); HBD_CUDA_CHECK(cudaMemset(d_dacc_n, 0, sizeof(unsigned) * size_t(n_ch))); }
    if (need > dacc_pitch) {
        if (sync_groups()) return HBD_ERR_CUDA;
        HBD_CUDA_CHECK(grow_rows(&d_dacc, &dacc_pitch, (need + 15) & ~size_t(15), size_t(n_ch), dacc_pitch, stream));
    }
    return HBD_OK;
}

extern "C" int hbd_set_demod_accumulate(hbd_decoder* h, int on)
{
    HBD_CHECK_H(h);
    std::lock_guard<std::mutex> l(h->mtx);
    h->demod_acc_on = on != 0;
    return HBD_OK;
}

// frames of channels [ch0, ch0 + nc) into the device scratch, then one copy of sizes + one of payloads
static size_t fetch_frames(hbd_decoder* h, int ch0, int nc, bool spectrum, float zoom, int resolution, int type_size,
                           unsigned char* out, size_t pitch, unsigned* sizes)
{
    if (type_size != 1 && type_size != 2 && type_size != 4) return 0;
    if (resolution < 0) resolution = 0;
    cudaSetDevice(h->device);
    if (h->sync_groups() || cudaStreamSynchronize(h->stream) != cudaSuccess) return 0;
    const size_t hdr = spectrum ? size_t(kSpectrumHeaderBytes) : size_t(kDemodHeaderBytes);
    const size_t max_n = spectrum ? size_t(h->fft_n) : std::max<size_t>(h->dacc_pitch, 1);
    const size_t need = hdr + std::min<size_t>(max_n, size_t(resolution)) * size_t(type_size);
    if (h->ensure_frames(need)) return 0;
    int nl = 0;
    cudaError_t e;
    if (spectrum) {
        SpectrumFrameArgs a{};
        a.state = h->d_state; a.power = h->d_power; a.fft_n = h->fft_n; a.fs_dec = h->fs_in / h->factor; a.zoom = zoom; a.resolution = resolution; a.type_size = type_size;
        a.out = h->d_frames; a.out_pitch = h->frames_pitch; a.sizes = h->d_frame_sizes; a.ch0 = ch0;
        e = launch_spectrum_frames(a, nc, h->stream, &nl);
    } else {
        if (!h->d_dacc || !h->d_dacc_n) return 0;
        DemodFrameArgs a{};
        a.acc = h->d_dacc; a.acc_pitch = h->dacc_pitch; a.acc_n = h->d_dacc_n; a.resolution = resolution; a.type_size = type_size;
        a.out = h->d_frames; a.out_pitch = h->frames_pitch; a.sizes = h->d_frame_sizes; a.ch0 = ch0;
        e = launch_demod_frames(a, nc, h->stream, &nl);
    }
    h->launches += unsigned(nl);
    if (e != cudaSuccess) { h->set_error(cudaGetErrorString(e)); return 0; }
    if (cudaMemcpyAsync(h->h_frame_sizes.data(), h->d_frame_sizes, sizeof(unsigned) * size_t(nc), cudaMemcpyDeviceToHost, h->stream) != cudaSuccess) return 0;
    if (cudaStreamSynchronize(h->stream) != cudaSuccess) return 0;
    size_t longest = 0;
    for (int i = 0; i < nc; ++i) longest = std::max<size_t>(longest, h->h_frame_sizes[size_t(i)]);
    if (sizes) for (int i = 0; i < nc; ++i) sizes[i] = h->h_frame_sizes[size_t(i)];
    if (out && pitch && longest) {
        const size_t w = std::min(pitch, longest);
        if (cudaMemcpy2D(out, pitch, h->d_frames, h->frames_pitch, w, size_t(nc), cudaMemcpyDeviceToHost) != cudaSuccess) return 0;
    }
    return longest;
}

extern "C" {
size_t hbd_get_spectrum_frame(hbd_decoder* h, int ch, float zoom, int resolution, int type_size, unsigned char* out, size_t cap)
{
    if (!h || ch < 0 || ch >= h->n_ch) return 0;
    std::lock_guard<std::mutex> l(h->mtx);
    return fetch_frames(h, ch, 1, true, zoom, resolution, type_size, out, cap, nullptr);
}
size_t hbd_get_spectrum_frames(hbd_decoder* h, float zoom, int resolution, int type_size, unsigned char* out, size_t pitch, unsigned* sizes)
{
    if (!h) return 0;
    std::lock_guard<std::mutex> l(h->mtx);
    return fetch_frames(h, 0, h->n_ch, true, zoom, resolution, type_size, out, pitch, sizes);
}
size_t hbd_get_demod_frame(hbd_decoder* h, int ch, int resolution, int type_size, unsigned char* out, size_t cap)
{
    if (!h || ch < 0 || ch >= h->n_ch) return 0;
    std::lock_guard<std::mutex> l(h->mtx);
    return fetch_frames(h, ch, 1, false, 0.f, resolution, type_size, out, cap, nullptr);
}
size_t hbd_get_demod_frames(hbd_decoder* h, int resolution, int type_size, unsigned char* out, size_t pitch, unsigned* sizes)
{
    if (!h) return 0;
    std::lock_guard<std::mutex> l(h->mtx);
    return fetch_frames(h, 0, h->n_ch, false, 0.f, resolution, type_size, out, pitch, sizes);
}
} // extern "C"

// AFC loop closed on the GPU for all channels (see afc_retune_kernel); returns the number of channels retuned
int hbd_afc_retune(hbd_decoder* h, double min_abs_hz, double* applied_out)
{
    HBD_CHECK_H(h);
    std::lock_guard<std::mutex> l(h->mtx);
    cudaSetDevice(h->device);
    if (h->sync_groups()) return HBD_ERR_CUDA;
    const size_t n = size_t(h->n_ch);
    double* d_applied = nullptr;
    if (cudaMalloc((void**)&d_applied, n * sizeof(double)) != cudaSuccess) return HBD_ERR_NOMEM;
    std::vector<double> applied(n, 0.0);
    cudaError_t e = launch_afc_retune(h->d_state, h->n_ch, min_abs_hz, h->fs_in / h->factor, d_applied, h->fft_n, h->stream);
    ++h->launches;
    if (e == cudaSuccess) e = cudaMemcpyAsync(applied.data(), d_applied, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(d_applied);
    if (e != cudaSuccess) { h->set_error(cudaGetErrorString(e)); return HBD_ERR_CUDA; }
    int count = 0;
    for (size_t c = 0; c < n; ++c) {
        if (applied[c] != 0.0) { h->hc[c].nco_freq += applied[c]; h->nco_any = true; ++count; }
        if (applied_out) applied_out[c] = applied[c];
    }
    return count;
}
// all channels in one device->host copy: out[ch*6 + {0..5}] = correction, shift, noise floor, noise variance, peak l, peak r
size_t hbd_get_stats_batch(hbd_decoder* h, double* out, size_t cap_doubles)
{
    if (!h) return 0;
    std::lock_guard<std::mutex> l(h->mtx);
    const size_t n = size_t(h->n_ch);
    if (!out || cap_doubles < 6 * n) return 6 * n;
    cudaSetDevice(h->device);
    h->sync_groups();
    cudaStreamSynchronize(h->stream);
    std::vector<ChanState> st(n);
    if (cudaMemcpy(st.data(), h->d_state, n * sizeof(ChanState), cudaMemcpyDeviceToHost) != cudaSuccess) return 0;
    for (size_t c = 0; c < n; ++c) {
        out[6 * c + 0] = st[c].afc_correction; out[6 * c + 1] = st[c].afc_shift_hz;
        out[6 * c + 2] = st[c].afc_noise_floor; out[6 * c + 3] = st[c].afc_noise_var;
        out[6 * c + 4] = st[c].gui_left; out[6 * c + 5] = st[c].gui_right;
    }
    return 6 * n;
}

size_t hbd_get_spectrum_info(hbd_decoder* h, int ch, hbd_spectrum_info* info, float* power, size_t cap)
{
    if (!h || ch < 0 || ch >= h->n_ch || !info) return 0;
    std::vector<float> p(size_t(h->fft_n));
    const size_t n = hbd_get_power_spectrum(h, ch, p.data(), p.size());
    memset(info, 0, sizeof(*info));
    if (!n) return 0; // Decoder.h:818-819
    info->min_ = *std::min_element(p.begin(), p.end());
    info->max_ = *std::max_element(p.begin(), p.end());
    int pl = 0, pr = 0;
    hbd_get_peaks(h, ch, &pl, &pr);
    info->peak_left_ = std::abs(pl); info->peak_left_valid_ = pl > 0;
    info->peak_right_ = std::abs(pr); info->peak_right_valid_ = pr > 0;
    hbd_get_noise_floor(h, ch, &info->noise_floor_, &info->noise_variance_);
    info->sampling_rate_ = hbd_get_decimated_sampling_rate(h);
    info->shift_ = hbd_get_shift(h, ch);
    if (power && cap) memcpy(power, p.data(), std::min(cap, n) * sizeof(float));
    return n;
}

size_t hbd_debug_stage(hbd_decoder* h, int ch, int stage, float* out, size_t cap)
{
    if (!h || ch < 0 || ch >= h->n_ch) return 0;
    std::lock_guard<std::mutex> l(h->mtx);
    h->sync_ctrs();
    const HostChan& x = h->hc[size_t(ch)];
    switch (stage) {
    case HBD_STAGE_DECIMATED:
        if (!h->d_rec_dec) return 0;
        return fetch_floats(h, h->d_rec_dec + size_t(ch) * h->rec_pitch, 2 * size_t(x.last_n2), out, cap);
    case HBD_STAGE_FILTERED:
        if (!h->d_rec_filt) return 0;
        return fetch_floats(h, h->d_rec_filt + size_t(ch) * h->rec_pitch, 2 * size_t(x.last_nf), out, cap);
    case HBD_STAGE_DEMOD:
        if (!h->d_demod) return 0;
        return fetch_floats(h, h->d_demod + size_t(ch) * h->demod_pitch, size_t(x.last_nf), out, cap);
    case HBD_STAGE_LPTAPS:
        return fetch_floats(h, h->d_lptaps + size_t(ch) * kLpMaxTaps, x.lp_ntaps, out, cap);
    case HBD_STAGE_PENDING: {
        ChanState st; if (fetch_state(h, ch, &st)) return 0;
        if (!h->d_slicer) return 0;
        return fetch_floats(h, h->d_slicer + size_t(ch) * h->slicer_pitch, st.slicer_n, out, cap);
    }
    case HBD_STAGE_BITS: {
        if (!h->d_rec_bits) return 0;
        cudaSetDevice(h->device); h->sync_groups(); cudaStreamSynchronize(h->stream);
        unsigned nb = 0;
        cudaMemcpy(&nb, h->d_rec_bits_n + ch, 4, cudaMemcpyDeviceToHost);
        nb = std::min(nb, h->rec_bits_pitch);
        if (out && cap && nb) {
            std::vector<unsigned char> b(nb);
            cudaMemcpy(b.data(), h->d_rec_bits + size_t(ch) * h->rec_bits_pitch, nb, cudaMemcpyDeviceToHost);
            for (size_t i = 0; i < std::min<size_t>(nb, cap); ++i) out[i] = float(b[i]);
        }
        return nb;
    }
    default: return 0;
    }
}

size_t hbd_design_lowpass(float rel_width, float trans, size_t input_size, size_t current_taps, float* out, size_t cap)
{
    std::vector<float> taps;
    const size_t T = design_lowpass(rel_width, trans, input_size, current_taps, taps);
    if (out && T != current_taps) memcpy(out, taps.data(), std::min(cap, taps.size()) * sizeof(float));
    return T;
}

int hbd_extract_sentence(const char* stream, size_t n, char* callsign, char* data, char* crc, size_t cap, size_t* rest)
{
    SentenceMatch m;
    if (!extract_sentence(std::string(stream, n), m)) return 0;
    auto put = [cap](char* dst, const std::string& s) { if (dst && cap) { const size_t k = std::min(cap - 1, s.size()); memcpy(dst, s.data(), k); dst[k] = 0; } };
    put(callsign, m.callsign); put(data, m.data); put(crc, m.crc);
    if (rest) *rest = m.rest_offset;
    return 1;
}

// the text layer alone (test hook, no GPU): TextChannel::feed over n_chunks pushes of raw characters; `out` receives
// "<CRC-valid sentences, one per line>\x1e<last sentence>\x1e<text stream>"; returns the size of that report
size_t hbd_text_replay(const unsigned char* chars, const size_t* chunk_sizes, size_t n_chunks, char* out, size_t cap)
{
    if (!chars || !chunk_sizes) return 0;
    TextChannel tc;
    hbd_result_record pr; memset(&pr, 0, sizeof(pr));
    size_t off = 0;
    for (size_t c = 0; c < n_chunks; ++c) { tc.feed(chars + off, chunk_sizes[c], 0, SentenceSink(), true, pr); off += chunk_sizes[c]; }
    const std::string rep = tc.sentences_all(pr) + "\x1e" + tc.last_sentence + "\x1e" + tc.text_stream;
    if (out && cap) memcpy(out, rep.data(), std::min(cap, rep.size()));
    return rep.size();
}

void hbd_crc16(const char* s, size_t n, char out[5])
{
    const std::string r = crc16_hex(std::string(s, n));
    memcpy(out, r.data(), 4); out[4] = 0;
}

} // extern "C"
