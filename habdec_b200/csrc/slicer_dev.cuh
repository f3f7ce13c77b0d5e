// RTTY bit slicer + UART deframer as WARP-level device code (one warp per channel), used by the
// fused tail kernel (tail.cu).
//
//   SymbolExtractor<float>::operator()/findFlipPoints/findFirstFlipPoint
//                                   code/Decoder/SymbolExtractor.h:129-241 (+ helpers :32-63)
//   RTTY<bool>::operator()          code/Decoder/RTTY.h:77-137
//
// The reference algorithm is sequential and data dependent (edge-timed slicing: scan for the
// first position whose left/right window means differ in sign, scan on until they agree again,
// take the arg-max of the mean difference in between).  Here the 32 lanes evaluate 32
// consecutive candidate positions at once and a ballot finds the first one that ends each scan
// phase; the arg-max is a shuffle reduction with lowest-index tie break (std::max_element
// semantics).  All decisions are bit-exact restatements:
//   * window sums are accumulated left to right in float from 0.0f (std::accumulate order),
//   * the arg-max weight is |(int)(avg_r - avg_l)|: in the reference build the unqualified
//     abs() at SymbolExtractor.h:212 binds to ::abs(int) (verified against oracle/_ref),
//   * run length = (size_t)round(float(len) / float(spb)), bit = mean(segment) > 0.
// The long segment sum only decides a sign, so it is summed in parallel and re-done
// sequentially only when the parallel sum is too close to zero to be certain.
//
// UART: see uart_feed_run / uart_replay below (shift register + run-length coded backlog).
//
// `v` is a generic pointer: the pending samples normally sit in shared memory (staged by the
// tail kernel), and in global memory only when a channel has more pending samples than the
// staging area holds.
#pragma once
#include "hbd_common.cuh"

namespace hbd {

__device__ __forceinline__ int sgn3(float v) { return (0.0f < v) - (v < 0.0f); }

// Left/right window sums around position i (SymbolExtractor.h:51-63, FlipPointAvrg): left-to-right float sums from
// 0.0f over [i-R, i) and [i, i+R), clipped to [0, n).  Loads are issued eight at a time ahead of the dependent add
// chain.  Out-of-range elements are added as +0.0f, which leaves a running sum that started at +0.0f bit-identical
// (such a sum is never -0.0f).
struct WinSums { float sl, sr; int len_l, len_r; };

__device__ __forceinline__ WinSums window_sums(const float* __restrict__ v, int n, int i, int R)
{
    WinSums w;
    const int lo = max(i - R, 0), hi = min(i + R, n);
    w.len_l = i - lo; w.len_r = hi - i;
    float sl = 0.f, sr = 0.f;
    for (int j0 = 0; j0 < R; j0 += 8) {
        float xa[8], xb[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int ka = i - R + j0 + u, kb = i + j0 + u;
            xa[u] = (j0 + u < R && ka >= 0) ? v[ka] : 0.f;
            xb[u] = (j0 + u < R && kb < n) ? v[kb] : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) { sl = __fadd_rn(sl, xa[u]); sr = __fadd_rn(sr, xb[u]); }
    }
    w.sl = sl; w.sr = sr;
    return w;
}

// sgn(sum / len) without the division unless the quotient could underflow
__device__ __forceinline__ int sgn_mean(float s, int len)
{
    if (fabsf(s) >= 1e-30f) return (0.0f < s) - (s < 0.0f);
    return sgn3(__fdiv_rn(s, float(len)));
}

// |(int)(avg_r - avg_l)|, the reference's arg-max weight (::abs(int) binding, SymbolExtractor.h:212)
__device__ __forceinline__ int flip_weight(const WinSums& w)
{
    // |avg_r - avg_l| <= (|sr| + |sl|) / min(len): below 1 the truncation gives 0 whatever the roundings are
    if (fabsf(w.sl) + fabsf(w.sr) < 0.9f * float(min(w.len_l, w.len_r))) return 0;
    const int d = (int)__fsub_rn(__fdiv_rn(w.sr, float(w.len_r)), __fdiv_rn(w.sl, float(w.len_l)));
    return d < 0 ? -d : d;
}

// SymbolExtractor.h:162-224.  Returns 0 for "none".  Warp-uniform result.
__device__ __noinline__ int next_flip(const float* __restrict__ v, int n, int start, int spb, int R, int lane)
{
    if (n - start < spb) return 0;
    const int p0 = start + R, limit = n - spb;
    int first = 0, p_end = 0;
    // phase 1: first position whose two means differ in sign
    for (int base = p0;; base += 32) {
        const int q = base + lane;
        const bool abort = (q > p0) && (q >= limit);
        bool stop = false;
        if (!abort && q < n) {
            const WinSums w = window_sums(v, n, q, R);
            stop = sgn_mean(w.sl, w.len_l) != sgn_mean(w.sr, w.len_r);
        }
        const unsigned ma = __ballot_sync(0xffffffffu, abort), ms = __ballot_sync(0xffffffffu, stop);
        const unsigned any = ma | ms;
        if (any) {
            const int f = __ffs(any) - 1;
            if ((ma >> f) & 1u) return 0;
            first = base + f;
            break;
        }
    }
    // phase 2: first later position whose means agree in sign again
    for (int base = first + 1;; base += 32) {
        const int q = base + lane;
        const bool abort = q >= limit;
        bool stop = false;
        if (!abort) {
            const WinSums w = window_sums(v, n, q, R);
            stop = sgn_mean(w.sl, w.len_l) == sgn_mean(w.sr, w.len_r);
        }
        const unsigned ma = __ballot_sync(0xffffffffu, abort), ms = __ballot_sync(0xffffffffu, stop);
        const unsigned any = ma | ms;
        if (any) {
            const int f = __ffs(any) - 1;
            if ((ma >> f) & 1u) return 0;
            p_end = base + f;
            break;
        }
    }
    // arg-max of |(int)(r - l)| over [first, p_end), first maximum wins
    int best_w = -2147483647 - 1, best_i = first;
    bool have = false;
    for (int base = first; base < p_end; base += 32) {
        const int q = base + lane;
        int w = -2147483647 - 1;
        if (q < p_end) w = flip_weight(window_sums(v, n, q, R));
        int wi = q;
        // the weight is almost always 0 everywhere (discriminator swings are << 1 rad): then the first index wins
        if (__all_sync(0xffffffffu, w <= 0)) { w = 0; wi = base; }
        else {
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                const int ow = __shfl_xor_sync(0xffffffffu, w, o), oi = __shfl_xor_sync(0xffffffffu, wi, o);
                if (ow > w || (ow == w && oi < wi)) { w = ow; wi = oi; }
            }
        }
        if (!have || w > best_w) { best_w = w; best_i = wi; have = true; }
    }
    return best_i;
}

// ---- mask-driven search -------------------------------------------------------------------------------------
// The predicates evaluated by the scans depend only on the position (v, n and R are fixed during a call), so the
// tail kernel evaluates them for EVERY position with all its threads at once (slicer_build_masks) and the
// sequential flip search becomes find-first-set / find-first-clear on bit masks:
//   A[q]  = sgn(mean_l(q)) != sgn(mean_r(q))
//   NZ[q] = the arg-max weight |(int)(mean_r - mean_l)| may be non-zero (exact weights are then recomputed)
// Words below w_begin are already in place (the tail kernel's cache of earlier calls).
__device__ __forceinline__ void slicer_build_masks(const float* __restrict__ v, int n, int R, unsigned* maskA, unsigned* maskN,
                                                   int w_begin, int warp, int n_warps, int lane)
{
    const int n_words = (n + 31) >> 5;
    for (int w = w_begin + warp; w < n_words; w += n_warps) {
        const int q = w * 32 + lane;
        bool a = false, nz = false;
        if (q >= R && q < n) {
            const WinSums ws = window_sums(v, n, q, R);
            a = sgn_mean(ws.sl, ws.len_l) != sgn_mean(ws.sr, ws.len_r);
            nz = !(fabsf(ws.sl) + fabsf(ws.sr) < 0.9f * float(min(ws.len_l, ws.len_r)));
        }
        const unsigned ma = __ballot_sync(0xffffffffu, a), mn = __ballot_sync(0xffffffffu, nz);
        if (lane == 0) { maskA[w] = ma; maskN[w] = mn; }
    }
}

// first q in [from, to) whose mask bit equals want_set, -1 if none (warp-uniform)
__device__ __forceinline__ int find_bit(const unsigned* __restrict__ mask, int from, int to, bool want_set, int lane)
{
    if (from >= to) return -1;
    for (int w0 = from >> 5; w0 * 32 < to; w0 += 32) {
        const int w = w0 + lane;
        unsigned m = 0;
        if (w * 32 < to) {
            m = mask[w];
            if (!want_set) m = ~m;
            if (w == (from >> 5)) m &= 0xffffffffu << (from & 31);
            const int hi = to - w * 32;
            if (hi < 32) m &= (1u << hi) - 1u;
        }
        const unsigned b = __ballot_sync(0xffffffffu, m != 0u);
        if (b) {
            const int l = __ffs(b) - 1;
            const unsigned mm = __shfl_sync(0xffffffffu, m, l);
            return (w0 + l) * 32 + __ffs(mm) - 1;
        }
    }
    return -1;
}

// next_flip() on the masks.  Same result as the scanning version.
__device__ __forceinline__ int next_flip_masked(const float* __restrict__ v, const unsigned* __restrict__ maskA,
                                                const unsigned* __restrict__ maskN, int n, int start, int spb, int R, int lane)
{
    if (n - start < spb) return 0;
    const int p0 = start + R, limit = n - spb;
    int first;
    if (p0 < n && ((maskA[p0 >> 5] >> (p0 & 31)) & 1u)) first = p0;   // the first candidate is tested before the range check
    else {
        first = find_bit(maskA, p0 + 1, limit, true, lane);
        if (first < 0) return 0;
    }
    const int p_end = find_bit(maskA, first + 1, limit, false, lane);
    if (p_end < 0) return 0;
    if (find_bit(maskN, first, p_end, true, lane) < 0) return first;  // all weights 0: the first index wins
    int best_w = -2147483647 - 1, best_i = first;
    bool have = false;
    for (int base = first; base < p_end; base += 32) {
        const int q = base + lane;
        int w = -2147483647 - 1;
        if (q < p_end) w = flip_weight(window_sums(v, n, q, R));
        int wi = q;
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            const int ow = __shfl_xor_sync(0xffffffffu, w, o), oi = __shfl_xor_sync(0xffffffffu, wi, o);
            if (ow > w || (ow == w && oi < wi)) { w = ow; wi = oi; }
        }
        if (!have || w > best_w) { best_w = w; best_i = wi; have = true; }
    }
    return best_i;
}

// sign of the left-to-right float sum of v[a..b) decided exactly: parallel sum, sequential fallback
__device__ __forceinline__ bool segment_mean_positive(const float* __restrict__ v, int a, int b, int lane)
{
    float s = 0.f, sa = 0.f;
    for (int k = a + lane; k < b; k += 32) { const float x = v[k]; s += x; sa += fabsf(x); }
#pragma unroll
    for (int o = 16; o; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); sa += __shfl_xor_sync(0xffffffffu, sa, o); }
    const float len = float(b - a);
    // both summation orders are within (len * 2^-24 * sum|v|) of the exact sum; 2^-21 leaves 4x margin,
    // and the quotient sum/len cannot underflow to zero while |sum| clears this bound
    const float bound = len * sa * 4.76837158203125e-7f + 1e-30f;
    if (fabsf(s) > bound) return s > 0.f;
    float seq = 0.f;
    if (lane == 0) for (int k = a; k < b; ++k) seq = __fadd_rn(seq, v[k]);
    seq = __shfl_sync(0xffffffffu, seq, 0);
    return __fdiv_rn(seq, len) > 0.f;
}

// decoded characters of this call, collected per channel and appended to the device-wide log in one go
struct CharSink {
    unsigned char* buf;   // shared memory, kCharBuf bytes (one warp writes, lane 0 only)
    int n;
    uint2* log; unsigned* ctl; unsigned log_mask; unsigned call_seq; unsigned ch;   // ctl: LogCtl words
    unsigned char* ring; unsigned ring_total;   // SSDV: the channel's raw-character ring (null: off) and its append count
};
constexpr unsigned kRawRingMask = 4096u - 1u;    // == kSsdvRing - 1 (ssdv.cuh)
constexpr int kCharBuf = 64;

__device__ __forceinline__ void sink_flush(CharSink& s, int lane)
{
    // warp-uniform: s.n is kept identical in all lanes
    if (s.n == 0) return;
    unsigned pos = 0, tail = 0;
    if (lane == 0) {
        pos = atomicAdd(s.ctl + kCtlCharHead, unsigned(s.n));
        tail = *reinterpret_cast<volatile unsigned*>(s.ctl + kCtlCharTail);
    }
    pos = __shfl_sync(0xffffffffu, pos, 0);
    tail = __shfl_sync(0xffffffffu, tail, 0);
    __syncwarp();
    // never overwrite entries the host has not read yet: a full ring drops the characters and counts them
    // (the host reports the loss, api.cu collect_locked; it drains early enough that this does not happen, log_pressure)
    if (pos + unsigned(s.n) - tail > s.log_mask + 1u) {
        if (lane == 0) atomicAdd(s.ctl + kCtlCharOvf, unsigned(s.n));
    } else {
        for (int i = lane; i < s.n; i += 32) s.log[(pos + unsigned(i)) & s.log_mask] = make_uint2(s.ch, (s.call_seq << 8) | unsigned(s.buf[i]));
    }
    if (s.ring) {
        for (int i = lane; i < s.n; i += 32) s.ring[(s.ring_total + unsigned(i)) & kRawRingMask] = s.buf[i];
        s.ring_total += unsigned(s.n);
    }
    __syncwarp();
    s.n = 0;
}

// samples per bit and window radius; false when the slicer does not run at all this call
__device__ __forceinline__ bool slicer_geometry(int n, double fs, double baud, int& spb, int& R)
{
    if (!fs || !baud) return false;
    if (double(n) < fs / baud * 3) return false;           // SymbolExtractor.h:134
    spb = int(size_t(round(fs / baud)));                    // :90
    R = max(4, int(spb / 4));                               // :170
    return true;
}

// ---- UART deframer (RTTY<bool>::operator(), RTTY.h:77-137) -----------------------------------------------------------
// The reference keeps every bit since the last decoded character (bits_) and rescans all of them on every call with the
// framing in force at that moment.  While the framing does not change, the verdict of a position whose frame was
// completely available never changes, so the same characters come out of a shift register that holds the undecided
// suffix only (win / have, < one frame).  What the shift register cannot reproduce is a framing CHANGE (rtty_bits /
// rtty_stops setters between calls, or the step from "not configured" to configured): the reference then re-evaluates
// the whole backlog under the new framing.  So the backlog is kept as well, run-length coded in HBM (one entry per
// slicer run; after a decode it shrinks to what the reference keeps: nothing, or the one stop bit a fractional stop
// count leaves behind), and replayed through the shift register when the host flags a change (ChanState::uart_rescan).
struct UartFraming {
    bool on;        // false: nbits == 0 && nstops == 0 (RTTY.h:79-80), or a frame that does not fit the 64-bit window
    int nbits, stop_chk, need, adv;
};
__device__ __forceinline__ UartFraming uart_framing(int nbits, float nstops)
{
    UartFraming f;
    f.on = (nbits != 0 || nstops != 0.f) && nbits >= 0 && nbits <= 16 && nstops >= 0.f && nstops <= 8.f;
    f.nbits = nbits;
    f.stop_chk = int(ceilf(nstops));          // stop bits inspected: s = 0 .. while s < nstops
    f.need = 1 + nbits + f.stop_chk;          // bits that must be available at a position
    f.adv = 1 + nbits + int(nstops);          // i += nstops_ truncates (size_t += float)
    return f;
}
struct UartState {
    unsigned long long win; int have;         // undecided suffix of the bit stream, LSB = oldest
    unsigned short* runs; unsigned n_runs;    // backlog since the last decoded character: (count << 1 | bit) per slicer run
    unsigned ovf;                             // the backlog outgrew kUartRunsCap (its newest part is missing)
};

// feed `cnt` copies of `bit` (warp-uniform scalar code; lane 0 commits the side effects)
__device__ __forceinline__ void uart_feed_run(UartState& u, const UartFraming& f, bool bit, int cnt, CharSink& sink, int lane)
{
    int keep = cnt;                            // bits of this run that stay in the backlog
    if (f.on) {
        for (int c = 0; c < cnt; ++c) {
            u.win |= (unsigned long long)(bit ? 1u : 0u) << u.have;
            ++u.have;
            while (u.have >= f.need) {
                const unsigned stops = unsigned(u.win >> (1 + f.nbits)) & ((1u << f.stop_chk) - 1u);
                const bool ok = ((u.win & 1ull) == 0ull) && stops == ((1u << f.stop_chk) - 1u);
                if (ok) {
                    const unsigned char cc = (unsigned char)((u.win >> 1) & ((1ull << f.nbits) - 1ull));
                    if (lane == 0) sink.buf[sink.n] = cc;
                    if (++sink.n == kCharBuf) sink_flush(sink, lane);
                    u.win >>= f.adv; u.have -= f.adv;
                    // RTTY.h:133-134: everything up to the last stop bit of this character is erased.  What is left in the
                    // window (0 or 1 bit: the last inspected stop bit when nstops is fractional) is the bit just fed.
                    u.n_runs = 0; u.ovf = 0;
                    keep = cnt - c - 1 + u.have;
                } else {
                    u.win >>= 1; u.have -= 1;
                }
            }
        }
    }
    if (keep > 0) {
        if (u.n_runs < kUartRunsCap) { if (lane == 0) u.runs[u.n_runs] = (unsigned short)((min(keep, 32767) << 1) | (bit ? 1 : 0)); ++u.n_runs; }
        else u.ovf = 1;
    }
}

// the framing changed: re-evaluate the backlog from its start, as the reference's next rtty_() does.  In place: every
// replayed run is read before it is fed, and feeding it appends at most one entry.
__device__ __forceinline__ void uart_replay(UartState& u, const UartFraming& f, CharSink& sink, int lane)
{
    const unsigned n_old = u.n_runs;
    const unsigned was_ovf = u.ovf;
    u.n_runs = 0; u.win = 0ull; u.have = 0;
    for (unsigned r = 0; r < n_old; ++r) {
        const unsigned e = u.runs[r];          // same address in every lane
        __syncwarp();
        uart_feed_run(u, f, (e & 1u) != 0u, int(e >> 1), sink, lane);
        __syncwarp();
    }
    if (was_ovf) { u.win = 0ull; u.have = 0; u.ovf = 1; }   // a gap follows the kept part: start afresh behind it
}

// One warp slices the pending samples v[0..n) of a channel.  Returns the number of samples to erase from the
// front (0 if no flip point was found).  UART state goes through `u`; characters go to `sink`.
// maskA/maskN: position masks from slicer_build_masks, or null (scan on the fly).
__device__ __forceinline__ int slice_channel(const float* __restrict__ v, int n, int spb, int R, const unsigned* maskA, const unsigned* maskN,
                                             int nbits, float nstops, UartState& u, bool& rescan, CharSink& sink,
                                             unsigned char* rec_bits, unsigned& rec_n, unsigned rec_cap, int lane)
{
    const UartFraming f = uart_framing(nbits, nstops);
    int last = 0, off = 0;
    for (;;) {
        const int flip = maskA ? next_flip_masked(v, maskA, maskN, n, off, spb, R, lane) : next_flip(v, n, off, spb, R, lane);
        if (flip == 0) break;
        const bool bit = segment_mean_positive(v, last, flip, lane);
        const int cnt = int(size_t(roundf(__fdiv_rn(float(flip - last), float(spb)))));
        last = off = flip;
        if (cnt <= 0) continue;
        if (rec_bits) { for (int c = 0; c < cnt; ++c) { if (lane == 0 && rec_n < rec_cap) rec_bits[rec_n] = bit; ++rec_n; } }
        if (rescan) { uart_replay(u, f, sink, lane); rescan = false; }   // Decoder.h:562-566: rtty_() runs when new symbols arrive
        uart_feed_run(u, f, bit, cnt, sink, lane);
    }
    return min(last, n);
}

} // namespace hbd
