// K3 -- RTTY bit slicer + UART deframer, ONE WARP PER CHANNEL.
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
// UART: the reference rescans its whole bit vector on every call; verdicts for positions
// whose frame was fully available never change, so the same characters come out of a
// shift-register automaton that carries < one frame of bits between calls (bounded state).
#include "hbd_common.cuh"
#include "slicer.cuh"

namespace hbd {

constexpr int kSlicerWarps = 4;

__device__ __forceinline__ int sgn3(float v) { return (0.0f < v) - (v < 0.0f); }

struct Means { float l, r; };

// SymbolExtractor.h:51-63 (FlipPointAvrg)
__device__ __forceinline__ Means window_means(const float* __restrict__ v, int n, int i, int R)
{
    const int lo = max(i - R, 0), hi = min(i + R, n);
    float sl = 0.f, sr = 0.f;
    for (int k = lo; k < i; ++k) sl = __fadd_rn(sl, v[k]);
    for (int k = i; k < hi; ++k) sr = __fadd_rn(sr, v[k]);
    Means m;
    m.l = __fdiv_rn(sl, float(i - lo));
    m.r = __fdiv_rn(sr, float(hi - i));
    return m;
}

// SymbolExtractor.h:162-224.  Returns 0 for "none".  Warp-uniform result.
__device__ int next_flip(const float* __restrict__ v, int n, int start, int spb, int R, int lane)
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
            const Means m = window_means(v, n, q, R);
            stop = sgn3(m.l) != sgn3(m.r);
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
            const Means m = window_means(v, n, q, R);
            stop = sgn3(m.l) == sgn3(m.r);
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
        if (q < p_end) {
            const Means m = window_means(v, n, q, R);
            const int d = (int)__fsub_rn(m.r, m.l);
            w = d < 0 ? -d : d;
        }
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
__device__ bool segment_mean_positive(const float* __restrict__ v, int a, int b, int lane)
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

__global__ void __launch_bounds__(kSlicerWarps * 32)
slicer_kernel(SlicerArgs a)
{
    const int lane = threadIdx.x & 31;
    const int chl = blockIdx.x * kSlicerWarps + (threadIdx.x >> 5);
    if (chl >= a.n_channels) return;
    const int ch = a.ch0 + chl;
    ChanState& st = a.state[ch];
    if (st.n_filtered == 0) return; // the reference only reaches the slicer after a low-pass/demod pass
    float* v = a.slicer + (size_t)ch * a.slicer_pitch;
    int n = int(st.slicer_n);

    const double fs = a.fs_dec, baud = st.baud;
    if (!fs || !baud) return;
    if (double(n) < fs / baud * 3) return;                 // SymbolExtractor.h:134
    const int spb = int(size_t(round(fs / baud)));          // :90
    const int R = max(4, int(spb / 4));                     // :170

    // UART automaton state (RTTY.h:77-137, see header)
    const int nbits = st.rtty_bits;
    const float nstops = st.rtty_stops;
    const bool uart_on = (nbits != 0 || nstops != 0.f) && nbits <= 16 && nstops <= 8.f;
    const int stop_chk = int(ceilf(nstops));               // stop bits inspected: s = 0 .. while s < nstops
    const int need = 1 + nbits + stop_chk;                  // bits that must be available at a position
    const int adv = 1 + nbits + int(nstops);                // i += nstops_ truncates (size_t += float)
    unsigned long long win = st.uart_win;                   // pending bits, LSB first
    int have = int(st.uart_n);
    unsigned char* rec_bits = a.rec_bits ? a.rec_bits + (size_t)ch * a.rec_bits_pitch : nullptr;
    unsigned rec_n = a.rec_bits ? a.rec_bits_n[ch] : 0;

    int last = 0, off = 0;
    bool any = false;
    for (;;) {
        const int flip = next_flip(v, n, off, spb, R, lane);
        if (flip == 0) break;
        any = true;
        const bool bit = segment_mean_positive(v, last, flip, lane);
        int cnt = int(size_t(roundf(__fdiv_rn(float(flip - last), float(spb)))));
        last = off = flip;
        // feed `cnt` copies of `bit` (warp-uniform scalar code; lane 0 commits the side effects)
        for (int c = 0; c < cnt; ++c) {
            if (rec_bits) { if (lane == 0 && rec_n < a.rec_bits_pitch) rec_bits[rec_n] = bit; ++rec_n; }
            if (!uart_on) continue;
            win |= (unsigned long long)(bit ? 1u : 0u) << have;
            ++have;
            while (have >= need) {
                const unsigned stops = unsigned(win >> (1 + nbits)) & ((1u << stop_chk) - 1u);
                const bool ok = ((win & 1ull) == 0ull) && stops == ((1u << stop_chk) - 1u);
                if (ok) {
                    const unsigned char cc = (unsigned char)((win >> 1) & ((1ull << nbits) - 1ull));
                    if (lane == 0) {
                        const unsigned pos = atomicAdd(a.log_head, 1u);
                        a.log[pos & (kLogCap - 1u)] = make_uint2(unsigned(ch), (a.call_seq << 8) | unsigned(cc));
                    }
                    win >>= adv; have -= adv;
                } else {
                    win >>= 1; have -= 1;
                }
            }
        }
    }
    if (lane == 0) {
        st.uart_win = win;
        st.uart_n = unsigned(have);
        if (a.rec_bits) a.rec_bits_n[ch] = rec_n;
    }
    if (!any) return;
    // erase consumed samples: SymbolExtractor.h:156-157
    const int erase = min(last, n);
    for (int base = 0; base < n - erase; base += 32) {
        const int k = base + lane;
        float x = 0.f;
        if (k < n - erase) x = v[k + erase];
        __syncwarp();
        if (k < n - erase) v[k] = x;
        __syncwarp();
    }
    if (lane == 0) st.slicer_n = unsigned(n - erase);
}

cudaError_t launch_slicer(const SlicerArgs& a, cudaStream_t stream, int* launches)
{
    const int grid = (a.n_channels + kSlicerWarps - 1) / kSlicerWarps;
    static bool configured = false;
    if (!configured) { cudaFuncSetAttribute(slicer_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared); configured = true; }
    slicer_kernel<<<grid, kSlicerWarps * 32, 0, stream>>>(a);
    if (launches) ++*launches;
    return cudaGetLastError();
}

} // namespace hbd
