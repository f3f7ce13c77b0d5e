// K2 -- everything after the first decimator, one CTA (128 threads) per channel, ONE kernel:
//
//   stage-2 FIR decimator     code/Decoder/Decimator.h:99-146 (second entry of Decoder.h:286-320)
//   DC removal (optional)     code/Decoder/Decoder.h:450-459
//   FFT frame assembly        code/Decoder/Decoder.h:467-473   (spectrum taps the stream BEFORE the low-pass)
//   batch-of-256 gate + AFC tick  Decoder.h:492-509, >160 kS/s cut-off Decoder.h:522-527
//   low-pass FIR              code/Decoder/FirFilter.h:117-169 (taps designed on the host, host_tail.cpp)
//   FM/FSK discriminator      code/Decoder/FSK2_Demod.h:30-42  (carry kept PER CHANNEL, not per thread)
//   slicer input append       code/Decoder/SymbolExtractor.h:108-125 (3e4 safety vent included)
//   bit slicer + UART         SymbolExtractor.h:129-241, RTTY.h:77-137 (warp 0, slicer_dev.cuh)
//
// This kernel runs in the shadow of the HBM-bound K1 (other channel group / other stream) with only a couple of
// CTAs per SM, while K1 keeps the memory system saturated: what costs time here is the number of DEPENDENT global
// round trips, not flops.  So the whole chain stays in shared memory:
//   round trip 1   plan (kernel argument when uniform) -> state, stage-1 window, low-pass history + pending,
//                  taps: all issued at once (cp.async, no registers held)
//   round trip 2   the channel's pending slicer samples (their count is in the state); overlaps the FIR work
//   then           stage 2 -> [DC] -> FFT frame -> low-pass -> discriminator -> slicer -> UART, all on shared memory
//   finally        write-only stores of the new histories / queues / state
// Both FIRs are register tiled: a thread owns TWO consecutive outputs and slides a small sample window through
// registers, so shared memory is read once per sample pair (LDS.128) instead of once per multiply, and every
// complex-by-real multiply-add is one packed FFMA2:
//   stage 2 (M2 = 4):  per tap group of 4:  2 LDS.128 samples + 1 LDS.128 taps -> 8 FFMA2
//   low-pass:          per tap pair:        1 LDS.128 samples + 1 LDS.64 taps  ->  4 FFMA2
// The stage-2 window tile is padded by 2 samples every 16 so that the 64-byte thread stride is bank-conflict free.
#include "hbd_common.cuh"
#include "slicer_dev.cuh"
#include "afc_dev.cuh"
#include "tail.cuh"
#include <algorithm>
#include <cstdlib>

namespace hbd {

constexpr int kTailThreads = 128;
constexpr int kTile = 2 * kTailThreads; // outputs per tile (two per thread) == kLpBatch
static_assert(kTile == kLpBatch, "one low-pass batch per stage-2 tile");

__device__ __forceinline__ float2 cmul_conj_ieee(float2 a, float2 b) // a * conj(b), separately rounded products (no FMA)
{
    float2 r;
    r.x = __fadd_rn(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y));
    r.y = __fadd_rn(__fmul_rn(a.y, b.x), -__fmul_rn(a.x, b.y));
    return r;
}

__host__ __device__ __forceinline__ int pad16(int s) { return s + 2 * (s >> 4); } // 2 float2 of padding per 16 samples

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async8(void* dst_smem, const void* src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_addr(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(void* dst_smem, const void* src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_addr(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// out of line on purpose: the float64 state machine would otherwise raise the register count of the whole kernel
__device__ __noinline__ void afc_step_staged(ChanState& staged, ChanState& global_state, double fs_dec, int n_fft)
{
    afc_step(staged, fs_dec, n_fft);
    afc_store(global_state, staged);
}

__global__ void __launch_bounds__(kTailThreads + 32)
tail_kernel(TailArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* s_x = reinterpret_cast<float2*>(smem_raw);            // stage-2 window tile (padded layout)
    float2* s_q = s_x + a.xw;                                     // low-pass input: [skew | history | pending | new]
    float2* s_f = s_q + a.qcap;                                   // filtered tile (+1 previous sample)
    float*  s_h2 = reinterpret_cast<float*>(s_f + kTile + 2);     // stage-2 taps (shifted, zero padded)
    float*  s_hl = s_h2 + a.h2cap;                                // low-pass taps (shifted, zero padded)
    float*  s_v  = s_hl + a.hlcap;                                // slicer samples: [old pending | this call's]
    unsigned* s_maskA = reinterpret_cast<unsigned*>(s_v + a.sv_cap);   // flip-search position masks over s_v (slicer_dev.cuh)
    unsigned* s_maskN = s_maskA + a.mask_words;
    unsigned char* s_chars = reinterpret_cast<unsigned char*>(s_maskN + a.mask_words);
    __shared__ ChanState s_st;
    __shared__ float2 sh_dc_wp;

    const int ch = a.ch0 + blockIdx.x, tid = threadIdx.x, lane = tid & 31;
    const ChanPlan pl = a.uniform ? a.uplan : a.plan[ch];
    const float2* s1 = a.s1 + (size_t)ch * a.s1_pitch;
    float2* s1n = a.s1_next + (size_t)ch * a.s1_pitch;
    if (pl.flags & 1u) { // fewer than `factor` samples queued (Decoder.h:429-430): only carry the stage-2 history over
        if (a.M2 > 1 && tid < kTailThreads) for (int i = tid; i < a.T2 - 1; i += kTailThreads) s1n[kS1Hist - (a.T2 - 1) + i] = s1[kS1Hist - (a.T2 - 1) + i];
        return;
    }
    if (tid >= kTailThreads) {
        // Warp 4, the AFC warp.  AFC<float>::process on a call that completes no FFT frame (15 of 16 calls at 256 decimated
        // samples per call): the reference re-runs it on the stale spectrum (Decoder.h:501-509); only the state machine
        // moves, on the statistics K4 cached for that spectrum.  One thread steps it while the four signal warps run the
        // FIRs and the slicer, so K4 is not launched at all for such a call.  The warp joins the two barriers of round
        // trip 1 (the staged state) and retires; barriers after that count the remaining warps only.
        __syncthreads();
        cp_async_wait_all();
        __syncthreads();
        if (tid == kTailThreads) {
            const unsigned total = pl.dec_pending + pl.n2;
            const unsigned have0 = s_st.fft_have, fn = unsigned(a.fft_n);
            const unsigned take = (have0 < fn) ? hbd_min_u(fn - have0, pl.n2) : 0u;
            const bool done = take && have0 + take >= fn;
            if (total >= unsigned(kLpBatch) && !done) afc_step_staged(s_st, a.state[ch], a.fs_dec, a.fft_n);
        }
        return;
    }

    ChanState& gst = a.state[ch];
    const unsigned n2 = pl.n2;
    float2* dq = a.decq + (size_t)ch * a.dq_pitch;
    const int T = int(pl.lp_ntaps);                      // 0 until the first design (host, Decoder.h:536-538)
    const int hist = T > 0 ? T - 1 : 0;
    const unsigned dec_pending = pl.dec_pending;
    const int T2 = a.T2, M2 = a.M2;
    // stage-2 geometry (M2 == 4): window starts `lead` samples before the tile, lead = T2-1 rounded up to 4
    const int lead = (M2 == 4) ? ((T2 - 1 + 3) & ~3) : (T2 - 1);
    const int skew2 = lead - (T2 - 1);
    const int n_blocks = (skew2 + T2 + 3) / 4;
    // Low-pass queue in HBM (one row per channel): slots [0, kLpHist) mirror the head of the reference's work buffer
    // (FirFilter.h:139-160: [T-1 history | the inputs of the last call]), the samples waiting for a 256-batch follow at
    // kLpHist.  In shared memory: s_q[qskew ..] = history, then the queue; qskew keeps the LDS.128 of the filter aligned.
    const int qskew = (T > 0 && (hist & 1)) ? 1 : 0;
    const int n_pairs = (qskew + T + 2) / 2;

    // ---- round trip 1 -------------------------------------------------------------------------------------
    for (int i = tid; i < a.qcap; i += kTailThreads) s_q[i] = make_float2(0.f, 0.f); // zero taps must meet finite samples
    __syncthreads();
    {
        const unsigned* src = reinterpret_cast<const unsigned*>(&gst);
        unsigned* dst = reinterpret_cast<unsigned*>(&s_st);
        for (int i = tid; i < int(sizeof(ChanState) / 4); i += kTailThreads) cp_async4(dst + i, src + i);
    }
    auto load_window = [&](unsigned k0, unsigned nk) {
        if (M2 > 1) {
            const long long x0 = (long long)k0 * M2 - lead;
            const int win = int(nk) * M2 + lead;
            if (M2 == 4) {   // two samples per copy: x0, the window length, the row start and the pad layout are all even
                for (int i = 2 * tid; i < win; i += 2 * kTailThreads) cp_async16(&s_x[pad16(i)], &s1[kS1Hist + x0 + i]);
            }
            else         { for (int i = tid; i < win; i += kTailThreads) cp_async8(&s_x[i], &s1[kS1Hist + x0 + i]); }
        }
    };
    load_window(0, hbd_min_u(kTile, n2));
    for (int i = tid; i < hist + int(dec_pending); i += kTailThreads) cp_async8(&s_q[qskew + i], &dq[i < hist ? i : kLpHist + (i - hist)]);
    cp_async_commit();
    if (M2 == 4) { for (int u = tid; u < 4 * n_blocks; u += kTailThreads) s_h2[u] = (u >= skew2 && u - skew2 < T2) ? a.taps2[u - skew2] : 0.f; }
    else if (M2 > 1) { for (int i = tid; i < T2; i += kTailThreads) s_h2[i] = a.taps2[i]; }
    {
        const float* taps = a.lptaps + (size_t)ch * kLpMaxTaps;
        for (int u = tid; u < 2 * n_pairs; u += kTailThreads) s_hl[u] = (u >= qskew && u - qskew < T) ? taps[u - qskew] : 0.f;
    }
    cp_async_wait_all();
    __syncthreads();

    // ---- batch gate (Decoder.h:492-527) and slicer vent (SymbolExtractor.h:116-120) ------------------------
    const unsigned total = dec_pending + n2;
    unsigned nf = 0, new_pending = total;
    bool tick = false;
    if (total >= unsigned(kLpBatch)) {
        tick = true;
        if (a.fs_dec > 4 * 40e3) new_pending = 0;     // everything queued is dropped, nothing decoded
        else { nf = total - total % unsigned(kLpBatch); new_pending = total - nf; }
    }
    unsigned n_old = s_st.slicer_n;
    if (nf && n_old > unsigned(kSlicerVent)) n_old = 0; // checked before the append, only when there is something to append
    const bool smem_mode = nf && (n_old + nf <= unsigned(a.sv_cap));
    float* v_g = a.slicer + (size_t)ch * a.slicer_pitch;
    // Position masks of the pending samples.  A position's two bits depend on v[q-R, q+R) only, so the bits of every
    // position whose right window was complete in the call that built them stay true while the samples stay; the kernel
    // keeps them in HBM (shifted by what the slicer erased) and a call evaluates only the positions behind them -- the
    // appended samples plus one radius -- instead of all n_old + nf.  Whole words are reused; the build restarts at the
    // word that holds the first unknown position.
    unsigned* mask_g = a.maskc ? a.maskc + (size_t)ch * (2 * kMaskWords) : nullptr;
    unsigned m_words = 0;
    if (smem_mode && mask_g && n_old == s_st.slicer_n) {
        int spb0 = 0, R0 = 0;
        if (slicer_geometry(int(n_old + nf), a.fs_dec, s_st.baud, spb0, R0) && R0 == s_st.mask_R && n_old > unsigned(R0))
            m_words = hbd_min_u(hbd_min_u(s_st.mask_valid, n_old - unsigned(R0)) >> 5, kMaskWords);
    }
    // ---- round trip 2: pending slicer samples and their cached masks (consumed after the FIRs) ---------------
    if (smem_mode) {
        for (unsigned i = tid; i < n_old; i += kTailThreads) cp_async4(&s_v[i], &v_g[i]);
        for (unsigned i = tid; i < m_words; i += kTailThreads) { cp_async4(&s_maskA[i], &mask_g[i]); cp_async4(&s_maskN[i], &mask_g[kMaskWords + i]); }
    }
    cp_async_commit();

    const unsigned fft_have0 = s_st.fft_have;
    const unsigned fft_n = unsigned(a.fft_n);
    const unsigned fft_take = (fft_have0 < fft_n) ? hbd_min_u(fft_n - fft_have0, n2) : 0u;
    const bool frame_done = fft_take && fft_have0 + fft_take >= fft_n;   // K4 transforms it (and steps the AFC) in this call
    const bool dc = s_st.dc_remove != 0;
    float2 carry_prev = make_float2(s_st.demod_last_re, s_st.demod_last_im);
    const bool primed = s_st.demod_primed != 0;
    float* dlast = a.demod_last ? a.demod_last + (size_t)ch * a.demod_pitch : nullptr;

    int qpos = qskew;              // s_q index where the current low-pass history starts
    unsigned have_q = dec_pending; // samples queued behind the history
    unsigned produced = 0;         // low-pass outputs so far in this call

    for (unsigned k0 = 0; k0 < n2; k0 += kTile) {
        const unsigned nk = hbd_min_u(kTile, n2 - k0);
        float2* ynew = s_q + qpos + hist + have_q;  // where this tile's decimated samples go
        if (k0) { // later tiles of a long call: the window was not prefetched
            __syncthreads();
            load_window(k0, nk);
            cp_async_commit();
            cp_async_wait_all();
            __syncthreads();
        }
        // ---- stage 2 ------------------------------------------------------------------------------------------
        if (M2 == 4) {
            // y[k] = sum_t x[4k - (T2-1) + t] h[t]; the tile holds x from 4*k0 - lead (a multiple of 4 samples, keeps
            // LDS.128 aligned); taps are applied as h'[u] = h[u - skew2] (zero for u < skew2)
            const int k = 2 * tid; // first of this thread's two outputs
            if (k < int(nk)) {
                float2 acc0 = make_float2(0.f, 0.f), acc1 = make_float2(0.f, 0.f);
                float2 b0[4], b1[4];
                // The padded tile in units of float4 (two samples): a quad of 4 samples q sits at F(q) = 9 (q >> 2) + 2 (q & 3)
                // (16 samples + 2 pad slots = 9 float4 per group).  Output k starts at quad k; block `blk` brings quad
                // k + blk + 1.  Four running indices, one per position in the group, each stepping one group per round of
                // four blocks: no address arithmetic in the loop beyond four adds per round.
                const float4* sx4 = reinterpret_cast<const float4*>(s_x);
                const float4* sh4 = reinterpret_cast<const float4*>(s_h2);
                auto F = [](int q) { return 9 * (q >> 2) + 2 * (q & 3); };
                {
                    const float4 p0 = sx4[F(k)], p1 = sx4[F(k) + 1];
                    b0[0] = make_float2(p0.x, p0.y); b0[1] = make_float2(p0.z, p0.w); b0[2] = make_float2(p1.x, p1.y); b0[3] = make_float2(p1.z, p1.w);
                }
                auto step = [&](int f, int blk) {
                    const float4 q0 = sx4[f], q1 = sx4[f + 1];
                    b1[0] = make_float2(q0.x, q0.y); b1[1] = make_float2(q0.z, q0.w); b1[2] = make_float2(q1.x, q1.y); b1[3] = make_float2(q1.z, q1.w);
                    const float4 h4 = sh4[blk];                // taps for tile offsets 4*blk .. 4*blk+3
                    const float hh[4] = {h4.x, h4.y, h4.z, h4.w};
#pragma unroll
                    for (int c = 0; c < 4; ++c) { acc0 = cfma(b0[c], hh[c], acc0); acc1 = cfma(b1[c], hh[c], acc1); }
#pragma unroll
                    for (int c = 0; c < 4; ++c) b0[c] = b1[c];
                };
                int f0 = F(k + 1), f1 = F(k + 2), f2 = F(k + 3), f3 = F(k + 4);
                int blk = 0;
#pragma unroll 1
                for (; blk + 4 <= n_blocks; blk += 4) {
                    step(f0, blk); step(f1, blk + 1); step(f2, blk + 2); step(f3, blk + 3);
                    f0 += 9; f1 += 9; f2 += 9; f3 += 9;
                }
                for (; blk < n_blocks; ++blk) step(F(k + blk + 1), blk);
                ynew[k] = acc0;
                if (k + 1 < int(nk)) ynew[k + 1] = acc1;
            }
        } else if (M2 > 1) { // M2 == 2 (69 taps): half-rate inputs (<= 512 kS/s), simple form
            for (int k = tid; k < int(nk); k += kTailThreads) {
                const float2* w = s_x + k * M2;
                float2 acc = make_float2(0.f, 0.f);
                for (int t = 0; t < T2; ++t) acc = cfma(w[t], s_h2[t], acc);
                ynew[k] = acc;
            }
        } else {
            for (unsigned k = tid; k < nk; k += kTailThreads) ynew[k] = s1[kS1Hist + k0 + k];
        }
        // stage-2 history for the next call: the last T2-1 stage-1 samples, taken from the last tile's window
        if (M2 > 1 && k0 + nk == n2) {
            const int win = int(nk) * M2 + lead;
            // The reference runs the stage IN PLACE (Decoder.h:441-446): when Decimator.h:141-143 copies the last T2-1
            // inputs, the n2 outputs already sit on the head of the same buffer.  For short calls (n1 - (T2-1) < n2,
            // i.e. n1 < 184 at M2 = 4) the head of the history therefore holds stage-2 OUTPUTS, not stage-1 samples.
            const int p0 = int(n2) * M2 - (T2 - 1);          // position of history sample 0 in this call's stage-1 stream
            const bool in_place_overlap = p0 >= 0 && p0 < int(n2);
            if (in_place_overlap) __syncthreads();            // ynew[] of the (single) tile is complete
            for (int i = tid; i < T2 - 1; i += kTailThreads) {
                const int w = win - (T2 - 1) + i;
                float2 v = s_x[M2 == 4 ? pad16(w) : w];
                if (in_place_overlap && p0 + i < int(n2)) v = ynew[p0 + i];
                s1n[kS1Hist - (T2 - 1) + i] = v;
            }
        }
        __syncthreads();

        // ---- DC removal: sequential recurrence re-seeded from the first sample of the call ----------------------
        if (dc) {
            if (tid == 0) {
                float2 wp = k0 ? sh_dc_wp : make_float2(__fmul_rn(.97f, ynew[0].x), __fmul_rn(.97f, ynew[0].y));
                for (unsigned i = 0; i < nk; ++i) {
                    const float2 x = ynew[i];
                    const float2 w = make_float2(__fadd_rn(x.x, __fmul_rn(.97f, wp.x)), __fadd_rn(x.y, __fmul_rn(.97f, wp.y)));
                    ynew[i] = make_float2(__fadd_rn(w.x, -wp.x), __fadd_rn(w.y, -wp.y));
                    wp = w;
                }
                sh_dc_wp = wp;
            }
            __syncthreads();
        }
        // ---- per-call copy of the decimated block (parity tests) and FFT frame (first fft_take samples) ---------
        if (a.rec_decimated) {
            float2* rec = a.rec_decimated + (size_t)ch * a.rec_pitch;
            for (unsigned k = tid; k < nk; k += kTailThreads) rec[k0 + k] = ynew[k];
        }
        if (k0 < fft_take) {
            float2* fb = a.fftbuf + (size_t)ch * fft_n + fft_have0;
            for (unsigned k = tid; k < nk && k0 + k < fft_take; k += kTailThreads) fb[k0 + k] = ynew[k];
        }
        have_q += nk;
        // above 160 kS/s nothing is decoded and the whole queue is dropped (Decoder.h:522-527): the tiles of a long call must
        // not pile up in the shared-memory queue (a factor-1 call at MS/s rates brings hundreds of them)
        if (tick && nf == 0) have_q = 0;

        // ---- low-pass FIR + discriminator on one batch of 256 --------------------------------------------------
        if (produced < nf && have_q >= unsigned(kLpBatch)) {
            const float2* q = s_q + (qpos - qskew);  // even index: LDS.128 stays aligned
            const int o = 2 * tid; // outputs o, o+1; output o uses samples q[o + qskew + t]
            float2 acc0 = make_float2(0.f, 0.f), acc1 = make_float2(0.f, 0.f);
            float2 w0, w1;
            {
                const float4 p = *reinterpret_cast<const float4*>(&q[o]);
                w0 = make_float2(p.x, p.y); w1 = make_float2(p.z, p.w);
            }
#pragma unroll 4
            for (int pr = 0; pr < n_pairs; ++pr) {
                const float4 p = *reinterpret_cast<const float4*>(&q[o + 2 * pr + 2]);
                const float2 w2 = make_float2(p.x, p.y), w3 = make_float2(p.z, p.w);
                const float2 hp = *reinterpret_cast<const float2*>(&s_hl[2 * pr]); // taps for offsets 2pr, 2pr+1
                acc0 = cfma(w0, hp.x, acc0); acc0 = cfma(w1, hp.y, acc0);
                acc1 = cfma(w1, hp.x, acc1); acc1 = cfma(w2, hp.y, acc1);
                w0 = w2; w1 = w3;
            }
            s_f[o + 1] = acc0;
            s_f[o + 2] = acc1;
            if (tid == 0) s_f[0] = (produced == 0) ? (primed ? carry_prev : acc0) : carry_prev;
            cp_async_wait_all();       // the pending slicer samples have landed (smem mode)
            __syncthreads();
            const float2 prev = s_f[o];
            const float2 p0 = cmul_conj_ieee(acc0, prev), p1 = cmul_conj_ieee(acc1, acc0);
            const float d0 = atan2f(p0.y, p0.x), d1 = atan2f(p1.y, p1.x);
            float* pend = (smem_mode ? s_v : v_g) + n_old;
            pend[produced + o] = d0; pend[produced + o + 1] = d1;
            if (dlast) { dlast[produced + o] = d0; dlast[produced + o + 1] = d1; }
            if (a.rec_filtered) {
                a.rec_filtered[(size_t)ch * a.rec_pitch + produced + o] = acc0;
                a.rec_filtered[(size_t)ch * a.rec_pitch + produced + o + 1] = acc1;
            }
            // the reference leaves this call's inputs behind the history in its work buffer (FirFilter.h:149-152); a later,
            // LONGER filter takes them for history (:139-147 do not clear a buffer that is large enough), so they are kept too
            {
                const int mi = hist + int(produced) + o;
                const float2* qin = s_q + qpos + hist;
                if (mi < kLpHist) dq[mi] = qin[o];
                if (mi + 1 < kLpHist) dq[mi + 1] = qin[o + 1];
            }
            carry_prev = s_f[kTile]; // last filtered sample of this batch (same value in every thread)
            produced += kLpBatch;
            have_q -= kLpBatch;
            qpos += kLpBatch;
        }
        // ---- more tiles follow: slide the queue back to the front of s_q ---------------------------------------
        if (k0 + nk < n2 && qpos != qskew) {
            const int n_move = hist + int(have_q);
            for (int m0 = 0; m0 < n_move; m0 += kTailThreads) {
                const int m = m0 + tid;
                __syncthreads();
                float2 t = make_float2(0.f, 0.f);
                if (m < n_move) t = s_q[qpos + m];
                __syncthreads();
                if (m < n_move) s_q[qskew + m] = t;
            }
            qpos = qskew;
        }
    }
    __syncthreads();

    // ---- persist: low-pass history + unfiltered remainder (write only) ---------------------------------------
    if (!(tick && nf == 0)) { // the >160 kS/s cut-off drops the queue and leaves the history alone
        const int n_keep = hist + int(have_q);
        for (int m = tid; m < n_keep; m += kTailThreads) dq[m < hist ? m : kLpHist + (m - hist)] = s_q[qpos + m];
    }

    // ---- slicer: position masks by all threads, then the sequential search + UART on warp 0 ----------------------
    int spb = 0, R = 0;
    const int n_sl = int(n_old + nf);
    const bool slicing = nf && slicer_geometry(n_sl, a.fs_dec, s_st.baud, spb, R);
    if (slicing && smem_mode) {
        slicer_build_masks(s_v, n_sl, R, s_maskA, s_maskN, int(m_words), tid >> 5, kTailThreads / 32, lane);
        __syncthreads();
    }
    if (tid < 32) {
        unsigned slicer_n = s_st.slicer_n;
        unsigned mask_valid = s_st.mask_valid;
        int mask_R = s_st.mask_R;
        UartState us{s_st.uart_win, int(s_st.uart_n), a.uart_runs + (size_t)ch * kUartRunsCap, s_st.uart_runs_n, s_st.uart_ovf};
        bool rescan = s_st.uart_rescan != 0;
        if (nf) {
            const int n = n_sl;
            const float* v = smem_mode ? s_v : v_g;
            CharSink sink{s_chars, 0, a.log, a.log_ctl, a.log_mask, a.call_seq, unsigned(ch), nullptr, 0u};
            if (a.ssdv_ring) { sink.ring = a.ssdv_ring + (size_t)ch * (kRawRingMask + 1u); sink.ring_total = a.ssdv_total[ch]; }
            unsigned char* rec_bits = a.rec_bits ? a.rec_bits + (size_t)ch * a.rec_bits_pitch : nullptr;
            unsigned rec_n = a.rec_bits ? a.rec_bits_n[ch] : 0;
            int erase = 0;
            if (slicing)
                erase = slice_channel(v, n, spb, R, smem_mode ? s_maskA : nullptr, smem_mode ? s_maskN : nullptr, s_st.rtty_bits,
                                      s_st.rtty_stops, us, rescan, sink, rec_bits, rec_n, a.rec_bits_pitch, lane);
            sink_flush(sink, lane);
            if (sink.ring && lane == 0) a.ssdv_total[ch] = sink.ring_total;
            if (a.rec_bits && lane == 0) a.rec_bits_n[ch] = rec_n;
            // erase consumed samples (SymbolExtractor.h:156-157) / write the queue back
            const int keep = n - erase;
            if (n_old != s_st.slicer_n) mask_valid = 0;     // vented: the cached masks describe samples that are gone
            if (slicing) {
                // masks of the positions that keep their value, moved down by the erased samples: [0, keep - R) of the new queue
                const int nv = keep - R;
                if (smem_mode && mask_g && nv > 0 && n <= int(kMaskWords * 32u)) {
                    const int nvw = (nv + 31) >> 5, nw = (n + 31) >> 5, ws = erase >> 5;
                    const unsigned bs = unsigned(erase) & 31u;
                    for (int w = (erase ? 0 : int(m_words)) + lane; w < nvw; w += 32) {
                        const unsigned la = s_maskA[w + ws], ln = s_maskN[w + ws];
                        const unsigned ha = (w + ws + 1 < nw) ? s_maskA[w + ws + 1] : 0u, hn = (w + ws + 1 < nw) ? s_maskN[w + ws + 1] : 0u;
                        mask_g[w] = __funnelshift_r(la, ha, bs);
                        mask_g[kMaskWords + w] = __funnelshift_r(ln, hn, bs);
                    }
                    mask_valid = unsigned(nv); mask_R = R;
                } else mask_valid = 0;
            }
            if (smem_mode) {
                const int from = erase ? 0 : int(n_old);    // nothing erased: only the new samples are missing in HBM
                for (int k = from + lane; k < keep; k += 32) v_g[k] = s_v[erase + k];
            } else if (erase) {
                for (int base = 0; base < keep; base += 32) {
                    const int k = base + lane;
                    float x = 0.f;
                    if (k < keep) x = v_g[k + erase];
                    __syncwarp();
                    if (k < keep) v_g[k] = x;
                    __syncwarp();
                }
            }
            slicer_n = unsigned(keep);
        }
        if (lane == 0) {
            gst.dec_pending = new_pending;
            gst.afc_tick = (tick && frame_done) ? 1u : 0u;   // with a new frame the step follows the transform, in K4
            gst.n_filtered = nf;
            if (fft_take) {
                gst.fft_have = fft_have0 + fft_take;
                if (fft_have0 + fft_take >= fft_n) gst.fft_ready = 1;
            }
            if (nf) {
                gst.demod_last_re = carry_prev.x;
                gst.demod_last_im = carry_prev.y;
                gst.demod_primed = 1;
                gst.demod_n = nf;
                gst.slicer_n = slicer_n;
                gst.mask_valid = mask_valid;
                gst.mask_R = mask_R;
                gst.uart_win = us.win;
                gst.uart_n = unsigned(us.have);
                gst.uart_runs_n = us.n_runs;
                gst.uart_ovf = us.ovf;
                gst.uart_rescan = rescan ? 1u : 0u;
            }
        }
    }
}

// shared-memory layout of one launch (all sizes in elements of the respective arrays)
static void tail_layout(TailArgs& a, size_t* bytes)
{
    const int w2 = (a.M2 == 4) ? pad16(kTile * 4 + ((a.T2 - 1 + 3) & ~3) + 8) : (a.M2 > 1 ? kTile * a.M2 + a.T2 + 8 : 4);
    a.xw = (w2 + 3) & ~3;
    const int Tm = a.max_lp_taps > 0 ? a.max_lp_taps : 1;
    a.qcap = (1 + (Tm - 1) + 2 * kLpBatch + 8 + 3) & ~3;
    a.h2cap = (a.T2 + 8 + 3) & ~3;
    a.hlcap = (Tm + 8 + 3) & ~3;
    a.sv_cap = (a.sv_want + 3) & ~3;
    a.mask_words = ((a.sv_cap + 31) >> 5) + 1;       // per mask: one bit per staged slicer sample
    *bytes = size_t(a.xw + a.qcap + kTile + 2) * 8 + size_t(a.h2cap + a.hlcap + a.sv_cap + 2 * a.mask_words) * 4 + kCharBuf;
}

cudaError_t launch_tail(TailArgs a, int n_channels, cudaStream_t stream, int* launches)
{
    size_t smem = 0;
    tail_layout(a, &smem);
    // measurement hook: HBD_TAIL_PAD_KB=<n> pads the request so that fewer tail CTAs share an SM with K1
    static const size_t pad = [] { const char* e = getenv("HBD_TAIL_PAD_KB"); return e ? size_t(atoi(e)) * 1024 : size_t(0); }();
    smem += pad;
    static size_t configured_smem = 0;
    if (configured_smem < smem) {
        const size_t want = std::max<size_t>(smem, 64 * 1024);
        cudaError_t e = cudaFuncSetAttribute(tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)want);
        if (e != cudaSuccess) return e;
        cudaFuncSetAttribute(tail_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        configured_smem = want;
    }
    tail_kernel<<<n_channels, kTailThreads + 32, smem, stream>>>(a);   // 4 signal warps + the AFC warp
    if (launches) ++*launches;
    return cudaGetLastError();
}

} // namespace hbd
