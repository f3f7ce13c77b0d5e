// K2 -- everything between the first decimator and the bit slicer, one CTA per channel.
//
//   stage-2 FIR decimator     code/Decoder/Decimator.h:99-146 (second entry of Decoder.h:286-320)
//   DC removal (optional)     code/Decoder/Decoder.h:450-459
//   FFT frame assembly        code/Decoder/Decoder.h:467-473   (spectrum taps the stream BEFORE the low-pass)
//   batch-of-256 gate + AFC tick  Decoder.h:492-509, >160 kS/s cut-off Decoder.h:522-527
//   low-pass FIR              code/Decoder/FirFilter.h:117-169 (taps designed on the host, host_tail.cpp)
//   FM/FSK discriminator      code/Decoder/FSK2_Demod.h:30-42  (carry kept PER CHANNEL, not per thread)
//   slicer input append       code/Decoder/SymbolExtractor.h:108-125 (3e4 safety vent included)
// and, as a separate tiny kernel, the stage-1 carry update (Decimator.h:141-143 + Decoder.h:432-435).
//
// The data here is 1/64 .. 1/256 of the input rate, so this kernel is FP32/LDS bound and runs in the shadow
// of K1 (other channel group, other stream).  Both FIRs are register tiled: a thread owns TWO consecutive
// outputs and slides a small sample window through registers, so shared memory is read once per sample pair
// (LDS.128) instead of once per multiply:
//   stage 2 (M2 = 4):  per tap group of 4:  2 LDS.128 samples + 1 LDS.128 taps -> 16 FFMA
//   low-pass:          per tap pair:        1 LDS.128 samples + 1 LDS.64 taps  ->  8 FFMA
// The stage-2 window tile is padded by 2 samples every 16 so that the 64-byte thread stride is bank-conflict free.
#include "hbd_common.cuh"
#include "tail.cuh"

namespace hbd {

constexpr int kTailThreads = 128;
constexpr int kTile = 2 * kTailThreads; // outputs per tile (two per thread)

__device__ __forceinline__ float2 cmul_conj_ieee(float2 a, float2 b) // a * conj(b), separately rounded products (no FMA)
{
    float2 r;
    r.x = __fadd_rn(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y));
    r.y = __fadd_rn(__fmul_rn(a.y, b.x), -__fmul_rn(a.x, b.y));
    return r;
}

__host__ __device__ __forceinline__ int pad16(int s) { return s + 2 * (s >> 4); } // 2 float2 of padding per 16 samples

__device__ __forceinline__ void fma2(float2& acc, float2 x, float h)
{
    acc.x = fmaf(x.x, h, acc.x);
    acc.y = fmaf(x.y, h, acc.y);
}

// ---- stage-1 carry for the NEXT call: last (T1-1 + r') samples of [carry | chunk] ---------------------------
__global__ void __launch_bounds__(128)
carry_kernel(const ChanPlan* __restrict__ plan, const float2* __restrict__ chunk_base, size_t chunk_pitch, float2* carry_base, int T1, int ch0)
{
    const int ch = ch0 + blockIdx.x, tid = threadIdx.x;
    const ChanPlan pl = plan[ch];
    const unsigned r_next = (pl.r + pl.n) - pl.consumed;
    const int keep = T1 - 1 + int(r_next);                 // <= kCarryCap (host checked)
    float2* carry = carry_base + (size_t)ch * kCarryCap + kCarryCap;
    const float2* chunk = chunk_base + (size_t)ch * chunk_pitch;
    // two passes through registers: source and destination may overlap inside the carry
    float2 tmp[kCarryCap / 128];
#pragma unroll
    for (int u = 0; u < kCarryCap / 128; ++u) {
        const int i = tid + u * 128;
        const long long j = (long long)pl.n - keep + i;
        tmp[u] = (i < keep) ? ((j < 0) ? carry[j] : chunk[j]) : make_float2(0.f, 0.f);
    }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < kCarryCap / 128; ++u) {
        const int i = tid + u * 128;
        if (i < keep) carry[-keep + i] = tmp[u];
    }
}

__global__ void __launch_bounds__(kTailThreads)
tail_kernel(TailArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* s_x = reinterpret_cast<float2*>(smem_raw);                // tile input window (padded layout for stage 2)
    float*  s_h = reinterpret_cast<float*>(s_x + a.smem_window);      // taps (zero padded to a multiple of 4)
    float2* s_f = reinterpret_cast<float2*>(s_h + kLpMaxTaps + 7);    // filtered tile (+1 previous sample)
    __shared__ unsigned sh_total, sh_nf, sh_slicer_base;

    const int ch = a.ch0 + blockIdx.x, tid = threadIdx.x;
    const ChanPlan pl = a.plan[ch];
    ChanState& st = a.state[ch];
    if (pl.flags & 1u) return; // fewer than `factor` samples queued: Decoder.h:429-430

    const unsigned n1 = pl.n1, n2 = pl.n2;
    float2* s1 = a.s1 + (size_t)ch * a.s1_pitch;
    float2* dq = a.decq + (size_t)ch * a.dq_pitch;
    const unsigned dec_pending = st.dec_pending;
    float2* ynew = dq + kLpHist + dec_pending; // where this call's decimated samples go

    // ---- stage 2 ------------------------------------------------------------------------------------------
    if (a.M2 == 4) {
        const int T2 = a.T2;
        const int lead = (T2 - 1 + 3) & ~3;                          // 140 for T2 = 139
        const int skew = lead - (T2 - 1);                            // leading samples that get a zero tap
        const int n_blocks = (skew + T2 + 3) / 4;
        // taps stored shifted by `skew` and zero padded: s_h[u] = h[u - skew]
        for (int u = tid; u < 4 * n_blocks; u += kTailThreads) s_h[u] = (u >= skew && u - skew < T2) ? a.taps2[u - skew] : 0.f;
        for (unsigned k0 = 0; k0 < n2; k0 += kTile) {
            const unsigned nk = hbd_min_u(kTile, n2 - k0);
            // y[k] = sum_t x[4k - (T2-1) + t] h[t]; x index 0 is s1[kS1Hist].  The tile holds x from
            // 4*k0 - (T2-1) rounded DOWN to a multiple of 4 samples (keeps LDS.128 aligned).
            const long long x0 = 4LL * k0 - lead;
            const int win = int(nk) * 4 + lead;                      // samples 0 .. 4*nk + lead - 1
            __syncthreads();
            for (int i = tid; i < win; i += kTailThreads) s_x[pad16(i)] = s1[kS1Hist + x0 + i];
            __syncthreads();
            const int k = 2 * tid; // first of this thread's two outputs
            if (k < int(nk)) {
                // with the window shifted by `skew`, output k uses tile samples 4k + skew + t; taps are applied
                // as h'[u] = h[u - skew] (zero for u < skew) over u = 0 .. lead+? so blocks stay 4-aligned
                float2 acc0 = make_float2(0.f, 0.f), acc1 = make_float2(0.f, 0.f);
                float2 b0[4], b1[4];
                {
                    const float4 p0 = *reinterpret_cast<const float4*>(&s_x[pad16(4 * k)]);
                    const float4 p1 = *reinterpret_cast<const float4*>(&s_x[pad16(4 * k + 2)]);
                    b0[0] = make_float2(p0.x, p0.y); b0[1] = make_float2(p0.z, p0.w); b0[2] = make_float2(p1.x, p1.y); b0[3] = make_float2(p1.z, p1.w);
                }
                for (int blk = 0; blk < n_blocks; ++blk) {
                    const float4 q0 = *reinterpret_cast<const float4*>(&s_x[pad16(4 * (k + blk + 1))]);
                    const float4 q1 = *reinterpret_cast<const float4*>(&s_x[pad16(4 * (k + blk + 1) + 2)]);
                    b1[0] = make_float2(q0.x, q0.y); b1[1] = make_float2(q0.z, q0.w); b1[2] = make_float2(q1.x, q1.y); b1[3] = make_float2(q1.z, q1.w);
                    // taps for tile offsets 4*blk .. 4*blk+3 (one broadcast LDS.128)
                    const float4 h4 = *reinterpret_cast<const float4*>(&s_h[4 * blk]);
                    const float hh[4] = {h4.x, h4.y, h4.z, h4.w};
#pragma unroll
                    for (int c = 0; c < 4; ++c) { fma2(acc0, b0[c], hh[c]); fma2(acc1, b1[c], hh[c]); }
#pragma unroll
                    for (int c = 0; c < 4; ++c) b0[c] = b1[c];
                }
                ynew[k0 + k] = acc0;
                if (k + 1 < int(nk)) ynew[k0 + k + 1] = acc1;
            }
        }
        __syncthreads();
    } else if (a.M2 > 1) { // M2 == 2 (69 taps): half-rate inputs (<= 512 kS/s), simple form
        const int T2 = a.T2, M2 = a.M2;
        for (int i = tid; i < T2; i += kTailThreads) s_h[i] = a.taps2[i];
        for (unsigned k0 = 0; k0 < n2; k0 += kTile) {
            const unsigned nk = hbd_min_u(kTile, n2 - k0);
            const int win = int(nk) * M2 + T2 - M2;
            const long long x0 = (long long)k0 * M2 - (T2 - 1);
            __syncthreads();
            for (int i = tid; i < win; i += kTailThreads) s_x[i] = s1[kS1Hist + x0 + i];
            __syncthreads();
            for (int k = tid; k < int(nk); k += kTailThreads) {
                const float2* w = s_x + k * M2;
                float2 acc = make_float2(0.f, 0.f);
                for (int t = 0; t < T2; ++t) fma2(acc, w[t], s_h[t]);
                ynew[k0 + k] = acc;
            }
        }
        __syncthreads();
    } else {
        for (unsigned k = tid; k < n2; k += kTailThreads) ynew[k] = s1[kS1Hist + k];
    }
    if (a.M2 > 1) {
        // history for the next call: last T2-1 stage-1 samples (source may overlap when n1 < T2-1)
        const int T2 = a.T2;
        float2 keep2[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int i = tid + u * kTailThreads;
            keep2[u] = (i < T2 - 1) ? s1[kS1Hist + (long long)n1 - (T2 - 1) + i] : make_float2(0.f, 0.f);
        }
        __syncthreads();
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int i = tid + u * kTailThreads;
            if (i < T2 - 1) s1[kS1Hist - (T2 - 1) + i] = keep2[u];
        }
    }
    __syncthreads();

    // ---- DC removal, sequential recurrence re-seeded from the first sample of the call ---------------------
    if (st.dc_remove && tid == 0 && n2) {
        float2 wp = make_float2(__fmul_rn(.97f, ynew[0].x), __fmul_rn(.97f, ynew[0].y));
        for (unsigned i = 0; i < n2; ++i) {
            const float2 x = ynew[i];
            const float2 w = make_float2(__fadd_rn(x.x, __fmul_rn(.97f, wp.x)), __fadd_rn(x.y, __fmul_rn(.97f, wp.y)));
            ynew[i] = make_float2(__fadd_rn(w.x, -wp.x), __fadd_rn(w.y, -wp.y));
            wp = w;
        }
    }
    __syncthreads();

    // ---- per-call copy of the decimated block for parity tests / GUI (optional) -----------------------------
    if (a.rec_decimated) {
        float2* rec = a.rec_decimated + (size_t)ch * a.rec_pitch;
        for (unsigned k = tid; k < n2; k += kTailThreads) rec[k] = ynew[k];
    }

    // ---- FFT frame: the first min(4096 - have, n2) samples of this call ------------------------------------
    {
        const unsigned have = st.fft_have;
        if (have < kFftN && n2) {
            const unsigned take = hbd_min_u(kFftN - have, n2);
            float2* fb = a.fftbuf + (size_t)ch * kFftN;
            for (unsigned i = tid; i < take; i += kTailThreads) fb[have + i] = ynew[i];
            if (tid == 0) {
                st.fft_have = have + take;
                if (have + take >= kFftN) st.fft_ready = 1;
            }
        }
    }

    // ---- batch gate ---------------------------------------------------------------------------------------
    if (tid == 0) {
        const unsigned total = dec_pending + n2;
        unsigned nf = 0;
        st.afc_tick = 0;
        if (total >= kLpBatch) {
            st.afc_tick = 1;
            if (a.fs_dec > 4 * 40e3) {
                st.dec_pending = 0; // Decoder.h:522-527: everything queued is dropped, nothing decoded
            } else {
                nf = total - total % kLpBatch;
                st.dec_pending = total - nf;
            }
        } else {
            st.dec_pending = total;
        }
        st.n_filtered = nf;
        sh_total = total;
        sh_nf = nf;
        // slicer vent, SymbolExtractor.h:116-120: checked before the append, only when there is something to append
        unsigned base = st.slicer_n;
        if (nf) {
            if (base > unsigned(kSlicerVent)) base = 0;
            st.slicer_n = base + nf;
        }
        sh_slicer_base = base;
    }
    __syncthreads();
    const unsigned total = sh_total, nf = sh_nf;
    if (!nf) return;

    // ---- low-pass FIR + discriminator ---------------------------------------------------------------------
    const int T = st.lp_ntaps;
    const float* taps = a.lptaps + (size_t)ch * kLpMaxTaps;
    float* pend = a.slicer + (size_t)ch * a.slicer_pitch + sh_slicer_base;
    float* dlast = a.demod_last ? a.demod_last + (size_t)ch * a.demod_pitch : nullptr;
    float2 carry_prev = make_float2(st.demod_last_re, st.demod_last_im);
    const bool primed = st.demod_primed != 0;
    // y[i] = sum_t q[i + t] h[t]; the tile is loaded from an even sample index so LDS.128 stays aligned
    const long long qbase = (long long)kLpHist - (T - 1);
    const int qskew = int(qbase & 1);          // tile sample 0 is q[-qskew]
    const float2* q = dq + qbase - qskew;
    const int n_pairs = (qskew + T + 2) / 2;
    // taps stored shifted by qskew and zero padded: s_h[u] = h[u - qskew]
    __syncthreads();
    for (int u = tid; u < 2 * n_pairs; u += kTailThreads) s_h[u] = (u >= qskew && u - qskew < T) ? taps[u - qskew] : 0.f;

    for (unsigned i0 = 0; i0 < nf; i0 += kTile) { // nf is a multiple of 256 == kTile
        __syncthreads();
        for (int i = tid; i < kTile + T + 2; i += kTailThreads) s_x[i] = q[i0 + i];
        __syncthreads();
        const int o = 2 * tid; // outputs o, o+1; output o uses tile samples o + qskew + t
        float2 acc0 = make_float2(0.f, 0.f), acc1 = make_float2(0.f, 0.f);
        float2 w0, w1;
        {
            const float4 p = *reinterpret_cast<const float4*>(&s_x[o]);
            w0 = make_float2(p.x, p.y); w1 = make_float2(p.z, p.w);
        }
#pragma unroll 4
        for (int pr = 0; pr < n_pairs; ++pr) {
            const float4 p = *reinterpret_cast<const float4*>(&s_x[o + 2 * pr + 2]);
            const float2 w2 = make_float2(p.x, p.y), w3 = make_float2(p.z, p.w);
            const float2 hp = *reinterpret_cast<const float2*>(&s_h[2 * pr]); // taps for tile offsets 2pr, 2pr+1
            const float ha = hp.x, hb = hp.y;
            fma2(acc0, w0, ha); fma2(acc0, w1, hb);
            fma2(acc1, w1, ha); fma2(acc1, w2, hb);
            w0 = w2; w1 = w3;
        }
        s_f[o + 1] = acc0;
        s_f[o + 2] = acc1;
        if (tid == 0) s_f[0] = (i0 == 0) ? (primed ? carry_prev : acc0) : carry_prev;
        __syncthreads();
        const float2 prev = s_f[o];
        const float2 p0 = cmul_conj_ieee(acc0, prev), p1 = cmul_conj_ieee(acc1, acc0);
        const float d0 = atan2f(p0.y, p0.x), d1 = atan2f(p1.y, p1.x);
        pend[i0 + o] = d0; pend[i0 + o + 1] = d1;   // (the queue base can be odd: no vector store)
        if (dlast) { dlast[i0 + o] = d0; dlast[i0 + o + 1] = d1; }
        if (a.rec_filtered) {
            a.rec_filtered[(size_t)ch * a.rec_pitch + i0 + o] = acc0;
            a.rec_filtered[(size_t)ch * a.rec_pitch + i0 + o + 1] = acc1;
        }
        carry_prev = s_f[kTile]; // last filtered sample of this tile (same value in every thread)
    }
    __syncthreads();
    if (tid == 0) {
        st.demod_last_re = carry_prev.x;
        st.demod_last_im = carry_prev.y;
        st.demod_primed = 1;
    }

    // ---- queue shuffle: new low-pass history + unfiltered remainder ----------------------------------------
    {
        const int hist = T - 1;
        const unsigned rem = total - nf;
        const int n_move = hist + int(rem);
        // element m of the new front region [kLpHist-hist, kLpHist+rem) comes from old index m + nf
        constexpr int kPer = (kLpHist + kLpBatch + kTailThreads - 1) / kTailThreads;
        float2 tmp[kPer];
#pragma unroll
        for (int u = 0; u < kPer; ++u) {
            const int m = tid + u * kTailThreads;
            tmp[u] = (m < n_move) ? dq[kLpHist - hist + m + nf] : make_float2(0.f, 0.f);
        }
        __syncthreads();
#pragma unroll
        for (int u = 0; u < kPer; ++u) {
            const int m = tid + u * kTailThreads;
            if (m < n_move) dq[kLpHist - hist + m] = tmp[u];
        }
    }
}

cudaError_t launch_carry(const ChanPlan* plan, const float2* chunk, size_t chunk_pitch, float2* carry, int T1, int ch0, int n_channels,
                         cudaStream_t stream, int* launches)
{
    static bool configured = false;
    if (!configured) { cudaFuncSetAttribute(carry_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared); configured = true; }
    carry_kernel<<<n_channels, 128, 0, stream>>>(plan, chunk, chunk_pitch, carry, T1, ch0);
    if (launches) ++*launches;
    return cudaGetLastError();
}

cudaError_t launch_tail(const TailArgs& a, int n_channels, cudaStream_t stream, int* launches)
{
    const size_t smem = size_t(a.smem_window) * 8 + size_t(kLpMaxTaps + 7) * 4 + size_t(kTile + 1) * 8;
    static size_t configured_smem = 0;
    if (configured_smem < smem) {
        cudaError_t e = cudaFuncSetAttribute(tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        cudaFuncSetAttribute(tail_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        configured_smem = smem;
    }
    tail_kernel<<<n_channels, kTailThreads, smem, stream>>>(a);
    if (launches) ++*launches;
    return cudaGetLastError();
}

int tail_smem_window(int M2, int T2)
{
    const int w2 = (M2 > 1) ? pad16(kTile * M2 + T2 + 8) : 0;
    const int wl = kTile + kLpMaxTaps + 8;
    return ((w2 > wl ? w2 : wl) + 8 + 3) & ~3; // multiple of 4 float2: keeps the tap array 16-byte aligned
}

} // namespace hbd
