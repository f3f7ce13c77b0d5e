// K2 -- everything between the first decimator and the bit slicer, one CTA per channel.
//
//   stage-2 FIR decimator     code/Decoder/Decimator.h:99-146 (second entry of Decoder.h:286-320)
//   DC removal (optional)     code/Decoder/Decoder.h:450-459
//   FFT frame assembly        code/Decoder/Decoder.h:467-473   (spectrum taps the stream BEFORE the low-pass)
//   batch-of-256 gate + AFC tick  Decoder.h:492-509, >160 kS/s cut-off Decoder.h:522-527
//   low-pass FIR              code/Decoder/FirFilter.h:117-169 (taps designed on the host, lp_design.cpp)
//   FM/FSK discriminator      code/Decoder/FSK2_Demod.h:30-42  (carry kept PER CHANNEL, not per thread)
//   slicer input append       code/Decoder/SymbolExtractor.h:108-125 (3e4 safety vent included)
//   stage-1 carry update      code/Decoder/Decimator.h:141-143 + Decoder.h:432-435 (unconsumed remainder)
//
// The data here is 1/64 .. 1/256 of the input rate, so this kernel is latency/FP32 bound and tiny
// next to K1; it is organised for exactness of the stream bookkeeping rather than for bandwidth.
#include "hbd_common.cuh"
#include "tail.cuh"

namespace hbd {

constexpr int kTailThreads = 256;
constexpr int kTile = 256; // outputs per tile

__device__ __forceinline__ float2 cmul_conj_ieee(float2 a, float2 b) // a * conj(b), separately rounded products (no FMA)
{
    float2 r;
    r.x = __fadd_rn(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y));
    r.y = __fadd_rn(__fmul_rn(a.y, b.x), -__fmul_rn(a.x, b.y));
    return r;
}

__global__ void __launch_bounds__(kTailThreads)
tail_kernel(TailArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* s_x = reinterpret_cast<float2*>(smem_raw);                // tile input window
    float*  s_h = reinterpret_cast<float*>(s_x + a.smem_window);      // taps
    float2* s_f = reinterpret_cast<float2*>(s_h + kLpMaxTaps + 7);    // filtered tile (+1 previous sample)
    __shared__ unsigned sh_total, sh_nf, sh_slicer_base;

    const int ch = blockIdx.x, tid = threadIdx.x;
    const ChanPlan pl = a.plan[ch];
    ChanState& st = a.state[ch];

    // ---- stage-1 carry for the NEXT call: last (T1-1 + r') samples of [carry | chunk] -------------------
    {
        const unsigned r_next = (pl.r + pl.n) - pl.consumed;
        const int keep = a.T1 - 1 + int(r_next);                 // <= kCarryCap (host checked)
        float2* carry = a.carry + (size_t)ch * kCarryCap + kCarryCap;
        const float2* chunk = a.chunk + (size_t)ch * a.chunk_pitch;
        // two passes through registers: source and destination may overlap inside the carry
        float2 tmp[(kCarryCap + kTailThreads - 1) / kTailThreads];
        int cnt = 0;
        for (int i = tid; i < keep; i += kTailThreads) {
            const long long j = (long long)pl.n - keep + i;
            tmp[cnt++] = (j < 0) ? ((j >= -kCarryCap) ? carry[j] : make_float2(0.f, 0.f)) : chunk[j];
        }
        __syncthreads();
        cnt = 0;
        for (int i = tid; i < keep; i += kTailThreads) carry[-keep + i] = tmp[cnt++];
    }
    if (pl.flags & 1u) return; // fewer than `factor` samples queued: Decoder.h:429-430

    const unsigned n1 = pl.n1, n2 = pl.n2;
    float2* s1 = a.s1 + (size_t)ch * a.s1_pitch;
    float2* dq = a.decq + (size_t)ch * a.dq_pitch;
    const unsigned dec_pending = st.dec_pending;
    float2* ynew = dq + kLpHist + dec_pending; // where this call's decimated samples go

    // ---- stage 2 ------------------------------------------------------------------------------------------
    if (a.M2 > 1) {
        const int T2 = a.T2, M2 = a.M2;
        for (int i = tid; i < T2; i += kTailThreads) s_h[i] = a.taps2[i];
        for (unsigned k0 = 0; k0 < n2; k0 += kTile) {
            const unsigned nk = hbd_min_u(kTile, n2 - k0);
            const int win = int(nk) * M2 + T2 - 1 - (M2 - 1); // samples needed: first window start .. last output end
            __syncthreads();
            // y[k] = sum_t x[k*M2 - (T2-1) + t] h[t]; x index 0 is s1[kS1Hist]
            const long long x0 = (long long)k0 * M2 - (T2 - 1);
            for (int i = tid; i < win; i += kTailThreads) s_x[i] = s1[kS1Hist + x0 + i];
            __syncthreads();
            if (tid < int(nk)) {
                const float2* w = s_x + tid * M2;
                float re = 0.f, im = 0.f;
                for (int t = 0; t < T2; ++t) {
                    re = fmaf(w[t].x, s_h[t], re);
                    im = fmaf(w[t].y, s_h[t], im);
                }
                ynew[k0 + tid] = make_float2(re, im);
            }
        }
        __syncthreads();
        // history for the next call: last T2-1 stage-1 samples (source may overlap when n1 < T2-1)
        float2 keep2[2];
        int c = 0;
        for (int i = tid; i < T2 - 1; i += kTailThreads) keep2[c++] = s1[kS1Hist + (long long)n1 - (T2 - 1) + i];
        __syncthreads();
        c = 0;
        for (int i = tid; i < T2 - 1; i += kTailThreads) s1[kS1Hist - (T2 - 1) + i] = keep2[c++];
    } else {
        for (unsigned k = tid; k < n2; k += kTailThreads) ynew[k] = s1[kS1Hist + k];
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
    for (int i = tid; i < T; i += kTailThreads) s_h[i] = taps[i];
    float* pend = a.slicer + (size_t)ch * a.slicer_pitch + sh_slicer_base;
    float* dlast = a.demod_last ? a.demod_last + (size_t)ch * a.demod_pitch : nullptr;
    float2 carry_prev = make_float2(st.demod_last_re, st.demod_last_im);
    const bool primed = st.demod_primed != 0;
    const float2* q = dq + kLpHist - (T - 1); // q[i + t], i = output index

    for (unsigned i0 = 0; i0 < nf; i0 += kTile) {
        __syncthreads();
        for (int i = tid; i < kTile + T - 1; i += kTailThreads) s_x[i] = q[i0 + i];
        __syncthreads();
        float re = 0.f, im = 0.f;
        {
            const float2* w = s_x + tid;
            for (int t = 0; t < T; ++t) {
                re = fmaf(w[t].x, s_h[t], re);
                im = fmaf(w[t].y, s_h[t], im);
            }
        }
        const float2 f = make_float2(re, im);
        s_f[tid + 1] = f;
        if (tid == 0) s_f[0] = (i0 == 0) ? (primed ? carry_prev : f) : carry_prev;
        __syncthreads();
        const float2 prev = s_f[tid];
        const float2 pr = cmul_conj_ieee(f, prev);
        const float d = atan2f(pr.y, pr.x);
        pend[i0 + tid] = d;
        if (dlast) dlast[i0 + tid] = d;
        if (a.rec_filtered) a.rec_filtered[(size_t)ch * a.rec_pitch + i0 + tid] = f;
        carry_prev = s_f[kTile]; // last filtered sample of this tile (same value in every thread)
        // (s_f is rewritten only after the next two barriers)
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
        float2 tmp[(kLpHist + kLpBatch + kTailThreads - 1) / kTailThreads];
        int c = 0;
        for (int m = tid; m < n_move; m += kTailThreads) tmp[c++] = dq[kLpHist - hist + m + nf];
        __syncthreads();
        c = 0;
        for (int m = tid; m < n_move; m += kTailThreads) dq[kLpHist - hist + m] = tmp[c++];
    }
}

cudaError_t launch_tail(const TailArgs& a, int n_channels, cudaStream_t stream, int* launches)
{
    const size_t smem = size_t(a.smem_window) * 8 + size_t(kLpMaxTaps + 7) * 4 + size_t(kTile + 1) * 8;
    cudaError_t e = cudaFuncSetAttribute(tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    tail_kernel<<<n_channels, kTailThreads, smem, stream>>>(a);
    if (launches) ++*launches;
    return cudaGetLastError();
}

int tail_smem_window(int M2, int T2)
{
    const int w2 = (M2 > 1) ? kTile * M2 + T2 : 0;
    const int wl = kTile + kLpMaxTaps;
    return (w2 > wl ? w2 : wl) + 8;
}

} // namespace hbd
