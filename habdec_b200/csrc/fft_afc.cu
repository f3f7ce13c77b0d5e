// K4 -- spectrum FFT + AFC, a small grid of CTAs each owning a strided set of channels.
//
//   FFT::operator() + swap_half     code/Decoder/FFT.cpp:77-99   (FFTW3f forward c2c, unnormalised, no window)
//   FftPower / ComputeVariance / FindPeaks   code/Decoder/AFC.h:225-329
//   AFC<float>::process state machine        code/Decoder/AFC.h:92-184, Average.h:34-70
//
// FFT: 4096 = 16 x 16 x 16.  Three register-resident radix-16 passes (each a 4x4 radix-4
// butterfly), 256 threads x 16 points, two exchanges through shared memory (second one
// padded so both the write and the transposed read are conflict free).  Twiddles come from a
// table evaluated in float64 on the host, so the result stays within ~1e-7 relative of a
// float64 DFT (the reference's FFTW is not in the image: parity at this boundary is
// judged against float64).  No cuFFT.
//
// AFC: the reference re-runs FftPower/FindPeaks on the same stale spectrum on every call;
// those results are cached per spectrum and only the per-call state machine (four moving
// averages, detection, stability, correction) is stepped, in float64 with explicitly
// unfused multiplies/adds so that it matches the CPU's arithmetic.
#include "hbd_common.cuh"
#include "fft_afc.cuh"
#include <algorithm>

namespace hbd {

constexpr int kFftThreads = 256;

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

// forward 4-point DFT, natural order out
__device__ __forceinline__ void dft4(float2& a0, float2& a1, float2& a2, float2& a3)
{
    const float2 t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3), t3 = csub(a1, a3);
    a0 = cadd(t0, t2);
    a2 = csub(t0, t2);
    a1 = make_float2(t1.x + t3.y, t1.y - t3.x);
    a3 = make_float2(t1.x - t3.y, t1.y + t3.x);
}

// forward 16-point DFT in registers; output X[c + 4d] is left in v[4c + d]
__device__ __forceinline__ void dft16(float2 (&v)[16])
{
#pragma unroll
    for (int b = 0; b < 4; ++b) dft4(v[b], v[4 + b], v[8 + b], v[12 + b]);
    // W16^e = exp(-2 pi i e / 16)
    const float c1 = 0.92387953251128673848f, s1 = 0.38268343236508978178f, r2 = 0.70710678118654752440f;
    const float2 w1 = make_float2(c1, -s1), w2 = make_float2(r2, -r2), w3 = make_float2(s1, -c1);
    const float2 w6 = make_float2(-r2, -r2), w9 = make_float2(-c1, s1);
    v[5] = cmul(v[5], w1);   // c=1,b=1
    v[6] = cmul(v[6], w2);   // c=1,b=2
    v[7] = cmul(v[7], w3);   // c=1,b=3
    v[9] = cmul(v[9], w2);   // c=2,b=1
    v[10] = make_float2(v[10].y, -v[10].x); // c=2,b=2: W16^4 = -i
    v[11] = cmul(v[11], w6); // c=2,b=3
    v[13] = cmul(v[13], w3); // c=3,b=1
    v[14] = cmul(v[14], w6); // c=3,b=2
    v[15] = cmul(v[15], w9); // c=3,b=3
#pragma unroll
    for (int c = 0; c < 4; ++c) dft4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
}
// index of DFT bin k (0..15) inside v[] after dft16
__device__ __forceinline__ constexpr int bin16(int k) { return 4 * (k & 3) + (k >> 2); }

template <typename T>
__device__ __forceinline__ T warp_sum(T v)
{
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ double block_sum(double v, double* scratch)
{
    v = warp_sum(v);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) scratch[w] = v;
    __syncthreads();
    double t = (threadIdx.x < kFftThreads / 32) ? scratch[threadIdx.x] : 0.0;
    if (w == 0) { t = warp_sum(t); if (l == 0) scratch[0] = t; }
    __syncthreads();
    return scratch[0];
}

// arg-max with first-index tie break over (value, index) pairs held one per thread
__device__ void block_argmax(float v, int i, float* sv, int* si, float& ov, int& oi)
{
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        const float ov2 = __shfl_xor_sync(0xffffffffu, v, o);
        const int oi2 = __shfl_xor_sync(0xffffffffu, i, o);
        if (ov2 > v || (ov2 == v && oi2 < i)) { v = ov2; i = oi2; }
    }
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) { sv[w] = v; si[w] = i; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float bv = sv[0]; int bi = si[0];
        for (int k = 1; k < kFftThreads / 32; ++k)
            if (sv[k] > bv || (sv[k] == bv && si[k] < bi)) { bv = sv[k]; bi = si[k]; }
        sv[0] = bv; si[0] = bi;
    }
    __syncthreads();
    ov = sv[0]; oi = si[0];
}

// ---- Average<T> (Average.h:39-55) -------------------------------------------------------------------------------
__device__ __forceinline__ double avg_get_d(double sum, unsigned cnt) { return cnt ? __ddiv_rn(sum, double(cnt)) : sum; }
__device__ __forceinline__ double avg_add_d(double& sum, unsigned& cnt, unsigned cap, double val)
{
    const double g = avg_get_d(sum, cnt);
    const double diff = __dsub_rn(g, val);
    if (cnt == cap) sum = __dadd_rn(__dmul_rn(g, double(cap - 1)), val);
    else { ++cnt; sum = __dadd_rn(sum, val); }
    return diff;
}
__device__ __forceinline__ double avg_get_i(int sum, unsigned cnt) { return cnt ? __ddiv_rn(double(sum), double(cnt)) : double(sum); }
__device__ __forceinline__ double avg_add_i(int& sum, unsigned& cnt, unsigned cap, int val)
{
    const double g = avg_get_i(sum, cnt);
    const double diff = __dsub_rn(g, double(val));
    if (cnt == cap) sum = int(__dadd_rn(__dmul_rn(g, double(cap - 1)), double(val))); // truncating assignment
    else { ++cnt; sum += val; }
    return diff;
}

__device__ void afc_step(ChanState& st, double fs_dec, int n_fft)
{
    if (!st.have_spectrum || !st.spec_ok) { st.afc_correction = 0; return; } // AFC.h:96-100
    st.afc_noise_floor = st.spec_nf;
    st.afc_noise_var = st.spec_nv;
    avg_add_d(st.nf_sum, st.nf_cnt, 100, st.spec_nf);
    avg_add_d(st.nv_sum, st.nv_cnt, 100, st.spec_nv);
    int p1 = st.spec_p1, p2 = st.spec_p2;
    const float thr = float(__dadd_rn(avg_get_d(st.nf_sum, st.nf_cnt), __dmul_rn(3.0, fabs(avg_get_d(st.nv_sum, st.nv_cnt)))));
    const bool d1 = st.spec_p1_val > thr, d2 = st.spec_p2_val > thr;
    bool stable_l = false, stable_r = false;
    if (d1 && d2) {
        if (p2 < p1) { const int t = p1; p1 = p2; p2 = t; }
        if (avg_add_i(st.pl_sum, st.pl_cnt, 4, p1) <= 2.0) stable_l = true;
        if (avg_add_i(st.pr_sum, st.pr_cnt, 4, p2) <= 2.0) stable_r = true;
    }
    const double la = avg_get_i(st.pl_sum, st.pl_cnt), ra = avg_get_i(st.pr_sum, st.pr_cnt);
    st.gui_left = 0;
    if (d1) st.gui_left = stable_l ? int(la) : int(-la);
    st.gui_right = 0;
    if (d2) st.gui_right = stable_r ? int(ra) : int(-ra);
    if (stable_l && stable_r) {
        const int pl = int(round(la)), pr = int(round(ra));
        const int dist = pr - pl;
        const double hz_per_bin = __ddiv_rn(fs_dec, double(n_fft));
        st.afc_shift_hz = __dmul_rn(hz_per_bin, double(dist));
        const double mid = double(pl + dist / 2);
        const double err = __dsub_rn(mid, double(n_fft) / 2);
        if (4 < fabs(err)) st.afc_correction = __dmul_rn(hz_per_bin, err);
    }
}

// 4096-point forward DFT of x[stride * n + offset] (n = 0 .. 4095) by 256 threads: three register-resident radix-16
// passes, two exchanges through s_a (16*16*17 float2).  tw[e * tw_step] = exp(-2 pi i e / 4096).  Bin k is handed
// to store(k, value).
template <typename Store>
__device__ __forceinline__ void fft4096(const float2* __restrict__ x, int stride, int offset, const float2* __restrict__ tw, int tw_step,
                                        float2* s_a, Store store)
{
    const int t = threadIdx.x;
    float2 v[16];
    // pass 1: n = 256*n1 + t, DFT over n1 -> k1; twiddle W_4096^(t*k1); A[k1][t]
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = x[(size_t)stride * (256 * i + t) + offset];
    dft16(v);
#pragma unroll
    for (int k1 = 0; k1 < 16; ++k1) {
        float2 y = v[bin16(k1)];
        if (k1) y = cmul(y, tw[(t * k1) * tw_step]);
        s_a[k1 * 256 + t] = y;
    }
    __syncthreads();
    // pass 2: thread (k1, m2): n2 = 16*m1 + m2, DFT over m1 -> j1; twiddle W_256^(m2*j1); B[k1][j1][m2] (row pitch 17)
    {
        const int k1 = t >> 4, m2 = t & 15;
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = s_a[k1 * 256 + 16 * i + m2];
        dft16(v);
        __syncthreads();
#pragma unroll
        for (int j1 = 0; j1 < 16; ++j1) {
            float2 y = v[bin16(j1)];
            if (j1) y = cmul(y, tw[(16 * m2 * j1) * tw_step]);
            s_a[k1 * (16 * 17) + j1 * 17 + m2] = y;
        }
    }
    __syncthreads();
    // pass 3: thread (k1, j1): DFT over m2 -> j2; X[k1 + 16*j1 + 256*j2]
    {
        const int k1 = t >> 4, j1 = t & 15;
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = s_a[k1 * (16 * 17) + j1 * 17 + i];
        dft16(v);
        __syncthreads();
#pragma unroll
        for (int j2 = 0; j2 < 16; ++j2) store(k1 + 16 * j1 + 256 * j2, v[bin16(j2)]);
    }
}

// N = 4096 (the reference's fft_bins_cnt_, Decoder.h:163) or 16384 (the 16k-bin spectrum of BASELINE configs[1]:
// four 4096-point sub-transforms of the decimated-by-4 phases + one radix-4 combining pass, all in shared memory).
template <int N>
__global__ void __launch_bounds__(kFftThreads)
fft_afc_kernel(FftArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* s_a = reinterpret_cast<float2*>(smem_raw);          // exchange area of the 4096-point transform
    float2* s_f = s_a + 16 * 16 * 17;                           // N == 16384: the four sub-spectra [4][4096]
    float* s_p = reinterpret_cast<float*>(s_a + 16 * 16 * 17);  // N == 4096: power spectrum (4096 floats)
    __shared__ double s_red[kFftThreads / 32];
    __shared__ float s_av[kFftThreads / 32];
    __shared__ int s_ai[kFftThreads / 32];
    __shared__ int s_bad;

    __shared__ unsigned char s_todo[kFftThreads];
    const int t = threadIdx.x;
    // CTA b owns channels b, b + G, b + 2G, ... (G = gridDim.x, at most kFftThreads of them).  Phase A: thread j looks at
    // the j-th owned channel; a channel without a complete frame only needs the per-call AFC step (the common case:
    // 15 of 16 calls at 256 decimated samples per call), done here by one thread per channel, all channels in parallel.
    // Phase B: channels with a complete frame are transformed one after the other by the whole CTA.  A small grid with
    // in-CTA ownership keeps the number of 50 KB-shared-memory CTAs that must find room next to the resident K1 small.
    {
        const int j_ch = int(blockIdx.x) + t * int(gridDim.x);
        unsigned char todo = 0;
        if (j_ch < a.n_channels) {
            ChanState& sj = a.state[a.ch0 + j_ch];
            if (sj.fft_ready) todo = 1;
            else if (sj.afc_tick) { afc_step(sj, a.fs_dec, N); sj.afc_tick = 0; }
        }
        s_todo[t] = todo;
    }
    __syncthreads();
    for (int j = 0; int(blockIdx.x) + j * int(gridDim.x) < a.n_channels; ++j) {
    if (!s_todo[j]) continue;
    const int ch = a.ch0 + int(blockIdx.x) + j * int(gridDim.x);
    ChanState& st = a.state[ch];
    const bool do_tick = st.afc_tick != 0;
    __syncthreads();   // everybody has read the flags thread 0 clears at the end of the previous channel

    {
        const float2* x = a.fftbuf + (size_t)ch * N;
        const float2* __restrict__ tw = a.twiddle; // tw[e] = exp(-2 pi i e / N)
        float2* spec = a.spectrum + (size_t)ch * N;
        float* pw = a.power + (size_t)ch * N;
        if (t == 0) s_bad = 0;
        int bad = 0;
        // FftPower of one bin (AFC.h:236-286), spectrum out with the halves swapped (FFT.cpp:77-87)
        auto emit = [&](int k, float2 z) {
            const int i = (k + N / 2) & (N - 1);
            spec[i] = z;
            if (z.x != z.x || z.y != z.y || isinf(z.x) || isinf(z.y)) bad = 1;
            float p = __fadd_rn(__fmul_rn(z.x, z.x), __fmul_rn(z.y, z.y)) / float(N);
            p = __fmul_rn(p, p);
            p = float(__ddiv_rn(double(p), a.fs_dec));
            p = __fmul_rn(10.0f, log10f(p));
            if (N == 4096) s_p[i] = p; else pw[i] = p;
        };
        if (N == 4096) {
            fft4096(x, 1, 0, tw, 1, s_a, emit);
        } else {
            for (int r = 0; r < 4; ++r) {
                fft4096(x, 4, r, tw, 4, s_a, [&](int k, float2 z) { s_f[r * 4096 + k] = z; });
                __syncthreads();
            }
            // X[k + 4096 q] = sum_r W_N^(r k) F_r[k] W_4^(r q)
            for (int k = t; k < 4096; k += kFftThreads) {
                float2 y0 = s_f[k], y1 = cmul(s_f[4096 + k], tw[k]), y2 = cmul(s_f[8192 + k], tw[2 * k]), y3 = cmul(s_f[12288 + k], tw[3 * k]);
                dft4(y0, y1, y2, y3);
                emit(k, y0); emit(k + 4096, y1); emit(k + 8192, y2); emit(k + 12288, y3);
            }
        }
        if (bad) s_bad = 1;
        __syncthreads();
        const float* P = (N == 4096) ? s_p : pw;   // N == 16384: the dB values are re-read from HBM/L2 (written by this CTA)
        int bad2 = 0;
        if (!s_bad) {
            for (int i = t; i < N; i += kFftThreads) {
                const float p = P[i];
                if (N == 4096) pw[i] = p;
                if (p != p || isinf(p)) bad2 = 1;
            }
        }
        if (bad2) s_bad = 1;
        __syncthreads();
        const bool ok = !s_bad;
        if (ok) {
            // noise floor = mean, "variance" = standard deviation, both float64 (AFC.h:103-104,225-232)
            double s = 0;
            for (int i = t; i < N; i += kFftThreads) s += double(P[i]);
            const double nf = block_sum(s, s_red) / double(N);
            double q = 0;
            for (int i = t; i < N; i += kFftThreads) { const double d = double(P[i]) - nf; q += d * d; }
            const double nv = sqrt(block_sum(q, s_red) / double(N));
            // FindPeaks (AFC.h:290-329)
            float bv = -INFINITY; int bi = 0x7fffffff;
            for (int i = t; i < N; i += kFftThreads) { const float p = P[i]; if (p > bv) { bv = p; bi = i; } }
            float p1v; int p1;
            block_argmax(bv, bi, s_av, s_ai, p1v, p1);
            const float rel_sep = float(500.0f / a.fs_dec);
            int sep = int(round(double(rel_sep) * double(N)));
            sep = max(8, sep);
            const int lo = max(p1 - 2 * sep, 0), hi = min(p1 + 2 * sep, N);
            bv = -INFINITY; bi = 0x7fffffff;
            for (int i = lo + t; i < hi; i += kFftThreads) {
                const float p = P[i];
                if (abs(i - p1) > sep / 2 && p > bv) { bv = p; bi = i; }
            }
            float p2v; int p2;
            block_argmax(bv, bi, s_av, s_ai, p2v, p2);
            if (t == 0) {
                if (!(p2v > P[0])) { p2 = 0; p2v = P[0]; } // running best starts at v[0], index 0
                if (p2 < p1) { const int ti = p1; p1 = p2; p2 = ti; const float tv = p1v; p1v = p2v; p2v = tv; }
                st.spec_nf = nf; st.spec_nv = nv;
                st.spec_p1 = p1; st.spec_p2 = p2; st.spec_p1_val = p1v; st.spec_p2_val = p2v;
            }
        }
        if (t == 0) {
            st.spec_ok = ok ? 1 : 0;
            st.have_spectrum = 1;
            st.fft_ready = 0;
            st.fft_have = 0;
        }
        __syncthreads();
    }
    if (do_tick && t == 0) {
        afc_step(st, a.fs_dec, N);
        st.afc_tick = 0;
    }
    }
}

template <int N>
static cudaError_t launch_fft_n(const FftArgs& a, int n_channels, cudaStream_t stream, int* launches)
{
    const size_t smem = size_t(16 * 16 * 17) * 8 + (N == 4096 ? size_t(4096) * 4 : size_t(16384) * 8);
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(fft_afc_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        cudaFuncSetAttribute(fft_afc_kernel<N>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        configured = true;
    }
    static int n_sms = 0;
    if (!n_sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, dev); if (n_sms < 1) n_sms = 1; }
    int grid = std::min(n_channels, 4 * n_sms);
    grid = std::max(grid, (n_channels + kFftThreads - 1) / kFftThreads);   // a CTA owns at most kFftThreads channels
    if (grid < 1) return cudaSuccess;
    FftArgs b = a; b.n_channels = n_channels;
    fft_afc_kernel<N><<<grid, kFftThreads, smem, stream>>>(b);
    if (launches) ++*launches;
    return cudaGetLastError();
}

cudaError_t launch_fft_afc(const FftArgs& a, int n_channels, cudaStream_t stream, int* launches)
{
    if (a.fft_n == 16384) return launch_fft_n<16384>(a, n_channels, stream, launches);
    return launch_fft_n<4096>(a, n_channels, stream, launches);
}

// AFC::resetFrequencyCorrection (AFC.h:188-194), one thread per call
__global__ void afc_reset_kernel(ChanState* state, int ch, double corr, double fs_dec, int n_fft)
{
    ChanState& st = state[ch];
    // the reference divides fft_samples_.size() by its sampling rate: both are 0 before the first spectrum
    const double bins_per_hz = st.have_spectrum ? __ddiv_rn(double(n_fft), fs_dec) : __ddiv_rn(0.0, 0.0);
    const double l = __dsub_rn(avg_get_i(st.pl_sum, st.pl_cnt), __dmul_rn(corr, bins_per_hz));
    const double r = __dsub_rn(avg_get_i(st.pr_sum, st.pr_cnt), __dmul_rn(corr, bins_per_hz));
    st.pl_sum = int(fmax(0.0, l)); st.pl_cnt = 1;
    st.pr_sum = int(fmax(0.0, r)); st.pr_cnt = 1;
    st.afc_correction = 0;
}

cudaError_t launch_afc_reset(ChanState* state, int ch, double corr, double fs_dec, int n_fft, cudaStream_t stream)
{
    afc_reset_kernel<<<1, 1, 0, stream>>>(state, ch, corr, fs_dec, n_fft);
    return cudaGetLastError();
}

// The retune decision of DECODER_THREAD (code/websocketServer/main.cpp:248-265) for every channel at once: where
// |frequency_correction| exceeds min_abs_hz the correction is handed to the caller (who adds it to the channel's
// NCO) and the AFC is reset exactly like resetFrequencyCorrection(); applied[ch] = 0 elsewhere.
__global__ void afc_retune_kernel(ChanState* state, int n_ch, double min_abs_hz, double fs_dec, double* applied, int n_fft)
{
    const int ch = blockIdx.x * blockDim.x + threadIdx.x;
    if (ch >= n_ch) return;
    ChanState& st = state[ch];
    const double corr = st.afc_correction;
    if (!(min_abs_hz < fabs(corr))) { applied[ch] = 0.0; return; }
    applied[ch] = corr;
    const double bins_per_hz = st.have_spectrum ? __ddiv_rn(double(n_fft), fs_dec) : __ddiv_rn(0.0, 0.0);
    const double l = __dsub_rn(avg_get_i(st.pl_sum, st.pl_cnt), __dmul_rn(corr, bins_per_hz));
    const double r = __dsub_rn(avg_get_i(st.pr_sum, st.pr_cnt), __dmul_rn(corr, bins_per_hz));
    st.pl_sum = int(fmax(0.0, l)); st.pl_cnt = 1;
    st.pr_sum = int(fmax(0.0, r)); st.pr_cnt = 1;
    st.afc_correction = 0;
}

cudaError_t launch_afc_retune(ChanState* state, int n_ch, double min_abs_hz, double fs_dec, double* applied, int n_fft, cudaStream_t stream)
{
    afc_retune_kernel<<<(n_ch + 127) / 128, 128, 0, stream>>>(state, n_ch, min_abs_hz, fs_dec, applied, n_fft);
    return cudaGetLastError();
}

} // namespace hbd
