// K5 -- the step right after the path: the websocket server's binary payloads, produced per channel on the GPU so
// that the existing web client can be served from the batch decoder with one small device->host copy per request.
//
//   SpectrumToStream / DemodToStream / ShrinkVector   code/websocketServer/habdec_ws_protocol.cpp:338-429
//   SpectrumInfoHeader / DemodHeader / Serialize*      code/websocketServer/NetTransport.h:29-102
//   CompressedVector (min/max, float -> u8 / u16)      code/websocketServer/CompressedVector.h:48-55, CompressedVector.cpp:72-116
//   Decoder::getSpectrumInfo                           code/Decoder/Decoder.h:814-836
//   demod accumulation (last 50 symbols)               code/websocketServer/main.cpp:267-282
//
// Integer / byte work, bit-exact against the reference's serialisation code (tests/golden/wire_frames.npz):
//   zoom slice [zb, ze) of the 4096 dB bins, peaks shifted (and invalidated when they leave the slice),
//   nearest-lower-index shrink to `resolution` bins (float index arithmetic as written in the reference),
//   min / max of what is left, values mapped to (x - min) / (max - min) in the reference's float/double mix and
//   truncated to u8 / u16 (f32 passes through), header in front.  One CTA per channel; min/max by block reduction.
#include "wire.cuh"

namespace hbd {

constexpr int kWireThreads = 256;

__device__ __forceinline__ void block_minmax(float& mn, float& mx, float* s_mn, float* s_mx)
{
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) { s_mn[w] = mn; s_mx[w] = mx; }
    __syncthreads();
    mn = s_mn[0]; mx = s_mx[0];
    for (int i = 1; i < kWireThreads / 32; ++i) { mn = fminf(mn, s_mn[i]); mx = fmaxf(mx, s_mx[i]); }
}

// CompressedVector<T>::copyValues(const std::vector<float>&, min, max), CompressedVector.cpp:72-116
__device__ __forceinline__ unsigned quantise(float x, double mn, double mx, float scale)
{
    const float t = float(double(x) - mn);                 // float(rhs_v - i_min)
    const float r = float(double(t) / (mx - mn));          // ... / (i_max - i_min), stored back into a float
    const float scaled = __fmul_rn(r, scale);              // rhs_v * numeric_limits<T>::max()
    // float -> integer conversion of the x86-64 build (cvttss2si): NaN / out of range give 0x80000000
    const int iv = (scaled != scaled || scaled >= 2147483648.0f || scaled < -2147483648.0f) ? (-2147483647 - 1) : int(scaled);
    return unsigned(iv);
}

__device__ __forceinline__ void store_value(unsigned char* body, size_t i, float x, double mn, double mx, int type_size)
{
    if (type_size == 4) {
        const unsigned u = __float_as_uint(x);
        body[4 * i] = u & 0xff; body[4 * i + 1] = (u >> 8) & 0xff; body[4 * i + 2] = (u >> 16) & 0xff; body[4 * i + 3] = u >> 24;
    } else if (type_size == 2) {
        const unsigned u = quantise(x, mn, mx, 65535.0f) & 0xffffu;
        body[2 * i] = u & 0xff; body[2 * i + 1] = u >> 8;
    } else {
        body[i] = (unsigned char)(quantise(x, mn, mx, 255.0f) & 0xffu);
    }
}

__device__ __forceinline__ void put_i32(unsigned char* p, int v) { const unsigned u = unsigned(v); p[0] = u & 0xff; p[1] = (u >> 8) & 0xff; p[2] = (u >> 16) & 0xff; p[3] = u >> 24; }
__device__ __forceinline__ void put_f32(unsigned char* p, float v) { put_i32(p, __float_as_int(v)); }

__global__ void __launch_bounds__(kWireThreads)
spectrum_frame_kernel(SpectrumFrameArgs a)
{
    __shared__ float s_mn[kWireThreads / 32], s_mx[kWireThreads / 32];
    const int slot = blockIdx.x, ch = a.ch0 + slot, tid = threadIdx.x;
    const ChanState& st = a.state[ch];
    unsigned char* out = a.out + size_t(slot) * a.out_pitch;
    if (!st.have_spectrum) { if (tid == 0) a.sizes[slot] = 0; return; }   // getSpectrumInfo: empty vector, no message
    const size_t n = size_t(a.fft_n);
    const float* v = a.power + size_t(ch) * n;
    const float zoom = fminf(fmaxf(a.zoom, 0.01f), 0.99f);
    const size_t zb = size_t(__fmul_rn(zoom / 2, float(n)));
    const size_t ze = size_t(__fmul_rn(1.0f - zoom / 2, float(n)));
    const size_t m = ze - zb;                                             // bins left after the two erase() calls
    int pl = abs(st.gui_left), pr = abs(st.gui_right);
    int plv = st.gui_left > 0, prv = st.gui_right > 0;
    pl -= int(zb);
    if (pl < 0 || size_t(pl) > m) { pl = 0; plv = 0; }
    pr -= int(zb);
    if (pr < 0 || size_t(pr) > m) { pr = 0; prv = 0; }
    const bool shrink = size_t(a.resolution) < m;                        // int compared as size_t, like the reference
    const size_t out_n = shrink ? size_t(a.resolution) : m;
    if (shrink) {
        pl = int(double(pl) * a.resolution / double(m));
        pr = int(double(pr) * a.resolution / double(m));
    }
    if (out_n == 0) { if (tid == 0) a.sizes[slot] = 0; return; }
    auto src_index = [&](size_t i) -> size_t {
        if (!shrink) return zb + i;
        const float i_0_1 = __fdiv_rn(float(i), float(out_n));           // float(i) / new_size
        return zb + size_t(__fmul_rn(i_0_1, float(m)));                   // size_t I = i_0_1 * vec.size()
    };
    float mn = INFINITY, mx = -INFINITY;
    for (size_t i = tid; i < out_n; i += kWireThreads) { const float x = v[src_index(i)]; mn = fminf(mn, x); mx = fmaxf(mx, x); }
    block_minmax(mn, mx, s_mn, s_mx);
    unsigned char* body = out + kSpectrumHeaderBytes;
    for (size_t i = tid; i < out_n; i += kWireThreads) store_value(body, i, v[src_index(i)], double(mn), double(mx), a.type_size);
    if (tid == 0) {
        put_i32(out + 0, kSpectrumHeaderBytes);
        put_f32(out + 4, float(st.afc_noise_floor));
        put_f32(out + 8, float(st.afc_noise_var));
        put_f32(out + 12, float(a.fs_dec));
        put_f32(out + 16, float(st.afc_shift_hz));
        put_i32(out + 20, pl); put_i32(out + 24, pr); put_i32(out + 28, plv); put_i32(out + 32, prv);
        put_f32(out + 36, mn); put_f32(out + 40, mx);
        put_i32(out + 44, a.type_size); put_i32(out + 48 + 0, int(out_n));
        a.sizes[slot] = unsigned(kSpectrumHeaderBytes + out_n * size_t(a.type_size));
    }
}

// main.cpp:267-282: after every process() the last demodulated block is appended and the front trimmed to 50 symbols
__global__ void __launch_bounds__(128)
demod_acc_kernel(DemodAccArgs a)
{
    const int ch = a.ch0 + blockIdx.x, tid = threadIdx.x;
    const ChanState& st = a.state[ch];
    const unsigned k = st.demod_n;                                        // size of Decoder::demodulated_ (stale blocks are re-appended)
    float* acc = a.acc + size_t(ch) * a.acc_pitch;
    const unsigned len = a.acc_n[ch];
    const double sym = st.baud;
    const size_t max_sz = size_t(a.fs_dec / sym * 50);                    // getDecimatedSamplingRate() / getSymbolRate() * 50
    const float* d = a.demod + size_t(ch) * a.demod_pitch;
    const size_t total = size_t(len) + k;
    const size_t drop = total > max_sz ? total - max_sz : 0;              // erase(begin, begin + size - max_sz)
    const size_t new_len = total - drop;
    // new[j] = (j + drop < len) ? old[j + drop] : d[j + drop - len]; ascending chunks, reads before writes
    for (size_t j0 = 0; j0 < new_len; j0 += 128) {
        const size_t j = j0 + tid;
        float x = 0.f;
        if (j < new_len) { const size_t s = j + drop; x = s < len ? acc[s] : d[s - len]; }
        __syncthreads();
        if (j < new_len) acc[j] = x;
        __syncthreads();
    }
    if (tid == 0) a.acc_n[ch] = unsigned(new_len);
}

__global__ void __launch_bounds__(kWireThreads)
demod_frame_kernel(DemodFrameArgs a)
{
    __shared__ float s_mn[kWireThreads / 32], s_mx[kWireThreads / 32];
    const int slot = blockIdx.x, ch = a.ch0 + slot, tid = threadIdx.x;
    unsigned char* out = a.out + size_t(slot) * a.out_pitch;
    const size_t n = a.acc_n[ch];
    const bool shrink = size_t(a.resolution) < n;
    const size_t out_n = shrink ? size_t(a.resolution) : n;
    if (n == 0 || out_n == 0) { if (tid == 0) a.sizes[slot] = 0; return; }
    const float* v = a.acc + size_t(ch) * a.acc_pitch;
    auto src_index = [&](size_t i) -> size_t {
        if (!shrink) return i;
        return size_t(__fmul_rn(__fdiv_rn(float(i), float(out_n)), float(n)));
    };
    float mn = INFINITY, mx = -INFINITY;
    for (size_t i = tid; i < out_n; i += kWireThreads) { const float x = v[src_index(i)]; mn = fminf(mn, x); mx = fmaxf(mx, x); }
    block_minmax(mn, mx, s_mn, s_mx);
    unsigned char* body = out + kDemodHeaderBytes;
    for (size_t i = tid; i < out_n; i += kWireThreads) store_value(body, i, v[src_index(i)], double(mn), double(mx), a.type_size);
    if (tid == 0) {
        put_i32(out + 0, kDemodHeaderBytes); put_f32(out + 4, mn); put_f32(out + 8, mx);
        put_i32(out + 12, a.type_size); put_i32(out + 16, int(out_n));
        a.sizes[slot] = unsigned(kDemodHeaderBytes + out_n * size_t(a.type_size));
    }
}

cudaError_t launch_spectrum_frames(const SpectrumFrameArgs& a, int n_channels, cudaStream_t stream, int* launches)
{
    spectrum_frame_kernel<<<n_channels, kWireThreads, 0, stream>>>(a);
    if (launches) ++*launches;
    return cudaGetLastError();
}
cudaError_t launch_demod_accumulate(const DemodAccArgs& a, int n_channels, cudaStream_t stream, int* launches)
{
    demod_acc_kernel<<<n_channels, 128, 0, stream>>>(a);
    if (launches) ++*launches;
    return cudaGetLastError();
}
cudaError_t launch_demod_frames(const DemodFrameArgs& a, int n_channels, cudaStream_t stream, int* launches)
{
    demod_frame_kernel<<<n_channels, kWireThreads, 0, stream>>>(a);
    if (launches) ++*launches;
    return cudaGetLastError();
}

} // namespace hbd
