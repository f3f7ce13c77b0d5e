// K5 launch interface (see wire.cu): websocket wire formats produced on the GPU
#pragma once
#include "hbd_common.cuh"

namespace hbd {

constexpr int kSpectrumHeaderBytes = 52; // SpectrumInfoHeader, NetTransport.h:29-47 (13 x 4 bytes)
constexpr int kDemodHeaderBytes = 20;    // DemodHeader, NetTransport.h:50-57

struct SpectrumFrameArgs {
    const ChanState* state;
    const float* power; int fft_n; // [channel][fft_n] dB spectrum
    double fs_dec;
    float zoom; int resolution; int type_size;   // 1: u8, 2: u16, 4: f32
    unsigned char* out; size_t out_pitch;        // [channel slot][out_pitch] bytes
    unsigned* sizes;                             // [channel slot] bytes written (0: no spectrum yet)
    int ch0;                                     // first channel; slot = blockIdx.x
};
cudaError_t launch_spectrum_frames(const SpectrumFrameArgs& a, int n_channels, cudaStream_t stream, int* launches);

struct DemodAccArgs {
    const ChanState* state;
    const float* demod; size_t demod_pitch;      // last call's discriminator output [channel][demod_pitch]
    float* acc; size_t acc_pitch; unsigned* acc_n; // accumulated samples [channel][acc_pitch], counts
    double fs_dec;
    int ch0;
};
cudaError_t launch_demod_accumulate(const DemodAccArgs& a, int n_channels, cudaStream_t stream, int* launches);

struct DemodFrameArgs {
    const float* acc; size_t acc_pitch; const unsigned* acc_n;
    int resolution; int type_size;
    unsigned char* out; size_t out_pitch; unsigned* sizes;
    int ch0;
};
cudaError_t launch_demod_frames(const DemodFrameArgs& a, int n_channels, cudaStream_t stream, int* launches);

} // namespace hbd
