// K4 launch interface (see fft_afc.cu)
#pragma once
#include "hbd_common.cuh"

namespace hbd {

struct FftArgs {
    ChanState* state;
    const float2* fftbuf;   // [channel][fft_n] frame being collected
    float2* spectrum;       // [channel][fft_n] last fft-shifted spectrum (getFFT)
    float* power;           // [channel][fft_n] last power spectrum in dB (getPowerSpectrum)
    const float2* twiddle;  // [fft_n] exp(-2 pi i e / fft_n) + [16][256] pass-1 + [16][16] pass-2 twiddles of the 4096-point transform (float64-evaluated, host)
    double fs_dec;
    int ch0;                // first channel of this launch
    int fft_n;              // 4096 (reference) or 16384
    int n_channels;         // channels of this launch (set by launch_fft_afc)
};

cudaError_t launch_fft_afc(const FftArgs& a, int n_channels, cudaStream_t stream, int* launches);
cudaError_t launch_afc_reset(ChanState* state, int ch, double corr, double fs_dec, int n_fft, cudaStream_t stream);
cudaError_t launch_afc_retune(ChanState* state, int n_ch, double min_abs_hz, double fs_dec, double* applied, int n_fft, cudaStream_t stream);

} // namespace hbd
