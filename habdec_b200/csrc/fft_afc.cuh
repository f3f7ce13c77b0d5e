// K4 launch interface (see fft_afc.cu)
#pragma once
#include "hbd_common.cuh"

namespace hbd {

struct FftArgs {
    ChanState* state;
    const float2* fftbuf;   // [channel][kFftN] frame being collected
    float2* spectrum;       // [channel][kFftN] last fft-shifted spectrum (getFFT)
    float* power;           // [channel][kFftN] last power spectrum in dB (getPowerSpectrum)
    const float2* twiddle;  // [kFftN] exp(-2 pi i e / kFftN), evaluated in float64 on the host
    double fs_dec;
    int ch0;                // first channel of this launch
};

cudaError_t launch_fft_afc(const FftArgs& a, int n_channels, cudaStream_t stream, int* launches);
cudaError_t launch_afc_reset(ChanState* state, int ch, double corr, double fs_dec, cudaStream_t stream);
cudaError_t launch_afc_retune(ChanState* state, int n_ch, double min_abs_hz, double fs_dec, double* applied, cudaStream_t stream);

} // namespace hbd
