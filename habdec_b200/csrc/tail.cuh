// K2 launch interface (see tail.cu)
#pragma once
#include "hbd_common.cuh"

namespace hbd {

struct TailArgs {
    const ChanPlan* plan;
    ChanState* state;
    int ch0;                  // first channel of this launch (channel groups run on their own streams)
    // stage 2
    float2* s1; size_t s1_pitch;
    const float* taps2; int M2, T2;
    // decimated queue: [kLpHist history slots | pending]
    float2* decq; size_t dq_pitch;
    double fs_dec;
    // FFT frame buffers [channel][kFftN]
    float2* fftbuf;
    // low-pass taps [channel][kLpMaxTaps]
    const float* lptaps;
    // slicer pending samples [channel][slicer_pitch]
    float* slicer; size_t slicer_pitch;
    // last call's discriminator output for getDemodulated() [channel][demod_pitch] (may be null)
    float* demod_last; size_t demod_pitch;
    // optional per-call stage recordings for parity tests [channel][rec_pitch]
    float2* rec_decimated; float2* rec_filtered; size_t rec_pitch;
    int smem_window; // float2 slots of the tile window (tail_smem_window)
};

cudaError_t launch_tail(const TailArgs& a, int n_channels, cudaStream_t stream, int* launches);
// stage-1 carry (history + unconsumed remainder) for the next call; must run after K1 of the same call
cudaError_t launch_carry(const ChanPlan* plan, const float2* chunk, size_t chunk_pitch, float2* carry, int T1, int ch0, int n_channels,
                         cudaStream_t stream, int* launches);
int tail_smem_window(int M2, int T2);

} // namespace hbd
