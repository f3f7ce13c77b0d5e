// K1 launch interface (see decim1.cu)
#pragma once
#include "hbd_common.cuh"

namespace hbd {

#ifndef HBD_K1_WARPS
#define HBD_K1_WARPS 8
#endif
constexpr int kDecimWarps = HBD_K1_WARPS; // warps per CTA; each warp runs its own TMA ring

struct DecimArgs {
    const float2* chunk;       // pushed samples, [channel][chunk_pitch] cf32, chunk[j] for j in [0, n)
    size_t chunk_pitch;        // in samples, even (16-byte rows)
    const float2* carry;       // [channel][kCarryCap], right aligned: sample j (< 0) at carry[kCarryCap + j]
    float2* s1;                // stage-1 output stream, [channel][s1_pitch]; outputs start at s1_hist
    size_t s1_pitch;
    int s1_hist;
    const ChanPlan* plan;      // per channel
    const float* taps;         // T floats (device)
    int ch0;                   // first channel of this launch
    int n_channels;            // channels in this launch
    int stretches_per_channel; // ceil(max superblocks / sb_per_stretch)
    int sb_per_stretch;        // owned superblocks per work item
};

// M == 1 means "no decimator" (copy).  `launches` is incremented per kernel launched.
cudaError_t launch_decim1(const DecimArgs& a, int M, int T, unsigned max_n1, int n_sms, cudaStream_t stream, int* launches);
int decim1_sb_per_stretch(int M);

} // namespace hbd
