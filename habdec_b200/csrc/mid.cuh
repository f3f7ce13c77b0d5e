// Middle stages of a cascaded decimation plan (see mid.cu)
#pragma once
#include "hbd_common.cuh"

namespace hbd {

struct MidArgs {
    const ChanPlan* plan; ChanPlan uplan; int uniform;
    const float2* in;      // [channel][in_pitch]: history slots [0, in_hist), this call's samples from in_hist on
    float2* in_next;       // buffer that receives the stage's history for the next call (same layout; may equal `in`)
    size_t in_pitch; int in_hist;
    float2* out;           // [channel][out_pitch]: outputs from out_hist on
    size_t out_pitch; int out_hist;
    const float* taps;     // T floats (device)
    int M, T;
    unsigned div_in;       // samples of this stage's input per call = plan.consumed / div_in
};

cudaError_t launch_mid_stage(const MidArgs& a, int n_channels, cudaStream_t stream, int* launches);

} // namespace hbd
