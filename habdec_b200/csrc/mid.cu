// Middle stages of a cascaded decimation plan.
//
// Decoder::setupDecimationStagesBW (code/Decoder/Decoder.h:336-412) keeps dividing by up to 256 until the rate is under
// the limit, so a 20 MS/s input and a 5 kHz limit give FOUR decimators: (64,348t) (4,139t) (8,54t) (2,69t).  K1 runs the
// first one at the input rate, the tail kernel the last one; the ones in between see at most 1/64 of the input rate and
// run here, one CTA per channel, straight from the definition (Decimator.h:99-146):
//     y[k] = sum_t buf[k M + t] h[t],   buf = [T-1 history | input]
// including the reference's in-place quirk: the stages share one work buffer (Decoder.h:441-446), so when the history
// (the last T-1 inputs, Decimator.h:141-143) is copied, the outputs already sit on the head of the input -- for short
// calls (n_in - (T-1) < n_out) the head of the next history therefore holds OUTPUTS.
#include "mid.cuh"

namespace hbd {

constexpr int kMidThreads = 256;

__global__ void __launch_bounds__(kMidThreads) mid_stage_kernel(MidArgs a)
{
    const int ch = blockIdx.x, tid = threadIdx.x;
    const ChanPlan pl = a.uniform ? a.uplan : a.plan[ch];
    const float2* in = a.in + (size_t)ch * a.in_pitch + a.in_hist;        // in[j], j in [-(T-1), n_in)
    float2* in_next = a.in_next + (size_t)ch * a.in_pitch + a.in_hist;
    float2* out = a.out + (size_t)ch * a.out_pitch + a.out_hist;
    const int T = a.T, M = a.M;
    if (pl.flags & 1u) {   // nothing consumed this call: only carry the history over to the buffer of the next call
        if (in_next != in) for (int i = tid; i < T - 1; i += kMidThreads) in_next[i - (T - 1)] = in[i - (T - 1)];
        return;
    }
    const int n_in = int(pl.consumed / a.div_in), n_out = n_in / M;
    extern __shared__ float s_taps[];
    for (int t = tid; t < T; t += kMidThreads) s_taps[t] = a.taps[t];
    __syncthreads();
    for (int k = tid; k < n_out; k += kMidThreads) {
        const float2* w = in + (long long)k * M - (T - 1);
        float2 acc = make_float2(0.f, 0.f);
        for (int t = 0; t < T; ++t) acc = cfma(w[t], s_taps[t], acc);
        out[k] = acc;
    }
    // history for the next call: buf[n_in .. n_in + T-1) of [history | input] == input positions p = n_in - (T-1) + i
    float2 keep[2];
    const int p0 = n_in - (T - 1);
    __syncthreads();       // the outputs this CTA wrote are visible to it (in-place quirk reads them back)
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const int i = tid + u * kMidThreads;
        if (i < T - 1) {
            const int p = p0 + i;
            keep[u] = (p >= 0 && p < n_out) ? out[p] : in[p];   // p < 0: older history (chunks shorter than T-1: streaming semantics)
        }
    }
    __syncthreads();       // every read of the old history is done before it is overwritten (in_next may be `in`)
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const int i = tid + u * kMidThreads;
        if (i < T - 1) in_next[i - (T - 1)] = keep[u];
    }
}

cudaError_t launch_mid_stage(const MidArgs& a, int n_channels, cudaStream_t stream, int* launches)
{
    if (a.T - 1 > 2 * kMidThreads || a.T - 1 > a.in_hist) return cudaErrorInvalidValue;
    mid_stage_kernel<<<n_channels, kMidThreads, size_t(a.T) * sizeof(float), stream>>>(a);
    if (launches) ++*launches;
    return cudaGetLastError();
}

} // namespace hbd
