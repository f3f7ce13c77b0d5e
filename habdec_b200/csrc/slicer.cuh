// K3 launch interface (see slicer.cu)
#pragma once
#include "hbd_common.cuh"

namespace hbd {

struct SlicerArgs {
    ChanState* state;
    float* slicer; size_t slicer_pitch;  // pending discriminator samples [channel][slicer_pitch]
    // decoded characters go to one device-wide append log (ring of kLogCap entries, monotonic head):
    // entry = (channel, call_seq << 8 | char).  Kernels of successive calls run in stream order, so the log is
    // sorted by call and, per channel, by time.
    uint2* log; unsigned* log_head; unsigned call_seq;
    unsigned char* rec_bits;             // optional: every emitted bit [channel][rec_bits_pitch]
    unsigned* rec_bits_n;
    unsigned rec_bits_pitch;
    double fs_dec;
    int ch0;                             // first channel of this launch
    int n_channels;                      // channels in this launch
};

cudaError_t launch_slicer(const SlicerArgs& a, cudaStream_t stream, int* launches);

} // namespace hbd
