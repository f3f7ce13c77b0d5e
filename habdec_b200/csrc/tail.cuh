// K2 launch interface (see tail.cu)
#pragma once
#include "hbd_common.cuh"

namespace hbd {

struct TailArgs {
    const ChanPlan* plan;     // per-channel plans (read only when !uniform)
    ChanPlan uplan;           // the plan of every channel when uniform (saves a dependent global load)
    int uniform;
    ChanState* state;
    int ch0;                  // first channel of this launch (channel groups run on their own streams)
    // stage 2
    float2* s1; float2* s1_next; size_t s1_pitch;  // this call's stage-1 stream; the next call's buffer (receives the stage-2 history)
    const float* taps2; int M2, T2;
    // decimated queue: [kLpHist slots mirroring the head of the reference's low-pass work buffer: history, then the last
    // call's inputs | samples pending in front of the low-pass]
    float2* decq; size_t dq_pitch;
    double fs_dec;
    // FFT frame buffers [channel][fft_n]
    float2* fftbuf; int fft_n;
    // low-pass taps [channel][kLpMaxTaps]
    const float* lptaps;
    int max_lp_taps;          // largest lp_ntaps of any channel (sizes the shared-memory queue)
    // slicer pending samples [channel][slicer_pitch]
    float* slicer; size_t slicer_pitch;
    int sv_want;              // slicer samples to stage in shared memory (channels with more run from HBM)
    // decoded characters go to one device-wide append log (ring of log_mask + 1 entries, monotonic head in
    // log_ctl[kCtlCharHead]): entry = (channel, call_seq << 8 | char).  Kernels of successive calls run in stream order,
    // so the log is sorted by call and, per channel, by time.
    uint2* log; unsigned* log_ctl; unsigned log_mask; unsigned call_seq;
    unsigned* maskc;             // cached slicer position masks [channel][2][kMaskWords] (null: rebuilt every call)
    unsigned short* uart_runs;   // UART backlog [channel][kUartRunsCap] (slicer_dev.cuh)
    // SSDV packet sync (ssdv.cu): per-channel raw-character ring [channel][kSsdvRing] + append counts; null = off
    unsigned char* ssdv_ring; unsigned* ssdv_total;
    // last call's discriminator output for getDemodulated() [channel][demod_pitch] (may be null)
    float* demod_last; size_t demod_pitch;
    // optional per-call stage recordings for parity tests [channel][rec_pitch]
    float2* rec_decimated; float2* rec_filtered; size_t rec_pitch;
    unsigned char* rec_bits; unsigned* rec_bits_n; unsigned rec_bits_pitch;  // every emitted bit [channel][rec_bits_pitch]
    // shared-memory layout, filled by launch_tail
    int xw, qcap, h2cap, hlcap, sv_cap, mask_words;
};

cudaError_t launch_tail(TailArgs a, int n_channels, cudaStream_t stream, int* launches);


} // namespace hbd
