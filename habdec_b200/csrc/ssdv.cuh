// K6 launch interface (see ssdv.cu): SSDV packet sync on the GPU
#pragma once
#include "hbd_common.cuh"

namespace hbd {

constexpr unsigned kSsdvRing   = 4096;  // raw-character ring per channel (bytes, power of two)
constexpr unsigned kSsdvPkt    = 256;   // SSDV_PKT_SIZE
constexpr unsigned kSsdvLogCap = 4096;  // accepted-packet log entries (ring) between host drains

// one accepted packet: where it starts in the channel's raw character stream, and its (corrected) bytes
struct SsdvLogEntry {
    unsigned ch;
    unsigned seq;       // call that completed the window (24 bits, like the character log)
    unsigned pos;       // stream index of the sync byte (mod 2^32)
    int      errors;    // symbols the Reed-Solomon decoder corrected
    unsigned char data[kSsdvPkt];
};

struct SsdvScanArgs {
    const unsigned char* ring;   // [channel][kSsdvRing]
    const unsigned* total;       // [channel] raw characters appended so far (written by the tail kernel)
    unsigned* scanned;           // [channel] window starts below this index have been examined
    SsdvLogEntry* log;           // accepted packets (ring of kSsdvLogCap entries)
    unsigned* ctl;               // LogCtl words: kCtlSsdvHead (monotonic), kCtlSsdvTail, kCtlSsdvOvf, kCtlSsdvRingOvf
    unsigned call_seq;
    int ch0;
};
cudaError_t launch_ssdv_scan(const SsdvScanArgs& a, int n_channels, cudaStream_t stream, int* launches);

// the packet test alone for n candidate windows: windows[n][256] corrected in place, verdict[i] = 0 / -1
cudaError_t launch_ssdv_check(unsigned char* windows, int n, int* verdict, int* errors, cudaStream_t stream, int* launches);

} // namespace hbd
