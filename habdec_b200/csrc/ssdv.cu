// K6 -- SSDV packet sync, the consumer of the raw UART characters (Decoder.h:572-573 hands every call's characters to
// SSDV_wraper_t::push, code/Decoder/ssdv_wrapper.cpp:37-148).
//
// What the reference does per channel, sequentially: keep a byte buffer, find the first 0x55, and once 256 bytes are
// there ask fsphil/ssdv's ssdv_dec_is_packet (ssdv_wrapper.cpp:66) whether a packet starts at it -- a CRC-32 test, and
// for anything that is not already clean a Reed-Solomon (255,223) decode followed by the CRC again.  On a noisy channel
// one character in 256 is a 0x55, so with thousands of channels the packet test is the expensive part (tens of
// microseconds of scalar CPU work per candidate) while the buffer bookkeeping is trivial.
//
// Split used here: every window the wrapper can ever test is 256 CONSECUTIVE characters of the channel's stream that
// start at a 0x55 (host_tail.cpp, SsdvChannel, shows why), so the verdict is a pure function of the stream position.
// The tail kernel appends each channel's raw characters to a per-channel ring; this kernel, one warp per channel, tests
// every 0x55-started window as soon as its 256th byte has arrived and appends the ACCEPTED packets (corrected bytes +
// stream position) to a device-wide log.  The host replays the wrapper's buffer automaton call by call and looks the
// verdicts up by position; rejected candidates never leave the GPU.
//
// ssdv_dec_is_packet itself is third-party code that is absent from /root/reference (un-vendored submodule
// fsphil/ssdv @ 1de34b9): what is implemented is its PUBLISHED algorithm -- packet layout and sanity checks of ssdv.c,
// CRC-32 (reflected 0xEDB88320), and Phil Karn's CCSDS Reed-Solomon codec rs8.c (GF(256) polynomial 0x187, first root
// 112, primitive element 11, 32 roots; syndromes -> Berlekamp-Massey -> Chien -> Forney, same accept / reject rule
// "number of roots == degree of the locator").  Parity is checked against oracle/ssdv_published.h (unpinned below that).
//
// Warp-level mapping: lane i owns syndrome i (Horner over the 255 symbols); in Berlekamp-Massey lane j owns
// coefficient j of the locator and of the shift register (coefficient 32 is kept by every lane), the discrepancy is a
// shuffle + xor-reduction; the Chien search tests 32 field elements per step; Forney handles one error per lane.
#include "ssdv.cuh"

namespace hbd {

constexpr int kSsdvWarps = 4;            // warps (= channels / candidate windows) per CTA
constexpr int kNN = 255, kNRoots = 32, kFcr = 112, kPrim = 11, kIprim = 116;

struct SsdvTables {
    unsigned char exp[512];   // alpha^i for i in [0, 510): a sum of two logarithms needs no reduction
    unsigned char log[256];   // log[0] = 255
    unsigned crc[256];
};
struct SsdvScratch {          // per warp
    unsigned char pkt[kSsdvPkt];
    unsigned char syn[kNRoots];       // syndromes, polynomial form
    unsigned char lam[kNRoots + 1];   // locator, polynomial form
    unsigned char omega[kNRoots];
    unsigned char root[kNRoots];
    unsigned char pad[27];
};

__device__ __forceinline__ void build_tables(SsdvTables& t)
{
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
        unsigned x = unsigned(i);
        for (int k = 0; k < 8; ++k) x = (x & 1u) ? (x >> 1) ^ 0xEDB88320u : x >> 1;
        t.crc[i] = x;
    }
    if (threadIdx.x == 0) {
        unsigned x = 1;
        for (int i = 0; i < kNN; ++i) {
            t.exp[i] = (unsigned char)x; t.exp[i + kNN] = (unsigned char)x; t.log[x] = (unsigned char)i;
            x <<= 1; if (x & 0x100u) x ^= 0x187u;
        }
        t.exp[510] = t.exp[0]; t.exp[511] = t.exp[1];
        t.log[0] = 255;
    }
    __syncthreads();
}

__device__ __forceinline__ unsigned gmul(const SsdvTables& t, unsigned a, unsigned b)
{
    return (a && b) ? t.exp[t.log[a] + t.log[b]] : 0u;
}
// a * alpha^e, e in [0, 255)
__device__ __forceinline__ unsigned gmul_pow(const SsdvTables& t, unsigned a, unsigned e) { return a ? t.exp[t.log[a] + e] : 0u; }

__device__ __forceinline__ unsigned warp_xor(unsigned v) { return __reduce_xor_sync(0xffffffffu, v); }

// CRC-32 over pkt[1 .. 1+len) against the 4 big-endian bytes that follow; lane 0 walks the table, the verdict is broadcast
__device__ __forceinline__ bool crc_ok(const SsdvTables& t, const unsigned char* pkt, int len, int lane)
{
    int ok = 0;
    if (lane == 0) {
        unsigned crc = 0xFFFFFFFFu;
        for (int i = 1; i <= len; ++i) crc = (crc >> 8) ^ t.crc[(crc ^ pkt[i]) & 0xFFu];
        crc ^= 0xFFFFFFFFu;
        const unsigned want = (unsigned(pkt[len + 1]) << 24) | (unsigned(pkt[len + 2]) << 16) | (unsigned(pkt[len + 3]) << 8) | unsigned(pkt[len + 4]);
        ok = crc == want;
    }
    return __shfl_sync(0xffffffffu, ok, 0) != 0;
}

// decode_rs_8(data[255]) of the published codec: corrects in place, returns the number of corrected symbols, -1 if the
// word is uncorrectable.  Warp collective; `data` and the scratch live in shared memory.
__device__ int rs_decode_warp(const SsdvTables& t, unsigned char* data, SsdvScratch& w, int lane)
{
    // syndromes: lane i evaluates the word at alpha^((FCR+i)*PRIM)
    const unsigned beta = unsigned((kFcr + lane) * kPrim) % kNN;
    unsigned s = data[0];
    for (int j = 1; j < kNN; ++j) s = unsigned(data[j]) ^ gmul_pow(t, s, beta);
    if (__ballot_sync(0xffffffffu, s != 0) == 0u) return 0;
    w.syn[lane] = (unsigned char)s;

    // Berlekamp-Massey: lane j holds lambda[j] and b[j]; coefficient 32 is replicated
    unsigned lam = lane == 0 ? 1u : 0u, b = lam, lam32 = 0u;   // b[32] is never read (t[i+1] uses b[i], i <= 31)
    int el = 0;
    for (int r = 1; r <= kNRoots; ++r) {
        const unsigned sv = __shfl_sync(0xffffffffu, s, (r - 1 - lane) & 31);
        const unsigned discr = warp_xor(lane < r ? gmul(t, lam, sv) : 0u);
        unsigned bprev = __shfl_up_sync(0xffffffffu, b, 1);
        if (lane == 0) bprev = 0u;
        const unsigned b31 = __shfl_sync(0xffffffffu, b, 31);
        if (discr == 0u) {
            b = bprev;
        } else {
            const unsigned tl = lam ^ gmul(t, discr, bprev);
            const unsigned tl32 = lam32 ^ gmul(t, discr, b31);
            if (2 * el <= r - 1) {
                el = r - el;
                const unsigned inv = unsigned(kNN - t.log[discr]);   // in [1, 255]; exp[] covers log + 255
                b = lam ? t.exp[t.log[lam] + inv] : 0u;
            } else {
                b = bprev;
            }
            lam = tl; lam32 = tl32;
        }
    }
    const unsigned nz = __ballot_sync(0xffffffffu, lam != 0u);
    const int deg = lam32 ? 32 : (nz ? 31 - __clz(nz) : 0);
    w.lam[lane] = (unsigned char)lam;
    if (lane == 0) w.lam[32] = (unsigned char)lam32;
    __syncwarp();

    // Chien search: lambda(alpha^i) for i = 1..255, 32 values of i at a time
    int count = 0;
    for (int base = 1; base <= kNN; base += 32) {
        const int i = base + lane;
        unsigned q = 1u;
        if (i <= kNN) {
            unsigned e = 0;   // (i * j) mod 255
            for (int j = 1; j <= deg; ++j) {
                e += unsigned(i); if (e >= unsigned(kNN)) e -= unsigned(kNN);
                q ^= gmul_pow(t, w.lam[j], e);
            }
        }
        const unsigned hit = __ballot_sync(0xffffffffu, i <= kNN && q == 0u);
        if (i <= kNN && q == 0u) {
            const int slot = count + __popc(hit & ((1u << lane) - 1u));
            if (slot < kNRoots) w.root[slot] = (unsigned char)i;
        }
        count += __popc(hit);
    }
    if (count != deg) return -1;
    __syncwarp();

    // omega = syndrome polynomial * lambda mod x^deg; lane i computes coefficient i
    {
        unsigned o = 0u;
        if (lane < deg)
            for (int j = 0; j <= lane; ++j) o ^= gmul(t, w.syn[lane - j], w.lam[j]);
        w.omega[lane] = (unsigned char)o;
    }
    __syncwarp();

    // Forney: one error per lane
    if (lane < count) {
        const unsigned rt = w.root[lane];
        unsigned num1 = 0u, den = 0u, e = 0u;   // e = (i * rt) mod 255
        const int top = min(deg, kNRoots - 1) & ~1;
        for (int i = 0; i < deg || i <= top; ++i) {
            if (i < deg) num1 ^= gmul_pow(t, w.omega[i], e);
            if (!(i & 1) && i <= top) den ^= gmul_pow(t, w.lam[i + 1], e);
            e += rt; if (e >= unsigned(kNN)) e -= unsigned(kNN);
        }
        if (num1 != 0u) {
            const unsigned num2_log = (rt * unsigned(kFcr - 1)) % unsigned(kNN);
            const unsigned den_log = den ? unsigned(t.log[den]) : 255u;   // the published code divides by "index of 0" = 255
            const unsigned loc = (rt * unsigned(kIprim) + unsigned(kNN) - 1u) % unsigned(kNN);
            data[loc] ^= t.exp[(unsigned(t.log[num1]) + num2_log + unsigned(kNN) - den_log) % unsigned(kNN)];
        }
    }
    __syncwarp();
    return count;
}

// ssdv_dec_is_packet on w.pkt (already loaded, 256 bytes).  Returns true when a packet starts here; w.pkt then holds
// the corrected packet and errors the corrected symbol count.
__device__ bool is_packet_warp(const SsdvTables& t, SsdvScratch& w, int& errors, int lane)
{
    unsigned char* pkt = w.pkt;
    if (lane == 0) pkt[0] = 0x55;
    __syncwarp();
    const unsigned t1 = pkt[1];
    int type = 0xFF, payload = 0;
    errors = 0;
    if (t1 == 0x67u) {
        payload = 256 - 15 - 4;
        if (crc_ok(t, pkt, 15 + payload - 1, lane)) type = 1;
    } else if (t1 == 0x66u) {
        payload = 256 - 15 - 4 - 32;
        if (crc_ok(t, pkt, 15 + payload - 1, lane)) type = 0;
    }
    if (type == 0xFF) {
        payload = 256 - 15 - 4 - 32;
        __syncwarp();
        if (lane == 0) pkt[1] = 0x66;
        __syncwarp();
        const int n = rs_decode_warp(t, pkt + 1, w, lane);
        if (n < 0) return false;
        errors = n;
        if (crc_ok(t, pkt, 15 + payload - 1, lane)) type = 0;
    }
    if (type == 0xFF) return false;
    // sanity checks on the header (ssdv_dec_header of the published library)
    if (int(pkt[1]) - 0x66 != type) return false;
    if (pkt[9] == 0 || pkt[10] == 0) return false;
    const unsigned mcu_id = (unsigned(pkt[13]) << 8) | pkt[14];
    if (mcu_id != 0xFFFFu) {
        unsigned mcu_count = unsigned(pkt[9]) * unsigned(pkt[10]);
        if (pkt[11] & 2u) mcu_count *= 2u;
        if (pkt[11] & 1u) mcu_count *= 2u;
        mcu_count &= 0xFFFFu;                         // uint16_t field
        if (mcu_id >= mcu_count) return false;
        if (unsigned(pkt[12]) >= unsigned(payload)) return false;
    }
    return true;
}

__global__ void __launch_bounds__(kSsdvWarps * 32) ssdv_scan_kernel(SsdvScanArgs a, int n_channels)
{
    __shared__ SsdvTables tab;
    __shared__ __align__(16) SsdvScratch scratch[kSsdvWarps];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int slot = blockIdx.x * kSsdvWarps + warp;
    // cheap exit for the whole CTA when none of its channels has a complete window to examine
    bool work = false;
    unsigned total = 0, from = 0;
    if (slot < n_channels) {
        total = a.total[a.ch0 + slot];
        from = a.scanned[a.ch0 + slot];
        work = total - from >= kSsdvPkt;   // stream indices wrap mod 2^32; total - from is the distance
    }
    if (!__syncthreads_or(work ? 1 : 0)) return;
    build_tables(tab);
    if (!work) return;

    const int ch = a.ch0 + slot;
    const unsigned char* ring = a.ring + size_t(ch) * kSsdvRing;
    SsdvScratch& w = scratch[warp];
    const unsigned end = total - (kSsdvPkt - 1u);     // window starts in [from, end) are complete
    if (total - from > kSsdvRing) {                    // the ring no longer holds [from, total)
        if (lane == 0) atomicAdd(a.ctl + kCtlSsdvRingOvf, 1u);
        from = total - kSsdvRing;
    }
    for (unsigned base = from; int(end - base) > 0; base += 32u) {
        const unsigned p = base + unsigned(lane);
        const bool cand = int(end - p) > 0 && ring[p & (kSsdvRing - 1u)] == 0x55;
        unsigned hits = __ballot_sync(0xffffffffu, cand);
        while (hits) {
            const int k = __ffs(hits) - 1;
            hits &= hits - 1u;
            const unsigned q = base + unsigned(k);
            for (int i = lane; i < int(kSsdvPkt); i += 32) w.pkt[i] = ring[(q + unsigned(i)) & (kSsdvRing - 1u)];
            __syncwarp();
            int errors = 0;
            if (is_packet_warp(tab, w, errors, lane)) {
                unsigned at = 0, tail = 0;
                if (lane == 0) { at = atomicAdd(a.ctl + kCtlSsdvHead, 1u); tail = *reinterpret_cast<volatile unsigned*>(a.ctl + kCtlSsdvTail); }
                at = __shfl_sync(0xffffffffu, at, 0);
                tail = __shfl_sync(0xffffffffu, tail, 0);
                if (at - tail >= kSsdvLogCap) {        // full: never overwrite an entry the host has not read (counted, reported)
                    if (lane == 0) atomicAdd(a.ctl + kCtlSsdvOvf, 1u);
                } else {
                    SsdvLogEntry& e = a.log[at & (kSsdvLogCap - 1u)];
                    if (lane == 0) { e.ch = unsigned(ch); e.seq = a.call_seq; e.pos = q; e.errors = errors; }
                    __syncwarp();
                    for (int i = lane; i < int(kSsdvPkt / 4); i += 32)
                        reinterpret_cast<unsigned*>(e.data)[i] = reinterpret_cast<const unsigned*>(w.pkt)[i];
                }
            }
            __syncwarp();
        }
    }
    if (lane == 0) a.scanned[ch] = end;
}

__global__ void __launch_bounds__(kSsdvWarps * 32) ssdv_check_kernel(unsigned char* windows, int n, int* verdict, int* errors)
{
    __shared__ SsdvTables tab;
    __shared__ __align__(16) SsdvScratch scratch[kSsdvWarps];
    build_tables(tab);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int idx = blockIdx.x * kSsdvWarps + warp;
    if (idx >= n) return;
    SsdvScratch& w = scratch[warp];
    unsigned char* src = windows + size_t(idx) * kSsdvPkt;
    for (int i = lane; i < int(kSsdvPkt); i += 32) w.pkt[i] = src[i];
    __syncwarp();
    int err = 0;
    const bool ok = is_packet_warp(tab, w, err, lane);
    __syncwarp();
    if (ok) for (int i = lane; i < int(kSsdvPkt); i += 32) src[i] = w.pkt[i];
    if (lane == 0) { verdict[idx] = ok ? 0 : -1; errors[idx] = err; }
}

cudaError_t launch_ssdv_scan(const SsdvScanArgs& a, int n_channels, cudaStream_t stream, int* launches)
{
    ssdv_scan_kernel<<<(n_channels + kSsdvWarps - 1) / kSsdvWarps, kSsdvWarps * 32, 0, stream>>>(a, n_channels);
    if (launches) ++*launches;
    return cudaGetLastError();
}

cudaError_t launch_ssdv_check(unsigned char* windows, int n, int* verdict, int* errors, cudaStream_t stream, int* launches)
{
    if (n <= 0) return cudaSuccess;
    ssdv_check_kernel<<<(n + kSsdvWarps - 1) / kSsdvWarps, kSsdvWarps * 32, 0, stream>>>(windows, n, verdict, errors);
    if (launches) ++*launches;
    return cudaGetLastError();
}

} // namespace hbd
