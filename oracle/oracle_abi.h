/* TEST INFRASTRUCTURE ONLY -- the C ABI shared by the two CPU oracles.
 *
 *   prefix ref_  : oracle/_ref/libhabdec_ref.so  -- the UNMODIFIED reference
 *                  habdec::Decoder<float> compiled from /root/reference/code
 *                  (oracle/ref_harness.cpp + oracle/Makefile)
 *   prefix orc_  : oracle/libhabdec_oracle.so    -- our CPU restatement
 *                  (oracle/habdec_oracle.cpp)
 *
 * Both export the same functions so tests can drive either one.  Nothing in
 * the product (habdec_b200/, include/) may include or link this.
 */
#ifndef HBD_ORACLE_ABI_H
#define HBD_ORACLE_ABI_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* per-call stage arrays that can be recorded (appended call after call) */
enum hbo_stage {
    HBO_STAGE_DECIMATED = 0, /* cf32, output of the decimator chain (+DC removal), Decoder.h:440-459 */
    HBO_STAGE_FILTERED  = 1, /* cf32, output of the low-pass FIR, Decoder.h:532-542 */
    HBO_STAGE_DEMOD     = 2, /* f32, discriminator output, Decoder.h:546-555 */
    HBO_STAGE_FFT       = 3, /* cf32, most recent fft-shifted spectrum (not appended), Decoder.h:479-489 */
    HBO_STAGE_POWER     = 4, /* f32, most recent AFC power spectrum in dB (not appended), AFC.h:236-286 */
    HBO_STAGE_LPTAPS    = 5, /* f32, current low-pass taps, FirFilter.h:173-209 */
    HBO_STAGE_PENDING   = 6, /* f32, slicer samples still pending after the call, SymbolExtractor.h:156-157 */
    HBO_STAGE_BITS      = 7, /* f32 (0/1), every bit the slicer emitted so far (appended) */
    HBO_STAGE_RAWCHARS  = 8  /* f32 (byte values), every raw UART char so far (port only; ref returns 0) */
};

typedef struct hbo_config {
    double baud;        /* Decoder::baud            Decoder.h:656 */
    int    rtty_bits;   /* Decoder::rtty_bits       Decoder.h:671 */
    float  rtty_stops;  /* Decoder::rtty_stops      Decoder.h:686 */
    float  lowpass_bw;  /* Decoder::lowpass_bw      Decoder.h:238 */
    float  lowpass_trans;/* Decoder::lowpass_trans  Decoder.h:252 */
    int    dec_factor;  /* setupDecimationStagesFactor Decoder.h:268 */
    int    dc_remove;   /* Decoder::dc_remove       Decoder.h:701 */
    int    record;      /* 1: append stage arrays on every call */
    int    fft_bins;    /* 0 / 4096: the reference's fft_bins_cnt_ (Decoder.h:163); 16384: the 16k-bin extension of
                           BASELINE configs[1] -- the reference object is run with that member patched at run time */
} hbo_config;

/* AFC / spectrum scalars after the most recent call */
typedef struct hbo_afc_info {
    double frequency_correction;
    double shift_hz;
    double noise_floor;
    double noise_variance;
    int    peak_left;
    int    peak_right;
} hbo_afc_info;

#define HBO_DECL(P) \
    void*  P##_create(const hbo_config* cfg); \
    void   P##_destroy(void* h); \
    /* pushSamples(iq[n] interleaved cf32, fs) followed by operator()() */ \
    void   P##_push_process(void* h, const float* iq, size_t n_complex, double fs); \
    /* concat of everything character_callback_ would have delivered */ \
    size_t P##_chars(void* h, char* out, size_t cap); \
    /* getRTTY() */ \
    size_t P##_rtty(void* h, char* out, size_t cap); \
    /* getLastSentence() */ \
    size_t P##_last_sentence(void* h, char* out, size_t cap); \
    /* '\n'-joined "callsign,data*crc" of every sentence_callback_ (CRC-valid) */ \
    size_t P##_sentences(void* h, char* out, size_t cap); \
    /* returns the number of floats available; copies min(cap, available) */ \
    size_t P##_stage(void* h, int stage, float* out, size_t cap_floats); \
    void   P##_afc(void* h, hbo_afc_info* out); \
    void   P##_reset_frequency_correction(void* h, double corr); \
    /* CPU baseline: n_threads decoders, one per OS thread, each decoding \
       iq + t*stride_complex .. (+n_complex) in `chunk`-sized pushes, `reps` \
       passes; returns wall seconds of the slowest thread, total printable \
       chars decoded in *o_chars */ \
    double P##_bench(const hbo_config* cfg, int n_threads, const float* iq, size_t n_complex, \
                     size_t stride_complex, size_t chunk, double fs, int reps, uint64_t* o_chars);

HBO_DECL(ref)
HBO_DECL(orc)

/* ---- websocket wire formats (the step right after the path) -------------------------
 * PWR_ payload of "cmd::power:res=R,zoom=Z" and DEM_ payload of "cmd::demod:res=R":
 *   code/websocketServer/habdec_ws_protocol.cpp:338-429, NetTransport.h:29-102, CompressedVector.cpp:72-116
 * ref_*: header + quantisation by the reference's own SerializeSpectrum / SerializeDemodulation /
 *        CompressedVector (compiled from /root/reference), zoom / peak shift / ShrinkVector restated from
 *        habdec_ws_protocol.cpp (that file needs boost and cannot be compiled here);
 * orc_*: everything restated. */
typedef struct hbo_spectrum_meta {
    double noise_floor, noise_variance, sampling_rate, shift;
    int    peak_left, peak_right;   /* signed GUI peaks as returned by Decoder::getPeaks (negative = not stable) */
} hbo_spectrum_meta;
#define HBO_WIRE_DECL(P) \
    size_t P##_spectrum_frame(const float* power, size_t n, const hbo_spectrum_meta* meta, float zoom, int resolution, \
                              int type_size, unsigned char* out, size_t cap); \
    size_t P##_demod_frame(const float* demod, size_t n, int resolution, int type_size, unsigned char* out, size_t cap);
HBO_WIRE_DECL(ref)
HBO_WIRE_DECL(orc)

/* port only: the sentence layer alone (std::regex, like sentence_extract.cpp:58-98) and the CRC */
int  orc_extract_sentence(const char* stream, size_t n, char* callsign, char* data, char* crc, size_t cap, size_t* rest_offset);
void orc_crc16(const char* s, size_t n, char out[5]);

#ifdef __cplusplus
}
#endif
#endif
