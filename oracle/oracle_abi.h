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

/* ---- SSDV packet sync + bookkeeping: SSDV_wraper_t::push, ssdv_wrapper.cpp:37-148 --------
 * One record per packet the wrapper accepted (== one ssdv_callback_, Decoder.h:631-632), in order. */
typedef struct hbo_ssdv_event {
    uint32_t call;        /* index of the push_process / ssdv_push call whose push() returned true */
    uint16_t image_id, packet_id, width, height;
    uint16_t set_size;    /* packets filed under (callsign, image_id) once this one is in (ssdv_wrapper.cpp:105-141) */
    uint16_t reserved;
    uint32_t set_crc32;   /* CRC-32 of those packets' 256 bytes each, concatenated in packet-id order */
    char     callsign[8];
} hbo_ssdv_event;

#define HBO_DECL(P) \
    /* events so far: returns the count, copies min(cap, count) */ \
    size_t P##_ssdv_events(void* h, hbo_ssdv_event* out, size_t cap); \
    /* hand raw characters straight to the decoder's SSDV wrapper (what Decoder.h:572-573 does), one push() */ \
    void   P##_ssdv_push(void* h, const uint8_t* chars, size_t n); \
    /* the packets filed under (callsign, image_id), 256 bytes each in packet-id order; returns the byte count */ \
    size_t P##_ssdv_image(void* h, const char* callsign, int image_id, uint8_t* out, size_t cap); \
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
    /* run-time setter between two calls: which = 0 baud(double), 1 rtty_bits(size_t), 2 rtty_stops(float), 3 dc_remove(bool), \
       4 lowpass_bw(float), 5 lowpass_trans(float) (Decoder.h:238-257, 654-706), 6 setupDecimationStagesBW(double) (:336-412) */ \
    void   P##_set_param(void* h, int which, double value); \
    /* CPU baseline: n_threads decoders, one per OS thread, each decoding \
       iq + t*stride_complex .. (+n_complex) in `chunk`-sized pushes, `reps` \
       passes; returns wall seconds of the slowest thread, total printable \
       chars decoded in *o_chars */ \
    double P##_bench(const hbo_config* cfg, int n_threads, const float* iq, size_t n_complex, \
                     size_t stride_complex, size_t chunk, double fs, int reps, uint64_t* o_chars); \
    /* parity checker for whole batches: n_channels decoders (each on its OWN fresh OS thread, n_threads at a time), \
       channel c decodes the periodic stream iq + c*stride_complex (period ring_n samples) in `chunk`-sized pushes, \
       chunks first_chunk .. first_chunk+n_chunks-1 of the stream; chars_out + c*pitch receives the printable \
       characters (count in chars_len[c], truncated to pitch), sent_out likewise the '\n'-joined CRC-valid sentences */ \
    void   P##_run_ring(const hbo_config* cfg, int n_threads, const float* iq, size_t n_channels, size_t stride_complex, \
                        size_t ring_n, size_t chunk, size_t first_chunk, size_t n_chunks, double fs, \
                        char* chars_out, size_t chars_pitch, uint32_t* chars_len, char* sent_out, size_t sent_pitch, uint32_t* sent_len);

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

/* the published-algorithm restatement of fsphil/ssdv's packet test (oracle/ssdv_published.h), exported by the port
 * library: verdict 0 = packet (corrected in place), -1 = not a packet; *errors = corrected symbols */
int  hbo_ssdv_is_packet(uint8_t pkt[256], int* errors);
/* test-vector builder: a well-formed packet (type 0 = with FEC, 1 = no FEC); payload bytes are taken from `payload`
 * (205 / 237 bytes used) */
void hbo_ssdv_make_packet(uint8_t out[256], int type, const char* callsign, int image_id, int packet_id, int width16,
                          int height16, int flags, int mcu_offset, int mcu_id, const uint8_t* payload);

/* port only: the sentence layer alone (std::regex, like sentence_extract.cpp:58-98) and the CRC */
int  orc_extract_sentence(const char* stream, size_t n, char* callsign, char* data, char* crc, size_t cap, size_t* rest_offset);
void orc_crc16(const char* s, size_t n, char out[5]);

#ifdef __cplusplus
}
#endif
#endif
