// TEST INFRASTRUCTURE ONLY (oracle build).
// Stand-in for the un-vendored submodule fsphil/ssdv @ 1de34b9cf0c71803fb189ed2ad3e81ebb95ed93c
// (reference .gitmodules:1-3; code/ssdv is empty in the checkout).  Declares the
// symbols code/Decoder/ssdv_wrapper.cpp uses (:46,66,92,103,123-124,164-171).
//  * packet detection and header decode (ssdv_dec_is_packet, ssdv_dec_header) are restated from the
//    library's published algorithm in oracle/ssdv_published.h, so that the reference's own
//    SSDV_wraper_t::push (sync scan + packet bookkeeping) runs for real in oracle/_ref;
//  * the JPEG decoder (ssdv_dec_feed / ssdv_dec_get_jpeg) stays a stub: image output is UNPINNED and
//    out of scope, the raw character path stays the reference's own.
#pragma once
#include "../../ssdv_published.h"
typedef struct { int unused; } ssdv_t;
static inline char ssdv_dec_init(ssdv_t*) { return 0; }
static inline char ssdv_dec_set_buffer(ssdv_t*, uint8_t*, size_t) { return 0; }
static inline char ssdv_dec_feed(ssdv_t*, uint8_t*) { return 0; }
static inline char ssdv_dec_get_jpeg(ssdv_t*, uint8_t**, size_t*) { return 0; }
