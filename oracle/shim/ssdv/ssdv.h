// TEST INFRASTRUCTURE ONLY (oracle build).
// Stub for the un-vendored submodule fsphil/ssdv @ 1de34b9cf0c71803fb189ed2ad3e81ebb95ed93c
// (reference .gitmodules:1-3; code/ssdv is empty in the checkout).  Declares the
// symbols code/Decoder/ssdv_wrapper.cpp uses (:46,66,92,103,123-124,164-171);
// the packet detector always answers "not a packet", so SSDV image output is
// UNPINNED while the raw character path stays the reference's own.
#pragma once
#include <stdint.h>
#include <stddef.h>
#define SSDV_PKT_SIZE 256
typedef struct {
    uint8_t type; uint32_t callsign; char callsign_s[7]; uint8_t image_id;
    uint16_t packet_id; uint16_t width; uint16_t height; uint8_t eoi; uint8_t quality;
    uint8_t mcu_mode; uint8_t mcu_offset; uint16_t mcu_id; uint16_t mcu_count;
} ssdv_packet_info_t;
typedef struct { int unused; } ssdv_t;
static inline char ssdv_dec_is_packet(uint8_t*, int*) { return -1; }
static inline void ssdv_dec_header(ssdv_packet_info_t*, uint8_t*) {}
static inline char ssdv_dec_init(ssdv_t*) { return 0; }
static inline char ssdv_dec_set_buffer(ssdv_t*, uint8_t*, size_t) { return 0; }
static inline char ssdv_dec_feed(ssdv_t*, uint8_t*) { return 0; }
static inline char ssdv_dec_get_jpeg(ssdv_t*, uint8_t**, size_t*) { return 0; }
