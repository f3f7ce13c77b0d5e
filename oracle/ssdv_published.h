/* TEST INFRASTRUCTURE ONLY -- CPU restatement of the two entry points of the un-vendored
 * third-party dependency fsphil/ssdv that habdec's SSDV packet sync calls:
 *
 *     ssdv_dec_is_packet(uint8_t* packet, int* errors)   call site code/Decoder/ssdv_wrapper.cpp:66
 *     ssdv_dec_header(ssdv_packet_info_t*, uint8_t*)     call site code/Decoder/ssdv_wrapper.cpp:92
 *
 * PARITY UNPINNED: the submodule is pinned at fsphil/ssdv @ 1de34b9cf0c71803fb189ed2ad3e81ebb95ed93c
 * (reference .gitmodules, .SUBMODULES.json: "the pinned commit is not served") and is absent from
 * /root/reference, there is no network, and the reference holds no SSDV test vector.  What follows restates the
 * PUBLISHED algorithm of that library (ssdv.c: packet layout, CRC-32, sanity checks; rs8.c: Phil Karn's
 * CCSDS RS(255,223) codec, field polynomial 0x187, first consecutive root 112, primitive element 11, 32 roots):
 *
 *   packet (256 bytes): [0] sync 0x55, [1] type 0x66 (+FEC) / 0x67 (no FEC), [2..5] base-40 callsign (big endian),
 *   [6] image id, [7..8] packet id, [9] width/16, [10] height/16, [11] flags 00qqqeMM, [12] mcu offset,
 *   [13..14] mcu index, payload (205 / 237 bytes), CRC-32 (big endian) over bytes 1 .. end of payload,
 *   and for type 0x66 32 Reed-Solomon parity bytes over bytes 1..223.
 *
 * The JPEG decoder half of the library (ssdv_dec_feed / ssdv_dec_get_jpeg) is NOT restated: image output stays
 * out of scope ("only the scan/bookkeeping is gradable").
 */
#ifndef HBD_ORACLE_SSDV_PUBLISHED_H
#define HBD_ORACLE_SSDV_PUBLISHED_H

#include <stdint.h>
#include <stddef.h>
#include <string.h>

#define SSDV_PKT_SIZE          256
#define SSDV_PKT_SIZE_HEADER   15
#define SSDV_PKT_SIZE_CRC      4
#define SSDV_PKT_SIZE_RSCODES  32
#define SSDV_TYPE_INVALID      0xFF
#define SSDV_TYPE_NORMAL       0x00
#define SSDV_TYPE_NOFEC        0x01

typedef struct {
    uint8_t type; uint32_t callsign; char callsign_s[7]; uint8_t image_id;
    uint16_t packet_id; uint16_t width; uint16_t height; uint8_t eoi; uint8_t quality;
    uint8_t mcu_mode; uint8_t mcu_offset; uint16_t mcu_id; uint16_t mcu_count;
} ssdv_packet_info_t;

/* ---- GF(256), x^8+x^7+x^2+x+1 ------------------------------------------------------------------ */
enum { SSDVP_NN = 255, SSDVP_NROOTS = 32, SSDVP_FCR = 112, SSDVP_PRIM = 11, SSDVP_IPRIM = 116 };

typedef struct { uint8_t exp[256]; uint8_t log[256]; int ready; } ssdvp_gf_t;

static inline const ssdvp_gf_t* ssdvp_gf(void)
{
    static ssdvp_gf_t g;   /* racy only in the benign "written twice with the same values" way */
    if (!g.ready) {
        unsigned x = 1;
        for (int i = 0; i < SSDVP_NN; ++i) {
            g.exp[i] = (uint8_t)x; g.log[x] = (uint8_t)i;
            x <<= 1; if (x & 0x100) x ^= 0x187;
        }
        g.exp[255] = 0; g.log[0] = 255;   /* log 0 = "A0" */
        g.ready = 1;
    }
    return &g;
}
static inline int ssdvp_modnn(int x) { while (x >= SSDVP_NN) { x -= SSDVP_NN; x = (x >> 8) + (x & SSDVP_NN); } return x; }

/* decode_rs_8(data[255], no erasures, no padding): corrects in place, returns the number of corrected symbols or
 * -1 when the word is uncorrectable.  Syndromes -> Berlekamp-Massey -> Chien search -> Forney, in the formulation of
 * the published codec (index-form shift register b, root step IPRIM) so that the accept / reject verdict on
 * uncorrectable words is the same. */
static inline int ssdvp_rs_decode(uint8_t* data)
{
    const ssdvp_gf_t* g = ssdvp_gf();
    const int A0 = SSDVP_NN;
    int s[SSDVP_NROOTS], lambda[SSDVP_NROOTS + 1], b[SSDVP_NROOTS + 1], t[SSDVP_NROOTS + 1], omega[SSDVP_NROOTS + 1];
    int reg[SSDVP_NROOTS + 1], root[SSDVP_NROOTS], loc[SSDVP_NROOTS];

    for (int i = 0; i < SSDVP_NROOTS; ++i) s[i] = data[0];
    for (int j = 1; j < SSDVP_NN; ++j)
        for (int i = 0; i < SSDVP_NROOTS; ++i)
            s[i] = s[i] == 0 ? data[j] : data[j] ^ g->exp[ssdvp_modnn(g->log[s[i]] + (SSDVP_FCR + i) * SSDVP_PRIM)];
    int syn_error = 0;
    for (int i = 0; i < SSDVP_NROOTS; ++i) { syn_error |= s[i]; s[i] = g->log[s[i]]; }   /* index form */
    if (!syn_error) return 0;

    memset(lambda, 0, sizeof(lambda));
    lambda[0] = 1;
    for (int i = 0; i <= SSDVP_NROOTS; ++i) b[i] = g->log[lambda[i]];
    int r = 0, el = 0;
    while (++r <= SSDVP_NROOTS) {
        int discr = 0;
        for (int i = 0; i < r; ++i)
            if (lambda[i] != 0 && s[r - i - 1] != A0) discr ^= g->exp[ssdvp_modnn(g->log[lambda[i]] + s[r - i - 1])];
        discr = g->log[discr];
        if (discr == A0) {
            memmove(&b[1], b, SSDVP_NROOTS * sizeof(b[0])); b[0] = A0;
        } else {
            t[0] = lambda[0];
            for (int i = 0; i < SSDVP_NROOTS; ++i)
                t[i + 1] = b[i] != A0 ? lambda[i + 1] ^ g->exp[ssdvp_modnn(discr + b[i])] : lambda[i + 1];
            if (2 * el <= r - 1) {
                el = r - el;
                for (int i = 0; i <= SSDVP_NROOTS; ++i) b[i] = lambda[i] == 0 ? A0 : ssdvp_modnn(g->log[lambda[i]] - discr + SSDVP_NN);
            } else {
                memmove(&b[1], b, SSDVP_NROOTS * sizeof(b[0])); b[0] = A0;
            }
            memcpy(lambda, t, sizeof(lambda));
        }
    }
    int deg_lambda = 0;
    for (int i = 0; i <= SSDVP_NROOTS; ++i) { lambda[i] = g->log[lambda[i]]; if (lambda[i] != A0) deg_lambda = i; }

    memcpy(&reg[1], &lambda[1], SSDVP_NROOTS * sizeof(reg[0]));
    int count = 0;
    for (int i = 1, k = SSDVP_IPRIM - 1; i <= SSDVP_NN; ++i, k = ssdvp_modnn(k + SSDVP_IPRIM)) {
        int q = 1;
        for (int j = deg_lambda; j > 0; --j)
            if (reg[j] != A0) { reg[j] = ssdvp_modnn(reg[j] + j); q ^= g->exp[reg[j]]; }
        if (q != 0) continue;
        root[count] = i; loc[count] = k;
        if (++count == deg_lambda) break;
    }
    if (deg_lambda != count) return -1;

    const int deg_omega = deg_lambda - 1;
    for (int i = 0; i <= deg_omega; ++i) {
        int tmp = 0;
        for (int j = i; j >= 0; --j)
            if (s[i - j] != A0 && lambda[j] != A0) tmp ^= g->exp[ssdvp_modnn(s[i - j] + lambda[j])];
        omega[i] = g->log[tmp];
    }
    for (int j = count - 1; j >= 0; --j) {
        int num1 = 0;
        for (int i = deg_omega; i >= 0; --i)
            if (omega[i] != A0) num1 ^= g->exp[ssdvp_modnn(omega[i] + i * root[j])];
        const int num2 = g->exp[ssdvp_modnn(root[j] * (SSDVP_FCR - 1) + SSDVP_NN)];
        int den = 0;
        const int top = (deg_lambda < SSDVP_NROOTS - 1 ? deg_lambda : SSDVP_NROOTS - 1) & ~1;
        for (int i = top; i >= 0; i -= 2)
            if (lambda[i + 1] != A0) den ^= g->exp[ssdvp_modnn(lambda[i + 1] + i * root[j])];
        if (num1 != 0)
            data[loc[j]] ^= g->exp[ssdvp_modnn(g->log[num1] + g->log[num2] + SSDVP_NN - g->log[den])];
    }
    return count;
}

/* systematic encoder (used by the test-vector generator only): parity[32] over data[223] */
static inline void ssdvp_rs_encode(const uint8_t* data, uint8_t* parity)
{
    const ssdvp_gf_t* g = ssdvp_gf();
    int gen[SSDVP_NROOTS + 1];   /* generator polynomial, poly form, gen[NROOTS] = 1 */
    gen[0] = 1;
    for (int i = 0, root = SSDVP_FCR * SSDVP_PRIM; i < SSDVP_NROOTS; ++i, root += SSDVP_PRIM) {
        gen[i + 1] = 1;
        for (int j = i; j > 0; --j)
            gen[j] = gen[j] != 0 ? gen[j - 1] ^ g->exp[ssdvp_modnn(g->log[gen[j]] + root)] : gen[j - 1];
        gen[0] = g->exp[ssdvp_modnn(g->log[gen[0]] + root)];
    }
    memset(parity, 0, SSDVP_NROOTS);
    for (int i = 0; i < SSDVP_NN - SSDVP_NROOTS; ++i) {
        const int fb = data[i] ^ parity[0];
        memmove(parity, parity + 1, SSDVP_NROOTS - 1);
        parity[SSDVP_NROOTS - 1] = 0;
        if (fb != 0)
            for (int j = 0; j < SSDVP_NROOTS; ++j)
                if (gen[SSDVP_NROOTS - 1 - j] != 0) parity[j] ^= g->exp[ssdvp_modnn(g->log[fb] + g->log[gen[SSDVP_NROOTS - 1 - j]])];
    }
}

/* CRC-32 (reflected 0xEDB88320, init and final xor 0xFFFFFFFF), bit by bit like the library */
static inline uint32_t ssdvp_crc32(const uint8_t* d, size_t length)
{
    uint32_t crc = 0xFFFFFFFFu;
    for (; length; --length) {
        uint32_t x = (crc ^ *d++) & 0xFF;
        for (int i = 0; i < 8; ++i) x = (x & 1) ? (x >> 1) ^ 0xEDB88320u : x >> 1;
        crc = (crc >> 8) ^ x;
    }
    return crc ^ 0xFFFFFFFFu;
}

static inline void ssdvp_decode_callsign(char* callsign, uint32_t code)
{
    char* c = callsign;
    *c = '\0';
    if (code > 0xF423FFFFu) return;
    for (; code; ++c, code /= 40) {
        const unsigned s = code % 40;
        if (s == 0) *c = '-';
        else if (s < 11) *c = (char)('0' + s - 1);
        else if (s < 14) *c = '-';
        else *c = (char)('A' + s - 14);
    }
    *c = '\0';
}

static inline void ssdv_dec_header(ssdv_packet_info_t* p, uint8_t* packet)
{
    p->type = (uint8_t)(packet[1] - 0x66);
    p->callsign = ((uint32_t)packet[2] << 24) | ((uint32_t)packet[3] << 16) | ((uint32_t)packet[4] << 8) | packet[5];
    ssdvp_decode_callsign(p->callsign_s, p->callsign);
    p->image_id = packet[6];
    p->packet_id = (uint16_t)((packet[7] << 8) | packet[8]);
    p->width = (uint16_t)(packet[9] << 4);
    p->height = (uint16_t)(packet[10] << 4);
    p->eoi = (packet[11] >> 2) & 1;
    p->quality = ((packet[11] >> 3) & 7) ^ 4;
    p->mcu_mode = packet[11] & 0x03;
    p->mcu_offset = packet[12];
    p->mcu_id = (uint16_t)((packet[13] << 8) | packet[14]);
    p->mcu_count = (uint16_t)(packet[9] * packet[10]);
    if (p->mcu_mode & 2) p->mcu_count *= 2;
    if (p->mcu_mode & 1) p->mcu_count *= 2;
}

static inline int ssdvp_crc_ok(const uint8_t* pkt, unsigned crcdata)
{
    const uint32_t x = ssdvp_crc32(&pkt[1], crcdata);
    const unsigned i = 1 + crcdata;
    return x == ((uint32_t)pkt[i + 3] | ((uint32_t)pkt[i + 2] << 8) | ((uint32_t)pkt[i + 1] << 16) | ((uint32_t)pkt[i] << 24));
}

/* 0: a valid packet starts here (corrected in place, *errors = corrected symbols); -1: not a packet */
static inline char ssdv_dec_is_packet(uint8_t* packet, int* errors)
{
    uint8_t pkt[SSDV_PKT_SIZE];
    uint8_t type = SSDV_TYPE_INVALID;
    unsigned payload = 0;
    memcpy(pkt, packet, SSDV_PKT_SIZE);   /* testing is destructive: work on a copy */
    pkt[0] = 0x55;

    if (pkt[1] == 0x66 + SSDV_TYPE_NOFEC) {
        payload = SSDV_PKT_SIZE - SSDV_PKT_SIZE_HEADER - SSDV_PKT_SIZE_CRC;
        if (errors) *errors = 0;
        if (ssdvp_crc_ok(pkt, SSDV_PKT_SIZE_HEADER + payload - 1)) type = SSDV_TYPE_NOFEC;
    } else if (pkt[1] == 0x66 + SSDV_TYPE_NORMAL) {
        payload = SSDV_PKT_SIZE - SSDV_PKT_SIZE_HEADER - SSDV_PKT_SIZE_CRC - SSDV_PKT_SIZE_RSCODES;
        if (errors) *errors = 0;
        if (ssdvp_crc_ok(pkt, SSDV_PKT_SIZE_HEADER + payload - 1)) type = SSDV_TYPE_NORMAL;
    }
    if (type == SSDV_TYPE_INVALID) {
        /* a NORMAL packet with correctable errors? */
        payload = SSDV_PKT_SIZE - SSDV_PKT_SIZE_HEADER - SSDV_PKT_SIZE_CRC - SSDV_PKT_SIZE_RSCODES;
        pkt[1] = 0x66 + SSDV_TYPE_NORMAL;
        const int n = ssdvp_rs_decode(&pkt[1]);
        if (n < 0) return -1;
        if (errors) *errors = n;
        if (ssdvp_crc_ok(pkt, SSDV_PKT_SIZE_HEADER + payload - 1)) type = SSDV_TYPE_NORMAL;
    }
    if (type == SSDV_TYPE_INVALID) return -1;

    ssdv_packet_info_t p;
    ssdv_dec_header(&p, pkt);
    if (p.type != type) return -1;
    if (p.width == 0 || p.height == 0) return -1;
    if (p.mcu_id != 0xFFFF) {
        if (p.mcu_id >= p.mcu_count) return -1;
        if (p.mcu_offset >= payload) return -1;
    }
    memcpy(packet, pkt, SSDV_PKT_SIZE);
    return 0;
}

#endif
