// TEST INFRASTRUCTURE ONLY -- CPU restatement of habdec's SSDV packet sync and bookkeeping,
// SSDV_wraper_t::push (code/Decoder/ssdv_wrapper.cpp:37-148, state in ssdv_wrapper.h:37-72), as driven by
// Decoder::process (Decoder.h:572-573: one push per call that produced raw characters).
// PARITY: pinned against the reference's own SSDV_wraper_t compiled into oracle/_ref (tests/test_oracle.py);
// the packet test it calls (fsphil/ssdv, absent) is the published-algorithm restatement of
// oracle/ssdv_published.h in BOTH builds, i.e. unpinned below this file.  JPEG decode is not restated.
#pragma once
#include <cstdint>
#include <cstring>
#include <map>
#include <string>
#include <utility>
#include <vector>

#include "oracle_abi.h"
#include "ssdv_published.h"

namespace hbo_ssdv {

struct Packet { ssdv_packet_info_t header; uint8_t data[256]; };
using ImageKey = std::pair<std::string, uint16_t>;

inline hbo_ssdv_event make_event(uint32_t call, const ssdv_packet_info_t& hd, const std::vector<const uint8_t*>& set_in_order)
{
    hbo_ssdv_event e;
    memset(&e, 0, sizeof(e));
    e.call = call;
    e.image_id = hd.image_id; e.packet_id = hd.packet_id; e.width = hd.width; e.height = hd.height;
    e.set_size = uint16_t(set_in_order.size());
    std::vector<uint8_t> cat;
    for (const uint8_t* p : set_in_order) cat.insert(cat.end(), p, p + 256);
    e.set_crc32 = ssdvp_crc32(cat.data(), cat.size());
    strncpy(e.callsign, hd.callsign_s, sizeof(e.callsign) - 1);
    return e;
}

struct Port {
    std::vector<uint8_t> buff;          // ssdv_wrapper.h:38
    long packet_begin = -1;             // :39 (int there; sizes here never reach 2^31)
    std::map<ImageKey, std::map<uint16_t, Packet>> packets;   // :62, a set ordered by packet id (:53-57)
    ImageKey last_key{"", 0};           // :82
    std::vector<hbo_ssdv_event> events;

    // the scan `while (++packet_begin_ < buff_.size() && buff_[packet_begin_] != 0x55)` (:54, :70)
    void scan_sync() { while (++packet_begin < long(buff.size()) && buff[size_t(packet_begin)] != 0x55) {} }

    bool push(const uint8_t* chars, size_t n, uint32_t call)
    {
        buff.insert(buff.end(), chars, chars + n);                       // :42-45
        if (buff.size() < SSDV_PKT_SIZE) return false;                  // :47-48
        if (packet_begin == -1) {                                        // :51-61
            scan_sync();
            if (packet_begin == long(buff.size())) { buff.clear(); packet_begin = -1; return false; }
        }
        if (buff.size() - size_t(packet_begin) < SSDV_PKT_SIZE) return false;   // :63-64
        int errors = 0;
        if (ssdv_dec_is_packet(buff.data() + packet_begin, &errors) != 0) {     // :66-85
            scan_sync();
            if (packet_begin == long(buff.size())) { buff.clear(); packet_begin = -1; }
            else { buff.erase(buff.begin(), buff.begin() + packet_begin); packet_begin = 0; }
            return false;
        }
        Packet p;                                                        // :89-93
        memcpy(p.data, buff.data() + packet_begin, 256);
        buff.erase(buff.begin() + packet_begin, buff.begin() + packet_begin + SSDV_PKT_SIZE);
        packet_begin = -1;
        ssdv_dec_header(&p.header, p.data);

        const ImageKey key(p.header.callsign_s, p.header.image_id);     // :105-141
        auto it = packets.find(key);
        if (it == packets.end() || it->second.empty()) {
            packets[key].clear();
            packets[key][p.header.packet_id] = p;
        } else {
            auto& set = it->second;
            const bool exists = set.count(p.header.packet_id) != 0;
            const Packet& last = set.rbegin()->second;                  // highest packet id so far (:125-126)
            if (exists || last.header.height != p.header.height || last.header.width != p.header.width) set.clear();
            set[p.header.packet_id] = p;
        }
        last_key = key;                                                 // make_jpeg :171
        std::vector<const uint8_t*> order;
        for (auto& kv : packets[key]) order.push_back(kv.second.data);
        events.push_back(make_event(call, p.header, order));
        return true;
    }

    size_t image(const std::string& callsign, int image_id, uint8_t* out, size_t cap) const
    {
        auto it = packets.find(ImageKey(callsign, uint16_t(image_id)));
        if (it == packets.end()) return 0;
        size_t n = 0;
        for (auto& kv : it->second) { if (out && n + 256 <= cap) memcpy(out + n, kv.second.data, 256); n += 256; }
        return n;
    }
};

} // namespace hbo_ssdv
