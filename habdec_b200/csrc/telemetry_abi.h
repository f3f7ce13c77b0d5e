// internal: lets api.cu feed an attached tracker from inside the sentence layer (telemetry_abi.cpp)
#pragma once
#include <string>
struct hbd_tracker;
namespace hbd {
// SentenceCallback for channel `ch`; returns 1 (record filed), 0 (rejected), -1 (the reference would have thrown)
int tracker_feed(hbd_tracker* t, int ch, const std::string& callsign, const std::string& data, const std::string& crc);
}
