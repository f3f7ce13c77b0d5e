// Telemetry layer behind the sentence callback, host C++: a CRC-valid sentence becomes a
// MinTelemetry record, a distance/elevation from the station and per-channel running statistics.  See telemetry.cpp.
#pragma once
#include <cstdint>
#include <map>
#include <string>
#include <vector>

namespace hbd {

// outcome of a parse step: the reference either returns a value, returns an empty optional, or lets a
// std::invalid_argument / std::out_of_range escape (stoi / stof / string::at); callers need to tell the three apart
enum ParseStatus { PARSE_THROW = -1, PARSE_NONE = 0, PARSE_OK = 1 };

struct Telemetry {                     // sondehub::MinTelemetry, code/sondehub/sondehub_uploader.h:12-21
    std::string payload_callsign, datetime;
    int frame = 0;
    float lat = 0, lon = 0, alt = 0;
};

struct GpsDistance {                   // habdec::GpsDistance, code/common/GpsDistance.h:7-14
    double dist_line_ = 0, dist_circle_ = 0, dist_radians_ = 0, elevation_ = 0, bearing_ = 0;
};

ParseStatus parse_sentence_time(const std::string& s, int& hour, int& minute, float& second);
ParseStatus parse_gps_pos(const std::string& s, float& out);
// now_unix < 0: read the system clock
std::string timestamp_from_hms(int hour, int minute, float second, long long now_unix);
ParseStatus parse_sentence(const std::string& sentence_without_crc, long long now_unix, Telemetry& out);
GpsDistance calc_gps_distance(double lat1, double lon1, double alt1, double lat2, double lon2, double alt2);
// payload of the "cmd::info:tracking_telemetry=" message, websocketServer/main.cpp:326-331
std::string tracking_payload(const Telemetry& t);

// GLOBALS::STATS + sentences_map_ of one receiver (GLOBALS.h:66-73,63), kept per channel here
struct TelemetryChannel {
    std::map<int, std::string> sentences_map;    // frame id -> "callsign,data*crc"
    unsigned num_ok_ = 0;
    GpsDistance D_;
    double dist_max_ = 0, elev_min_ = 90.0;
    double last_sentence_mono = -1;              // steady-clock seconds of the last parsed sentence (< 0: since creation)
    std::vector<Telemetry> pending;              // records since the last poll

    // SentenceCallback (websocketServer/main.cpp:292-366) minus network and file side effects
    ParseStatus on_sentence(const std::string& callsign, const std::string& data, const std::string& crc,
                            float st_lat, float st_lon, float st_alt, long long now_unix, double now_mono);
    // payload of "cmd::info:stats=..." (habdec_ws_protocol.cpp:486-498); age < 0 leaves the ",age:" field out
    std::string stats_payload(float st_lat, float st_lon, float st_alt, long long age_s) const;
};

} // namespace hbd
