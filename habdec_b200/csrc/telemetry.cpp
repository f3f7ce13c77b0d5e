// Telemetry layer: what the reference's websocket server does with every CRC-valid sentence
// (SentenceCallback, code/websocketServer/main.cpp:292-366): parse it into a MinTelemetry record
// (habdec::parse_sentence, code/common/sentence_parse.cpp:148-199), compute distance / elevation / bearing from the
// station (habdec::CalcGpsDistance, code/common/GpsDistance.cpp:21-84) and keep running statistics
// (GLOBALS::STATS, websocketServer/GLOBALS.h:66-73).  It is scalar string work on a few sentences per second and
// channel, so it runs on the host next to the sentence layer (host_tail.cpp), once per sentence callback.
//
// The reference leans on std::regex, std::stoi/stof and ostream formatting; here the same results come from a small
// hand-written scanner, strtol/strtof with the std:: conversion rules spelled out, and printf("%g").  Three-way
// outcomes (value / empty optional / exception) are reported as ParseStatus.
#include "telemetry.h"

#include <cerrno>
#include <chrono>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <ctime>

namespace hbd {
namespace {

inline bool is_digit(char c) { return c >= '0' && c <= '9'; }

// std::stoi: strtol base 10; nothing converted -> invalid_argument, ERANGE or outside int -> out_of_range
bool to_int(const std::string& s, int& out)
{
    const char* p = s.c_str(); char* end = nullptr;
    errno = 0;
    const long v = strtol(p, &end, 10);
    if (end == p || errno == ERANGE || v < long(INT_MIN) || v > long(INT_MAX)) return false;
    out = int(v);
    return true;
}

// std::stof: strtof; nothing converted -> invalid_argument, ERANGE (overflow and underflow) -> out_of_range
bool to_float(const std::string& s, float& out)
{
    const char* p = s.c_str(); char* end = nullptr;
    errno = 0;
    const float v = strtof(p, &end);
    if (end == p || errno == ERANGE) return false;
    out = v;
    return true;
}

// `os << x` of a float/double with default flags and precision
std::string fmt_g(double v) { char b[48]; snprintf(b, sizeof b, "%g", v); return b; }

std::string pad2(const std::string& s) { return s.size() >= 2 ? s : std::string(2 - s.size(), '0') + s; }

std::vector<std::string> split_commas(const std::string& text)
{
    std::vector<std::string> out;
    size_t a = 0;
    for (size_t i = 0; i <= text.size(); ++i)
        if (i == text.size() || text[i] == ',') { out.emplace_back(text, a, i - a); a = i + 1; }
    return out;
}

} // namespace

// Whole-string match of  dd [x] dd [x] [ dd [ . d+ ] ]  with x = any one non-digit (sentence_parse.cpp:53-56).
// The optional separators cannot backtrack into anything else (the next item always starts with a digit), so one
// left-to-right pass decides the match.
ParseStatus parse_sentence_time(const std::string& s, int& hour, int& minute, float& second)
{
    const size_t n = s.size();
    size_t i = 0;
    auto two_digits = [&](size_t at) { return at + 1 < n && is_digit(s[at]) && is_digit(s[at + 1]); };
    if (!two_digits(i)) return PARSE_NONE;
    const size_t h0 = i; i += 2;
    if (i < n && !is_digit(s[i])) ++i;
    if (!two_digits(i)) return PARSE_NONE;
    const size_t m0 = i; i += 2;
    if (i < n && !is_digit(s[i])) ++i;
    size_t s0 = i, s1 = i;
    if (two_digits(i)) {
        i += 2;
        if (i + 1 < n && s[i] == '.' && is_digit(s[i + 1])) { i += 2; while (i < n && is_digit(s[i])) ++i; }
        s1 = i;
    }
    if (i != n) return PARSE_NONE;
    hour = (s[h0] - '0') * 10 + (s[h0 + 1] - '0');
    minute = (s[m0] - '0') * 10 + (s[m0 + 1] - '0');
    second = 0;
    if (s1 > s0 && !to_float(s.substr(s0, s1 - s0), second)) return PARSE_THROW;
    return PARSE_OK;
}

// dd.dddd / ddd.dddd are decimal degrees, ddmm.mmmm / dddmm.mmmm are NMEA degrees+minutes; anything else is 0
// (sentence_parse.cpp:103-145).  The position of the '.' is counted after a leading '-'; the decimal form converts
// the whole string (sign included), the NMEA forms convert the unsigned part and apply the sign afterwards.
ParseStatus parse_gps_pos(const std::string& s, float& out)
{
    if (s.empty()) return PARSE_THROW;                       // string::at(0)
    const bool neg = s[0] == '-';
    const std::string body = neg ? s.substr(1) : s;
    const size_t dot = body.find('.');
    out = 0;
    if (dot == 2 || dot == 3) return to_float(s, out) ? PARSE_OK : PARSE_THROW;
    if (dot == 4 || dot == 5) {
        float v;
        if (!to_float(body, v)) return PARSE_THROW;
        const float degs = std::trunc(v / 100);
        const float mins = v - 100.0f * degs;
        v = degs + mins / 60.0f;
        out = (neg ? -1.0f : 1.0f) * v;
        return PARSE_OK;
    }
    return PARSE_OK;
}

// Today's UTC date + the sentence's H:M:S; a 23h sentence seen just after midnight belongs to yesterday, a 0h
// sentence seen just before midnight to tomorrow (sentence_parse.cpp:72-98).  Hours and minutes are zero padded to
// two characters, seconds are the default float format padded to two characters ("07", "59.25", but "5.5").
std::string timestamp_from_hms(int hour, int minute, float second, long long now_unix)
{
    if (now_unix < 0)
        now_unix = (long long)std::chrono::duration_cast<std::chrono::seconds>(std::chrono::system_clock::now().time_since_epoch()).count();
    long long day = now_unix / 86400;
    if (now_unix % 86400 < 0) --day;
    const int sys_hour = int((now_unix - day * 86400) / 3600);
    if (hour == 23 && sys_hour == 0) --day;
    else if (hour == 0 && sys_hour == 23) ++day;
    const time_t t = time_t(day * 86400);
    struct tm g;
    gmtime_r(&t, &g);
    char b[96];
    snprintf(b, sizeof b, "%d-%02d-%02dT%s:%s:%sZ", g.tm_year + 1900, g.tm_mon + 1, g.tm_mday, pad2(std::to_string(hour)).c_str(),
             pad2(std::to_string(minute)).c_str(), pad2(fmt_g(second)).c_str());
    return b;
}

ParseStatus parse_sentence(const std::string& sentence, long long now_unix, Telemetry& out)
{
    const std::vector<std::string> tok = split_commas(sentence);
    if (tok.size() < 6) return PARSE_NONE;
    // callsign = what follows the first run of '$'; a callsign that ENDS in that run makes the reference's
    // string::at() throw (sentence_parse.cpp:159-166)
    std::string callsign = tok[0];
    size_t d = callsign.find('$');
    if (d != std::string::npos) {
        while (d < callsign.size() && callsign[d] == '$') ++d;
        if (d == callsign.size()) return PARSE_THROW;
        callsign.erase(0, d);
    }
    int frame; float alt, lat, lon;
    if (!to_int(tok[1], frame) || !to_float(tok[5], alt)) return PARSE_THROW;
    { const ParseStatus r = parse_gps_pos(tok[3], lat); if (r != PARSE_OK) return r; }
    { const ParseStatus r = parse_gps_pos(tok[4], lon); if (r != PARSE_OK) return r; }
    if (!lat && !lon) return PARSE_NONE;                     // no GPS fix
    int hh, mm; float ss;
    { const ParseStatus r = parse_sentence_time(tok[2], hh, mm, ss); if (r != PARSE_OK) return r; }
    out.payload_callsign = callsign;
    out.datetime = timestamp_from_hms(hh, mm, ss, now_unix);
    out.frame = frame; out.lat = lat; out.lon = lon; out.alt = alt;
    return PARSE_OK;
}

// Spherical earth (r = 6371 km): bearing and central angle from the atan2 form of the great-circle formulas, then
// the plane triangle (earth centre, station, payload) gives elevation (sine rule) and slant range (cosine rule).
// Same operation order as GpsDistance.cpp:21-84 so that the doubles come out identical with the same libm.
GpsDistance calc_gps_distance(double lat1, double lon1, double alt1, double lat2, double lon2, double alt2)
{
    const double R = 6371000.0;
    const double rad = M_PI / 180.0;
    lat1 *= rad; lat2 *= rad; lon1 *= rad; lon2 *= rad;
    const double dl = lon2 - lon1;
    const double y = cos(lat2) * sin(dl);
    const double x = (cos(lat1) * sin(lat2)) - (sin(lat1) * cos(lat2) * cos(dl));
    double bearing = atan2(y, x);
    const double chord = sqrt((y * y) + (x * x));
    const double dot = (sin(lat1) * sin(lat2)) + (cos(lat1) * cos(lat2) * cos(dl));
    const double angle = atan2(chord, dot);
    const double r1 = R + alt1, r2 = R + alt2;
    const double up = (cos(angle) * r2) - r1;
    const double along = sin(angle) * r2;
    GpsDistance g;
    g.dist_circle_ = angle * R;
    g.dist_radians_ = angle;
    g.dist_line_ = sqrt((r1 * r1) + (r2 * r2) - 2 * r2 * r1 * cos(angle));
    g.elevation_ = atan2(up, along) / rad;
    if (bearing < 0) bearing += 2 * M_PI;
    g.bearing_ = bearing / rad;
    return g;
}

std::string tracking_payload(const Telemetry& t)
{
    return t.payload_callsign + "," + t.datetime + "," + fmt_g(t.lat) + "," + fmt_g(t.lon) + "," + fmt_g(t.alt);
}

ParseStatus TelemetryChannel::on_sentence(const std::string& callsign, const std::string& data, const std::string& crc,
                                          float st_lat, float st_lon, float st_alt, long long now_unix, double now_mono)
{
    Telemetry t;
    const ParseStatus r = parse_sentence(callsign + "," + data, now_unix, t);
    if (r != PARSE_OK) return r;
    last_sentence_mono = now_mono;
    sentences_map[t.frame] = callsign + "," + data + "*" + crc;      // a repeated frame id replaces, num_ok_ counts ids
    num_ok_ = unsigned(sentences_map.size());
    if (st_lat) {                                                    // main.cpp:358: only with a station latitude
        D_ = calc_gps_distance(st_lat, st_lon, st_alt, t.lat, t.lon, t.alt);
        if (D_.dist_line_ > dist_max_) dist_max_ = D_.dist_line_;
        if (D_.elevation_ < elev_min_) elev_min_ = D_.elevation_;
    }
    pending.push_back(std::move(t));
    return PARSE_OK;
}

std::string TelemetryChannel::stats_payload(float st_lat, float st_lon, float st_alt, long long age_s) const
{
    std::string s = "cmd::info:stats=ok:" + std::to_string(num_ok_) + ",dist_line:" + fmt_g(D_.dist_line_) + ",dist_circ:" + fmt_g(D_.dist_circle_)
                  + ",max_dist:" + fmt_g(dist_max_) + ",min_elev:" + fmt_g(elev_min_) + ",lat:" + fmt_g(st_lat) + ",lon:" + fmt_g(st_lon)
                  + ",alt:" + fmt_g(st_alt);
    if (age_s >= 0) s += ",age:" + std::to_string(age_s);
    return s;
}

} // namespace hbd
