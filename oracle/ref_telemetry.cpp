// TEST INFRASTRUCTURE ONLY: the reference's own telemetry layer -- habdec::parse_sentence / parse_sentence_time /
// parse_gps_pos / timestamp_from_HMS (/root/reference/code/common/sentence_parse.cpp:50-199) and
// habdec::CalcGpsDistance (code/common/GpsDistance.cpp:21-84) -- compiled from the sources where they lie and driven
// line by line from stdin.  Built into oracle/_ref/telemetry_ref (oracle/Makefile).
//
// timestamp_from_HMS reads std::chrono::system_clock::now(); to pin its midnight window the executable defines
// clock_gettime itself (an executable's symbols come first in the dynamic lookup, so libstdc++'s call lands here):
// a "NOW <unix seconds>" line freezes CLOCK_REALTIME, "NOW -1" hands it back to the kernel.
//
// The two message payloads the websocket server formats from these values (websocketServer/main.cpp:324-331
// "tracking_telemetry", habdec_ws_protocol.cpp:486-498 "stats") and the STATS update of SentenceCallback
// (main.cpp:339-366, GLOBALS.h:66-73) live in files that need boost/cpr and cannot be compiled here; this harness
// applies the same ostream insertions to the reference's own values (default-formatted `<<` of int/float/double).
//
// Protocol (tab separated, one answer line per request line):
//   NOW\t<t>                         -> OK
//   TIME\t<str>                      -> NONE | <h>\t<m>\t<sec %a>        | THROW
//   POS\t<str>                       -> <float %a>                        | THROW
//   STAMP\t<h>\t<m>\t<sec>           -> <timestamp>
//   SENT\t<sentence without crc>     -> NONE | <callsign>\t<datetime>\t<frame>\t<lat %a>\t<lon %a>\t<alt %a>\t<tracking payload> | THROW
//   DIST\t<6 doubles>                -> <5 doubles %a>
//   STATION\t<lat>\t<lon>\t<alt>     -> OK          (floats, GLOBALS::PARAMS station_*_)
//   CB\t<callsign>\t<data>\t<crc>    -> the STATS after SentenceCallback: <num_ok>\t<5 x D_ %a>\t<dist_max %a>\t<elev_min %a>\t<stats payload without age> | NONE | THROW
#include <time.h>
#include <dlfcn.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <cmath>
#include <cstdio>
#include <iostream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#include "common/sentence_parse.h"
#include "common/GpsDistance.h"

static long long g_fake_now = -1;

extern "C" int clock_gettime(clockid_t id, struct timespec* ts)
{
    if (id == CLOCK_REALTIME && g_fake_now >= 0) { ts->tv_sec = time_t(g_fake_now); ts->tv_nsec = 250000000; return 0; }
    return int(syscall(SYS_clock_gettime, id, ts));
}

namespace {
std::vector<std::string> fields(const std::string& line)
{
    std::vector<std::string> f; size_t a = 0;
    for (;;) { size_t b = line.find('\t', a); if (b == std::string::npos) { f.push_back(line.substr(a)); break; } f.push_back(line.substr(a, b - a)); a = b + 1; }
    return f;
}
struct Stats {                       // GLOBALS::STATS, GLOBALS.h:66-73
    unsigned num_ok_ = 0; habdec::GpsDistance D_; double dist_max_ = 0; double elev_min_ = 90.0;
};
}

int main()
{
    std::string line;
    float st_lat = 0, st_lon = 0, st_alt = 0;
    Stats stats; std::map<int, std::string> sentences_map;
    while (std::getline(std::cin, line)) {
        auto f = fields(line);
        try {
            if (f[0] == "NOW" && f.size() == 2) { g_fake_now = atoll(f[1].c_str()); printf("OK\n"); }
            else if (f[0] == "TIME" && f.size() == 2) {
                auto r = habdec::parse_sentence_time(f[1]);
                if (!r) printf("NONE\n"); else printf("%d\t%d\t%a\n", std::get<0>(*r), std::get<1>(*r), double(std::get<2>(*r)));
            }
            else if (f[0] == "POS" && f.size() == 2) printf("%a\n", double(habdec::parse_gps_pos(f[1])));
            else if (f[0] == "STAMP" && f.size() == 4) printf("%s\n", habdec::timestamp_from_HMS(atoi(f[1].c_str()), atoi(f[2].c_str()), strtof(f[3].c_str(), nullptr)).c_str());
            else if (f[0] == "SENT" && f.size() == 2) {
                auto r = habdec::parse_sentence(f[1]);
                if (!r) printf("NONE\n");
                else {
                    std::stringstream s;             // websocketServer/main.cpp:326-331
                    s << r->payload_callsign << "," << r->datetime << "," << r->lat << "," << r->lon << "," << r->alt;
                    printf("%s\t%s\t%d\t%a\t%a\t%a\t%s\n", r->payload_callsign.c_str(), r->datetime.c_str(), r->frame, double(r->lat), double(r->lon), double(r->alt), s.str().c_str());
                }
            }
            else if (f[0] == "DIST" && f.size() == 7) {
                double v[6]; for (int i = 0; i < 6; ++i) v[i] = strtod(f[size_t(i) + 1].c_str(), nullptr);
                auto d = habdec::CalcGpsDistance(v[0], v[1], v[2], v[3], v[4], v[5]);
                printf("%a\t%a\t%a\t%a\t%a\n", d.dist_line_, d.dist_circle_, d.dist_radians_, d.elevation_, d.bearing_);
            }
            else if (f[0] == "STATION" && f.size() == 4) {
                st_lat = strtof(f[1].c_str(), nullptr); st_lon = strtof(f[2].c_str(), nullptr); st_alt = strtof(f[3].c_str(), nullptr);
                stats = Stats(); sentences_map.clear(); printf("OK\n");
            }
            else if (f[0] == "CB" && f.size() == 4) {
                // SentenceCallback, websocketServer/main.cpp:292-366 without the network / file side effects
                const std::string no_crc = f[1] + "," + f[2];
                auto r = habdec::parse_sentence(no_crc);
                if (!r) { printf("NONE\n"); continue; }
                sentences_map[r->frame] = f[1] + "," + f[2] + "*" + f[3];
                stats.num_ok_ = unsigned(sentences_map.size());
                if (st_lat) {
                    stats.D_ = habdec::CalcGpsDistance(st_lat, st_lon, st_alt, r->lat, r->lon, r->alt);
                    stats.dist_max_ = std::max(stats.dist_max_, stats.D_.dist_line_);
                    stats.elev_min_ = std::min(stats.elev_min_, stats.D_.elevation_);
                }
                std::stringstream s;                 // habdec_ws_protocol.cpp:486-494
                s << "cmd::info:stats=" << "ok:" << stats.num_ok_ << ",dist_line:" << stats.D_.dist_line_ << ",dist_circ:" << stats.D_.dist_circle_
                  << ",max_dist:" << stats.dist_max_ << ",min_elev:" << stats.elev_min_ << ",lat:" << st_lat << ",lon:" << st_lon << ",alt:" << st_alt;
                printf("%u\t%a\t%a\t%a\t%a\t%a\t%a\t%a\t%s\n", stats.num_ok_, stats.D_.dist_line_, stats.D_.dist_circle_, stats.D_.dist_radians_,
                       stats.D_.elevation_, stats.D_.bearing_, stats.dist_max_, stats.elev_min_, s.str().c_str());
            }
            else printf("BAD\n");
        } catch (const std::exception&) { printf("THROW\n"); }
        fflush(stdout);
    }
    return 0;
}
