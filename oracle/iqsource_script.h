// TEST INFRASTRUCTURE ONLY: one scripted session against an IQ file source, templated on the source type so that
// the reference's habdec::IQSource_File<float> (oracle/ref_iqsource.cpp, built into oracle/_ref/) and
// habdec_b200::IQSourceFile (tests/cpp/iqsource_ours.cpp) print the same transcript.
#pragma once
#include <complex>
#include <cstdint>
#include <cstdio>
#include <iostream>
#include <string>
#include <vector>

template <class Src>
int iqsource_script(const char* path)
{
    using std::cout; using std::endl;
    Src s;
    std::string file = path, got_file;
    double sr = 48000.0, got_sr = 0;
    bool f = false, t = true, got_b = true;
    int bogus = 0;
    cout << "set file_string " << s.setOption("file_string", &file) << endl;
    cout << "set sampling_rate_double " << s.setOption("sampling_rate_double", &sr) << endl;
    cout << "set realtime_bool " << s.setOption("realtime_bool", &f) << endl;
    cout << "set loop_bool " << s.setOption("loop_bool", &f) << endl;
    cout << "set bogus " << s.setOption("bogus_int", &bogus) << endl;
    cout << "get file_string " << s.getOption("file_string", &got_file) << " match " << (got_file == file) << endl;
    cout << "get sampling_rate_double " << s.getOption("sampling_rate_double", &got_sr) << " " << got_sr << endl;
    cout << "get realtime_bool " << s.getOption("realtime_bool", &got_b) << " " << got_b << endl;
    cout << "get loop_bool " << s.getOption("loop_bool", &got_b) << " " << got_b << endl;
    cout << "type " << s.type() << " rate " << s.samplingRate() << endl;
    std::vector<std::complex<float>> buf(1000);
    auto sum = [&](size_t n) { uint64_t h = 1469598103934665603ull; const unsigned char* p = reinterpret_cast<const unsigned char*>(buf.data());
                               for (size_t i = 0; i < n * sizeof(std::complex<float>); ++i) { h ^= p[i]; h *= 1099511628211ull; } return h; };
    cout << "running before init " << s.isRunning() << endl;
    cout << "get before init " << s.get(buf.data(), buf.size()) << endl;
    cout << "start before init " << s.start() << endl;
    cout << "init " << s.init() << " count " << s.count() << endl;
    cout << "running after init " << s.isRunning() << endl;
    cout << "get before start " << s.get(buf.data(), buf.size()) << endl;
    cout << "start " << s.start() << " running " << s.isRunning() << endl;
    for (int i = 0; i < 5; ++i) { const size_t n = s.get(buf.data(), buf.size()); cout << "get#" << i << " " << n << " fnv " << sum(n) << endl; }
    cout << "set loop_bool " << s.setOption("loop_bool", &t) << endl;
    for (int i = 5; i < 10; ++i) { const size_t n = s.get(buf.data(), buf.size()); cout << "get#" << i << " " << n << " fnv " << sum(n) << endl; }
    std::vector<std::complex<float>> big(4000);
    { const size_t n = s.get(big.data(), big.size()); cout << "get big " << n << endl; }
    { const size_t n = s.get(big.data(), big.size()); cout << "get big " << n << endl; }
    cout << "stop " << s.stop() << " running " << s.isRunning() << endl;
    cout << "get after stop " << s.get(buf.data(), buf.size()) << endl;
    return 0;
}
