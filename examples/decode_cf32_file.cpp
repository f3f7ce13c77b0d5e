// Minimal drop-in check of the C++ facade: decode a raw interleaved cf32 file the way the reference's
// DECODER_THREAD does (code/websocketServer/main.cpp:235-245: 65536-sample pushes, process after each).
//   g++ -std=c++17 -I include examples/decode_cf32_file.cpp -L habdec_b200 -lhabdec_b200 -Wl,-rpath,$PWD/habdec_b200 -o decode_cf32
//   ./decode_cf32 capture.cf32 2048000 300 8 2 256
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include "habdec_b200/Decoder.hpp"

int main(int argc, char** argv)
{
    if (argc < 3) { std::cerr << "usage: " << argv[0] << " file.cf32 sampling_rate [baud bits stops dec_factor]\n"; return 2; }
    const double fs = atof(argv[2]);
    habdec_b200::Decoder D;
    D.baud(argc > 3 ? atof(argv[3]) : 300);
    D.rtty_bits(argc > 4 ? atoi(argv[4]) : 8);
    D.rtty_stops(argc > 5 ? atof(argv[5]) : 2);
    D.lowpass_bw(1500); D.lowpass_trans(0.025f);
    D.setupDecimationStagesFactor(argc > 6 ? atoi(argv[6]) : 256);
    D.sentence_callback_ = [](std::string cs, std::string data, std::string crc) { std::cout << "\nSENTENCE " << cs << "," << data << "*" << crc << std::endl; };
    D.character_callback_ = [](std::string s) { std::cout << s << std::flush; };
    FILE* f = fopen(argv[1], "rb");
    if (!f) { perror("open"); return 1; }
    habdec_b200::IQVector v; v.samplingRate(fs);
    for (;;) {
        v.resize(256 * 256);
        const size_t n = fread(v.data(), sizeof(std::complex<float>), v.size(), f);
        if (!n) break;
        v.resize(n);
        D.pushSamples(v);
        D();
    }
    fclose(f);
    std::cout << "\nlast sentence: " << D.getLastSentence() << std::endl;
    return 0;
}
