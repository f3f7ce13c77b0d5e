// The reference's DECODER_THREAD (code/websocketServer/main.cpp:203-283) with habdec_b200's types swapped in:
// IQSource::get() -> IQVector -> Decoder::pushSamples -> Decoder::operator() -> callbacks / getters.
// Prints what the callbacks and getters deliver so that tests/test_gpu_cpp_facade.py can hold it against the oracle.
//   decoder_thread file.cf32 fs baud bits stops
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <memory>
#include "habdec_b200/Decoder.hpp"
#include "habdec_b200/IQSource.hpp"

int main(int argc, char** argv)
{
    if (argc < 6) return 2;
    const std::string path = argv[1];
    double fs = atof(argv[2]);
    bool no = false;
    habdec_b200::IQSourceFile src;
    src.quiet = true;
    src.setOption("file_string", &path);                 // IQSource_File.h:205-232 option names
    src.setOption("sampling_rate_double", &fs);
    src.setOption("realtime_bool", &no);
    src.setOption("loop_bool", &no);
    if (!src.init() || !src.start()) { std::cerr << "source\n"; return 1; }

    habdec_b200::Decoder D;                               // typedef habdec::Decoder<TReal> TDecoder (GLOBALS.h:36-37)
    D.livePrint(false);
    D.baud(atof(argv[3])); D.rtty_bits(atoi(argv[4])); D.rtty_stops(float(atof(argv[5])));   // main.cpp:544-553
    D.dc_remove(false);
    D.lowpass_bw(1500); D.lowpass_trans(0.025f);
    D.setupDecimationStagesFactor(256);
    std::string chars;
    size_t n_sent = 0;
    D.sentence_callback_ = [&](std::string cs, std::string data, std::string crc) { ++n_sent; std::cout << "SENT " << cs << "," << data << "*" << crc << "\n"; };
    D.character_callback_ = [&](std::string s) { chars += s; };

    habdec_b200::IQVector samples;                        // TIQVector, main.cpp:233-236
    samples.resize(256 * 256);
    samples.samplingRate(src.samplingRate());
    for (;;) {
        const size_t count = src.get(samples.data(), samples.size());   // main.cpp:238
        if (!count) break;
        samples.resize(count);
        D.pushSamples(samples);                           // main.cpp:243
        D();                                              // main.cpp:245
        samples.resize(256 * 256);
        if (count < samples.size()) break;
    }
    std::cout << "NSENT " << n_sent << "\n";
    std::cout << "LAST " << D.getLastSentence() << "\n";
    std::cout << "RATE " << D.getDecimatedSamplingRate() << " " << D.getDecimationFactor() << " " << D.getBinsCount() << "\n";
    int pl = 0, pr = 0; D.getPeaks(pl, pr);
    std::cout << "PEAKS " << pl << " " << pr << "\n";
    std::cout << "CHARS " << chars.size() << "\n" << chars << "\nEND\n";
    return 0;
}
