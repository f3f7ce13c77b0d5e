// The step in front of the decoder: the reference's IQ source interface and its cf32-file implementation,
// restated for this library (header only, host C++17, no CUDA needed).
//
//   habdec::IQSource            code/IQSource/IQSource.h:30-43          -> habdec_b200::IQSource
//   habdec::IQSource_File<float> code/IQSource/IQSource_File.h:40-262   -> habdec_b200::IQSourceFile
//
// Same contract: get() fills interleaved cf32, samplingRate(), stringly-typed setOption()/getOption()
// ("file_string", "sampling_rate_double", "realtime_bool", "loop_bool").  Behaviour that callers may rely on is
// kept, including the odd bits:
//   * count() = file size / 8; get() reads min(requested, count()) samples and returns what it got;
//   * EOF is noticed on the call AFTER the short read (stream eof flag): with loop the file is rewound and that
//     call already delivers data, without loop it returns 0 from then on (IQSource_File.h:150-165);
//   * real-time pacing sleeps size_t(read / fs * 1000) milliseconds per call (:177-181), on by default;
//   * setOption("file_string") stores the path but reports false (the if/else chain at :205-231 falls into its
//     "unknown option" branch for it); stop() leaves the source running (:96-102).
// WidebandFeeder / MultiFileFeeder below turn sources into the batched pushes of the decoder.
#pragma once
#include <algorithm>
#include <chrono>
#include <complex>
#include <fstream>
#include <iostream>
#include <memory>
#include <string>
#include <thread>
#include <vector>

namespace habdec_b200 {

class IQSource {
public:
    virtual ~IQSource() = default;
    virtual bool init() = 0;
    virtual bool start() = 0;
    virtual bool stop() = 0;
    virtual bool isRunning() const = 0;
    virtual std::string type() const = 0;
    virtual size_t count() const = 0;
    virtual size_t get(void* p_data, const size_t i_count) = 0;
    virtual double samplingRate() const = 0;
    virtual bool setOption(const std::string& option, const void* p_data) = 0;
    virtual bool getOption(const std::string& option, void* p_data) = 0;
};

class IQSourceFile : public IQSource {
public:
    using TComplex = std::complex<float>;
    bool quiet = false;   // the reference prints its EOF / short-read notes to stdout; tests switch that off

    bool init() override
    {
        std::ifstream probe(path_, std::ifstream::ate | std::ifstream::binary);
        const auto bytes = probe.tellg();
        count_ = bytes > 0 ? size_t(bytes) / sizeof(TComplex) : 0;
        file_.open(path_, std::ios::binary);
        return file_.is_open();
    }
    bool start() override { if (!file_.is_open()) return false; running_ = true; return true; }
    bool stop() override { if (!file_.is_open()) return false; running_ = true; return true; } // sic (IQSource_File.h:96-102)
    bool isRunning() const override { return file_.is_open() && running_; }
    std::string type() const override { return "File"; }
    size_t count() const override { return count_; }
    double samplingRate() const override { return sampling_rate_; }

    size_t get(void* p_data, const size_t i_count) override
    {
        if (!file_.is_open()) { if (!quiet) std::cout << "file_handle_ == 0" << std::endl; return 0; }
        if (!isRunning()) { if (!quiet) std::cout << "Not Running." << std::endl; return 0; }
        if (file_.eof()) {
            if (loop_) { if (!quiet) std::cout << path_ << " EOF. REWIND." << std::endl; file_.clear(); file_.seekg(0); }
            else { if (!quiet) std::cout << path_ << " EOF." << std::endl; return 0; }
        }
        file_.read(reinterpret_cast<char*>(p_data), std::streamsize(std::min(i_count, count_) * sizeof(TComplex)));
        const size_t read_count = size_t(file_.gcount()) / sizeof(TComplex);
        if (!file_ && read_count != i_count && !quiet)
            std::cout << "IQSource_File<T>::get() read less than desired: " << read_count << " of " << i_count << std::endl;
        if (realtime_) {
            const size_t wait = size_t(double(read_count) / sampling_rate_ * 1000);
            std::this_thread::sleep_for(std::chrono::duration<double, std::milli>(double(wait)));
        }
        return read_count;
    }

    bool setOption(const std::string& option, const void* p_data) override
    {
        if (option == "file_string") path_ = *static_cast<const std::string*>(p_data);
        if (option == "sampling_rate_double") sampling_rate_ = *static_cast<const double*>(p_data);
        else if (option == "realtime_bool") realtime_ = *static_cast<const bool*>(p_data);
        else if (option == "loop_bool") loop_ = *static_cast<const bool*>(p_data);
        else {
            if (!quiet) std::cout << "IQSource_File::setOption error. Unknown option: " << option << std::endl;
            return false;
        }
        return true;
    }
    bool getOption(const std::string& option, void* p_data) override
    {
        if (option == "file_string") *static_cast<std::string*>(p_data) = path_;
        if (option == "sampling_rate_double") *static_cast<double*>(p_data) = sampling_rate_;
        else if (option == "realtime_bool") *static_cast<bool*>(p_data) = realtime_;
        else if (option == "loop_bool") *static_cast<bool*>(p_data) = loop_;
        else {
            if (!quiet) std::cout << "IQSource_File::getOption error. Unknown option: " << option << std::endl;
            return false;
        }
        return true;
    }

private:
    bool running_ = false, realtime_ = true, loop_ = false;
    std::ifstream file_;
    std::string path_;
    double sampling_rate_ = 0;
    size_t count_ = 0;
};

// ---- batch feeders: what DECODER_THREAD's `src->get(buf, 65536); decoder.pushSamples(buf); decoder();` becomes
// ---- (code/websocketServer/main.cpp:235-245) when one process drives many channels -------------------------------

// N sources -> one host matrix [n_channels][chunk] per step (pushSamples(host_matrix, n, pitch, fs)).
// Channels whose source ran dry deliver zeros for the rest of the row; next() returns the longest read.
class MultiSourceFeeder {
public:
    explicit MultiSourceFeeder(size_t chunk = 256 * 256) : chunk_(chunk) {}
    void add(std::shared_ptr<IQSource> s) { src_.push_back(std::move(s)); rows_.resize(src_.size() * chunk_); }
    size_t channels() const { return src_.size(); }
    size_t chunk() const { return chunk_; }
    double samplingRate() const { return src_.empty() ? 0 : src_[0]->samplingRate(); }
    const std::complex<float>* matrix() const { return rows_.data(); }
    const std::vector<size_t>& counts() const { return got_; }
    size_t next()
    {
        got_.assign(src_.size(), 0);
        size_t longest = 0;
        for (size_t c = 0; c < src_.size(); ++c) {
            std::complex<float>* row = rows_.data() + c * chunk_;
            const size_t n = src_[c]->get(row, chunk_);
            std::fill(row + n, row + chunk_, std::complex<float>(0.f, 0.f));
            got_[c] = n;
            longest = std::max(longest, n);
        }
        return longest;
    }
private:
    size_t chunk_;
    std::vector<std::shared_ptr<IQSource>> src_;
    std::vector<std::complex<float>> rows_;
    std::vector<size_t> got_;
};

} // namespace habdec_b200
