// TEST INFRASTRUCTURE ONLY (oracle build) -- never linked into the product.
//
// Stand-in for the FFTW3 single-precision API subset that habdec's FFT wrapper
// uses (reference: code/Decoder/FFT.h:24,54-56; code/Decoder/FFT.cpp:33-37,
// 54-64,95).  FFTW is not installed in this image, so "parity at the FFT
// boundary" is UNPINNED: this shim evaluates the forward, unnormalised DFT in
// float64 (radix-2, per-index twiddles from libm) and rounds once to float32,
// i.e. it is the closest-to-exact answer any float FFT can be compared with.
#pragma once
#include <cstdlib>
#include <cmath>
#include <complex>
#include <vector>

typedef float fftwf_complex[2];
struct hbd_fftwf_plan_s { int n; fftwf_complex* in; fftwf_complex* out; };
typedef hbd_fftwf_plan_s* fftwf_plan;
#define FFTW_FORWARD  (-1)
#define FFTW_ESTIMATE (1U << 6)

static inline void* fftwf_malloc(size_t n) { return std::malloc(n); }
static inline void  fftwf_free(void* p)    { std::free(p); }

static inline fftwf_plan fftwf_plan_dft_1d(int n, fftwf_complex* in, fftwf_complex* out, int /*sign*/, unsigned /*flags*/)
{
    return new hbd_fftwf_plan_s{n, in, out};
}
static inline void fftwf_destroy_plan(fftwf_plan p) { delete p; }

// forward DFT, X[k] = sum_n x[n] exp(-2 pi i k n / N), float64 arithmetic
static inline void hbd_dft_f64(const float* in_iq, float* out_iq, int n)
{
    std::vector<std::complex<double>> a(n);
    for (int i = 0; i < n; ++i) a[i] = {double(in_iq[2 * i]), double(in_iq[2 * i + 1])};
    if (n > 0 && (n & (n - 1)) == 0) {
        for (int i = 1, j = 0; i < n; ++i) {
            int bit = n >> 1;
            for (; j & bit; bit >>= 1) j ^= bit;
            j ^= bit;
            if (i < j) std::swap(a[i], a[j]);
        }
        for (int len = 2; len <= n; len <<= 1) {
            const int half = len / 2;
            std::vector<std::complex<double>> w(half);
            for (int j = 0; j < half; ++j) {
                const double ang = -2.0 * M_PI * double(j) / double(len);
                w[j] = {std::cos(ang), std::sin(ang)};
            }
            for (int i = 0; i < n; i += len)
                for (int j = 0; j < half; ++j) {
                    const std::complex<double> u = a[i + j], v = a[i + j + half] * w[j];
                    a[i + j] = u + v;
                    a[i + j + half] = u - v;
                }
        }
    } else {
        std::vector<std::complex<double>> b(n);
        for (int k = 0; k < n; ++k) {
            std::complex<double> s = 0;
            for (int i = 0; i < n; ++i) {
                const double ang = -2.0 * M_PI * double((long long)k * i % n) / double(n);
                s += a[i] * std::complex<double>(std::cos(ang), std::sin(ang));
            }
            b[k] = s;
        }
        a.swap(b);
    }
    for (int i = 0; i < n; ++i) { out_iq[2 * i] = float(a[i].real()); out_iq[2 * i + 1] = float(a[i].imag()); }
}

static inline void fftwf_execute(fftwf_plan p)
{
    hbd_dft_f64(&p->in[0][0], &p->out[0][0], p->n);
}
