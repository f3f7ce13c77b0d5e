// TEST INFRASTRUCTURE ONLY: stand-in so that the reference's IQSource_File.h (which includes the SoapySDR header
// without using it, but relies on the standard headers it pulls in) compiles in an image without SoapySDR.
#pragma once
#include <algorithm>
#include <chrono>
#include <complex>
#include <string>
#include <vector>
