// TEST INFRASTRUCTURE ONLY: empty stand-in so that the reference's IQSource_File.h (which includes the SoapySDR
// header without using it) compiles in an image without SoapySDR.
#pragma once
