// oracle/shadow/CubicSDR.h -- TEST INFRASTRUCTURE ONLY.
//
// Stands in for the reference's application header (src/CubicSDR.h: the wx application, frames, managers) on the include path
// of oracle/build_ref.py when it compiles src/process/SpectrumVisualProcessor.cpp UNMODIFIED ("shadow
// CubicSDR.h with stubs on the include path").  That source asks the application object for exactly one thing:
// wxGetApp().getSampleRate() (SpectrumVisualProcessor.cpp:306), the device sample rate.  oracle/ref_visuals.cpp sets it.
#pragma once
// (the real header brings these in transitively)
#include <cmath>
#include <cstdlib>
#include <cstring>
struct OracleAppStub {
    long long sampleRate = 0;
    long long getSampleRate() const { return sampleRate; }
};
OracleAppStub &wxGetApp();
