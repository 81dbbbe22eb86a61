// oracle/shadow/CubicSDR.h -- TEST INFRASTRUCTURE ONLY.
//
// Stands in for the reference's application header (src/CubicSDR.h: the wx application, frames, managers) on the include path
// of oracle/build_ref.py when it compiles src/process/SpectrumVisualProcessor.cpp UNMODIFIED ("shadow
// CubicSDR.h with stubs on the include path") and src/audio/AudioThread.cpp.  The former asks the application object for
// exactly one thing, wxGetApp().getSampleRate() (SpectrumVisualProcessor.cpp:306), the device sample rate, which
// oracle/ref_visuals.cpp sets; the latter for the list of demodulators (AudioThread.cpp:349), empty here.
#pragma once
// (the real header brings these in transitively)
#include <cmath>
#include <cstdlib>
#include <cstring>
#include "DemodulatorInstance.h"          // oracle/shadow
struct OracleDemodMgrStub {               // AudioThread.cpp:349 asks for the list of demodulators: there are none here
    std::vector<DemodulatorInstancePtr> getDemodulators() { return {}; }
};
struct OracleAppStub {
    long long sampleRate = 0;
    OracleDemodMgrStub demodMgr;
    long long getSampleRate() const { return sampleRate; }
    OracleDemodMgrStub &getDemodMgr() { return demodMgr; }
};
OracleAppStub &wxGetApp();
