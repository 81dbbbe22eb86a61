// oracle/shadow/CubicSDR.h -- TEST INFRASTRUCTURE ONLY.
//
// Stands in for the reference's application header (src/CubicSDR.h: the wx application, frames, managers) on the include path
// of oracle/build_ref.py when it compiles src/process/SpectrumVisualProcessor.cpp UNMODIFIED ("shadow
// CubicSDR.h with stubs on the include path") and src/audio/AudioThread.cpp.  The former asks the application object for
// exactly one thing, wxGetApp().getSampleRate() (SpectrumVisualProcessor.cpp:306), the device sample rate, which
// oracle/ref_visuals.cpp sets; the second for the list of demodulators (AudioThread.cpp:349), empty here; and
// src/audio/AudioFile.cpp / AudioFileWAV.cpp ask its configuration for the recording directory; the demodulator threads
// (src/demod/DemodulatorWorkerThread.cpp, DemodulatorThread.cpp) notify the GUI and ask for solo mode (off).
#pragma once
// (the real header brings these in transitively)
#include <cmath>
#include <cstdlib>
#include <cstring>
#include "DemodulatorInstance.h"          // the reference's own (it compiles without the application)
struct OracleDemodMgrStub {
    std::vector<DemodulatorInstancePtr> demods;                  // filled by oracle/ref_demod.cpp's application-level driver
    DemodulatorInstancePtr current;
    // AudioThread.cpp:349, SDRPostThread.cpp:53,103: the list of demodulators
    std::vector<DemodulatorInstancePtr> getDemodulators() { return demods; }
    // SDRPostThread.cpp:70-86, 304: the demodulator the GUI has selected; DemodulatorThread.cpp:206-207, 320 (solo mode only, off here)
    void setActiveDemodulator(DemodulatorInstancePtr d) { current = d; }
    void setActiveDemodulatorByRawPointer(DemodulatorInstance *, bool) {}
    DemodulatorInstancePtr getCurrentModem() { return current; }
    // DemodulatorInstance.cpp:327-328, 524: the manager remembers the last squelch / mute setting for new demodulators
    void setLastSquelchLevel(float) {}
    void setLastSquelchEnabled(bool) {}
    void setLastMuted(bool) {}
};
struct OracleBookmarkMgrStub { void updateActiveList() {} };      // DemodulatorInstance.cpp:162, 298, 393, 456: a GUI list refresh
// DemodulatorInstance::startRecording builds a file name out of the user's label through wxString (:611-613)
class wxString {
    std::wstring w;
public:
    wxString(const std::wstring &s) : w(s) {}
    std::string ToStdString() const { return std::string(w.begin(), w.end()); }
};
struct OracleAppFrameStub {               // DemodulatorWorkerThread.cpp:70, DemodulatorThread.cpp:202: GUI notifications
    void notifyUpdateModemProperties() {}
    bool isUserDemodBusy() { return false; }
};
#include <string>
#include "CubicSDRDefs.h"                 // the reference's own (filePathSeparator for AudioFile.cpp; it needs nothing else)
struct OracleConfigStub {                 // AudioFile.cpp:18 / AudioFileWAV.cpp:187 ask the configuration for the recording directory
    std::string recordingPath;
    std::string getRecordingPath() { return recordingPath; }
    int getRecordingSquelchOption() const { return 0; }          // DemodulatorInstance.cpp:624-625 (AudioSinkFileThread::SQUELCH_RECORD_SILENCE, no time limit)
    int getRecordingFileTimeLimit() const { return 0; }
};
struct OracleAppStub {
    long long sampleRate = 0;
    OracleDemodMgrStub demodMgr;
    OracleBookmarkMgrStub bookmarkMgr;
    OracleAppFrameStub appFrame;
    OracleConfigStub config;
    long long frequency = 0;                                     // the tuner frequency (SDRPostThread.cpp:50, 79)
    long long getSampleRate() const { return sampleRate; }
    long long getFrequency() const { return frequency; }
    void setFrequency(long long f) { frequency = f; }
    bool getSoloMode() const { return false; }
    OracleDemodMgrStub &getDemodMgr() { return demodMgr; }
    OracleBookmarkMgrStub &getBookmarkMgr() { return bookmarkMgr; }
    OracleAppFrameStub *getAppFrame() { return &appFrame; }
    OracleConfigStub *getConfig() { return &config; }
};
OracleAppStub &wxGetApp();
