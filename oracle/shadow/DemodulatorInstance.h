// oracle/shadow/DemodulatorInstance.h -- TEST INFRASTRUCTURE ONLY.
//
// Stands in for the reference's src/demod/DemodulatorInstance.h (which needs the whole application) on the include path when
// oracle/build_ref.py compiles src/audio/AudioThread.cpp UNMODIFIED.  That source touches a demodulator instance in one place,
// AudioThread::setSampleRate of a controller thread (AudioThread.cpp:349-355), which the oracle never reaches with a
// demodulator: the two methods exist so that the lines compile.
#pragma once
#include <memory>
#include <vector>
class DemodulatorInstance {
public:
    int getOutputDevice() { return -1; }
    void setAudioSampleRate(int) {}
};
typedef std::shared_ptr<DemodulatorInstance> DemodulatorInstancePtr;
