// oracle/shadow/AppConfig.h -- TEST INFRASTRUCTURE ONLY.
//
// Stands in for the reference's src/AppConfig.h (wx paths, the XML configuration tree) on the include path when
// oracle/build_ref.py compiles src/sdr/SDRPostThread.cpp UNMODIFIED: src/sdr/SoapySDRThread.h, which SDRPostThread.h includes,
// declares a pointer to a DeviceConfig and nothing more.
#pragma once
#include <atomic>
class DeviceConfig;
class AppConfig;
// the real AppConfig.h includes util/DataTree.h, whose `using namespace std;` (DataTree.h:36) is what lets SDRPostThread.h write
// `atomic_bool` / `atomic_int` unqualified
using std::atomic_bool;
using std::atomic_int;
