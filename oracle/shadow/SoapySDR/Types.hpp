// oracle/shadow/SoapySDR/Types.hpp -- TEST INFRASTRUCTURE ONLY.
//
// SoapySDR is not installed here.  The reference's src/sdr/SoapySDRThread.h and SDRDeviceInfo.h -- which src/sdr/SDRPostThread.h
// includes, and through which oracle/build_ref.py compiles src/sdr/SDRPostThread.cpp UNMODIFIED -- only DECLARE members of these
// types; SDRPostThread.cpp itself uses none of them.  The names below are what those declarations need to parse.
#pragma once
#include <map>
#include <string>
#include <vector>
namespace SoapySDR {
typedef std::map<std::string, std::string> Kwargs;
class Range {
public:
    Range() : lo(0), hi(0) {}
    Range(double minimum, double maximum) : lo(minimum), hi(maximum) {}
    double minimum() const { return lo; }
    double maximum() const { return hi; }
private:
    double lo, hi;
};
typedef std::vector<Range> RangeList;
struct ArgInfo { std::string key, value, name; };
typedef std::vector<ArgInfo> ArgInfoList;
class Device;
class Stream;
}  // namespace SoapySDR
