// oracle/shadow/SoapySDR/Device.hpp -- TEST INFRASTRUCTURE ONLY: see Types.hpp (SoapySDR is not installed; nothing of it is used).
#pragma once
#include "Types.hpp"
