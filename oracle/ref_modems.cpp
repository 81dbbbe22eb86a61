// oracle/ref_modems.cpp -- TEST INFRASTRUCTURE ONLY.
//
// C entry points over the reference's OWN, UNMODIFIED modem classes.  oracle/build_ref.py compiles
// /root/reference/src/modules/modem/{Modem,ModemAnalog}.cpp and analog/Modem{FM,NBFM,FMStereo,AM,CW,LSB,USB,DSB,IQ}.cpp from
// where they lie (nothing is copied into this repository) into oracle/_ref/libmodems_ref.so together with this file; their
// liquid-dsp calls land in the reference's own liquid 1.5.0 binary through oracle/ref_modem_abi.c.  The tests use it to pin
// the restated modem glue of oracle/ref_chain.cpp (which the GPU parity tests compare against) to the source it restates:
// same decimated IQ in, bit-identical audio out (tests/test_oracle_modems.py).
//
// The registration below is what CubicSDRApp::OnInit does (src/CubicSDR.cpp:305-313): same names, same default rates.
#include "Modem.h"
#include "ModemFM.h"
#include "ModemNBFM.h"
#include "ModemFMStereo.h"
#include "ModemAM.h"
#include "ModemCW.h"
#include "ModemLSB.h"
#include "ModemUSB.h"
#include "ModemDSB.h"
#include "ModemIQ.h"

#include <cstring>
#include <mutex>

namespace {
std::once_flag g_registered;
void register_modems() {
    Modem::addModemFactory(ModemFM::factory, "FM", 200000);
    Modem::addModemFactory(ModemNBFM::factory, "NBFM", 12500);
    Modem::addModemFactory(ModemFMStereo::factory, "FMS", 200000);
    Modem::addModemFactory(ModemAM::factory, "AM", 6000);
    Modem::addModemFactory(ModemCW::factory, "CW", 500);
    Modem::addModemFactory(ModemLSB::factory, "LSB", 5400);
    Modem::addModemFactory(ModemUSB::factory, "USB", 5400);
    Modem::addModemFactory(ModemDSB::factory, "DSB", 5400);
    Modem::addModemFactory(ModemIQ::factory, "I/Q", 48000);
}
struct RefModem {
    Modem *modem = nullptr;
    ModemKit *kit = nullptr;
    ModemIQData in;
    AudioThreadInput out;
};
}  // namespace

extern "C" {

void *refmodem_create(const char *name) {
    std::call_once(g_registered, register_modems);
    Modem *m = Modem::makeModem(name);
    if (!m) return nullptr;
    RefModem *r = new RefModem;
    r->modem = m;
    return r;
}
int refmodem_default_rate(const char *name) {
    std::call_once(g_registered, register_modems);
    return Modem::getModemDefaultSampleRate(name);
}
int refmodem_type(void *h, char *buf, int cap) {                      // Modem::getType(): "analog" / "digital"
    const std::string t = static_cast<RefModem *>(h)->modem->getType();
    std::strncpy(buf, t.c_str(), cap > 0 ? cap - 1 : 0);
    if (cap > 0) buf[cap - 1] = 0;
    return (int)t.size();
}
int refmodem_check_sample_rate(void *h, long long sample_rate, int audio_rate) {
    return static_cast<RefModem *>(h)->modem->checkSampleRate(sample_rate, audio_rate);
}
void refmodem_write_setting(void *h, const char *key, const char *value) {
    static_cast<RefModem *>(h)->modem->writeSetting(key, value);
}
// (re)build the kit the way DemodulatorWorkerThread does (src/demod/DemodulatorWorkerThread.cpp:63-108): dispose the old
// one, buildKit(sampleRate, audioSampleRate)
int refmodem_build_kit(void *h, long long sample_rate, int audio_rate) {
    RefModem *r = static_cast<RefModem *>(h);
    if (r->kit) r->modem->disposeKit(r->kit);
    r->kit = r->modem->buildKit(sample_rate, audio_rate);
    r->modem->clearRebuildKit();
    return r->kit ? 0 : -1;
}
int refmodem_should_rebuild_kit(void *h) { return static_cast<RefModem *>(h)->modem->shouldRebuildKit() ? 1 : 0; }
int refmodem_use_signal_output(void *h) { return static_cast<RefModem *>(h)->modem->useSignalOutput() ? 1 : 0; }

// Modem::demodulate(kit, input, audioOut) on n decimated IQ samples {re, im}; returns the number of floats in audioOut->data
// (copied to `audio` up to `cap`), *channels / *out_rate as the modem left them in audioOut (ModemAnalog.cpp:81-93).
int refmodem_demodulate(void *h, const float *iq, unsigned n, long long sample_rate, float *audio, unsigned cap, int *channels,
                        int *out_rate) {
    RefModem *r = static_cast<RefModem *>(h);
    if (!r->kit) return -1;
    r->in.sampleRate = sample_rate;
    r->in.data.resize(n);
    if (n) std::memcpy(reinterpret_cast<void *>(r->in.data.data()), iq, (size_t)n * 2 * sizeof(float));
    r->out.data.clear();
    r->modem->demodulate(r->kit, &r->in, &r->out);
    const size_t m = r->out.data.size();
    if (audio && m) std::memcpy(audio, r->out.data.data(), (m < cap ? m : cap) * sizeof(float));
    if (channels) *channels = r->out.channels;
    if (out_rate) *out_rate = r->out.sampleRate;
    return (int)m;
}
void refmodem_destroy(void *h) {
    RefModem *r = static_cast<RefModem *>(h);
    if (!r) return;
    if (r->kit) r->modem->disposeKit(r->kit);
    delete r->modem;
    delete r;
}

}  // extern "C"
