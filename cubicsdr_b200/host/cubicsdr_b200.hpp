// cubicsdr_b200/host/cubicsdr_b200.hpp -- C++ host side above the C-ABI, mirroring CubicSDR's own interfaces for the
// hot path (same names, argument meaning and error behaviour), so that the reference's threads can call the engine
// where they call liquid-dsp today.  Header-only, C++11, depends only on include/sdr_b200.h.
//
// Mirrored reference interfaces (relative to /root/reference/src):
//   Modem / ModemKit / ModemIQData                 modules/modem/Modem.h:19-39,127-166, Modem.cpp:40-55
//   ModemAnalog / ModemFM / NBFM / AM / USB / LSB  modules/modem/ModemAnalog.*, modules/modem/analog/*
//   AudioThreadInput                               audio/AudioThread.h:16-51
//   DemodulatorThreadIQData / PostIQData           demod/DemodDefs.h:18-59
//   SDRPostThread::runSingleCH                     sdr/SDRPostThread.cpp:248-299
//   DemodulatorPreThread (mix + resample part)     demod/DemodulatorPreThread.cpp:154-220
//   SpectrumVisualProcessor / SpectrumVisualData   process/SpectrumVisualProcessor.h:14-23, .cpp:140-178,212-637
//
// Everything lives in namespace cubicsdr_b200 so it can be compiled next to the reference's classes; INTEGRATION.md
// shows the few lines that register these modems and swap the three call sites.
#pragma once
#include <cmath>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/sdr_b200.h"

namespace cubicsdr_b200 {

struct liquid_float_complex { float real, imag; };      // liquid.h:153 layout

// ---- engine singleton-ish holder: one per GPU, shared by every adapter ------------------------------------------------
class Engine {
public:
    explicit Engine(int device = 0, unsigned max_block = 1u << 21) : eng_(nullptr) {
        if (sdr_engine_create(device, max_block, nullptr, &eng_) != SDR_OK) throw std::runtime_error(sdr_last_error());
    }
    ~Engine() { if (eng_) sdr_engine_destroy(eng_); }
    Engine(const Engine &) = delete;
    Engine &operator=(const Engine &) = delete;
    sdr_engine *get() const { return eng_; }
private:
    sdr_engine *eng_;
};

// ---- block carriers (demod/DemodDefs.h:18-59) ---------------------------------------------------------------------------
class DemodulatorThreadIQData {
public:
    long long frequency = 0;
    long long sampleRate = 0;
    std::vector<liquid_float_complex> data;
};
typedef std::shared_ptr<DemodulatorThreadIQData> DemodulatorThreadIQDataPtr;

class DemodulatorThreadPostIQData {
public:
    std::vector<liquid_float_complex> data;
    long long sampleRate = 0;
    std::string modemName, modemType;
};

// ---- audio/AudioThread.h:16-51 (fields written on this path) --------------------------------------------------------------
class AudioThreadInput {
public:
    long long frequency = 0;
    int inputRate = 0;
    int sampleRate = 0;
    int channels = 0;
    float peak = 0;
    int type = 0;
    bool is_squelch_active = false;
    std::vector<float> data;
};

// ---- modules/modem/Modem.h:19-39 -------------------------------------------------------------------------------------------
class ModemKit {
public:
    ModemKit() : sampleRate(0), audioSampleRate(0), handle(nullptr) {}
    virtual ~ModemKit() {}
    long long sampleRate;
    int audioSampleRate;
    sdr_chan *handle;          // engine-side modem state (what the liquid objects + msresamp_rrrf are in the reference)
};
class ModemIQData {
public:
    std::vector<liquid_float_complex> data;
    long long sampleRate = 0;
};

class Modem;
typedef Modem *(*ModemFactoryFn)();

// ---- modules/modem/Modem.h:127-166: the plugin contract -----------------------------------------------------------------------
class Modem {
public:
    virtual ~Modem() {}
    virtual std::string getType() { return "analog"; }
    virtual std::string getName() = 0;
    virtual int getDefaultSampleRate() = 0;
    // Modem::checkSampleRate (ModemAnalog.cpp:14-19, ModemUSB.cpp:29-37)
    virtual int checkSampleRate(long long sampleRate, int /*audioSampleRate*/) {
        return sdr_modem_check_sample_rate(getName().c_str(), sampleRate);
    }
    // ModemAnalog::buildKit (ModemAnalog.cpp:21-33): runs on the DemodulatorWorkerThread
    virtual ModemKit *buildKit(long long sampleRate, int audioSampleRate) {
        ModemKit *kit = new ModemKit;
        kit->sampleRate = sampleRate;
        kit->audioSampleRate = audioSampleRate;
        if (sdr_modem_create(engine()->get(), getName().c_str(), sampleRate, audioSampleRate, kMaxSamplesPerCall, &kit->handle) != SDR_OK) {
            delete kit;
            return nullptr;                                  // the reference's callers treat a null kit as "not ready yet"
        }
        return kit;
    }
    // ModemAnalog::disposeKit (ModemAnalog.cpp:35-40): runs on the channel's DemodulatorThread
    virtual void disposeKit(ModemKit *kit) {
        if (!kit) return;
        if (kit->handle) sdr_chan_destroy(kit->handle);
        delete kit;
    }
    // Modem*::demodulate (e.g. ModemFM.cpp:26-39).  On an engine error audioOut->data is left empty, which
    // DemodulatorThread tolerates (it skips metering, DemodulatorThread.cpp:145, and pushes an empty buffer).
    virtual void demodulate(ModemKit *kit, ModemIQData *input, AudioThreadInput *audioOut) {
        if (!kit || !kit->handle || !input || !audioOut) return;
        const size_t bufSize = input->data.size();
        if (!bufSize) return;                                // ModemAnalog::initOutputBuffers: `if (!bufSize) return`
        const double ratio = double(kit->audioSampleRate) / double(kit->sampleRate);
        const size_t audio_out_size = 2 * ((size_t)std::ceil((double)bufSize * ratio) + 512);   // ModemAnalog.cpp:51 (x 2: FM stereo's two channels)
        resampledOutputData.resize(audio_out_size);
        unsigned numAudioWritten = 0;
        double level = 0.0;
        float peak = 0.f;
        if (sdr_modem_demodulate(kit->handle, reinterpret_cast<const sdr_cf32 *>(input->data.data()), (unsigned)bufSize,
                                 resampledOutputData.data(), (unsigned)audio_out_size, &numAudioWritten, &level, &peak) != SDR_OK) {
            audioOut->data.clear();
            return;
        }
        audioOut->channels = (getName() == "FMS" || getName() == "I/Q") ? 2 : 1;   // ModemAnalog.cpp:90-92; ModemFMStereo.cpp:264, ModemIQ.cpp:46
        audioOut->sampleRate = kit->audioSampleRate;
        audioOut->data.assign(resampledOutputData.begin(), resampledOutputData.begin() + numAudioWritten);
        audioOut->peak = peak;                               // DemodulatorThread.cpp:223-233 computes the same value
        lastLevelDb = level;                                 // DemodulatorThread.cpp:142-163
    }
    double getLastSignalLevelDb() const { return lastLevelDb; }

    // ---- registry: Modem::addModemFactory / makeModem / getModemDefaultSampleRate (Modem.cpp:40-63) ----------------------
    static void addModemFactory(ModemFactoryFn fn, const std::string &name, int defaultRate) {
        factories()[name] = fn;
        defaultRates()[name] = defaultRate;
    }
    static Modem *makeModem(const std::string &name) {
        auto it = factories().find(name);
        return it == factories().end() ? nullptr : it->second();
    }
    static int getModemDefaultSampleRate(const std::string &name) {
        auto it = defaultRates().find(name);
        return it == defaultRates().end() ? 0 : it->second;
    }
    // the engine every modem built through this registry uses (set once at start-up)
    static void setEngine(const std::shared_ptr<Engine> &e) { engine() = e; }

    static const unsigned kMaxSamplesPerCall = 1u << 16;

protected:
    static std::shared_ptr<Engine> &engine() { static std::shared_ptr<Engine> e; return e; }
    std::vector<float> resampledOutputData;
    double lastLevelDb = 0.0;

private:
    static std::map<std::string, ModemFactoryFn> &factories() { static std::map<std::string, ModemFactoryFn> m; return m; }
    static std::map<std::string, int> &defaultRates() { static std::map<std::string, int> m; return m; }
};

#define CUBICSDR_B200_MODEM(Class, Name, Rate)                                   \
    class Class : public Modem {                                                 \
    public:                                                                      \
        static Modem *factory() { return new Class; }                            \
        std::string getName() override { return Name; }                          \
        int getDefaultSampleRate() override { return Rate; }                     \
    };
CUBICSDR_B200_MODEM(ModemFM, "FM", 200000)        // analog/ModemFM.cpp:18-24
CUBICSDR_B200_MODEM(ModemNBFM, "NBFM", 12500)     // analog/ModemNBFM.cpp:18-24
CUBICSDR_B200_MODEM(ModemAM, "AM", 6000)          // analog/ModemAM.cpp:21-27
CUBICSDR_B200_MODEM(ModemUSB, "USB", 5400)        // analog/ModemUSB.cpp:19-41
CUBICSDR_B200_MODEM(ModemLSB, "LSB", 5400)        // analog/ModemLSB.cpp:19-41
CUBICSDR_B200_MODEM(ModemDSB, "DSB", 5400)        // analog/ModemDSB.cpp:15-25
CUBICSDR_B200_MODEM(ModemCW, "CW", 500)           // analog/ModemCW.cpp:99-112
CUBICSDR_B200_MODEM(ModemIQ, "I/Q", 48000)        // analog/ModemIQ.cpp:12-37
CUBICSDR_B200_MODEM(ModemFMStereo, "FMS", 200000) // analog/ModemFMStereo.cpp:15-40
#undef CUBICSDR_B200_MODEM

// what CubicSDR::OnInit does at CubicSDR.cpp:305-311
inline void registerModems(const std::shared_ptr<Engine> &e) {
    Modem::setEngine(e);
    Modem::addModemFactory(ModemFM::factory, "FM", 200000);
    Modem::addModemFactory(ModemNBFM::factory, "NBFM", 12500);
    Modem::addModemFactory(ModemAM::factory, "AM", 6000);
    Modem::addModemFactory(ModemLSB::factory, "LSB", 5400);
    Modem::addModemFactory(ModemUSB::factory, "USB", 5400);
    Modem::addModemFactory(ModemCW::factory, "CW", 500);
    Modem::addModemFactory(ModemDSB::factory, "DSB", 5400);
    Modem::addModemFactory(ModemIQ::factory, "I/Q", 48000);
    Modem::addModemFactory(ModemFMStereo::factory, "FMS", 200000);
}

// ---- SDRPostThread::runSingleCH (sdr/SDRPostThread.cpp:248-299): one block in, DC-blocked block + all attached work out ----
class SDRPostB200 {
public:
    explicit SDRPostB200(const std::shared_ptr<Engine> &e) : eng_(e) {}
    // Runs the DC blocker and every attached spectrum / pre-demod channel on the block.  Never blocks on a consumer
    // (the reference only try_push()es here); returns false on an engine error, in which case the block is dropped.
    bool runSingleCH(const DemodulatorThreadIQData &data_in) {
        if (data_in.data.empty()) return true;
        if (sdr_engine_submit(eng_->get(), reinterpret_cast<const sdr_cf32 *>(data_in.data.data()), (unsigned)data_in.data.size(),
                              data_in.frequency, data_in.sampleRate) != SDR_OK) return false;
        return sdr_engine_wait(eng_->get()) == SDR_OK;
    }
    // SDRPostThread::setChannelizerType (SDRPostThread.cpp:142-144): SDRPostPFBCH (1) makes blocks above 500 kSps take the
    // runPFBCH path; demodulators must then be created with the channel rate of channelPlan().  Call before the first block.
    bool setChannelizerType(int chType) { return sdr_engine_set_channelizer(eng_->get(), chType) == SDR_OK; }
    // numChannels (SDRThread::getOptimalChannelCount) and chanBw = sampleRate / numChannels (:408) for a wideband rate
    static void channelPlan(long long sampleRate, int &numChannels, long long &chanBw) {
        sdr_engine_channel_plan(sampleRate, &numChannels, &chanBw);
    }
    // Pipelined variant: submits the block and returns once the PREVIOUS block's results are on the host (false for the
    // very first block: nothing to fetch yet).  Mirrors the one-block lag the reference's queues between SDRPostThread
    // and the demodulator threads introduce anyway; fetchAudio / getFrame then read that previous block.
    bool runSingleCHPipelined(const DemodulatorThreadIQData &data_in, bool *have_results) {
        if (have_results) *have_results = false;
        if (data_in.data.empty()) return true;
        if (sdr_engine_submit(eng_->get(), reinterpret_cast<const sdr_cf32 *>(data_in.data.data()), (unsigned)data_in.data.size(),
                              data_in.frequency, data_in.sampleRate) != SDR_OK) return false;
        if (!primed_) { primed_ = true; return true; }
        if (sdr_engine_wait_block(eng_->get(), 1) != SDR_OK) return false;
        if (have_results) *have_results = true;
        return true;
    }
    // the DC-corrected block that runSingleCH pushes to the visual queues and the demodulator pipes (:284-298)
    bool fetchDemodDataOut(DemodulatorThreadIQData &out, long long frequency, long long sampleRate, size_t n) {
        out.frequency = frequency; out.sampleRate = sampleRate; out.data.resize(n);
        unsigned got = 0;
        if (sdr_engine_fetch_block(eng_->get(), reinterpret_cast<sdr_cf32 *>(out.data.data()), (unsigned)n, &got) != SDR_OK) return false;
        out.data.resize(got);
        return true;
    }
private:
    std::shared_ptr<Engine> eng_;
    bool primed_ = false;
};

// ---- DemodulatorPreThread's DSP (demod/DemodulatorPreThread.cpp:154-220) + DemodulatorThread's (DemodulatorThread.cpp:120-233)
//      for one demodulator, batched with all the others inside SDRPostB200::runSingleCH ---------------------------------------
class DemodulatorB200 {
public:
    DemodulatorB200(const std::shared_ptr<Engine> &e, const std::string &modemName, long long sampleRate, long long bandwidth,
                    int audioSampleRate, bool wantIQ = false)
        : eng_(e), ch_(nullptr), name_(modemName), audioRate_(audioSampleRate) {
        if (sdr_chan_create(e->get(), modemName.c_str(), sampleRate, bandwidth, audioSampleRate, &ch_) != SDR_OK)
            throw std::runtime_error(sdr_last_error());
        sdr_chan_get_bandwidth(ch_, &bandwidth_);
        sdr_chan_result_caps(ch_, &audioCap_, &iqCap_);          // most samples one block can yield: the fetch buffers' size
        sdr_chan_audio_channels(ch_, &channels_);                // 2 for "I/Q" (ModemIQ.cpp:46) and "FMS" (ModemFMStereo.cpp:264), else 1
        if (wantIQ) sdr_chan_want_iq(ch_, 1);
    }
    ~DemodulatorB200() { if (ch_) sdr_chan_destroy(ch_); }
    DemodulatorB200(const DemodulatorB200 &) = delete;
    DemodulatorB200 &operator=(const DemodulatorB200 &) = delete;
    void setFrequency(long long f) { sdr_chan_set_frequency(ch_, f); }     // DemodulatorInstance::setFrequency
    int getBandwidth() const { return bandwidth_; }
    // after SDRPostB200::runSingleCH: what DemodulatorThread would push to the audio queue for this block.
    // Returns false when the block was skipped for this channel (|shift| > 0.75 fs) or on error.
    bool fetchAudio(AudioThreadInput &ati, double *signalLevelDb = nullptr) {
        ati.data.resize(audioCap_);
        unsigned n = 0; double lvl = 0; float pk = 0; int skipped = 0;
        if (sdr_chan_fetch_audio(ch_, ati.data.data(), (unsigned)ati.data.size(), &n, &lvl, &pk, &skipped) != SDR_OK) { ati.data.clear(); return false; }
        ati.data.resize(n);
        ati.channels = channels_; ati.sampleRate = audioRate_; ati.inputRate = bandwidth_; ati.peak = pk;
        if (signalLevelDb) *signalLevelDb = lvl;
        int squelched = 0;                                       // DemodulatorThread.cpp:198, 235-238
        sdr_chan_fetch_meter(ch_, &signalLevel_, &signalFloor_, &signalCeil_, &squelched);
        ati.is_squelch_active = squelched != 0;
        // squelchBreak bookkeeping (:200-220); the visual cue / solo-mode lock that hang off it stay with the GUI host
        if (squelchEnabled_) {
            if (!squelched && !squelchBreak_) squelchBreak_ = true;
            else if (squelched && squelchBreak_) squelchBreak_ = false;
        }
        return !skipped;
    }
    // DemodulatorThread::setSquelchEnabled / setSquelchLevel / get* (DemodulatorThread.cpp:372-404)
    void setSquelchEnabled(bool on) { squelchEnabled_ = on; sdr_chan_set_squelch(ch_, on ? 1 : 0, squelchLevel_); }
    void setSquelchLevel(float db) { squelchEnabled_ = true; squelchLevel_ = db; sdr_chan_set_squelch(ch_, 1, db); }
    bool isSquelchEnabled() const { return squelchEnabled_; }
    float getSquelchLevel() const { return squelchLevel_; }
    float getSignalLevel() const { return signalLevel_; }
    float getSignalFloor() const { return signalFloor_; }
    float getSignalCeil() const { return signalCeil_; }
    bool getSquelchBreak() const { return squelchBreak_; }
    bool fetchPostIQ(DemodulatorThreadPostIQData &out) {                    // DemodulatorPreThread.cpp:211-217
        out.data.resize(iqCap_);
        unsigned n = 0;
        if (sdr_chan_fetch_iq(ch_, reinterpret_cast<sdr_cf32 *>(out.data.data()), (unsigned)out.data.size(), &n) != SDR_OK) { out.data.clear(); return false; }
        out.data.resize(n);
        out.sampleRate = bandwidth_; out.modemName = name_; out.modemType = "analog";
        return true;
    }
private:
    std::shared_ptr<Engine> eng_;
    sdr_chan *ch_;
    std::string name_;
    int bandwidth_ = 0, audioRate_, channels_ = 1;
    unsigned audioCap_ = 0, iqCap_ = 0;
    bool squelchEnabled_ = false, squelchBreak_ = false;
    float squelchLevel_ = -100.f, signalLevel_ = -100.f, signalFloor_ = -30.f, signalCeil_ = 30.f;
};

// ---- SpectrumVisualProcessor (process/SpectrumVisualProcessor.h:14-23) ---------------------------------------------------------
class SpectrumVisualData {
public:
    std::vector<float> spectrum_points;        // interleaved (x_i = i / fftSize, y_i)   (.cpp:560-568)
    std::vector<float> spectrum_hold_points;   // peak hold: not produced by the engine (GUI feature)
    double fft_ceiling = 0, fft_floor = 0;
    long long centerFreq = 0;
    int bandwidth = 0;
    std::vector<unsigned char> waterfall_line; // WaterfallPanel::step quantisation of the y values (panel/WaterfallPanel.cpp:64-70)
};

class SpectrumVisualProcessorB200 {
public:
    // fftSize = display points (setup(), .cpp:140-178); linesPerSecond = FFTDataDistributor::setLinesPerSecond, 0 = every frame
    SpectrumVisualProcessorB200(const std::shared_ptr<Engine> &e, unsigned fftSize, unsigned linesPerSecond = 0)
        : eng_(e), sp_(nullptr), fftSize_(fftSize) {
        if (sdr_spectrum_create(e->get(), fftSize, linesPerSecond, &sp_) != SDR_OK) throw std::runtime_error(sdr_last_error());
    }
    ~SpectrumVisualProcessorB200() { if (sp_) sdr_spectrum_destroy(sp_); }
    SpectrumVisualProcessorB200(const SpectrumVisualProcessorB200 &) = delete;
    SpectrumVisualProcessorB200 &operator=(const SpectrumVisualProcessorB200 &) = delete;
    void setFFTAverageRate(float r) { sdr_spectrum_set_average_rate(sp_, r); }
    unsigned getFFTSize() const { return fftSize_; }
    int getDesiredInputSize() const { return (int)(2 * fftSize_); }           // fftSizeInternal (.cpp:388)
    // frames produced from the block submitted last; each is what process() would distribute() (.cpp:626-632)
    unsigned framesReady() const { unsigned n = 0; sdr_spectrum_frames(sp_, &n); return n; }
    bool getFrame(unsigned idx, SpectrumVisualData &out, long long centerFreq, int bandwidth) {
        std::vector<float> y(fftSize_);
        out.waterfall_line.resize(fftSize_);
        double cf[2];
        if (sdr_spectrum_fetch(sp_, idx, y.data(), out.waterfall_line.data(), cf) != SDR_OK) return false;
        out.spectrum_points.resize(2 * (size_t)fftSize_);
        for (unsigned i = 0; i < fftSize_; ++i) {
            out.spectrum_points[2 * i] = (float)i / (float)fftSize_;
            out.spectrum_points[2 * i + 1] = y[i];
        }
        out.spectrum_hold_points.resize(0);
        out.fft_ceiling = cf[0];                                              // point_ceil / sf with sf = 1 (.cpp:626)
        out.fft_floor = cf[1];
        out.centerFreq = centerFreq; out.bandwidth = bandwidth;
        return true;
    }
private:
    std::shared_ptr<Engine> eng_;
    sdr_spectrum *sp_;
    unsigned fftSize_;
};

}  // namespace cubicsdr_b200
