// oracle/ref_demod.cpp -- TEST INFRASTRUCTURE ONLY.
//
// C entry points over the reference's OWN, UNMODIFIED demodulator: a real DemodulatorInstance with its DemodulatorPreThread
// (frequency shift + msresamp_crcf, src/demod/DemodulatorPreThread.cpp), DemodulatorWorkerThread (modem + kit + resampler
// construction), DemodulatorThread (Modem::demodulate, signal level, floor / ceiling trackers, squelch, peak,
// src/demod/DemodulatorThread.cpp) and the modem classes, all running on their own threads exactly as in the application.
// oracle/build_ref.py compiles src/demod/{DemodulatorInstance,DemodulatorPreThread,DemodulatorWorkerThread,DemodulatorThread}.cpp,
// src/audio/{AudioThread,AudioSinkThread,AudioSinkFileThread,AudioFile,AudioFileWAV}.cpp, src/IOThread.cpp, src/util/Timer.cpp,
// the modem sources and the reference's RtAudio (dummy API) from where they lie -- nothing is copied into this repository --
// into oracle/_ref/libdemod_ref.so; the application header is shadowed by oracle/shadow/CubicSDR.h (GUI notifications, solo mode
// off).  Compiled with -fno-access-control for two things the application does through an audio device this box does not have:
// the audio thread's sample rate (AudioThread::setupDevice) and the tap on the demodulator's "AudioSink" output.
//
// tests/test_oracle_demod.py pins the restated per-channel chain of oracle/ref_chain.cpp (RefChan: what every GPU parity test
// compares against) to it: the same wideband blocks in, the SAME BITS of audio, peak, level trackers and squelch decision out.
//
// Threads are asynchronous; the driver makes each block deterministic with sentinels: a block with no samples pushed behind
// the real one is popped by DemodulatorPreThread only after the real block's loop iteration has ended, and likewise an empty
// post-IQ item for DemodulatorThread (both sources skip empty items: DemodulatorPreThread.cpp:168, DemodulatorThread.cpp:92-95).
#include "CubicSDR.h"                     // oracle/shadow
#include "DemodulatorInstance.h"
#include "DemodulatorPreThread.h"
#include "DemodulatorThread.h"
#include "SDRPostThread.h"
#include "ModemFM.h"
#include "ModemNBFM.h"
#include "ModemFMStereo.h"
#include "ModemAM.h"
#include "ModemCW.h"
#include "ModemLSB.h"
#include "ModemUSB.h"
#include "ModemDSB.h"
#include "ModemIQ.h"

#include <chrono>
#include <cstring>
#include <mutex>
#include <thread>

OracleAppStub &wxGetApp() {
    static OracleAppStub app;
    return app;
}

namespace {
std::once_flag g_registered;
void register_modems() {                  // src/CubicSDR.cpp:305-313
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
struct Chain {
    DemodulatorInstance *inst = nullptr;
    AudioThreadInputQueuePtr sink;         // the demodulator thread's "AudioSink" output: every block's AudioThreadInput
    DemodulatorThreadPostInputQueuePtr post;
    long long rate = 0;
    long long last_center = 0;
};
template <class Q>
bool wait_empty(const Q &q, int timeout_ms) {
    const auto t0 = std::chrono::steady_clock::now();
    while (!q->empty()) {
        if (std::chrono::steady_clock::now() - t0 > std::chrono::milliseconds(timeout_ms)) return false;
        std::this_thread::sleep_for(std::chrono::microseconds(50));
    }
    return true;
}
DemodulatorThreadIQDataPtr make_block(const float *iq, unsigned n, long long rate, long long center) {
    auto p = std::make_shared<DemodulatorThreadIQData>();
    p->frequency = center;
    p->sampleRate = rate;
    p->data.resize(n);
    if (n) std::memcpy(reinterpret_cast<void *>(p->data.data()), iq, (size_t)n * 2 * sizeof(float));
    return p;
}
// Bring the demodulator to rest: the start-up above usually leaves `sampleRateChanged` / `bandwidthChanged` raised, so the next
// block makes DemodulatorPreThread order a second set of filters from its worker (DemodulatorPreThread.cpp:135-151), which
// arrive some blocks later -- whenever the worker happens to finish -- and replace the running resampler mid-stream.  In the
// application that is a start-up transient nobody hears; here it would make the comparison depend on thread timing.  So: let an
// empty block raise the order, wait for the worker with a command it ignores (it takes its next command only once the previous
// one has been answered, DemodulatorWorkerThread.cpp:23-50), let another empty block collect the answer, until nothing is
// pending.  The demodulator then starts the first real block with fresh filters, like the restated chain.
bool settle(Chain *c, long long center) {
    auto in = c->inst->getIQInputDataPipe();
    DemodulatorPreThread *p = c->inst->demodulatorPreThread;
    for (int round = 0; round < 16; ++round) {
        for (int k = 0; k < 2; ++k) in->push(make_block(nullptr, 0, c->rate, center));   // the first one's iteration has ended
        if (!wait_empty(in, 20000)) return false;                                          // when the second has been popped
        p->workerQueue->push(DemodulatorWorkerThreadCommand());                             // (type NULL: ignored by the worker)
        if (!wait_empty(p->workerQueue, 20000)) return false;
        for (int k = 0; k < 2; ++k) in->push(make_block(nullptr, 0, c->rate, center));
        if (!wait_empty(in, 20000)) return false;
        const bool pending = p->sampleRateChanged.load() || p->bandwidthChanged.load() || p->audioSampleRateChanged.load() ||
                             p->demodTypeChanged.load() || !p->workerResults->empty() || !p->workerQueue->empty();
        if (!pending && round > 0) return true;
    }
    return false;
}
}  // namespace

extern "C" {

// a demodulator the way the application makes one (DemodulatorMgr::newThread + the caller's setters, e.g.
// src/visual/WaterfallCanvas.cpp: setFrequency, setDemodulatorType, setBandwidth, run): returns once its modem is initialized
void *refsrc_demod_create(const char *modem, long long sample_rate, int bandwidth, int audio_rate, long long frequency, long long center) {
    std::call_once(g_registered, register_modems);
    Chain *c = new Chain;
    c->rate = sample_rate;
    c->last_center = center;
    c->inst = new DemodulatorInstance();
    c->inst->audioThread->sampleRate = audio_rate;              // what AudioThread::setupDevice takes from the device
    c->inst->setDemodulatorType(modem);
    c->inst->setBandwidth(bandwidth);
    c->inst->setFrequency(frequency);
    c->sink = std::make_shared<AudioThreadInputQueue>();
    c->sink->set_max_num_items(1u << 16);
    c->inst->demodulatorThread->setOutputQueue("AudioSink", c->sink);
    c->post = c->inst->pipeIQDemodData;
    c->inst->run();
    // the first blocks an application feeds are consumed while the worker thread builds the modem (timing dependent, nothing
    // comes out); here blocks WITHOUT samples do that job, so that the first real block meets an initialized demodulator at rest
    const auto t0 = std::chrono::steady_clock::now();
    while (!c->inst->isModemInitialized()) {
        c->inst->getIQInputDataPipe()->push(make_block(nullptr, 0, sample_rate, center));
        std::this_thread::sleep_for(std::chrono::milliseconds(1));
        if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(20)) { return nullptr; }
    }
    if (!settle(c, center)) return nullptr;
    return c;
}
void refsrc_demod_set_frequency(void *h, long long frequency) { static_cast<Chain *>(h)->inst->setFrequency(frequency); }
void refsrc_demod_set_squelch(void *h, int enabled, float level_db) {
    Chain *c = static_cast<Chain *>(h);
    if (enabled) c->inst->setSquelchLevel(level_db);            // (enables it, DemodulatorThread.cpp:387-392)
    c->inst->setSquelchEnabled(enabled != 0);
}
int refsrc_demod_bandwidth(void *h) {                           // the rate the kit was built at (Modem::checkSampleRate's answer)
    Chain *c = static_cast<Chain *>(h);
    ModemKit *k = c->inst->demodulatorPreThread->getModemKit();
    return k ? (int)k->sampleRate : 0;
}
// wait until the demodulator has finished everything in its input pipe, then hand back what came out of its "AudioSink" tap.
// The sentinels carry the rate and tuner frequency of the blocks the demodulator has been getting (a different rate would
// make DemodulatorPreThread order new filters, :86-91).
static int collect(Chain *c, float *audio, unsigned cap, unsigned *n_audio, int *channels, float *peak, int *squelched, float *meter) {
    auto in = c->inst->getIQInputDataPipe();
    DemodulatorPreThread *p = c->inst->demodulatorPreThread;
    const long long rate = p->currentSampleRate ? (long long)p->currentSampleRate : c->rate;
    const long long center = c->last_center;
    in->push(make_block(nullptr, 0, rate, center));             // sentinel: popped only once the blocks before it are done
    if (!wait_empty(in, 20000)) return -1;
    c->post->push(std::make_shared<DemodulatorThreadPostIQData>());   // sentinel for DemodulatorThread (no samples: skipped)
    if (!wait_empty(c->post, 20000)) return -1;
    c->inst->pipeAudioData->flush();                            // nobody plays it here (no audio device)
    *n_audio = 0;
    AudioThreadInputPtr a;
    int got = 1;
    while (c->sink->try_pop(a)) {
        if (!a) continue;
        const unsigned m = (unsigned)a->data.size();
        if (m) std::memcpy(audio, a->data.data(), (size_t)(m < cap ? m : cap) * sizeof(float));
        *n_audio = m;
        if (channels) *channels = a->channels;
        if (peak) *peak = a->peak;
        if (squelched) *squelched = a->is_squelch_active ? 1 : 0;
        got = 0;
    }
    meter[0] = c->inst->getSignalLevel();
    meter[1] = c->inst->getSignalFloor();
    meter[2] = c->inst->getSignalCeil();
    return got;
}
// one wideband block through the chain.  Returns 0 when an AudioThreadInput came out (audio: up to cap floats, *n_audio its
// size), 1 when the block produced none (the 0.75 fs skip rule, DemodulatorPreThread.cpp:161-164), -1 on a time-out.
// meter = {signalLevel, signalFloor, signalCeil} after the block.
int refsrc_demod_process(void *h, const float *iq, unsigned n, long long center, float *audio, unsigned cap, unsigned *n_audio,
                         int *channels, float *peak, int *squelched, float *meter) {
    Chain *c = static_cast<Chain *>(h);
    c->last_center = center;
    c->inst->getIQInputDataPipe()->push(make_block(iq, n, c->rate, center));
    return collect(c, audio, cap, n_audio, channels, peak, squelched, meter);
}
// the same without feeding a block: for a demodulator that the application-level driver below feeds through SDRPostThread
// (`center` = the frequency stamped on the blocks it gets: the tuner frequency, or its filterbank channel's centre)
int refsrc_demod_collect(void *h, long long center, float *audio, unsigned cap, unsigned *n_audio, int *channels, float *peak,
                         int *squelched, float *meter) {
    Chain *c = static_cast<Chain *>(h);
    c->last_center = center;
    return collect(c, audio, cap, n_audio, channels, peak, squelched, meter);
}

// ---- the ingress: a real SDRPostThread on its own thread, feeding the demodulators registered with it -------------------------
struct App {
    SDRPostThread post;
    std::thread *t = nullptr;
    SDRThreadIQDataQueuePtr in;
    DemodulatorThreadInputQueuePtr out, vis, active_vis;
};
// (SDRPostThread.cpp finds the demodulators through wxGetApp().getDemodMgr(): one application per process)
void *refsrc_app_create(int channelizer_type /* 1: firpfbch, 2: firpfbch2; used when a block says numChannels > 1 */) {
    App *a = new App;
    a->in = std::make_shared<SDRThreadIQDataQueue>(); a->in->set_max_num_items(64);
    a->out = std::make_shared<DemodulatorThreadInputQueue>(); a->out->set_max_num_items(1u << 16);
    a->vis = std::make_shared<DemodulatorThreadInputQueue>(); a->vis->set_max_num_items(1);
    a->active_vis = std::make_shared<DemodulatorThreadInputQueue>(); a->active_vis->set_max_num_items(1);
    a->post.setInputQueue("IQDataInput", a->in);                // src/CubicSDR.cpp:372-386 wires these
    a->post.setOutputQueue("IQDataOutput", a->out);
    a->post.setOutputQueue("IQVisualDataOutput", a->vis);
    a->post.setOutputQueue("IQActiveDemodVisualDataOutput", a->active_vis);
    a->post.setChannelizerType((SDRPostThreadChannelizerType)channelizer_type);
    wxGetApp().getDemodMgr().demods.clear();
    wxGetApp().getDemodMgr().current = nullptr;
    a->t = new std::thread(&SDRPostThread::threadMain, &a->post);
    return a;
}
// DemodulatorMgr::newThread + CubicSDRApp::notifyDemodulatorsChanged: the demodulator joins the manager's list
void refsrc_app_add_demod(void *ah, void *h) {
    App *a = static_cast<App *>(ah);
    Chain *c = static_cast<Chain *>(h);
    wxGetApp().getDemodMgr().demods.push_back(DemodulatorInstancePtr(c->inst, [](DemodulatorInstance *) {}));   // (owned by its Chain)
    a->post.notifyDemodulatorsChanged();
}
// one block from the device thread (SDRThreadIQData: frequency, sampleRate, numChannels, data).  Returns the number of samples
// of the block SDRPostThread put on "IQDataOutput" (single channel: the DC-blocked block; channelizer modes: the block as it
// came), copied to out; -1 on a time-out.
int refsrc_app_block(void *ah, const float *iq, unsigned n, long long center, long long rate, int num_channels, float *out, unsigned cap) {
    App *a = static_cast<App *>(ah);
    wxGetApp().setFrequency(center);
    auto b = std::make_shared<SDRThreadIQData>();
    b->frequency = center; b->sampleRate = rate; b->numChannels = num_channels; b->dcCorrected = true;
    b->data.resize(n);
    if (n) std::memcpy(reinterpret_cast<void *>(b->data.data()), iq, (size_t)n * 2 * sizeof(float));
    a->in->push(b);
    auto s = std::make_shared<SDRThreadIQData>();               // sentinel: no samples (skipped, SDRPostThread.cpp:176)
    s->frequency = center; s->sampleRate = rate; s->numChannels = num_channels;
    a->in->push(s);
    if (!wait_empty(a->in, 20000)) return -1;
    a->in->push(std::make_shared<SDRThreadIQData>(*s));         // ... and a second one: the first is popped before it is done with
    if (!wait_empty(a->in, 20000)) return -1;
    a->vis->flush(); a->active_vis->flush();
    int m = 0;
    DemodulatorThreadIQDataPtr o;
    while (a->out->try_pop(o)) {
        if (!o) continue;
        m = (int)o->data.size();
        if (m) std::memcpy(out, o->data.data(), (size_t)((unsigned)m < cap ? (unsigned)m : cap) * 2 * sizeof(float));
    }
    return m;
}
// which filterbank channel SDRPostThread routed demodulator j of its run list to, its centre frequency and the channel rate
int refsrc_app_route(void *ah, void *h, long long *channel_center, int *chan_bw) {
    App *a = static_cast<App *>(ah);
    Chain *c = static_cast<Chain *>(h);
    for (size_t j = 0; j < a->post.runDemods.size(); ++j)
        if (a->post.runDemods[j].get() == c->inst) {
            const int ch = a->post.numChannels > 1 ? a->post.demodChannel[j] : 0;
            if (channel_center) *channel_center = (a->post.numChannels > 1 && ch >= 0) ? a->post.chanCenters[(size_t)ch] : a->post.frequency;
            if (chan_bw) *chan_bw = a->post.numChannels > 1 ? a->post.chanBw : (int)a->post.sampleRate;
            return ch;
        }
    return -1;                                                  // not in the run list (inactive / out of range)
}
void refsrc_app_destroy(void *ah) {
    App *a = static_cast<App *>(ah);
    a->post.terminate();
    a->t->join();
    delete a->t;
    wxGetApp().getDemodMgr().demods.clear();
    wxGetApp().getDemodMgr().current = nullptr;
    delete a;
}

// diagnostics: {sampleRateChanged, bandwidthChanged, audioSampleRateChanged, modemSettingsChanged, shouldRebuildKit, resampler id}
void refsrc_demod_debug(void *h, long long *out) {
    Chain *c = static_cast<Chain *>(h);
    DemodulatorPreThread *p = c->inst->demodulatorPreThread;
    out[0] = p->sampleRateChanged.load(); out[1] = p->bandwidthChanged.load(); out[2] = p->audioSampleRateChanged.load();
    out[3] = p->modemSettingsChanged.load(); out[4] = p->cModem ? (p->cModem->shouldRebuildKit() ? 1 : 0) : -1;
    out[5] = (long long)(size_t)p->iqResampler; out[6] = (long long)(size_t)p->cModemKit; out[7] = p->currentSampleRate;
}
void refsrc_demod_destroy(void *h) {
    Chain *c = static_cast<Chain *>(h);
    c->inst->demodulatorThread->setOutputQueue("AudioSink", nullptr);
    c->inst->terminate();
    const auto t0 = std::chrono::steady_clock::now();
    while (!c->inst->isTerminated()) {                          // joins and deletes the threads as they end (:201-270)
        std::this_thread::sleep_for(std::chrono::milliseconds(1));
        if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(20)) break;
    }
    delete c->inst;
    delete c;
}

}  // extern "C"
