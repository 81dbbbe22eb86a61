// oracle/ref_audio.cpp -- TEST INFRASTRUCTURE ONLY.
//
// C entry points over the reference's OWN, UNMODIFIED audio mixer: the RtAudio callback `audioCallback` of
// src/audio/AudioThread.cpp:88-243 working on real AudioThread objects (a controller and the sources bound to it, each with its
// inputQueue / currentInput / audioQueuePtr / gain / active state).  The callback is a file-static function, so this driver
// takes the reference source into its own translation unit with one #include -- compiled from where it lies, nothing copied
// into this repository -- and oracle/build_ref.py compiles it with -fno-access-control so that the driver can set the
// controller's private sampleRate without opening an audio device (setSampleRate of a controller opens the stream).
// RtAudio is the reference's own external/rtaudio/RtAudio.cpp built with no audio API (its dummy API): no device is touched,
// the callback is invoked by hand exactly as RtAudio would (outputBuffer, nBufferFrames, status 0, userData = controller).
// tests/test_oracle_demod.py pins the restated mixer of oracle/ref_chain.cpp (what the device mixer is compared against)
// to it: same pushes and callbacks in, the same bits out.
#include "AudioThread.cpp"                // /root/reference/src/audio, through the include path
#include "AudioFileWAV.h"

namespace {
struct Mixer {
    AudioThread controller;
    std::vector<AudioThread *> sources;
    int device_id = 0;
};
int g_next_device = 1000;
}  // namespace

// (wxGetApp() of the shadowed application header is defined in oracle/ref_demod.cpp, same library: libdemod_ref.so)

extern "C" {

// the controller thread of one output device, registered where AudioThread::setActive looks for it (deviceController, keyed by
// device id; setupDevice does this for a real device, :383-431)
void *refsrc_mixer_create(int sample_rate) {
    Mixer *m = new Mixer;
    m->device_id = g_next_device++;
    m->controller.sampleRate = sample_rate;          // what setSampleRate leaves behind, without the stream
    AudioThread::deviceController[m->device_id] = &m->controller;
    return m;
}
// a demodulator's AudioThread on that device: the inputQueue DemodulatorInstance gives it, its gain, then setActive(true), which
// binds it to the controller (AudioThread.cpp:577-579)
int refsrc_mixer_add_source(void *h, float gain) {
    Mixer *m = static_cast<Mixer *>(h);
    AudioThread *s = new AudioThread;
    s->parameters.deviceId = (unsigned)m->device_id;
    s->inputQueue = std::make_shared<AudioThreadInputQueue>();
    s->inputQueue->set_max_num_items(1u << 20);
    s->setGain(gain);
    s->setActive(true);
    m->sources.push_back(s);
    return (int)m->sources.size() - 1;
}
void refsrc_mixer_set_gain(void *h, int i, float gain) { static_cast<Mixer *>(h)->sources[(size_t)i]->setGain(gain); }
void refsrc_mixer_set_active(void *h, int i, int active) { static_cast<Mixer *>(h)->sources[(size_t)i]->setActive(active != 0); }
void refsrc_mixer_push(void *h, int i, const float *data, unsigned n, int channels, float peak, int sample_rate) {
    auto a = std::make_shared<AudioThreadInput>();
    a->channels = channels;
    a->peak = peak;
    a->sampleRate = sample_rate;
    a->data.assign(data, data + n);
    static_cast<Mixer *>(h)->sources[(size_t)i]->inputQueue->push(a);
}
int refsrc_mixer_callback(void *h, float *out, unsigned n_frames) {
    return audioCallback(out, nullptr, n_frames, 0.0, 0, &static_cast<Mixer *>(h)->controller);
}
// ---- AudioFileWAV (src/audio/AudioFileWAV.cpp, AudioFile.cpp: compiled unmodified next to this file) --------------------------
// a recording in directory `dir` named `name`: every write is AudioFileWAV::writeToFile(input) (header on the first call, then
// the 16-bit payload, :122-165); close fixes the chunk sizes (:102-120).  The test reads dir/name.wav back.
void *refsrc_wav_open(const char *dir, const char *name) {
    wxGetApp().getConfig()->recordingPath = dir;
    AudioFileWAV *w = new AudioFileWAV;
    w->setOutputFileName(name);
    return w;
}
int refsrc_wav_write(void *h, const float *data, unsigned n, int channels, int sample_rate, float peak) {
    auto a = std::make_shared<AudioThreadInput>();
    a->channels = channels;
    a->sampleRate = sample_rate;
    a->peak = peak;
    a->data.assign(data, data + n);
    return static_cast<AudioFileWAV *>(h)->writeToFile(a) ? 0 : -1;
}
void refsrc_wav_close(void *h) {
    AudioFileWAV *w = static_cast<AudioFileWAV *>(h);
    w->closeFile();
    delete w;
}

void refsrc_mixer_destroy(void *h) {
    Mixer *m = static_cast<Mixer *>(h);
    for (AudioThread *s : m->sources) { m->controller.removeThread(s); delete s; }
    AudioThread::deviceController.erase(m->device_id);
    delete m;
}

}  // extern "C"
