// tests/cpp/host_adapter_test.cpp -- drives the C++ host adapters (cubicsdr_b200/host/cubicsdr_b200.hpp) the way
// CubicSDR's threads would, on inputs prepared by tests/test_host_adapters.py, and writes the outputs back for the
// Python side to compare with the oracle.  Reads like the reference's own call sites:
//   Modem::makeModem -> checkSampleRate -> buildKit -> demodulate            (DemodulatorWorkerThread.cpp:63-84, DemodulatorThread.cpp:140)
//   SDRPostThread::runSingleCH -> per-demod fetch, SpectrumVisualProcessor    (SDRPostThread.cpp:248-299)
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>

#include "../../cubicsdr_b200/host/cubicsdr_b200.hpp"

using namespace cubicsdr_b200;

template <typename T> static std::vector<T> slurp(const std::string &path) {
    std::ifstream f(path, std::ios::binary | std::ios::ate);
    if (!f) { std::cerr << "cannot open " << path << "\n"; std::exit(2); }
    size_t bytes = (size_t)f.tellg();
    std::vector<T> v(bytes / sizeof(T));
    f.seekg(0); f.read(reinterpret_cast<char *>(v.data()), bytes);
    return v;
}
template <typename T> static void dump(const std::string &path, const std::vector<T> &v) {
    std::ofstream f(path, std::ios::binary);
    f.write(reinterpret_cast<const char *>(v.data()), v.size() * sizeof(T));
}

int main(int argc, char **argv) {
    if (argc < 2) { std::cerr << "usage: host_adapter_test <dir>\n"; return 2; }
    const std::string dir = argv[1];
    auto eng = std::make_shared<Engine>(0, 1u << 17);
    registerModems(eng);

    // ---- S1: Modem plugin seam ---------------------------------------------------------------------------------------
    const char *names[5] = {"FM", "NBFM", "AM", "USB", "LSB"};
    for (int m = 0; m < 5; ++m) {
        Modem *modem = Modem::makeModem(names[m]);
        if (!modem) return 3;
        const int rate = modem->checkSampleRate(Modem::getModemDefaultSampleRate(names[m]), 48000);
        ModemKit *kit = modem->buildKit(rate, 48000);
        if (!kit) { std::cerr << "buildKit failed: " << sdr_last_error() << "\n"; return 4; }
        std::vector<float> all;
        for (int b = 0; b < 3; ++b) {
            ModemIQData in;
            in.sampleRate = rate;
            in.data = slurp<liquid_float_complex>(dir + "/iq_" + names[m] + "_" + std::to_string(b) + ".bin");
            AudioThreadInput ati;
            modem->demodulate(kit, &in, &ati);
            if (ati.channels != 1 || ati.sampleRate != 48000) return 5;
            all.insert(all.end(), ati.data.begin(), ati.data.end());
        }
        ModemIQData empty; AudioThreadInput untouched;
        modem->demodulate(kit, &empty, &untouched);            // empty input is a no-op, as in the reference
        if (!untouched.data.empty()) return 6;
        dump(dir + "/audio_modem_" + names[m] + ".bin", all);
        modem->disposeKit(kit);
        delete modem;
    }
    if (Modem::makeModem("QAM") != nullptr) return 7;           // not registered (a digital modem) -> nullptr, like Modem::makeModem
    { Modem *d = Modem::makeModem("DSB"); if (!d || d->getName() != "DSB") return 7; delete d; }
    { Modem *d = Modem::makeModem("FMS"); if (!d || d->getName() != "FMS" || d->checkSampleRate(50000, 48000) != 100000) return 8; delete d; }

    // ---- S3 + S2: block ingress, pre-demod + demod batched, spectrum ----------------------------------------------------
    SDRPostB200 post(eng);
    DemodulatorB200 fm(eng, "FM", 2400000, 200000, 48000, true);
    fm.setFrequency(100000000 + 300000);
    SpectrumVisualProcessorB200 spec(eng, 256);
    std::vector<float> audio, pts;
    std::vector<liquid_float_complex> postiq;
    for (int b = 0; b < 2; ++b) {
        DemodulatorThreadIQData blk;
        blk.frequency = 100000000; blk.sampleRate = 2400000;
        blk.data = slurp<liquid_float_complex>(dir + "/wide_" + std::to_string(b) + ".bin");
        if (!post.runSingleCH(blk)) { std::cerr << sdr_last_error() << "\n"; return 8; }
        AudioThreadInput ati; double lvl = 0;
        if (!fm.fetchAudio(ati, &lvl)) return 9;
        audio.insert(audio.end(), ati.data.begin(), ati.data.end());
        DemodulatorThreadPostIQData p;
        if (!fm.fetchPostIQ(p) || p.sampleRate != 200000) return 10;
        postiq.insert(postiq.end(), p.data.begin(), p.data.end());
        for (unsigned f = 0; f < spec.framesReady(); ++f) {
            SpectrumVisualData d;
            if (!spec.getFrame(f, d, 100000000, 2400000)) return 11;
            if (d.spectrum_points.size() != 512) return 12;
            if (b == 1 && f + 1 == spec.framesReady()) for (unsigned i = 0; i < 256; ++i) pts.push_back(d.spectrum_points[2 * i + 1]);
        }
    }
    // ---- pipelined ingress: the block just submitted stays on the GPU while the previous one is fetched; the audio of
    // blocks 0,1 re-run through a fresh channel must equal what the blocking loop produced (block 1 is drained at the end)
    {
        auto eng2 = std::make_shared<Engine>(0, 1u << 17);       // fresh engine: the DC blocker starts from rest, as above
        SDRPostB200 post2(eng2);
        DemodulatorB200 fm2(eng2, "FM", 2400000, 200000, 48000, false);
        fm2.setFrequency(100000000 + 300000);
        std::vector<float> audio2;
        for (int b = 0; b < 2; ++b) {
            DemodulatorThreadIQData blk;
            blk.frequency = 100000000; blk.sampleRate = 2400000;
            blk.data = slurp<liquid_float_complex>(dir + "/wide_" + std::to_string(b) + ".bin");
            bool have = false;
            if (!post2.runSingleCHPipelined(blk, &have)) { std::cerr << sdr_last_error() << "\n"; return 13; }
            if (have != (b > 0)) return 14;
            if (have) { AudioThreadInput a; if (!fm2.fetchAudio(a, nullptr)) return 15; audio2.insert(audio2.end(), a.data.begin(), a.data.end()); }
        }
        if (sdr_engine_wait_block(eng2->get(), 0) != SDR_OK) return 16;
        { AudioThreadInput a; if (!fm2.fetchAudio(a, nullptr)) return 17; audio2.insert(audio2.end(), a.data.begin(), a.data.end()); }
        dump(dir + "/audio_chan_FM_pipelined.bin", audio2);
    }
    // ---- channelizer mode (SDRPostThread::setChannelizerType(SDRPostPFBCH) -> runPFBCH): the demodulator is created with
    // the channel rate and keeps its absolute frequency; the engine routes it to its filterbank channel
    {
        auto eng3 = std::make_shared<Engine>(0, 1u << 17);
        SDRPostB200 post3(eng3);
        if (!post3.setChannelizerType(1)) return 18;
        int numChannels = 0; long long chanBw = 0;
        SDRPostB200::channelPlan(2400000, numChannels, chanBw);
        if (numChannels != 4 || chanBw != 600000) return 19;
        DemodulatorB200 fm3(eng3, "FM", chanBw, 200000, 48000, false);
        fm3.setFrequency(100000000 + 300000);
        std::vector<float> audio3;
        for (int b = 0; b < 2; ++b) {
            DemodulatorThreadIQData blk;
            blk.frequency = 100000000; blk.sampleRate = 2400000;
            blk.data = slurp<liquid_float_complex>(dir + "/wide_" + std::to_string(b) + ".bin");
            if (!post3.runSingleCH(blk)) { std::cerr << sdr_last_error() << "\n"; return 20; }   // the engine dispatches by mode
            AudioThreadInput a;
            if (!fm3.fetchAudio(a, nullptr)) return 21;
            audio3.insert(audio3.end(), a.data.begin(), a.data.end());
        }
        dump(dir + "/audio_chan_FM_channelizer.bin", audio3);
    }
    dump(dir + "/audio_chan_FM.bin", audio);
    dump(dir + "/postiq_chan_FM.bin", postiq);
    dump(dir + "/spec_last_points.bin", pts);
    std::cout << "host adapters ok\n";
    return 0;
}
