"""oracle/refbind.py -- ctypes binding of oracle/_ref/libsdr_ref.so.  TEST INFRASTRUCTURE ONLY.

The library is the reference's own liquid-dsp 1.5.0 binary plus the CubicSDR glue restated in oracle/ref_chain.cpp
(see oracle/build_ref.py), and of oracle/_ref/libmodems_ref.so, the reference's own unmodified modem classes over that
binary (RefModems / RefModem at the end), and of oracle/_ref/libvisuals_ref.so, its unmodified FFTDataDistributor /
SpectrumVisualProcessor / ScopeVisualProcessor / audioCallback / AudioFileWAV (RefVisuals / SrcDistrib / SrcSpectrum / SrcScope /
SrcMixer / SrcWav), and of oracle/_ref/libdemod_ref.so, its whole DemodulatorInstance (RefDemodLib / SrcDemod).  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_LIB = os.path.join(_HERE, "_ref", "libsdr_ref.so")


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class Ref:
    def __init__(self, path=REF_LIB):
        if not os.path.exists(path):
            import subprocess, sys
            subprocess.check_call([sys.executable, os.path.join(_HERE, "build_ref.py")])
        L = self.L = C.CDLL(path)
        vp, u, i, ll, f, d = C.c_void_p, C.c_uint, C.c_int, C.c_longlong, C.c_float, C.c_double
        L.ref_libversion.restype = C.c_char_p
        for name in ("ref_nco_create", "ref_ingress_create"):
            getattr(L, name).restype = vp
        L.ref_estimate_req_filter_len.restype = u; L.ref_estimate_req_filter_len.argtypes = [f, f]
        L.ref_kaiser_beta_As.restype = f; L.ref_kaiser_beta_As.argtypes = [f]
        L.ref_firdespm_halfband_as.argtypes = [u, f, vp]
        L.ref_firdes_kaiser.argtypes = [u, f, f, f, vp]
        L.ref_nco_destroy.argtypes = [vp]
        L.ref_nco_set_frequency.argtypes = [vp, f]
        L.ref_nco_raw_dtheta.restype = C.c_uint32; L.ref_nco_raw_dtheta.argtypes = [vp]
        L.ref_nco_raw_theta.restype = C.c_uint32; L.ref_nco_raw_theta.argtypes = [vp]
        L.ref_nco_sintab.argtypes = [vp, vp]
        L.ref_nco_mix_block.argtypes = [vp, i, vp, vp, u]
        L.ref_resamp_crcf_create.restype = vp; L.ref_resamp_crcf_create.argtypes = [f, u, f, f, u]
        L.ref_resamp_crcf_destroy.argtypes = [vp]
        L.ref_resamp_crcf_execute_block.restype = u; L.ref_resamp_crcf_execute_block.argtypes = [vp, vp, u, vp]
        L.ref_msresamp_crcf_create.restype = vp; L.ref_msresamp_crcf_create.argtypes = [f, f]
        L.ref_msresamp_crcf_destroy.argtypes = [vp]
        L.ref_msresamp_crcf_execute.restype = u; L.ref_msresamp_crcf_execute.argtypes = [vp, vp, u, vp]
        L.ref_msresamp_rrrf_create.restype = vp; L.ref_msresamp_rrrf_create.argtypes = [f, f]
        L.ref_msresamp_rrrf_destroy.argtypes = [vp]
        L.ref_msresamp_rrrf_execute.restype = u; L.ref_msresamp_rrrf_execute.argtypes = [vp, vp, u, vp]
        L.ref_firfilt_dc_blocker_create.restype = vp; L.ref_firfilt_dc_blocker_create.argtypes = [u, f]
        L.ref_firfilt_destroy.argtypes = [vp]
        L.ref_firfilt_block.argtypes = [vp, vp, u, vp]
        L.ref_iirfilt_lowpass_create.restype = vp; L.ref_iirfilt_lowpass_create.argtypes = [u, f]
        L.ref_iirfilt_destroy.argtypes = [vp]
        L.ref_iirfilt_block.argtypes = [vp, vp, u, vp]
        L.ref_firhilb_create.restype = vp; L.ref_firhilb_create.argtypes = [u, f]
        L.ref_firhilb_destroy.argtypes = [vp]
        L.ref_firhilb_c2r_block.argtypes = [vp, vp, u, vp, vp]
        L.ref_fft_forward.argtypes = [vp, vp, u]
        L.ref_ingress_destroy.argtypes = [vp]
        L.ref_ingress_block.argtypes = [vp, vp, u, vp]
        L.ref_modem_check_sample_rate.argtypes = [C.c_char_p, ll]
        L.ref_chan_create.restype = vp; L.ref_chan_create.argtypes = [C.c_char_p, ll, ll, i]
        L.ref_chan_destroy.argtypes = [vp]
        L.ref_chan_bandwidth.argtypes = [vp]
        L.ref_chan_meter.argtypes = [vp, vp, vp]; L.ref_chan_meter.restype = None
        L.ref_chan_set_squelch.argtypes = [vp, C.c_int, C.c_float]; L.ref_chan_set_squelch.restype = None
        L.ref_chan_process.argtypes = [vp, vp, u, ll, vp, vp, vp, vp, vp, vp, vp]
        L.ref_distrib_create.restype = vp; L.ref_distrib_create.argtypes = [u, u]
        L.ref_distrib_destroy.argtypes = [vp]
        L.ref_distrib_push.restype = u; L.ref_distrib_push.argtypes = [vp, vp, u, ll, ll, vp, u]
        L.ref_spectrum_create.restype = vp; L.ref_spectrum_create.argtypes = [u]
        L.ref_spectrum_destroy.argtypes = [vp]
        L.ref_spectrum_set_average_rate.argtypes = [vp, f]
        L.ref_spectrum_frame.argtypes = [vp, vp, u, vp, vp, vp, vp, vp]
        L.ref_bench_path.restype = d
        L.ref_bench_path.argtypes = [vp, u, u, u, ll, u, vp, vp, vp, i, u, u, vp]

    def version(self):
        return self.L.ref_libversion().decode()


class RefPfb:
    """SDRPostThread::runPFBCH channelizer: the reference binary's firpfbch_crcf analyzer (SDRPostThread.cpp:401-455)."""

    def __init__(self, ref, num_channels, m=4, as_db=60.0):
        self.ref, self.M = ref, num_channels
        ref.L.ref_pfb_create.restype = C.c_void_p
        ref.L.ref_pfb_create.argtypes = [C.c_uint, C.c_uint, C.c_float]
        ref.L.ref_pfb_execute.argtypes = [C.c_void_p, C.c_void_p, C.c_uint, C.c_void_p]
        ref.L.ref_pfb_destroy.argtypes = [C.c_void_p]
        self.h = C.c_void_p(ref.L.ref_pfb_create(num_channels, m, as_db))

    def execute(self, x):
        x = np.ascontiguousarray(x, np.complex64)
        assert x.size % self.M == 0
        y = np.zeros(x.size, np.complex64)
        self.ref.L.ref_pfb_execute(self.h, _p(x), x.size, _p(y))
        return y.reshape(-1, self.M)

    def close(self):
        self.ref.L.ref_pfb_destroy(self.h)


class RefPfb2:
    """SDRPostThread::runPFBCH2 channelizer: the reference binary's firpfbch2_crcf analyzer (SDRPostThread.cpp:457-511)."""

    def __init__(self, ref, num_channels, m=4, as_db=60.0):
        self.ref, self.M = ref, num_channels
        ref.L.ref_pfb2_create.restype = C.c_void_p
        ref.L.ref_pfb2_create.argtypes = [C.c_uint, C.c_uint, C.c_float]
        ref.L.ref_pfb2_execute.argtypes = [C.c_void_p, C.c_void_p, C.c_uint, C.c_void_p]
        ref.L.ref_pfb2_destroy.argtypes = [C.c_void_p]
        self.h = C.c_void_p(ref.L.ref_pfb2_create(num_channels, m, as_db))

    def execute(self, x):
        x = np.ascontiguousarray(x, np.complex64)
        assert x.size % (self.M // 2) == 0
        y = np.zeros(2 * x.size, np.complex64)
        self.ref.L.ref_pfb2_execute(self.h, _p(x), x.size, _p(y))
        return y.reshape(-1, self.M)

    def close(self):
        self.ref.L.ref_pfb2_destroy(self.h)


class RefIngress:
    """SDRPostThread::runSingleCH DC blocker (src/sdr/SDRPostThread.cpp:29,284)."""

    def __init__(self, ref):
        self.ref = ref
        self.h = C.c_void_p(ref.L.ref_ingress_create())

    def block(self, x):
        x = np.ascontiguousarray(x, np.complex64)
        y = np.zeros_like(x)
        self.ref.L.ref_ingress_block(self.h, _p(x), x.size, _p(y))
        return y

    def close(self):
        self.ref.L.ref_ingress_destroy(self.h)


class RefChan:
    """DemodulatorPreThread + Modem + DemodulatorThread metering for one channel (see ref_chain.cpp)."""

    def __init__(self, ref, modem, sample_rate, bandwidth, audio_rate=48000):
        self.ref = ref
        self.h = C.c_void_p(ref.L.ref_chan_create(modem.encode(), int(sample_rate), int(bandwidth), int(audio_rate)))
        assert self.h.value
        self.bandwidth = ref.L.ref_chan_bandwidth(self.h)
        self.sample_rate, self.audio_rate = sample_rate, audio_rate

    def process(self, block, shift):
        block = np.ascontiguousarray(block, np.complex64)
        n = block.size
        cap_iq = int(n * self.bandwidth / self.sample_rate) + 1024
        cap_a = 2 * int(n * self.audio_rate / self.sample_rate) + 4096      # "I/Q" yields two floats per decimated sample
        iq = np.zeros(cap_iq, np.complex64); dm = np.zeros(cap_iq, np.float32); au = np.zeros(cap_a, np.float32)
        niq, na, lvl, pk = C.c_uint(), C.c_uint(), C.c_double(), C.c_float()
        skipped = self.ref.L.ref_chan_process(self.h, _p(block), n, int(shift), _p(iq), C.byref(niq), _p(dm), _p(au), C.byref(na),
                                              C.byref(lvl), C.byref(pk))
        m3 = (C.c_float * 3)(); sq = C.c_int()
        self.ref.L.ref_chan_meter(self.h, m3, C.byref(sq))
        return dict(skipped=bool(skipped), iq=iq[:niq.value].copy(), demod=dm[:niq.value].copy(), audio=au[:na.value].copy(),
                    level_db=lvl.value, peak=pk.value, signal_level=m3[0], signal_floor=m3[1], signal_ceil=m3[2], squelched=bool(sq.value))

    def set_squelch(self, enabled, level_db):
        """DemodulatorThread::setSquelchEnabled / setSquelchLevel."""
        self.ref.L.ref_chan_set_squelch(self.h, 1 if enabled else 0, C.c_float(level_db))

    def close(self):
        self.ref.L.ref_chan_destroy(self.h)


class RefDistrib:
    """FFTDataDistributor (src/process/FFTDataDistributor.cpp:28-144)."""

    def __init__(self, ref, frame_len, lps):
        self.ref, self.n = ref, frame_len
        self.h = C.c_void_p(ref.L.ref_distrib_create(frame_len, lps))

    def push(self, block, sample_rate, frequency=0):
        block = np.ascontiguousarray(block, np.complex64)
        cap = (block.size + (1 << 20)) // self.n + 4             # the block's frames + those in what is still buffered from before
        out = np.zeros(cap * self.n, np.complex64)
        nf = self.ref.L.ref_distrib_push(self.h, _p(block), block.size, int(sample_rate), int(frequency), _p(out), cap)
        return out[:nf * self.n].reshape(nf, self.n)

    def set_frame_len(self, n):
        """FFTDataDistributor::setFFTSize: only the frame length changes; buffered samples and the line accumulator stay."""
        self.ref.L.ref_distrib_set_fft_size(self.h, int(n))
        self.n = int(n)

    def close(self):
        self.ref.L.ref_distrib_destroy(self.h)


class RefSpectrum:
    """SpectrumVisualProcessor non-view branch + WaterfallPanel::step."""

    def __init__(self, ref, fft_size):
        self.ref, self.fft_size = ref, fft_size
        self.h = C.c_void_p(ref.L.ref_spectrum_create(fft_size))

    def set_average_rate(self, r):
        self.ref.L.ref_spectrum_set_average_rate(self.h, C.c_float(r))

    def frame(self, frame):
        frame = np.ascontiguousarray(frame, np.complex64)
        N = 2 * self.fft_size
        pts = np.zeros(self.fft_size, np.float32); line = np.zeros(self.fft_size, np.uint8)
        res = np.zeros(N, np.float64); maa = np.zeros(N, np.float64); cf = np.zeros(2, np.float64)
        self.ref.L.ref_spectrum_frame(self.h, _p(frame), frame.size, _p(pts), _p(line), _p(res), _p(maa), _p(cf))
        return dict(points=pts, line=line, fft_result=res, maa=maa, ceil_floor=cf)

    def set_view(self, is_view, center_freq, bandwidth):
        """SpectrumVisualProcessor::setView(bView, centerFreq, bandwidth)."""
        self.ref.L.ref_spectrum_set_view.argtypes = [C.c_void_p, C.c_int, C.c_longlong, C.c_long]
        self.ref.L.ref_spectrum_set_view(self.h, 1 if is_view else 0, int(center_freq), int(bandwidth))

    def desired_input(self, sample_rate):
        self.ref.L.ref_spectrum_desired_input.argtypes = [C.c_void_p, C.c_longlong]
        self.ref.L.ref_spectrum_desired_input.restype = C.c_uint
        return int(self.ref.L.ref_spectrum_desired_input(self.h, int(sample_rate)))

    def view_frame(self, frame, frequency, sample_rate):
        """One process() call of the view branch on the frame the distributor handed over; None if no transform ran."""
        frame = np.ascontiguousarray(frame, np.complex64)
        N = 2 * self.fft_size
        pts = np.zeros(self.fft_size, np.float32); line = np.zeros(self.fft_size, np.uint8)
        res = np.zeros(N, np.float64); maa = np.zeros(N, np.float64); cf = np.zeros(2, np.float64)
        self.ref.L.ref_spectrum_view_frame.argtypes = [C.c_void_p, C.c_void_p, C.c_uint, C.c_longlong, C.c_longlong] + [C.c_void_p] * 5
        ok = self.ref.L.ref_spectrum_view_frame(self.h, _p(frame), frame.size, int(frequency), int(sample_rate), _p(pts), _p(line), _p(res), _p(maa), _p(cf))
        return dict(points=pts, line=line, fft_result=res, maa=maa, ceil_floor=cf) if ok else None

    def close(self):
        self.ref.L.ref_spectrum_destroy(self.h)


class RefMixer:
    """audioCallback (src/audio/AudioThread.cpp:88-243) restated in oracle/ref_chain.cpp."""

    def __init__(self, ref, sample_rate=48000):
        L = self.L = ref.L
        L.ref_mixer_create.restype = C.c_void_p; L.ref_mixer_create.argtypes = [C.c_int]
        L.ref_mixer_destroy.argtypes = [C.c_void_p]
        L.ref_mixer_add_source.argtypes = [C.c_void_p, C.c_float]
        L.ref_mixer_set_gain.argtypes = [C.c_void_p, C.c_int, C.c_float]
        L.ref_mixer_push.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_uint, C.c_int, C.c_float, C.c_int]
        L.ref_mixer_callback.argtypes = [C.c_void_p, C.c_void_p, C.c_uint]
        self.h = C.c_void_p(L.ref_mixer_create(int(sample_rate)))

    def add_source(self, gain=1.0):
        return self.L.ref_mixer_add_source(self.h, float(gain))

    def set_gain(self, i, gain):
        self.L.ref_mixer_set_gain(self.h, int(i), float(gain))

    def set_active(self, i, active):
        self.L.ref_mixer_set_active.argtypes = [C.c_void_p, C.c_int, C.c_int]
        self.L.ref_mixer_set_active(self.h, int(i), 1 if active else 0)

    def push(self, i, data, channels=1, peak=0.0, sample_rate=48000):
        d = np.ascontiguousarray(data, np.float32)
        self.L.ref_mixer_push(self.h, int(i), _p(d), d.size, int(channels), float(peak), int(sample_rate))

    def callback(self, n_frames):
        out = np.zeros(2 * n_frames, np.float32)
        self.L.ref_mixer_callback(self.h, _p(out), int(n_frames))
        return out

    def close(self):
        self.L.ref_mixer_destroy(self.h)


class RefScope:
    """ScopeVisualProcessor (src/process/ScopeVisualProcessor.cpp:14-217) restated in oracle/ref_chain.cpp."""

    def __init__(self, ref, fft_size=1024, max_samples=1024):
        L = self.L = ref.L
        L.ref_scope_create.restype = C.c_void_p; L.ref_scope_create.argtypes = [C.c_uint, C.c_uint]
        L.ref_scope_destroy.argtypes = [C.c_void_p]
        L.ref_scope_process.restype = C.c_uint
        L.ref_scope_process.argtypes = [C.c_void_p, C.c_void_p, C.c_uint, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        self.fft_size, self.max_samples = fft_size, max_samples
        self.h = C.c_void_p(L.ref_scope_create(fft_size, max_samples))

    def process(self, data, channels, sample_rate, input_rate):
        # with sampleRate > inputRate the reference reads past the end of its averages (ScopeVisualProcessor.cpp:194-203:
        # outSize = fftSize / 2 * sampleRate / inputRate points out of fftSize / 2 bins); the restatement would do the same
        assert sample_rate <= input_rate, "the reference's scope is undefined for an audio rate above the modem rate"
        d = np.ascontiguousarray(data, np.float32)
        wave = np.zeros(2 * max(self.max_samples, 1), np.float32); spec = np.zeros(self.fft_size, np.float32)
        ns = C.c_uint(); fc = np.zeros(2, np.float64)
        nw = self.L.ref_scope_process(self.h, _p(d), d.size, int(channels), int(sample_rate), int(input_rate), _p(wave), _p(spec), C.byref(ns), _p(fc))
        return wave[:2 * nw].copy(), spec[:2 * ns.value].copy(), fc

    def close(self):
        self.L.ref_scope_destroy(self.h)


def ref_wav_payload(ref, data, channels, peak):
    """AudioFileWAV::writePayloadToFileStream's 16-bit words for data[0, len)."""
    d = np.ascontiguousarray(data, np.float32)
    out = np.zeros(d.size + 2, np.int16)
    ref.L.ref_wav_payload.restype = C.c_uint
    ref.L.ref_wav_payload.argtypes = [C.c_void_p, C.c_uint, C.c_uint, C.c_int, C.c_float, C.c_void_p]
    n = ref.L.ref_wav_payload(_p(d), 0, d.size, int(channels), float(peak), _p(out))
    return out[:n].copy()


MODEMS_LIB = os.path.join(_HERE, "_ref", "libmodems_ref.so")


class RefModems:
    """oracle/_ref/libmodems_ref.so: the reference's own, unmodified Modem classes (oracle/ref_modems.cpp)."""

    def __init__(self, path=MODEMS_LIB):
        if not os.path.exists(path):
            import subprocess, sys
            subprocess.check_call([sys.executable, os.path.join(_HERE, "build_ref.py")])
        L = self.L = C.CDLL(path)
        vp, u, i, ll = C.c_void_p, C.c_uint, C.c_int, C.c_longlong
        L.refmodem_create.restype = vp; L.refmodem_create.argtypes = [C.c_char_p]
        L.refmodem_default_rate.argtypes = [C.c_char_p]
        L.refmodem_type.argtypes = [vp, C.c_char_p, i]
        L.refmodem_check_sample_rate.argtypes = [vp, ll, i]
        L.refmodem_write_setting.argtypes = [vp, C.c_char_p, C.c_char_p]; L.refmodem_write_setting.restype = None
        L.refmodem_build_kit.argtypes = [vp, ll, i]
        L.refmodem_should_rebuild_kit.argtypes = [vp]
        L.refmodem_use_signal_output.argtypes = [vp]
        L.refmodem_demodulate.argtypes = [vp, vp, u, ll, vp, u, vp, vp]
        L.refmodem_destroy.argtypes = [vp]; L.refmodem_destroy.restype = None

    def default_rate(self, name):
        return self.L.refmodem_default_rate(name.encode())


class RefModem:
    """One Modem made by Modem::makeModem(name) with its kit (Modem.h:127-166)."""

    def __init__(self, mods, name, sample_rate, audio_rate=48000):
        self.L = mods.L
        self.h = C.c_void_p(self.L.refmodem_create(name.encode()))
        assert self.h.value, "unknown modem %r" % name
        # DemodulatorWorkerThread.cpp:84-101: the kit is built at the rate checkSampleRate returns
        self.sample_rate = self.L.refmodem_check_sample_rate(self.h, int(sample_rate), int(audio_rate))
        self.audio_rate = audio_rate
        assert self.L.refmodem_build_kit(self.h, self.sample_rate, int(audio_rate)) == 0

    @property
    def type(self):
        buf = C.create_string_buffer(32)
        self.L.refmodem_type(self.h, buf, 32)
        return buf.value.decode()

    @property
    def use_signal_output(self):
        return bool(self.L.refmodem_use_signal_output(self.h))

    def write_setting(self, key, value):
        self.L.refmodem_write_setting(self.h, key.encode(), str(value).encode())
        if self.L.refmodem_should_rebuild_kit(self.h):
            assert self.L.refmodem_build_kit(self.h, self.sample_rate, int(self.audio_rate)) == 0

    def demodulate(self, iq):
        iq = np.ascontiguousarray(iq, np.complex64)
        cap = 4 * iq.size + 65536
        au = np.zeros(cap, np.float32)
        ch, rate = C.c_int(), C.c_int()
        n = self.L.refmodem_demodulate(self.h, _p(iq), iq.size, self.sample_rate, _p(au), cap, C.byref(ch), C.byref(rate))
        assert 0 <= n <= cap
        return au[:n].copy(), ch.value, rate.value

    def close(self):
        self.L.refmodem_destroy(self.h)


VISUALS_LIB = os.path.join(_HERE, "_ref", "libvisuals_ref.so")


class RefVisuals:
    """oracle/_ref/libvisuals_ref.so: the reference's own, unmodified FFTDataDistributor / SpectrumVisualProcessor /
    ScopeVisualProcessor (oracle/ref_visuals.cpp)."""

    def __init__(self, path=VISUALS_LIB):
        if not os.path.exists(path):
            import subprocess, sys
            subprocess.check_call([sys.executable, os.path.join(_HERE, "build_ref.py")])
        L = self.L = C.CDLL(path)
        vp, u, i, ll, f = C.c_void_p, C.c_uint, C.c_int, C.c_longlong, C.c_float
        L.refsrc_distrib_create.restype = vp; L.refsrc_distrib_create.argtypes = [u, u]
        L.refsrc_distrib_set_fft_size.argtypes = [vp, u]; L.refsrc_distrib_set_fft_size.restype = None
        L.refsrc_distrib_set_lines_per_second.argtypes = [vp, u]; L.refsrc_distrib_set_lines_per_second.restype = None
        L.refsrc_distrib_push.restype = u; L.refsrc_distrib_push.argtypes = [vp, vp, u, ll, ll, vp, u, vp]
        L.refsrc_distrib_destroy.argtypes = [vp]; L.refsrc_distrib_destroy.restype = None
        L.refsrc_spectrum_create.restype = vp; L.refsrc_spectrum_create.argtypes = [u, ll]
        L.refsrc_spectrum_set_view.argtypes = [vp, i, ll, C.c_long]; L.refsrc_spectrum_set_view.restype = None
        L.refsrc_spectrum_set_average_rate.argtypes = [vp, f]; L.refsrc_spectrum_set_average_rate.restype = None
        L.refsrc_spectrum_set_band.argtypes = [vp, ll, C.c_long]; L.refsrc_spectrum_set_band.restype = None
        L.refsrc_spectrum_set_hide_dc.argtypes = [vp, i]; L.refsrc_spectrum_set_hide_dc.restype = None
        L.refsrc_spectrum_desired_input.argtypes = [vp]
        L.refsrc_spectrum_frame.argtypes = [vp, vp, u, ll, ll, vp, u, vp, vp, vp]
        L.refsrc_spectrum_destroy.argtypes = [vp]; L.refsrc_spectrum_destroy.restype = None
        L.refsrc_scope_create.restype = vp; L.refsrc_scope_create.argtypes = [i]
        L.refsrc_scope_process.restype = u; L.refsrc_scope_process.argtypes = [vp, vp, u, i, i, i, i, vp, u, vp, u, vp, vp, vp]
        L.refsrc_scope_destroy.argtypes = [vp]; L.refsrc_scope_destroy.restype = None


class SrcDistrib:
    """The reference's FFTDataDistributor object: setInput / attachOutput / setFFTSize / setLinesPerSecond, then per block
    push + run() + drain (FFTVisualDataThread.cpp:40-84)."""

    def __init__(self, vis, frame_len, lps):
        self.L, self.n = vis.L, frame_len
        self.h = C.c_void_p(self.L.refsrc_distrib_create(frame_len, lps))

    def push(self, block, sample_rate, frequency=0):
        block = np.ascontiguousarray(block, np.complex64)
        out = np.zeros(block.size + 4 * self.n + 1048576, np.complex64)
        ns = C.c_uint()
        nf = self.L.refsrc_distrib_push(self.h, _p(block), block.size, int(sample_rate), int(frequency), _p(out), out.size, C.byref(ns))
        assert ns.value <= out.size and ns.value == nf * self.n
        return out[:ns.value].reshape(nf, self.n)

    def set_frame_len(self, n):
        self.L.refsrc_distrib_set_fft_size(self.h, int(n))
        self.n = int(n)

    def close(self):
        self.L.refsrc_distrib_destroy(self.h)


class SrcSpectrum:
    """The reference's SpectrumVisualProcessor object."""

    def __init__(self, vis, fft_size, device_sample_rate):
        self.L, self.fft_size = vis.L, fft_size
        self.h = C.c_void_p(self.L.refsrc_spectrum_create(fft_size, int(device_sample_rate)))

    def set_average_rate(self, r):
        self.L.refsrc_spectrum_set_average_rate(self.h, C.c_float(r))

    def set_band(self, center_freq, bandwidth):
        """Full-band display: setCenterFrequency(frequency) + setBandwidth(sampleRate), as AppFrame.cpp:1396-1399 does."""
        self.L.refsrc_spectrum_set_band(self.h, int(center_freq), int(bandwidth))

    def set_view(self, is_view, center_freq, bandwidth):
        self.L.refsrc_spectrum_set_view(self.h, 1 if is_view else 0, int(center_freq), int(bandwidth))

    def desired_input(self):
        return int(self.L.refsrc_spectrum_desired_input(self.h))

    def frame(self, frame, frequency, sample_rate):
        """One process() call; None if the processor distributed nothing."""
        frame = np.ascontiguousarray(frame, np.complex64)
        pts = np.zeros(self.fft_size, np.float32); cf = np.zeros(2, np.float64)
        center, bw = C.c_longlong(), C.c_int()
        m = self.L.refsrc_spectrum_frame(self.h, _p(frame), frame.size, int(frequency), int(sample_rate), _p(pts), pts.size, _p(cf),
                                         C.byref(center), C.byref(bw))
        if not m:
            return None
        assert m == self.fft_size
        return dict(points=pts, ceil_floor=cf, center=center.value, bandwidth=bw.value)

    def close(self):
        self.L.refsrc_spectrum_destroy(self.h)


class SrcScope:
    """The reference's ScopeVisualProcessor object (scope and spectrum both enabled, as constructed)."""

    def __init__(self, vis, fft_size=1024):
        self.L, self.fft_size = vis.L, fft_size
        self.h = C.c_void_p(self.L.refsrc_scope_create(fft_size))

    def process(self, data, channels, sample_rate, input_rate, type=0):
        d = np.ascontiguousarray(data, np.float32)
        wave = np.zeros(2 * d.size + 4096, np.float32); spec = np.zeros(2 * self.fft_size + 16, np.float32)
        ns, mode = C.c_uint(), C.c_int(-1); fc = np.zeros(2, np.float64)
        nw = self.L.refsrc_scope_process(self.h, _p(d), d.size, int(channels), int(sample_rate), int(input_rate), int(type),
                                         _p(wave), wave.size, _p(spec), spec.size, C.byref(ns), _p(fc), C.byref(mode))
        return wave[:nw].copy(), spec[:ns.value].copy(), fc, mode.value

    def close(self):
        self.L.refsrc_scope_destroy(self.h)


class SrcMixer:
    """The reference's audioCallback (src/audio/AudioThread.cpp:88-243) on real AudioThread objects (oracle/ref_audio.cpp, in
    libdemod_ref.so: pass a RefDemodLib)."""

    def __init__(self, demod_lib, sample_rate=48000):
        self.L = demod_lib.L
        self.h = C.c_void_p(self.L.refsrc_mixer_create(int(sample_rate)))

    def add_source(self, gain=1.0):
        return self.L.refsrc_mixer_add_source(self.h, float(gain))

    def set_gain(self, i, gain):
        self.L.refsrc_mixer_set_gain(self.h, int(i), float(gain))

    def set_active(self, i, active):
        self.L.refsrc_mixer_set_active(self.h, int(i), 1 if active else 0)

    def push(self, i, data, channels=1, peak=0.0, sample_rate=48000):
        d = np.ascontiguousarray(data, np.float32)
        self.L.refsrc_mixer_push(self.h, int(i), _p(d), d.size, int(channels), float(peak), int(sample_rate))

    def callback(self, n_frames):
        out = np.full(2 * n_frames, np.nan, np.float32)          # the callback zeroes the buffer itself
        assert self.L.refsrc_mixer_callback(self.h, _p(out), int(n_frames)) == 0
        return out

    def close(self):
        self.L.refsrc_mixer_destroy(self.h)


class SrcWav:
    """The reference's AudioFileWAV object writing directory/name.wav (oracle/ref_audio.cpp, in libdemod_ref.so)."""

    def __init__(self, demod_lib, directory, name):
        self.L = demod_lib.L
        self.path = os.path.join(directory, name + ".wav")
        self.h = C.c_void_p(self.L.refsrc_wav_open(str(directory).encode(), name.encode()))

    def write(self, data, channels, sample_rate, peak):
        d = np.ascontiguousarray(data, np.float32)
        assert self.L.refsrc_wav_write(self.h, _p(d), d.size, int(channels), int(sample_rate), float(peak)) == 0

    def close(self):
        """Returns (header bytes, payload as int16)."""
        self.L.refsrc_wav_close(self.h)
        raw = open(self.path, "rb").read()
        return raw[:44], np.frombuffer(raw[44:], np.int16).copy()


DEMOD_LIB = os.path.join(_HERE, "_ref", "libdemod_ref.so")


class RefDemodLib:
    """oracle/_ref/libdemod_ref.so: the reference's own, unmodified DemodulatorInstance on its own threads (oracle/ref_demod.cpp)."""

    def __init__(self, path=DEMOD_LIB):
        if not os.path.exists(path):
            import subprocess, sys
            subprocess.check_call([sys.executable, os.path.join(_HERE, "build_ref.py")])
        L = self.L = C.CDLL(path)
        vp, u, i, ll, f = C.c_void_p, C.c_uint, C.c_int, C.c_longlong, C.c_float
        L.refsrc_demod_create.restype = vp; L.refsrc_demod_create.argtypes = [C.c_char_p, ll, i, i, ll, ll]
        L.refsrc_demod_set_frequency.argtypes = [vp, ll]; L.refsrc_demod_set_frequency.restype = None
        L.refsrc_demod_set_squelch.argtypes = [vp, i, f]; L.refsrc_demod_set_squelch.restype = None
        L.refsrc_demod_bandwidth.argtypes = [vp]
        L.refsrc_demod_process.argtypes = [vp, vp, u, ll, vp, u, vp, vp, vp, vp, vp]
        L.refsrc_demod_destroy.argtypes = [vp]; L.refsrc_demod_destroy.restype = None
        L.refsrc_demod_collect.argtypes = [vp, ll, vp, u, vp, vp, vp, vp, vp]
        L.refsrc_app_create.restype = vp; L.refsrc_app_create.argtypes = [i]
        L.refsrc_app_add_demod.argtypes = [vp, vp]; L.refsrc_app_add_demod.restype = None
        L.refsrc_app_block.argtypes = [vp, vp, u, ll, ll, i, vp, u]
        L.refsrc_app_route.argtypes = [vp, vp, vp, vp]
        L.refsrc_app_destroy.argtypes = [vp]; L.refsrc_app_destroy.restype = None
        L.refsrc_mixer_create.restype = vp; L.refsrc_mixer_create.argtypes = [i]
        L.refsrc_mixer_add_source.argtypes = [vp, f]
        L.refsrc_mixer_set_gain.argtypes = [vp, i, f]; L.refsrc_mixer_set_gain.restype = None
        L.refsrc_mixer_set_active.argtypes = [vp, i, i]; L.refsrc_mixer_set_active.restype = None
        L.refsrc_mixer_push.argtypes = [vp, i, vp, u, i, f, i]; L.refsrc_mixer_push.restype = None
        L.refsrc_mixer_callback.argtypes = [vp, vp, u]
        L.refsrc_mixer_destroy.argtypes = [vp]; L.refsrc_mixer_destroy.restype = None
        L.refsrc_wav_open.restype = vp; L.refsrc_wav_open.argtypes = [C.c_char_p, C.c_char_p]
        L.refsrc_wav_write.argtypes = [vp, vp, u, i, i, f]
        L.refsrc_wav_close.argtypes = [vp]; L.refsrc_wav_close.restype = None


class SrcDemod:
    """One real DemodulatorInstance: modem `modem` of `bandwidth` Hz at absolute frequency `frequency`, fed wideband blocks at
    `sample_rate` whose tuner frequency is `center`."""

    def __init__(self, lib, modem, sample_rate, bandwidth, frequency, center, audio_rate=48000):
        self.L = lib.L
        self.sample_rate, self.audio_rate = sample_rate, audio_rate
        self.h = C.c_void_p(self.L.refsrc_demod_create(modem.encode(), int(sample_rate), int(bandwidth), int(audio_rate),
                                                       int(frequency), int(center)))
        assert self.h.value, "the reference's demodulator did not initialize"
        self.bandwidth = self.L.refsrc_demod_bandwidth(self.h)

    def set_frequency(self, frequency):
        self.L.refsrc_demod_set_frequency(self.h, int(frequency))

    def set_squelch(self, enabled, level_db):
        self.L.refsrc_demod_set_squelch(self.h, 1 if enabled else 0, C.c_float(level_db))

    def process(self, block, center):
        block = np.ascontiguousarray(block, np.complex64)
        cap = 2 * int(block.size * self.audio_rate / self.sample_rate) + 65536
        au = np.zeros(cap, np.float32)
        n, ch, pk, sq = C.c_uint(), C.c_int(), C.c_float(), C.c_int()
        m3 = (C.c_float * 3)()
        rc = self.L.refsrc_demod_process(self.h, _p(block), block.size, int(center), _p(au), cap, C.byref(n), C.byref(ch), C.byref(pk),
                                         C.byref(sq), m3)
        assert rc >= 0, "the reference's demodulator threads timed out"
        return dict(skipped=bool(rc), audio=au[:n.value].copy(), channels=ch.value, peak=pk.value, squelched=bool(sq.value),
                    signal_level=m3[0], signal_floor=m3[1], signal_ceil=m3[2])

    def collect(self, center):
        """What came out for the blocks an SrcApp fed this demodulator (center: the frequency stamped on those blocks)."""
        cap = 1 << 20
        au = np.zeros(cap, np.float32)
        n, ch, pk, sq = C.c_uint(), C.c_int(), C.c_float(), C.c_int()
        m3 = (C.c_float * 3)()
        rc = self.L.refsrc_demod_collect(self.h, int(center), _p(au), cap, C.byref(n), C.byref(ch), C.byref(pk), C.byref(sq), m3)
        assert rc >= 0, "the reference's demodulator threads timed out"
        return dict(skipped=bool(rc), audio=au[:n.value].copy(), channels=ch.value, peak=pk.value, squelched=bool(sq.value),
                    signal_level=m3[0], signal_floor=m3[1], signal_ceil=m3[2])

    def close(self):
        self.L.refsrc_demod_destroy(self.h)


class SrcApp:
    """The reference's SDRPostThread on its own thread (oracle/ref_demod.cpp): blocks in as the device thread delivers them,
    demodulators (SrcDemod) registered the way DemodulatorMgr holds them.  One per process at a time."""

    def __init__(self, lib, channelizer_type=1):
        self.L = lib.L
        self.h = C.c_void_p(self.L.refsrc_app_create(int(channelizer_type)))

    def add(self, demod):
        self.L.refsrc_app_add_demod(self.h, demod.h)

    def block(self, x, center, sample_rate, num_channels=1):
        """Returns what SDRPostThread put on its "IQDataOutput" queue for this block."""
        x = np.ascontiguousarray(x, np.complex64)
        out = np.zeros(x.size, np.complex64)
        m = self.L.refsrc_app_block(self.h, _p(x), x.size, int(center), int(sample_rate), int(num_channels), _p(out), out.size)
        assert m >= 0, "the reference's SDRPostThread timed out"
        return out[:m]

    def route(self, demod):
        """(filterbank channel or -1 when the demodulator is not in the run list, that channel's centre frequency, channel rate)."""
        cc, bw = C.c_longlong(), C.c_int()
        ch = self.L.refsrc_app_route(self.h, demod.h, C.byref(cc), C.byref(bw))
        return ch, cc.value, bw.value

    def close(self):
        self.L.refsrc_app_destroy(self.h)
