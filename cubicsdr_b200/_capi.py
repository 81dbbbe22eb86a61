"""ctypes binding of include/sdr_b200.h.

`Lib(path)` binds every entry point of the C-ABI.  The package binds the product library (built by nvcc for
sm_100a, cubicsdr_b200/libsdr_b200.so) and refuses to work without it: there is no CPU path.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PRODUCT_LIB = os.path.join(_HERE, "libsdr_b200.so")

# every symbol include/sdr_b200.h declares (tests/test_abi.py checks the header against this list and the .so)
SYMBOLS = [
    "sdr_last_error", "sdr_version", "sdr_is_device_build",
    "sdr_host_alloc", "sdr_host_free", "sdr_engine_host_buffer", "sdr_chan_result_caps", "sdr_chan_audio_channels", "sdr_modem_check_sample_rate2", "sdr_chan_set_squelch", "sdr_chan_fetch_meter",
    "sdr_engine_create", "sdr_engine_destroy", "sdr_engine_submit", "sdr_engine_submit_device", "sdr_engine_submit_device_async",
    "sdr_engine_wait",
    "sdr_engine_wait_block",
    "sdr_engine_fetch_block", "sdr_engine_launch_count", "sdr_engine_set_timing", "sdr_engine_get_timing",
    "sdr_chan_create", "sdr_chan_destroy", "sdr_chan_set_frequency", "sdr_chan_get_bandwidth", "sdr_chan_want_iq",
    "sdr_chan_fetch_audio", "sdr_chan_fetch_iq", "sdr_engine_fetch_audio_batch", "sdr_engine_view_audio_batch", "sdr_spectrum_view",
    "sdr_modem_check_sample_rate", "sdr_modem_create", "sdr_modem_demodulate",
    "sdr_spectrum_create", "sdr_spectrum_destroy", "sdr_spectrum_set_average_rate", "sdr_spectrum_set_view", "sdr_spectrum_desired_input", "sdr_spectrum_frames",
    "sdr_spectrum_fetch", "sdr_spectrum_fetch_all", "sdr_spectrum_fetch_debug",
    "sdr_engine_set_channelizer", "sdr_engine_set_dc_window", "sdr_engine_channel_plan", "sdr_engine_set_channel_block_max",
    "sdr_pfb_optimal_channel_count", "sdr_pfb_create", "sdr_pfb2_create", "sdr_pfb_destroy", "sdr_pfb_reset", "sdr_pfb_execute", "sdr_pfb_execute_device",
    "sdr_group_create", "sdr_group_create_rank", "sdr_group_unique_id", "sdr_group_destroy", "sdr_group_info", "sdr_group_engine",
    "sdr_group_chan_create", "sdr_group_spectrum_create", "sdr_group_submit", "sdr_group_wait", "sdr_group_wait_block", "sdr_group_set_host_results",
    "sdr_design_kaiser", "sdr_design_nco_table", "sdr_design_nco_step", "sdr_design_resampler",
    "sdr_design_halfband_branch", "sdr_design_am_dc_block", "sdr_design_hilbert_q", "sdr_design_butter6_sos", "sdr_design_cheby2_bp5_sos", "sdr_design_hilbert_q_as",
    "sdr_mixer_create", "sdr_mixer_destroy", "sdr_mixer_bind", "sdr_mixer_add_source", "sdr_mixer_set_gain", "sdr_mixer_set_active",
    "sdr_mixer_push", "sdr_mixer_render", "sdr_chan_fetch_pcm16", "sdr_scope_create", "sdr_scope_destroy", "sdr_scope_process",
]


class SdrError(RuntimeError):
    pass


class Lib:
    def __init__(self, path=PRODUCT_LIB):
        if not os.path.exists(path):
            raise SdrError(
                "%s is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  cubicsdr_b200 has no CPU fallback." % path)
        self.path = path
        L = self.L = C.CDLL(path)
        vp, u, i, ll, f, d = C.c_void_p, C.c_uint, C.c_int, C.c_longlong, C.c_float, C.c_double
        P = C.POINTER
        L.sdr_last_error.restype = C.c_char_p
        L.sdr_version.restype = C.c_char_p
        L.sdr_engine_create.argtypes = [i, u, vp, P(vp)]
        L.sdr_engine_destroy.argtypes = [vp]
        L.sdr_engine_submit.argtypes = [vp, vp, u, ll, ll]
        L.sdr_host_alloc.argtypes = [C.c_size_t, P(vp)]
        L.sdr_host_free.argtypes = [vp]
        L.sdr_engine_host_buffer.argtypes = [vp, P(vp), P(u)]
        L.sdr_chan_result_caps.argtypes = [vp, P(u), P(u)]
        L.sdr_chan_audio_channels.argtypes = [vp, P(i)]
        L.sdr_modem_check_sample_rate2.argtypes = [C.c_char_p, ll, i]
        L.sdr_chan_set_squelch.argtypes = [vp, i, f]
        L.sdr_chan_fetch_meter.argtypes = [vp, P(f), P(f), P(f), P(i)]
        L.sdr_engine_submit_device.argtypes = [vp, vp, u, ll, ll, i]
        L.sdr_engine_submit_device_async.argtypes = [vp, vp, u, ll, ll, i, vp]
        L.sdr_engine_wait.argtypes = [vp]
        L.sdr_engine_wait_block.argtypes = [vp, C.c_int]
        L.sdr_engine_fetch_block.argtypes = [vp, vp, u, P(u)]
        L.sdr_engine_launch_count.argtypes = [vp]
        L.sdr_engine_launch_count.restype = C.c_ulonglong
        L.sdr_engine_set_timing.argtypes = [vp, i]
        L.sdr_engine_get_timing.argtypes = [vp, vp, P(C.c_ulonglong)]
        L.sdr_chan_create.argtypes = [vp, C.c_char_p, ll, ll, i, P(vp)]
        L.sdr_chan_destroy.argtypes = [vp]
        L.sdr_chan_set_frequency.argtypes = [vp, ll]
        L.sdr_chan_get_bandwidth.argtypes = [vp, P(i)]
        L.sdr_chan_want_iq.argtypes = [vp, i]
        L.sdr_chan_fetch_audio.argtypes = [vp, vp, u, P(u), P(d), P(f), P(i)]
        L.sdr_chan_fetch_iq.argtypes = [vp, vp, u, P(u)]
        L.sdr_engine_fetch_audio_batch.argtypes = [vp, vp, u, vp, u, vp, vp, vp]
        L.sdr_engine_view_audio_batch.argtypes = [vp, vp, u, vp, vp, vp, vp]
        L.sdr_spectrum_view.argtypes = [vp, P(vp), P(vp), P(u)]
        L.sdr_modem_check_sample_rate.argtypes = [C.c_char_p, ll]
        L.sdr_modem_create.argtypes = [vp, C.c_char_p, ll, i, u, P(vp)]
        L.sdr_modem_demodulate.argtypes = [vp, vp, u, vp, u, P(u), P(d), P(f)]
        L.sdr_spectrum_create.argtypes = [vp, u, u, P(vp)]
        L.sdr_spectrum_destroy.argtypes = [vp]
        L.sdr_spectrum_set_average_rate.argtypes = [vp, f]
        L.sdr_spectrum_set_view.argtypes = [vp, i, ll, C.c_long]
        L.sdr_spectrum_desired_input.argtypes = [vp, ll, P(u)]
        L.sdr_spectrum_frames.argtypes = [vp, P(u)]
        L.sdr_spectrum_fetch.argtypes = [vp, u, vp, vp, vp]
        L.sdr_spectrum_fetch_debug.argtypes = [vp, u, vp, vp]
        L.sdr_engine_set_channelizer.argtypes = [vp, C.c_int]
        L.sdr_engine_set_dc_window.argtypes = [vp, C.c_int]
        L.sdr_engine_debug_dc_window.argtypes = [vp, C.POINTER(C.c_uint), C.POINTER(C.c_ulonglong)]
        L.sdr_engine_channel_plan.argtypes = [C.c_longlong, vp, vp]
        L.sdr_engine_set_channel_block_max.argtypes = [vp, u]
        L.sdr_pfb_optimal_channel_count.argtypes = [C.c_longlong]
        L.sdr_pfb_create.argtypes = [vp, u, u, C.c_float, u, C.POINTER(vp)]
        L.sdr_pfb2_create.argtypes = [vp, u, u, C.c_float, u, C.POINTER(vp)]
        L.sdr_pfb_destroy.argtypes = [vp]
        L.sdr_pfb_reset.argtypes = [vp]
        L.sdr_pfb_execute.argtypes = [vp, vp, u, vp]
        L.sdr_pfb_execute_device.argtypes = [vp, vp, u, vp]
        L.sdr_spectrum_fetch_all.argtypes = [vp, u, vp, vp, vp, P(u)]
        L.sdr_group_create.argtypes = [vp, i, u, P(vp)]
        L.sdr_group_create_rank.argtypes = [i, i, i, vp, u, P(vp)]
        L.sdr_group_unique_id.argtypes = [vp]
        L.sdr_group_destroy.argtypes = [vp]
        L.sdr_group_info.argtypes = [vp, P(i), P(i), P(i)]
        L.sdr_group_engine.argtypes = [vp, i, P(vp)]
        L.sdr_group_chan_create.argtypes = [vp, C.c_char_p, ll, ll, i, P(vp)]
        L.sdr_group_spectrum_create.argtypes = [vp, u, u, P(vp)]
        L.sdr_group_submit.argtypes = [vp, vp, u, ll, ll]
        L.sdr_group_wait.argtypes = [vp]
        L.sdr_group_wait_block.argtypes = [vp, i]
        L.sdr_group_set_host_results.argtypes = [vp, i]
        L.sdr_design_kaiser.argtypes = [u, f, f, f, vp]
        L.sdr_design_nco_table.argtypes = [vp]
        L.sdr_design_nco_step.argtypes = [f]
        L.sdr_design_nco_step.restype = C.c_uint32
        L.sdr_design_resampler.argtypes = [f, f, P(i), P(i), P(f), P(C.c_uint32), vp, vp]
        L.sdr_design_halfband_branch.argtypes = [i, f, vp]
        L.sdr_design_am_dc_block.argtypes = [vp]
        L.sdr_design_hilbert_q.argtypes = [vp]
        L.sdr_design_butter6_sos.argtypes = [vp]
        L.sdr_mixer_create.argtypes = [vp, C.c_int, C.c_uint, C.POINTER(vp)]
        L.sdr_mixer_destroy.argtypes = [vp]
        L.sdr_mixer_bind.argtypes = [vp, vp, C.c_float, C.POINTER(C.c_int)]
        L.sdr_mixer_add_source.argtypes = [vp, C.c_float, C.POINTER(C.c_int)]
        L.sdr_mixer_set_gain.argtypes = [vp, C.c_int, C.c_float]
        L.sdr_mixer_set_active.argtypes = [vp, C.c_int, C.c_int]
        L.sdr_mixer_push.argtypes = [vp, C.c_int, vp, C.c_uint, C.c_int, C.c_float, C.c_int]
        L.sdr_mixer_render.argtypes = [vp, C.c_uint, vp]
        L.sdr_chan_fetch_pcm16.argtypes = [vp, vp, C.c_uint, C.POINTER(C.c_uint)]
        L.sdr_scope_create.argtypes = [vp, vp, C.c_uint, C.c_uint, C.POINTER(vp)]
        L.sdr_scope_destroy.argtypes = [vp]
        L.sdr_scope_process.argtypes = [vp, vp, C.POINTER(C.c_uint), vp, C.POINTER(C.c_uint), vp]
        L.sdr_design_cheby2_bp5_sos.argtypes = [C.c_float, C.c_float, vp]
        L.sdr_design_hilbert_q_as.argtypes = [C.c_float, vp]

    def host_alloc(self, n_complex):
        """Pinned complex64 array of n samples (sdr_host_alloc); keep the returned array alive while it is in use."""
        p = C.c_void_p()
        self.check(self.L.sdr_host_alloc(int(n_complex) * 8, C.byref(p)))
        buf = (C.c_float * (2 * int(n_complex))).from_address(p.value)
        a = np.frombuffer(buf, dtype=np.complex64)
        return a, p.value

    def host_free(self, addr):
        self.L.sdr_host_free(C.c_void_p(addr))

    def check(self, rc):
        if rc != 0:
            raise SdrError("sdr_b200 error %d: %s" % (rc, self.L.sdr_last_error().decode()))

    @property
    def is_device_build(self):
        return bool(self.L.sdr_is_device_build())


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class Engine:
    """One engine per GPU (sdr_engine).  submit() replaces SDRPostThread::runSingleCH (SDRPostThread.cpp:248-299)."""

    def __init__(self, lib, device=0, max_block=1 << 20, stream=None, handle=None):
        self.lib = lib
        self.owned = handle is None
        if handle is None:
            handle = C.c_void_p()
            lib.check(lib.L.sdr_engine_create(device, max_block, C.c_void_p(stream) if stream else None, C.byref(handle)))
        self.h = handle
        self.max_block = max_block

    def close(self):
        if self.h and self.owned:
            self.lib.L.sdr_engine_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def submit(self, iq, center_freq, sample_rate):
        iq = np.ascontiguousarray(iq, dtype=np.complex64)
        self.lib.check(self.lib.L.sdr_engine_submit(self.h, _ptr(iq), iq.size, int(center_freq), int(sample_rate)))

    def submit_ptr(self, host_ptr, n, center_freq, sample_rate):
        """sdr_engine_submit on a raw host address (e.g. pinned memory from Lib.host_alloc)."""
        self.lib.check(self.lib.L.sdr_engine_submit(self.h, C.c_void_p(host_ptr), int(n), int(center_freq), int(sample_rate)))

    def host_buffer(self):
        """The pinned buffer the next submit takes without a staging copy, as a complex64 numpy view."""
        p, cap = C.c_void_p(), C.c_uint()
        self.lib.check(self.lib.L.sdr_engine_host_buffer(self.h, C.byref(p), C.byref(cap)))
        buf = (C.c_float * (2 * cap.value)).from_address(p.value)
        return np.frombuffer(buf, dtype=np.complex64)

    def submit_device(self, dev_ptr, n, center_freq, sample_rate, fetch_to_host=True):
        self.lib.check(self.lib.L.sdr_engine_submit_device(self.h, C.c_void_p(dev_ptr), n, int(center_freq), int(sample_rate),
                                                           1 if fetch_to_host else 0))

    def submit_device_async(self, dev_ptr, n, center_freq, sample_rate, fetch_to_host=True, ready_event=None):
        """ready_event: a cudaEvent_t handle (int) after which the block is valid, or None if it already is."""
        self.lib.check(self.lib.L.sdr_engine_submit_device_async(self.h, C.c_void_p(dev_ptr), n, int(center_freq), int(sample_rate),
                                                                 1 if fetch_to_host else 0,
                                                                 C.c_void_p(ready_event) if ready_event else None))

    def set_channelizer(self, mode):
        """0: SDRPostThread::runSingleCH for every block; 1: runPFBCH when the rate gives more than one channel."""
        self.lib.check(self.lib.L.sdr_engine_set_channelizer(self.h, int(mode)))

    def set_dc_window(self, mode):
        """DC blocker replay: 1 candidate windows, 0 warmed-up segments, -1 automatic (by channel load).  Same samples either way."""
        self.lib.check(self.lib.L.sdr_engine_set_dc_window(self.h, int(mode)))

    def dc_window_stats(self):
        """(stretches the window walker had to run serially, blocks that took the window path) so far"""
        s, b = C.c_uint(), C.c_ulonglong()
        self.lib.check(self.lib.L.sdr_engine_debug_dc_window(self.h, C.byref(s), C.byref(b)))
        return s.value, b.value

    def set_channel_block_max(self, n):
        """Most samples a demodulator gets per block (channelizer mode: block / numChannels); before creating channels."""
        self.lib.check(self.lib.L.sdr_engine_set_channel_block_max(self.h, int(n)))

    def channel_plan(self, sample_rate):
        """(numChannels, channel rate) of channelizer mode for this wideband rate."""
        m, r = C.c_int(), C.c_longlong()
        self.lib.check(self.lib.L.sdr_engine_channel_plan(int(sample_rate), C.byref(m), C.byref(r)))
        return m.value, r.value

    def wait(self):
        self.lib.check(self.lib.L.sdr_engine_wait(self.h))

    def wait_block(self, lag=0):
        """Wait for the host results of the latest (lag=0) or previous (lag=1) block only; fetches then read that block."""
        self.lib.check(self.lib.L.sdr_engine_wait_block(self.h, int(lag)))

    def fetch_block(self, n):
        out = np.zeros(n, np.complex64)
        got = C.c_uint()
        self.lib.check(self.lib.L.sdr_engine_fetch_block(self.h, _ptr(out), n, C.byref(got)))
        return out[:got.value]

    @property
    def launch_count(self):
        return int(self.lib.L.sdr_engine_launch_count(self.h))

    def fetch_audio_batch(self, chans, stride=1024):
        """(audio[n_chan, stride], n_audio[n_chan], level_db[n_chan], peak[n_chan]) for a list of Channel objects."""
        n = len(chans)
        key = (tuple(c.h.value for c in chans), stride)
        if getattr(self, "_batch_key", None) != key:
            self._batch_key = key
            self._batch_h = (C.c_void_p * n)(*[c.h for c in chans])
            self._batch_audio = np.zeros((n, stride), np.float32)
            self._batch_n = np.zeros(n, np.uint32)
            self._batch_lvl = np.zeros(n, np.float64)
            self._batch_pk = np.zeros(n, np.float32)
        self.lib.check(self.lib.L.sdr_engine_fetch_audio_batch(self.h, self._batch_h, n, _ptr(self._batch_audio), stride,
                                                               _ptr(self._batch_n), _ptr(self._batch_lvl), _ptr(self._batch_pk)))
        return self._batch_audio, self._batch_n, self._batch_lvl, self._batch_pk

    def view_audio_batch(self, chans):
        """Zero-copy batch fetch: (pointers[n_chan] as uint64, n_audio[n_chan], level_db[n_chan], peak[n_chan]); the audio stays in
        the engine's pinned result buffer (valid until two further blocks have been submitted)."""
        n = len(chans)
        key = (tuple(c.h.value for c in chans),)
        if getattr(self, "_view_key", None) != key:
            self._view_key = key
            self._view_h = (C.c_void_p * n)(*[c.h for c in chans])
            self._view_ptr = np.zeros(n, np.uint64)
            self._view_n = np.zeros(n, np.uint32)
            self._view_lvl = np.zeros(n, np.float64)
            self._view_pk = np.zeros(n, np.float32)
        self.lib.check(self.lib.L.sdr_engine_view_audio_batch(self.h, self._view_h, n, _ptr(self._view_ptr), _ptr(self._view_n),
                                                              _ptr(self._view_lvl), _ptr(self._view_pk)))
        return self._view_ptr, self._view_n, self._view_lvl, self._view_pk

    @staticmethod
    def audio_at(ptr, n):
        """numpy view of n floats at a pointer returned by view_audio_batch"""
        return np.frombuffer((C.c_float * int(n)).from_address(int(ptr)), dtype=np.float32)

    def set_timing(self, on=True):
        self.lib.check(self.lib.L.sdr_engine_set_timing(self.h, 1 if on else 0))

    def get_timing(self):
        """({'dc','spectrum','front','tail'} accumulated ms, blocks)"""
        ms = np.zeros(4, np.float64)
        nb = C.c_ulonglong()
        self.lib.check(self.lib.L.sdr_engine_get_timing(self.h, _ptr(ms), C.byref(nb)))
        return dict(dc=ms[0], spectrum=ms[1], front=ms[2], tail=ms[3]), int(nb.value)


class Channel:
    """sdr_chan: DemodulatorPreThread + Modem + DemodulatorThread metering for one demodulator."""

    def __init__(self, engine, modem, sample_rate, bandwidth, audio_rate=48000, frequency=0, want_iq=False, handle=None):
        self.engine, self.lib = engine, engine.lib
        h = handle
        if h is None:
            h = C.c_void_p()
            self.lib.check(self.lib.L.sdr_chan_create(engine.h, modem.encode(), int(sample_rate), int(bandwidth), int(audio_rate), C.byref(h)))
        self.h = h
        self.modem = modem
        bw = C.c_int()
        self.lib.check(self.lib.L.sdr_chan_get_bandwidth(h, C.byref(bw)))
        self.bandwidth = bw.value
        self.set_frequency(frequency)
        if want_iq:
            self.lib.check(self.lib.L.sdr_chan_want_iq(h, 1))
        ca, ci = C.c_uint(), C.c_uint()
        self.lib.check(self.lib.L.sdr_chan_result_caps(h, C.byref(ca), C.byref(ci)))
        self._audio = np.zeros(ca.value + 16, np.float32)
        self._iq = np.zeros(ci.value + 16, np.complex64)

    def set_frequency(self, f):
        self.lib.check(self.lib.L.sdr_chan_set_frequency(self.h, int(f)))

    def fetch_audio(self):
        n, lvl, pk, sk = C.c_uint(), C.c_double(), C.c_float(), C.c_int()
        self.lib.check(self.lib.L.sdr_chan_fetch_audio(self.h, _ptr(self._audio), self._audio.size, C.byref(n), C.byref(lvl),
                                                       C.byref(pk), C.byref(sk)))
        return self._audio[:n.value].copy(), lvl.value, pk.value, bool(sk.value)

    def set_squelch(self, enabled, level_db):
        self.lib.check(self.lib.L.sdr_chan_set_squelch(self.h, 1 if enabled else 0, C.c_float(level_db)))

    def fetch_meter(self):
        """(signal_level, signal_floor, signal_ceil, squelched) after the fetched block."""
        a, b, c_, s = C.c_float(), C.c_float(), C.c_float(), C.c_int()
        self.lib.check(self.lib.L.sdr_chan_fetch_meter(self.h, C.byref(a), C.byref(b), C.byref(c_), C.byref(s)))
        return a.value, b.value, c_.value, bool(s.value)

    def fetch_iq(self):
        n = C.c_uint()
        self.lib.check(self.lib.L.sdr_chan_fetch_iq(self.h, _ptr(self._iq), self._iq.size, C.byref(n)))
        return self._iq[:n.value].copy()

    def close(self):
        if self.h:
            self.lib.L.sdr_chan_destroy(self.h)
            self.h = None


class Spectrum:
    """sdr_spectrum: FFTDataDistributor framing + SpectrumVisualProcessor (non-view) + WaterfallPanel::step."""

    def __init__(self, engine, fft_size, lines_per_second=0, handle=None):
        self.engine, self.lib = engine, engine.lib
        h = handle
        if h is None:
            h = C.c_void_p()
            self.lib.check(self.lib.L.sdr_spectrum_create(engine.h, fft_size, lines_per_second, C.byref(h)))
        self.h = h
        self.fft_size = fft_size

    def set_average_rate(self, r):
        self.lib.check(self.lib.L.sdr_spectrum_set_average_rate(self.h, r))

    def set_view(self, is_view, center_freq=0, bandwidth=0):
        """SpectrumVisualProcessor::setView(bView, centerFreq, bandwidth): zoomed view from the next block on."""
        self.lib.check(self.lib.L.sdr_spectrum_set_view(self.h, 1 if is_view else 0, int(center_freq), int(bandwidth)))

    def desired_input(self, sample_rate):
        n = C.c_uint()
        self.lib.check(self.lib.L.sdr_spectrum_desired_input(self.h, int(sample_rate), C.byref(n)))
        return n.value

    @property
    def frames(self):
        n = C.c_uint()
        self.lib.check(self.lib.L.sdr_spectrum_frames(self.h, C.byref(n)))
        return n.value

    def fetch(self, idx):
        pts = np.zeros(self.fft_size, np.float32)
        line = np.zeros(self.fft_size, np.uint8)
        cf = np.zeros(2, np.float64)
        self.lib.check(self.lib.L.sdr_spectrum_fetch(self.h, idx, _ptr(pts), _ptr(line), _ptr(cf)))
        return pts, line, cf

    def fetch_all(self, cap_frames=64):
        """(points[n, fft_size], lines[n, fft_size], ceil_floor[n, 2]) of the last block, one call."""
        if getattr(self, "_all_cap", 0) != cap_frames:
            self._all_cap = cap_frames
            self._all_pts = np.zeros((cap_frames, self.fft_size), np.float32)
            self._all_lines = np.zeros((cap_frames, self.fft_size), np.uint8)
            self._all_cf = np.zeros((cap_frames, 2), np.float64)
        n = C.c_uint()
        self.lib.check(self.lib.L.sdr_spectrum_fetch_all(self.h, cap_frames, _ptr(self._all_pts), _ptr(self._all_lines),
                                                         _ptr(self._all_cf), C.byref(n)))
        m = min(n.value, cap_frames)
        return self._all_pts[:m], self._all_lines[:m], self._all_cf[:m]

    def view(self):
        """Zero-copy: (points[n_frames, fft_size] float32, lines[n_frames, fft_size] uint8) views of the pinned result buffer."""
        pp, pl, n = C.c_void_p(), C.c_void_p(), C.c_uint()
        self.lib.check(self.lib.L.sdr_spectrum_view(self.h, C.byref(pp), C.byref(pl), C.byref(n)))
        if not n.value:
            return np.zeros((0, self.fft_size), np.float32), np.zeros((0, self.fft_size), np.uint8)
        pts = np.frombuffer((C.c_float * (n.value * self.fft_size)).from_address(pp.value), dtype=np.float32).reshape(n.value, self.fft_size)
        lines = np.frombuffer((C.c_ubyte * (n.value * self.fft_size)).from_address(pl.value), dtype=np.uint8).reshape(n.value, self.fft_size)
        return pts, lines

    def fetch_debug(self, idx):
        mags = np.zeros(2 * self.fft_size, np.float32)
        maa = np.zeros(2 * self.fft_size, np.float32)
        self.lib.check(self.lib.L.sdr_spectrum_fetch_debug(self.h, idx, _ptr(mags), _ptr(maa)))
        return mags, maa

    def close(self):
        if self.h:
            self.lib.L.sdr_spectrum_destroy(self.h)
            self.h = None


class Group:
    """sdr_group: several GPUs behind the same boundary.  Group(lib, devices=[0, 1, ...]) is the one-process form;
    Group(lib, device=d, rank=r, world=w, nccl_id=bytes) the one-process-per-GPU form (nccl_id from Group.unique_id on rank 0)."""

    def __init__(self, lib, devices=None, max_block=1 << 20, device=0, rank=0, world=1, nccl_id=None):
        self.lib, self.max_block = lib, max_block
        h = C.c_void_p()
        if world > 1:
            idbuf = (C.c_char * 128).from_buffer_copy(nccl_id) if nccl_id is not None else None
            lib.check(lib.L.sdr_group_create_rank(int(device), int(rank), int(world), idbuf, max_block, C.byref(h)))
        else:
            devs = list(devices if devices is not None else [device])
            arr = (C.c_int * len(devs))(*devs)
            lib.check(lib.L.sdr_group_create(arr, len(devs), max_block, C.byref(h)))
        self.h = h
        nl, tot, tr = C.c_int(), C.c_int(), C.c_int()
        lib.check(lib.L.sdr_group_info(h, C.byref(nl), C.byref(tot), C.byref(tr)))
        self.n_local, self.total, self.transport = nl.value, tot.value, tr.value
        self.rank, self.world = (rank, world) if world > 1 else (0, 1)
        self._next = 0
        self.engines = []
        for k in range(self.n_local):
            eh = C.c_void_p()
            lib.check(lib.L.sdr_group_engine(h, k, C.byref(eh)))
            self.engines.append(Engine(lib, max_block=max_block, handle=eh))

    @staticmethod
    def unique_id(lib):
        buf = (C.c_char * 128)()
        lib.check(lib.L.sdr_group_unique_id(buf))
        return bytes(buf)

    def channel(self, modem, sample_rate, bandwidth, audio_rate=48000, frequency=0, want_iq=False):
        """Next channel of the group (placed on member c mod N); None if it lives on another rank."""
        ch = C.c_void_p()
        self.lib.check(self.lib.L.sdr_group_chan_create(self.h, modem.encode(), int(sample_rate), int(bandwidth), int(audio_rate), C.byref(ch)))
        owner = self._next % self.total
        self._next += 1
        if not ch.value:
            return None
        eng = self.engines[0] if self.world > 1 else self.engines[owner]
        return Channel(eng, modem, sample_rate, bandwidth, audio_rate, frequency, want_iq, handle=ch)

    def spectrum(self, fft_size, lines_per_second=0):
        sp = C.c_void_p()
        self.lib.check(self.lib.L.sdr_group_spectrum_create(self.h, fft_size, lines_per_second, C.byref(sp)))
        return Spectrum(self.engines[0], fft_size, lines_per_second, handle=sp) if sp.value else None

    def submit(self, iq, center_freq, sample_rate, n=None):
        """iq: complex64 array, a raw host address (int), or None on ranks that receive the block by broadcast."""
        if iq is None:
            ptr = None
        elif isinstance(iq, int):
            ptr = C.c_void_p(iq)
        else:
            iq = np.ascontiguousarray(iq, dtype=np.complex64)
            ptr, n = _ptr(iq), iq.size
        self.lib.check(self.lib.L.sdr_group_submit(self.h, ptr, int(n), int(center_freq), int(sample_rate)))

    def set_host_results(self, on=True):
        self.lib.check(self.lib.L.sdr_group_set_host_results(self.h, 1 if on else 0))

    def wait(self):
        self.lib.check(self.lib.L.sdr_group_wait(self.h))

    def wait_block(self, lag=0):
        self.lib.check(self.lib.L.sdr_group_wait_block(self.h, int(lag)))

    def close(self):
        if self.h:
            for e in self.engines:
                e.h = None
            self.lib.L.sdr_group_destroy(self.h)
            self.h = None


class Channelizer:
    """sdr_pfb: SDRPostThread::initPFBCH / runPFBCH -- firpfbch_crcf_create_kaiser(LIQUID_ANALYZER, M, 4, 60) + one
    analyzer_execute per M input samples (src/sdr/SDRPostThread.cpp:401-455)."""

    def __init__(self, engine, num_channels, max_frames, m=4, as_db=60.0, oversampled=False):
        """oversampled=True: firpfbch2_crcf (runPFBCH2): one output frame per num_channels / 2 input samples."""
        self.engine, self.lib = engine, engine.lib
        self.M, self.max_frames, self.osf = int(num_channels), int(max_frames), 2 if oversampled else 1
        h = C.c_void_p()
        create = self.lib.L.sdr_pfb2_create if oversampled else self.lib.L.sdr_pfb_create
        self.lib.check(create(engine.h, self.M, int(m), C.c_float(as_db), self.max_frames, C.byref(h)))
        self.h = h

    @staticmethod
    def optimal_channel_count(lib, sample_rate):
        return int(lib.L.sdr_pfb_optimal_channel_count(int(sample_rate)))

    def execute(self, x):
        """x: complex64[n], n a multiple of M -> complex64[n / M, M] (one row per input frame, M channel samples)."""
        x = np.ascontiguousarray(x, np.complex64)
        y = np.zeros(x.size * self.osf, np.complex64)
        self.lib.check(self.lib.L.sdr_pfb_execute(self.h, _ptr(x), x.size, _ptr(y)))
        return y.reshape(-1, self.M)

    def execute_device(self, d_in, n, d_out):
        self.lib.check(self.lib.L.sdr_pfb_execute_device(self.h, C.c_void_p(d_in), int(n), C.c_void_p(d_out)))

    def reset(self):
        self.lib.check(self.lib.L.sdr_pfb_reset(self.h))

    def close(self):
        if self.h:
            self.lib.L.sdr_pfb_destroy(self.h)
            self.h = None



class Mixer:
    """audioCallback of the reference (src/audio/AudioThread.cpp:88-243) on the device: bound demodulators' audio is queued after every
    block without leaving the GPU; render(n) is one callback for n stereo frames."""

    def __init__(self, eng, sample_rate=48000, max_frames=4096):
        self.lib, self.eng = eng.lib, eng
        h = C.c_void_p()
        self.lib.check(self.lib.L.sdr_mixer_create(eng.h, int(sample_rate), int(max_frames), C.byref(h)))
        self.h = h

    def bind(self, chan, gain=1.0):
        i = C.c_int()
        self.lib.check(self.lib.L.sdr_mixer_bind(self.h, chan.h, float(gain), C.byref(i)))
        return i.value

    def add_source(self, gain=1.0):
        i = C.c_int()
        self.lib.check(self.lib.L.sdr_mixer_add_source(self.h, float(gain), C.byref(i)))
        return i.value

    def set_gain(self, source, gain):
        self.lib.check(self.lib.L.sdr_mixer_set_gain(self.h, int(source), float(gain)))

    def set_active(self, source, active):
        self.lib.check(self.lib.L.sdr_mixer_set_active(self.h, int(source), 1 if active else 0))

    def push(self, source, data, channels=1, peak=0.0, sample_rate=48000):
        d = np.ascontiguousarray(data, np.float32)
        self.lib.check(self.lib.L.sdr_mixer_push(self.h, int(source), _ptr(d), d.size, int(channels), float(peak), int(sample_rate)))

    def render(self, n_frames):
        out = np.zeros(2 * n_frames, np.float32)
        self.lib.check(self.lib.L.sdr_mixer_render(self.h, int(n_frames), _ptr(out)))
        return out

    def close(self):
        if self.h:
            self.lib.L.sdr_mixer_destroy(self.h)
            self.h = None


class Scope:
    """ScopeVisualProcessor (src/process/ScopeVisualProcessor.cpp) on a channel's audio: waveform points and audio spectrum."""

    def __init__(self, eng, chan, fft_size=1024, max_samples=1024):
        self.lib, self.fft_size, self.max_samples = eng.lib, fft_size, max_samples
        h = C.c_void_p()
        self.lib.check(self.lib.L.sdr_scope_create(eng.h, chan.h, int(fft_size), int(max_samples), C.byref(h)))
        self.h = h

    def process(self):
        wave = np.zeros(2 * self.max_samples, np.float32); spec = np.zeros(self.fft_size, np.float32)
        nw, ns = C.c_uint(), C.c_uint()
        fc = np.zeros(2, np.float64)
        self.lib.check(self.lib.L.sdr_scope_process(self.h, _ptr(wave), C.byref(nw), _ptr(spec), C.byref(ns), _ptr(fc)))
        return wave[:2 * nw.value].copy(), spec[:2 * ns.value].copy(), fc

    def close(self):
        if self.h:
            self.lib.L.sdr_scope_destroy(self.h)
            self.h = None


def fetch_pcm16(chan, cap=1 << 16):
    """AudioFileWAV::writePayloadToFileStream's 16-bit words for the channel's last block."""
    out = np.zeros(cap, np.int16)
    n = C.c_uint()
    chan.lib.check(chan.lib.L.sdr_chan_fetch_pcm16(chan.h, _ptr(out), cap, C.byref(n)))
    return out[:n.value].copy()
