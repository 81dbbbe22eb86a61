"""Host-side filter design of the product (via the C-ABI introspection calls) against the reference binary."""
import ctypes as C

import numpy as np

from tests.parity import relrms


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def test_nco_table_and_frequency_word(product_lib, ref):
    q = C.c_void_p(ref.L.ref_nco_create())
    t_ref = np.zeros(1024, np.float32); t = np.zeros(1024, np.float32)
    ref.L.ref_nco_sintab(q, _p(t_ref)); product_lib.L.sdr_design_nco_table(_p(t))
    assert np.array_equal(t, t_ref)            # bit-identical sine table
    rng = np.random.default_rng(0)
    for f in list(rng.uniform(-10, 10, 500)) + [0.0, 0.3, np.pi, 2 * np.pi, -0.001, 6.2831855, 1e-9]:
        ref.L.ref_nco_set_frequency(q, C.c_float(f))
        assert ref.L.ref_nco_raw_dtheta(q) == product_lib.L.sdr_design_nco_step(C.c_float(f)), f
    ref.L.ref_nco_set_frequency(q, C.c_float(0.3))
    assert ref.L.ref_nco_raw_dtheta(q) == 205069584
    ref.L.ref_nco_destroy(q)


def test_kaiser_lowpass(product_lib, ref):
    for n, fc, As in [(31, 0.2, 60.0), (3585, 0.412 / 256, 60.0), (3585, 0.49 / 256, 60.0), (21, 0.25, 90.0)]:
        a = np.zeros(n, np.float32); b = np.zeros(n, np.float32)
        ref.L.ref_firdes_kaiser(n, fc, As, 0.0, _p(a)); product_lib.L.sdr_design_kaiser(n, fc, As, 0.0, _p(b))
        assert np.abs(a - b).max() <= 1e-6
    # SURVEY.md Appendix C pins
    a = np.zeros(31, np.float32); ref.L.ref_firdes_kaiser(31, 0.2, 60.0, 0.0, _p(a))
    assert abs(a[15] - 1.0) < 1e-7 and abs(a[14] - 0.748247385) < 1e-6 and abs(a[1] + 0.00256504235) < 1e-7


def test_halfband_prototypes_bit_exact(product_lib, ref):
    for m in (3, 5, 10):
        hf = np.zeros(4 * m + 1, np.float32)
        ref.L.ref_firdespm_halfband_as(m, 65.0, _p(hf))
        h1 = np.zeros(2 * m, np.float32)
        assert product_lib.L.sdr_design_halfband_branch(m, 60.0, _p(h1)) == 0
        # resamp2: h = 2 * hf (f0 = 0), branch taps = odd taps, reversed (symmetric)
        expect = (2.0 * hf[1::2])[::-1].astype(np.float32)
        assert np.array_equal(h1, expect)


def test_resampler_plans(product_lib, ref):
    rates = [0.08333333, 0.0125, 0.24, 3.84, 8.0, 200000 / 61.44e6, 6000 / 40e6, 5400 / 61.44e6, 48000 / 5400, 12500 / 20e6, 0.6, 1.5]
    for r in rates:
        it, S, ra, step = C.c_int(), C.c_int(), C.c_float(), C.c_uint32()
        pfb = np.zeros(256 * 14, np.float32); ms = np.zeros(16, np.int32)
        product_lib.L.sdr_design_resampler(C.c_float(r), 60.0, C.byref(it), C.byref(S), C.byref(ra), C.byref(step), _p(pfb), _p(ms))
        fc = min(np.float32(0.515) * np.float32(ra.value), np.float32(0.49))
        q = C.c_void_p(ref.L.ref_resamp_crcf_create(ra.value, 7, float(fc), 60.0, 256))
        n = 40
        x = np.zeros(n, np.complex64); x[0] = 1
        y = np.zeros(4 * n, np.complex64)
        ny = ref.L.ref_resamp_crcf_execute_block(q, _p(x), n, _p(y))
        pred = []
        phase = 0
        for k in range(n):                         # closed-form schedule the kernels use
            while phase < (1 << 24):
                pred.append(pfb[(phase >> 16) * 14 + k] if k < 14 else 0.0)
                phase += step.value
            phase -= 1 << 24
        assert ny == len(pred), r
        assert np.abs(np.array(pred, np.float32) - y.real[:ny]).max() < 3e-6, r
        ref.L.ref_resamp_crcf_destroy(q)
        assert list(ms[:S.value]) == ([10, 5] + [3] * 14)[:S.value]


def test_demod_filters(product_lib, ref):
    # AM DC-block FIR (firfilt_rrrf_create_dc_blocker(25, 30)) by impulse response
    f = C.c_void_p(ref.L.ref_firfilt_dc_blocker_create(25, 30.0))
    x = np.zeros(51, np.float32); x[0] = 1; y = np.zeros(51, np.float32)
    ref.L.ref_firfilt_block(f, _p(x), 51, _p(y)); ref.L.ref_firfilt_destroy(f)
    h = np.zeros(51, np.float32); product_lib.L.sdr_design_am_dc_block(_p(h))
    assert np.abs(h - y).max() < 1e-6 and abs(y[25] - 0.974644363) < 1e-6
    # Hilbert quadrature taps (firhilbf_create(5, 90)): imaginary impulse -> lsb output = quadrature taps
    hb = C.c_void_p(ref.L.ref_firhilb_create(5, 90.0))
    x = np.zeros(24, np.complex64); x[0] = 1j; l = np.zeros(24, np.float32); u = np.zeros(24, np.float32)
    ref.L.ref_firhilb_c2r_block(hb, _p(x), 24, _p(l), _p(u)); ref.L.ref_firhilb_destroy(hb)
    hq = np.zeros(10, np.float32); product_lib.L.sdr_design_hilbert_q(_p(hq))
    assert np.abs(l[1:20:2] - hq[::-1]).max() < 1e-6
    # Butterworth-6 low-pass (iirfilt_crcf_create_lowpass(6, 0.25)) by impulse response
    from scipy.signal import sosfilt
    iq = C.c_void_p(ref.L.ref_iirfilt_lowpass_create(6, 0.25))
    x = np.zeros(64, np.complex64); x[0] = 1; y = np.zeros(64, np.complex64)
    ref.L.ref_iirfilt_block(iq, _p(x), 64, _p(y)); ref.L.ref_iirfilt_destroy(iq)
    s = np.zeros(18, np.float32); product_lib.L.sdr_design_butter6_sos(_p(s))
    yy = sosfilt(s.reshape(3, 6).astype(np.float64), x.real.astype(np.float64))
    assert np.abs(yy - y.real).max() < 1e-6
    assert abs(y.real[0] - 0.0295882113) < 1e-7        # SURVEY.md A.8 probe


def test_fm_stereo_pilot_bandpass_is_the_reference_design_bit_for_bit(ref, product_lib):
    """iirfilt_crcf_create_prototype(CHEBY2, BANDPASS, SOS, 5, 19500 / bw, 19000 / bw, 1, 60) (ModemFMStereo.cpp:130-136): poles
    0.002 inside the unit circle, so the product's host-side design must give the reference's sections to the last bit -- over the
    whole range of modem rates -- and the Hilbert taps of firhilbf_create(5, 60)."""
    vp, f, u = C.c_void_p, C.c_float, C.c_uint
    ref.L.ref_iirdes.argtypes = [C.c_int, C.c_int, C.c_int, u, f, f, f, f, vp, vp]
    for bw in list(range(100000, 400001, 2477)) + [200000, 250000, 192000, 150000, 100000, 256000]:
        b = np.zeros(30, np.float32); a = np.zeros(30, np.float32)
        f0 = np.float32(19000) / np.float32(bw); fc = np.float32(19500) / np.float32(bw)
        assert ref.L.ref_iirdes(2, 2, 0, 5, fc, f0, 1.0, 60.0, _p(b), _p(a)) == 0
        o = np.zeros(30, np.float32)
        assert product_lib.L.sdr_design_cheby2_bp5_sos(fc, f0, _p(o)) == 0
        want = np.concatenate([np.concatenate([b[3 * i:3 * i + 3], a[3 * i:3 * i + 3]]) for i in range(5)])
        assert np.array_equal(want.view(np.uint32), o.view(np.uint32)), bw
    ref.L.ref_firhilb_r2c_block.argtypes = [vp, vp, u, vp]
    hb = C.c_void_p(ref.L.ref_firhilb_create(5, 60.0))
    x = np.zeros(24, np.float32); x[0] = 1; y = np.zeros(24, np.complex64)
    ref.L.ref_firhilb_r2c_block(hb, _p(x), 24, _p(y)); ref.L.ref_firhilb_destroy(hb)
    hq = np.zeros(10, np.float32); product_lib.L.sdr_design_hilbert_q_as(60.0, _p(hq))
    assert np.abs(y.imag[1:20:2] - hq[::-1]).max() < 1e-6 and y.real[10] == 1.0
