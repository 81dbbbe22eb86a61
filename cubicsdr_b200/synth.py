"""Synthetic wideband IQ for parity tests and bench.py.

Sum of modulated carriers at the channel centres + complex AWGN at -40 dBFS + a DC offset of (0.01, -0.02) so that
the DC blocker is exercised.  Total amplitude ~0.5 full scale.  Everything is generated with numpy's counter-based
Philox generator from `seed`, so the oracle and the engine see identical bytes.
"""
import numpy as np

MODEM_DEFAULT_BW = {"FM": 200000, "NBFM": 12500, "AM": 6000, "USB": 5400, "LSB": 5400, "FMS": 200000}


def block_size(sample_rate):
    """Reference block length (SDRThread::getOptimalElementCount / getOptimalChannelCount, SoapySDRThread.cpp:668-695):
    ceil(floor(fs/60)/nch)*nch with nch = ceil(fs/500000) rounded DOWN to even, at least 2 (1 when fs <= 500 kHz)."""
    if sample_rate <= 500000:
        nch = 1
    else:
        nch = int(np.ceil(sample_rate / 500000.0))
        if nch % 2:
            nch -= 1
        if nch < 2:
            nch = 2
    n = int(sample_rate // 60)
    return int(np.ceil(n / nch) * nch)


def channel_plan(kind, sample_rate, n_chan):
    """Channel list [(modem, bandwidth, offset_hz)] for the BASELINE.json configs."""
    chans = []
    if kind == "nbfm64":            # cfg3: 64 NBFM on a 12.5 kHz raster
        for k in range(n_chan):
            chans.append(("NBFM", 12500, int(-393750 + k * 12500)))
    elif kind == "mixed":           # cfg4 / cfg5: FM / AM / USB / LSB round robin on a raster that keeps FM channels clear
        raster = 150000 if sample_rate <= 40e6 else 110000
        order = ["FM", "AM", "USB", "LSB"]
        start = -(n_chan // 2) * raster
        for k in range(n_chan):
            m = order[k % 4]
            chans.append((m, MODEM_DEFAULT_BW[m], int(start + k * raster)))
    elif kind == "wbfm1":           # cfg1
        chans.append(("FM", 200000, 300000))
    else:
        raise ValueError(kind)
    lim = sample_rate // 2
    return [c for c in chans if abs(c[2]) <= lim]


def wideband(sample_rate, n_samples, chans, seed=1234, t0=0, dtype=np.complex64):
    """n_samples of the synthetic stream starting at absolute sample index t0."""
    rng = np.random.Generator(np.random.Philox(key=seed, counter=[0, 0, 0, t0 & 0xFFFFFFFFFFFFFFFF]))
    n = np.arange(t0, t0 + n_samples, dtype=np.float64)
    t = n / float(sample_rate)
    x = np.zeros(n_samples, np.complex128)
    amp = 0.5 / max(1.0, np.sqrt(len(chans)) * 1.5) if chans else 0.0
    for k, (modem, bw, off) in enumerate(chans):
        tone = 1000.0 + (400.0 if (k & 1) else 0.0)
        if modem == "FMS":
            # FM stereo multiplex: 0.45 (L + R) + 0.1 pilot + 0.45 (L - R) sin(2 wp t), L = sin(w1 t), R = 0.7 sin(w2 t), 75 kHz
            # deviation; the phase is the closed-form integral, so blocks generated separately join up
            w1, w2, wp = 2 * np.pi * tone, 2 * np.pi * (tone + 1650.0), 2 * np.pi * 19000.0
            A, P, g = 0.45, 0.1, 0.7
            integ = A * (-np.cos(w1 * t) / w1 - g * np.cos(w2 * t) / w2) - P * np.cos(wp * t) / wp
            for wk, gk in ((w1, 1.0), (w2, -g)):
                integ += A * gk * 0.5 * (np.sin((2 * wp - wk) * t) / (2 * wp - wk) - np.sin((2 * wp + wk) * t) / (2 * wp + wk))
            x += amp * np.exp(1j * (2 * np.pi * off * t + 2 * np.pi * 75000.0 * integ))
        elif modem in ("FM", "NBFM"):
            dev = 75000.0 if modem == "FM" else 2500.0
            ph = 2 * np.pi * off * t + (dev / tone) * np.sin(2 * np.pi * tone * t)
            x += amp * np.exp(1j * ph)
        elif modem == "AM":
            env = 1.0 + 0.8 * np.sin(2 * np.pi * tone * t)
            x += amp * 0.5 * env * np.exp(2j * np.pi * off * t)
        else:
            sgn = 1.0 if modem == "USB" else -1.0
            x += amp * 0.5 * (np.exp(2j * np.pi * (off + sgn * 700.0) * t) + np.exp(2j * np.pi * (off + sgn * 1900.0) * t))
    sigma = 0.01 / np.sqrt(2.0)
    x += sigma * (rng.standard_normal(n_samples) + 1j * rng.standard_normal(n_samples))
    x += (0.01 - 0.02j)
    return x.astype(dtype)
