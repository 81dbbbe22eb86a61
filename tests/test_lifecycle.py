"""Channel lifecycle next to a running stream: DemodulatorWorkerThread / DemodulatorMgr create and delete demodulators on
their own threads while SDRPostThread keeps pushing blocks (src/demod/DemodulatorWorkerThread.cpp:63-108,
src/sdr/SDRPostThread.cpp:238-298: the post thread only try_push()es and is never made to wait).

One thread submits blocks back to back; a second one creates and destroys channels as fast as it can.  The channels that
live through the whole run must produce exactly the audio of an undisturbed run, and (GPU tier) no submit call may take
long: creation and destruction use the stream-ordered allocator and fenced recycling instead of device-wide waits."""
import threading
import time

import numpy as np
import pytest

from cubicsdr_b200 import Channel, Engine, synth
from tests.parity import CENTER


def _run(lib, fs, n_blocks, keep, churn, with_churn):
    N = synth.block_size(fs)
    blocks = [synth.wideband(fs, N, keep, seed=99, t0=b * N) for b in range(4)]
    eng = Engine(lib, max_block=N)
    cs = [Channel(eng, m, fs, bw, 48000, frequency=CENTER + off) for m, bw, off in keep]
    stop = threading.Event()
    made = [0]
    errors = []

    def churner():
        try:
            k = 0
            while not stop.is_set():
                m, bw, off = churn[k % len(churn)]
                live = [Channel(eng, m, fs, bw, 48000, frequency=CENTER + off + 1000 * j) for j in range(3)]
                time.sleep(0.0005)
                for c in live:
                    c.close()
                made[0] += 3
                k += 1
        except Exception as ex:      # surfaced by the main thread
            errors.append(ex)

    th = threading.Thread(target=churner)
    if with_churn:
        th.start()
    lat, audio = [], [[] for _ in cs]
    for b in range(n_blocks):
        t0 = time.perf_counter()
        eng.submit(blocks[b % 4], CENTER, fs)
        lat.append(time.perf_counter() - t0)
        eng.wait_block(0)
        for i, c in enumerate(cs):
            audio[i].append(c.fetch_audio()[0])
    stop.set()
    if with_churn:
        th.join()
    for c in cs:
        c.close()
    eng.close()
    assert not errors, errors
    return [np.concatenate(a) for a in audio], np.array(lat), made[0]


KEEP = [("FM", 200000, 300000), ("AM", 6000, -450000), ("USB", 5400, 150000)]
CHURN = [("NBFM", 12500, -200000), ("FM", 200000, 600000), ("LSB", 5400, -700000)]


def test_create_destroy_while_streaming_emulator(emu_lib):
    quiet, _, _ = _run(emu_lib, 2400000, 12, KEEP, CHURN, False)
    busy, _, made = _run(emu_lib, 2400000, 12, KEEP, CHURN, True)
    assert made >= 3
    for a, b in zip(quiet, busy):
        assert np.array_equal(a, b)


@pytest.mark.gpu
def test_create_destroy_while_streaming_gpu(gpu_lib):
    fs = 10000000
    keep = synth.channel_plan("mixed", fs, 16)
    quiet, lat0, _ = _run(gpu_lib, fs, 150, keep, CHURN, False)
    busy, lat1, made = _run(gpu_lib, fs, 150, keep, CHURN, True)
    assert made >= 30, made                                     # the churn thread really ran next to the stream
    for a, b in zip(quiet, busy):
        assert np.array_equal(a, b)
    # the submitter is never held up by channel creation / destruction: no device-wide wait, no long lock
    lat1 = np.sort(lat1[5:])
    print("submit latency ms: undisturbed median %.3f max %.3f | with churn median %.3f p99 %.3f max %.3f (%d channels made)" % (
        1e3 * np.median(lat0[5:]), 1e3 * lat0[5:].max(), 1e3 * np.median(lat1), 1e3 * lat1[int(0.99 * len(lat1)) - 1], 1e3 * lat1[-1], made))
    # (bounds leave room for the interpreter's 5 ms thread switch interval: the churn thread is Python too)
    assert np.median(lat1) < 1.5e-3 and lat1[int(0.99 * len(lat1)) - 1] < 10e-3 and lat1[-1] < 50e-3
