"""The hold glide of the low-latency pull path (klatt_pull_core.cuh PullOsc::pitch, pull_manager.h syncToCounter): the reference
adds voicePitchInc to curFrame.voicePitch once per tick (src/frame.cpp:77).  A chunk of the time-parallel kernel starts in the
middle of a hold, so it seeks with glideExact and then adds like the reference: every tick's pitch must carry the reference's
roundings, bit for bit, wherever the chunk starts."""
import ctypes

import numpy as np
import pytest

from tests.hostsim import sim


def _reference_pitch(old, new, inc, F, ticks):
    """src/frame.cpp:49-52 (fade, src/utils.h:20-23), :44-47 (swap tick keeps the landing value), :76-78 (hold)."""
    out = np.empty(ticks, dtype=np.float64)
    cur = np.float64(old)
    o, n, inc = np.float64(old), np.float64(new), np.float64(inc)
    for c in range(ticks):
        if c == 0:
            cur = o                                            # pop tick: stale value
        elif c <= F:
            cur = o if np.isnan(n) else o + ((n - o) * (np.float64(min(c, F)) / np.float64(F)))
        elif c == F + 1:
            pass                                               # swap tick
        else:
            cur = cur + inc
        out[c] = cur
    return out


@pytest.mark.parametrize("old,new,inc,F", [
    (120.0, 133.7, 0.0011337868480725624, 110),     # 25 Hz/s glide at 22.05 kHz
    (99.99999999999999, 100.0, 1e-9, 1),            # increments far below an ulp of the sum ... that still add up
    (255.99999, 256.0, 3.0e-5, 64),                 # crosses a binade boundary during the hold
    (440.0, float("nan"), -0.0009070294784580499, 32),  # NaN target keeps the old pitch; falling glide
    (1e-3, 2e-3, 0.125, 8),                         # many binades
])
def test_pull_pitch_walk_is_the_reference_recurrence(old, new, inc, F):
    L = sim.lib()
    L.hostsim_pull_pitch_walk.restype = None
    L.hostsim_pull_pitch_walk.argtypes = [ctypes.c_double] * 3 + [ctypes.c_uint32] * 3 + [ctypes.c_void_p]
    ticks = 60000
    want = _reference_pitch(old, new, inc, F, ticks)
    rng = np.random.default_rng(F)
    starts = [0, 1, F, F + 1, F + 2, F + 3] + [int(x) for x in rng.integers(F + 2, ticks - 40, 40)]
    for c0 in starts:
        n = min(37, ticks - c0)
        got = np.empty(n, dtype=np.float64)
        L.hostsim_pull_pitch_walk(old, new, inc, F, c0, n, got.ctypes.data_as(ctypes.c_void_p))
        np.testing.assert_array_equal(got.view(np.uint64), want[c0:c0 + n].view(np.uint64), err_msg="chunk starting at tick %d" % c0)
