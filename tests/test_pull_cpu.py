"""CPU: the low-latency pull path -- the product's request-level frame manager
(nvspeechplayer_b200/csrc/pull_manager.h, reference src/frame.cpp:41-115 jumped from event to event) and the arithmetic
of klatt_pull_kernel (klatt_pull_core.cuh, emulated thread by thread by tests/hostsim) against the golden output of the
reference.  Sample counts, drain behaviour and getLastIndex must be identical; PCM meets the FP32 bar.  The CUDA kernel
itself (block scans included) is checked by tests/test_gpu_pull.py on the B200."""
import numpy as np
import pytest

from nvspeechplayer_b200 import workloads
from tests import parity, scenarios
from tests.hostsim import sim


def _player(sr, **kw):
    return sim.PullPlayer(sr, seed=scenarios.SEED, stream=scenarios.STREAM, **kw)


@pytest.mark.parametrize("name", sorted(scenarios.all_scenarios().keys()))
def test_pull_scenarios(golden_scenarios, name):
    sc = scenarios.all_scenarios()[name]
    pcm, counts, idx = scenarios.run_script(_player, sc)
    assert counts == list(golden_scenarios[name + "/counts"])
    assert idx == list(golden_scenarios[name + "/last_index"])
    parity.assert_f32_parity(pcm, golden_scenarios[name + "/pcm"], name)


@pytest.mark.parametrize("name", ["chunk_777", "purge_mid_fade", "drain_resume", "null_first", "fade_gt_hold"])
def test_pull_launch_limits(golden_scenarios, name):
    """A pull cut into several launches (tick limit) and launches cut by the segment limit render the same stream."""
    sc = scenarios.all_scenarios()[name]
    pcm, counts, idx = scenarios.run_script(lambda sr: _player(sr, max_ticks=300, max_segs=2), sc)
    assert counts == list(golden_scenarios[name + "/counts"])
    assert idx == list(golden_scenarios[name + "/last_index"])
    parity.assert_f32_parity(pcm, golden_scenarios[name + "/pcm"], name)


def test_pull_config1(golden_config1):
    """sampleIpa.txt as the NVDA audio thread pulls it: 8192 samples at a time."""
    g = golden_config1
    p = sim.PullPlayer(int(g["sample_rate"]), seed=int(g["philox_seed"]), stream=int(g["philox_stream"]))
    nul = g["is_null"]
    for j in range(len(g["min_dur"])):
        p.queue_frame(None if nul[j] else g["frames"][j], int(g["min_dur"][j]), int(g["fade_dur"][j]))
    chunks = []
    while True:
        c = p.synthesize(8192)
        if c.size == 0:
            break
        chunks.append(c.copy())
    p.close()
    pcm = np.concatenate(chunks)
    assert pcm.size == 288454
    w1, exact, snr, mx = parity.assert_f32_parity(pcm, g["pcm_philox"], "config1 via pulls")
    assert snr >= 80.0


@pytest.mark.parametrize("sr,pull", [(16000, 8192), (22050, 2048), (44100, 8192), (22050, 333)])
def test_pull_random_frames(port, sr, pull):
    secs = 2.0
    fr, m, f, nul, ux = workloads.random_stream(777, secs, sr)
    n = int(secs * sr)
    want = port.render(sr, fr, m, f, nul, ux, max_samples=n, noise=("philox", 9, 777))
    p = sim.PullPlayer(sr, seed=9, stream=777)
    for j in range(len(m)):
        p.queue_frame(None if nul[j] else fr[j], int(m[j]), int(f[j]), int(ux[j]))
    got = []
    left = n
    while left > 0:
        c = p.synthesize(min(pull, left))
        assert c.size == min(pull, left)
        got.append(c.copy())
        left -= c.size
    p.close()
    parity.assert_f32_parity(np.concatenate(got), want, "random@%d pulls of %d" % (sr, pull), within=0.9995)
