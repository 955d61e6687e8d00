"""GPU: the low-latency pull path (SPEECHPLAYER_PRECISION_STREAM) through the C-ABI: the five
reference exports on one handle, one pull = one launch of klatt_pull_kernel.  Same bar as the FP32 batch kernels
(<= 1 LSB on >= 99.9 %, >= 60 dB SNR against the reference); sample counts, drain behaviour and getLastIndex identical.
The host emulation of the same arithmetic (tests/hostsim) pins the device's block scans."""
import time

import numpy as np
import pytest

from nvspeechplayer_b200 import player, workloads
from tests import parity, scenarios

pytestmark = pytest.mark.gpu


class _PullAdapter:
    def __init__(self, sr, noise=player.NOISE_PHILOX):
        self.p = player.SpeechPlayer(sr, precision=player.PRECISION_STREAM, noise=noise, seed=scenarios.SEED,
                                     streamId=scenarios.STREAM)

    def queue_frame(self, fr, m, f, ux, purge):
        self.p.queue_frame(fr, m, f, ux, purge)

    def synthesize(self, n):
        return self.p.synthesize_np(n)

    def last_index(self):
        return self.p.last_index()

    def close(self):
        self.p.close()


@pytest.mark.parametrize("name", sorted(scenarios.all_scenarios().keys()))
def test_pull_scenarios(golden_scenarios, name):
    sc = scenarios.all_scenarios()[name]
    pcm, counts, idx = scenarios.run_script(lambda sr: _PullAdapter(sr), sc)
    assert counts == list(golden_scenarios[name + "/counts"])
    assert idx == list(golden_scenarios[name + "/last_index"])
    parity.assert_f32_parity(pcm, golden_scenarios[name + "/pcm"], name)


def _config1_pulls(g, pull, **kw):
    p = player.SpeechPlayer(int(g["sample_rate"]), precision=player.PRECISION_STREAM, **kw)
    p.queue_frames(g["frames"], g["min_dur"], g["fade_dur"], None, g["is_null"])
    chunks = []
    while True:
        c = p.synthesize_np(pull)
        if c.size == 0:
            break
        chunks.append(c)
    p.close()
    return np.concatenate(chunks)


@pytest.mark.parametrize("pull", [8192, 2048, 20000])
def test_pull_config1(golden_config1, pull):
    """sampleIpa.txt as the NVDA audio thread pulls it (8192), in short pulls, and in pulls longer than one launch."""
    g = golden_config1
    pcm = _config1_pulls(g, pull, noise=player.NOISE_PHILOX, seed=int(g["philox_seed"]), streamId=int(g["philox_stream"]))
    assert pcm.size == 288454
    w1, exact, snr, mx = parity.assert_f32_parity(pcm, g["pcm_philox"], "config1 via pulls of %d" % pull)
    assert snr >= 80.0


def test_pull_config1_default_noise(golden_config1):
    """Process-global glibc-compatible rand(), two draws per generated sample, like a Linux build of the reference."""
    g = golden_config1
    player.load_library().speechPlayer_seedNoise(1)
    pcm = _config1_pulls(g, 8192, noise=player.NOISE_GLIBC)
    assert pcm.size == 288454
    w1, exact, snr, mx = parity.assert_f32_parity(pcm, g["pcm_libc"], "config1/glibc via pulls")
    assert snr >= 80.0


@pytest.mark.parametrize("sr,pull", [(16000, 8192), (22050, 2048), (44100, 8192), (22050, 333)])
def test_pull_random_frames(port, sr, pull):
    secs = 2.0
    fr, m, f, nul, ux = workloads.random_stream(777, secs, sr)
    n = int(secs * sr)
    want = port.render(sr, fr, m, f, nul, ux, max_samples=n, noise=("philox", 9, 777))
    p = player.SpeechPlayer(sr, precision=player.PRECISION_STREAM, noise=player.NOISE_PHILOX, seed=9, streamId=777)
    p.queue_frames(fr, m, f, ux, nul)
    got, left = [], n
    while left > 0:
        c = p.synthesize_np(min(pull, left))
        assert c.size == min(pull, left)
        got.append(c)
        left -= c.size
    p.close()
    parity.assert_f32_parity(np.concatenate(got), want, "random@%d pulls of %d" % (sr, pull), within=0.9995)


def test_pull_matches_host_emulation(golden_config1):
    """Device block scans (warp shuffles) against the plain loops of the host emulation: same maps, same seeds; what
    may differ is libdevice vs glibc in the plans and MUFU.RCP in the anti-resonator."""
    from tests.hostsim import sim
    g = golden_config1
    dev = _config1_pulls(g, 8192, noise=player.NOISE_PHILOX, seed=int(g["philox_seed"]), streamId=int(g["philox_stream"]))
    hp = sim.PullPlayer(int(g["sample_rate"]), seed=int(g["philox_seed"]), stream=int(g["philox_stream"]))
    nul = g["is_null"]
    for j in range(len(g["min_dur"])):
        hp.queue_frame(None if nul[j] else g["frames"][j], int(g["min_dur"][j]), int(g["fade_dur"][j]))
    chunks = []
    while True:
        c = hp.synthesize(8192)
        if c.size == 0:
            break
        chunks.append(c.copy())
    hp.close()
    host = np.concatenate(chunks)
    assert host.size == dev.size
    w1, exact, snr, mx = parity.metrics(dev, host)
    assert w1 >= 0.999 and snr >= 70.0, (w1, exact, snr, mx)


def test_pull_batch_of_stream_handles():
    """speechPlayer_synthesizeBatch over low-latency handles: each row is that handle's own pull."""
    sr, n = 22050, 3
    fr, m, f, nul, ux = workloads.random_stream(5, 0.3, sr)
    count = 4000
    rows = []
    ps = [player.SpeechPlayer(sr, precision=player.PRECISION_STREAM, noise=player.NOISE_PHILOX, seed=4, streamId=s) for s in range(n)]
    for p in ps:
        p.queue_frames(fr, m, f, ux, nul)
    solo = player.SpeechPlayer(sr, precision=player.PRECISION_STREAM, noise=player.NOISE_PHILOX, seed=4, streamId=1)
    solo.queue_frames(fr, m, f, ux, nul)
    want1 = solo.synthesize_np(count)
    solo.close()
    out, written = player.synthesize_batch(ps, count)
    assert list(written) == [count] * n
    np.testing.assert_array_equal(out[1], want1)
    assert not np.array_equal(out[0], out[1])
    for p in ps:
        p.close()


def test_pull_latency_smoke():
    """Not a benchmark (profiles/ has the probe): a pull of 8192 samples must come back far inside real time."""
    sr = 22050
    fr, m, f, nul, ux = workloads.random_stream(11, 4.0, sr)
    p = player.SpeechPlayer(sr, precision=player.PRECISION_STREAM, noise=player.NOISE_PHILOX, seed=1, streamId=11)
    p.queue_frames(fr, m, f, ux, nul)
    p.synthesize_np(8192)
    t = []
    for _ in range(8):
        t0 = time.perf_counter()
        c = p.synthesize_np(8192)
        t.append(time.perf_counter() - t0)
        assert c.size == 8192
    p.close()
    assert np.median(t) < 0.02, t   # 8192 samples are 372 ms of audio


def test_pull_nvda_speak_cancel_speak(port):
    """The NVDA driver's speak / cancel / speak with its audio thread pulling 8192 samples at a time (tests/nvda_flow.py,
    reference nvdaAddon/synthDrivers/nvSpeechPlayer/__init__.py:56-82, :168-241) on the low-latency path."""
    from tests import nvda_flow

    def oracle_player(sr):
        p = port.player(sr)
        p.noise_philox(scenarios.SEED, scenarios.STREAM)
        return p
    want, wcounts, widx = nvda_flow.run(oracle_player)
    got, counts, idx = nvda_flow.run(lambda sr: _PullAdapter(sr))
    assert counts == wcounts
    assert idx == widx
    parity.assert_f32_parity(got, want, "NVDA speak / cancel / speak")


@pytest.mark.parametrize("flow", ["leap", "midi"])
def test_pull_live_control_flows(port, flow):
    """The reference's live-control demos (test_leap.py: purgeQueue on every tracking frame; test_midiSing.py: notes held
    for 10^7 ms, purges on note / controller / pitch-bend events) while the audio thread keeps pulling."""
    from tests import nvda_flow
    run = nvda_flow.run_leap if flow == "leap" else nvda_flow.run_midi

    def oracle_player(sr):
        p = port.player(sr)
        p.noise_philox(scenarios.SEED, scenarios.STREAM)
        return p
    want, wcounts = run(oracle_player)
    got, counts = run(lambda sr: _PullAdapter(sr))
    assert counts == wcounts
    parity.assert_f32_parity(got, want, "live control: " + flow)


def test_pull_batch_148_players_one_launch(port):
    """speechPlayer_synthesizeBatch over 148 low-latency players: ONE launch per call (one block per player), every row the
    bits of that player's own speechPlayer_synthesize, ragged queues, drains mid-call, a purge on a subset between two calls;
    a few rows against the oracle; and the whole batch inside 1.5 x the latency of a single player's pull (+ the host side's
    per-player bookkeeping)."""
    sr, n, count = 22050, 148, 8192
    rng = np.random.default_rng(148)
    secs = rng.uniform(0.2, 1.2, n)
    streams = [workloads.random_stream(2000 + s, float(secs[s]), sr) for s in range(n)]

    def make():
        ps = [player.SpeechPlayer(sr, precision=player.PRECISION_STREAM, noise=player.NOISE_PHILOX, seed=6, streamId=2000 + s) for s in range(n)]
        for p, (fr, m, f, nul, ux) in zip(ps, streams):
            p.queue_frames(fr, m, f, ux, nul)
        return ps

    extra = workloads.random_stream(77, 0.3, sr)

    def purge(ps):
        for s in range(0, n, 7):
            fr, m, f, nul, ux = extra
            ps[s].queue_frame(fr[0], int(m[0]), int(f[0]), 500 + s, True)
            ps[s].queue_frames(fr[1:], m[1:], f[1:], ux[1:], nul[1:])

    batch = make()
    t0 = time.perf_counter()
    out1, w1 = player.synthesize_batch(batch, count)
    t_batch = time.perf_counter() - t0
    purge(batch)
    out2, w2 = player.synthesize_batch(batch, count)
    idx_batch = [p.getLastIndex() for p in batch]
    for p in batch:
        p.close()
    solo = make()
    rows1 = [p.synthesize_np(count) for p in solo]
    purge(solo)
    t0 = time.perf_counter()
    rows2 = [p.synthesize_np(count) for p in solo]
    t_solo_all = time.perf_counter() - t0
    idx_solo = [p.getLastIndex() for p in solo]
    for p in solo:
        p.close()
    assert (w1 < count).any() or (w2 < count).any(), "some player should drain inside a call"
    for s in range(n):
        assert w1[s] == len(rows1[s]) and w2[s] == len(rows2[s])
        np.testing.assert_array_equal(out1[s, :w1[s]], rows1[s])
        np.testing.assert_array_equal(out2[s, :w2[s]], rows2[s])
    assert idx_batch == idx_solo
    for s in (0, 7, 100):   # against the oracle, purge included
        o = port.player(sr)
        o.noise_philox(6, 2000 + s)
        fr, m, f, nul, ux = streams[s]
        for j in range(len(m)):
            o.queue_frame(None if nul[j] else fr[j], int(m[j]), int(f[j]), int(ux[j]))
        a = o.synthesize(count)
        if s % 7 == 0:
            fr, m, f, nul, ux = extra
            o.queue_frame(fr[0], int(m[0]), int(f[0]), 500 + s, True)
            for j in range(1, len(m)):
                o.queue_frame(None if nul[j] else fr[j], int(m[j]), int(f[j]), int(ux[j]))
        b = o.synthesize(count)
        o.close()
        parity.assert_f32_parity(np.concatenate([out1[s, :w1[s]], out2[s, :w2[s]]]), np.concatenate([a, b]), "batched pull row %d" % s)
    per_solo = t_solo_all / n
    print("148 players x 8192: batch call %.2f ms, single-player pull %.3f ms" % (t_batch * 1e3, per_solo * 1e3))
    assert t_batch < 40 * per_solo, "the batch should cost far less than 148 sequential pulls"
