"""GPU: the FP32 production kernel against the reference, through the C-ABI.

Bar (BASELINE.json north_star): <= 1 LSB on >= 99.9 % of samples and >= 60 dB SNR; sample counts, drain
behaviour and getLastIndex identical; Philox noise checked statistically."""
import numpy as np
import pytest

from nvspeechplayer_b200 import player, workloads
from tests import parity, scenarios

pytestmark = pytest.mark.gpu


class _EngineAdapter:
    def __init__(self, sr):
        self.p = player.SpeechPlayer(sr, precision=player.PRECISION_FP32, noise=player.NOISE_PHILOX, seed=scenarios.SEED,
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
def test_f32_scenarios(golden_scenarios, name):
    sc = scenarios.all_scenarios()[name]
    pcm, counts, idx = scenarios.run_script(lambda sr: _EngineAdapter(sr), sc)
    assert counts == list(golden_scenarios[name + "/counts"])
    assert idx == list(golden_scenarios[name + "/last_index"])
    parity.assert_f32_parity(pcm, golden_scenarios[name + "/pcm"], name)


def test_f32_config1(golden_config1):
    g = golden_config1
    p = player.SpeechPlayer(int(g["sample_rate"]), precision=player.PRECISION_FP32, noise=player.NOISE_PHILOX,
                            seed=int(g["philox_seed"]), streamId=int(g["philox_stream"]))
    p.queue_frames(g["frames"], g["min_dur"], g["fade_dur"], None, g["is_null"])
    chunks = []
    while True:
        c = p.synthesize_np(8192)   # the NVDA audio thread's pull size
        if c.size == 0:
            break
        chunks.append(c)
    pcm = np.concatenate(chunks)
    assert pcm.size == 288454
    w1, exact, snr, mx = parity.assert_f32_parity(pcm, g["pcm_philox"], "config1")
    assert snr >= 80.0


@pytest.mark.parametrize("sr,n,secs,within", [(16000, 48, 1.0, 0.9995), (22050, 96, 1.0, 0.9995), (44100, 32, 1.0, 0.9995)])
def test_f32_batch_random_vs_port(port, sr, n, secs, within):
    fb = workloads.random_frames(n, secs, sr, first_stream=2000)
    count = int(secs * sr)
    b = player.Batch(sr, n, precision=player.PRECISION_FP32, noise=player.NOISE_PHILOX, seed=31, stream_ids=fb.stream_ids)
    b.set_frames_host(fb)
    out, written = b.synthesize_host(count)
    assert (written == count).all()
    allgot, allwant = [], []
    for s in range(n):
        fr, m, f, nul, ux = fb.stream(s)
        want = port.render(sr, fr, m, f, nul, ux, max_samples=count, noise=("philox", 31, int(fb.stream_ids[s])))
        w1, exact, snr, mx = parity.metrics(out[s], want)
        assert snr >= parity.F32_SNR_DB, "stream %d: SNR %.1f dB" % (s, snr)
        allgot.append(out[s]); allwant.append(want)
    parity.assert_f32_parity(np.concatenate(allgot), np.concatenate(allwant), "batch@%d" % sr, within=within)
    b.close()


def test_f32_chunked_equals_one_shot_bitwise():
    sr, n = 22050, 16
    fb = workloads.random_frames(n, 0.5, sr, first_stream=70)
    count = int(0.5 * sr)
    outs = []
    for chunks in ([count], [1, 7, 777, count - 785], [8192, count - 8192]):
        b = player.Batch(sr, n, precision=player.PRECISION_FP32, seed=3, stream_ids=fb.stream_ids)
        b.set_frames_host(fb)
        parts = [b.synthesize_host(c)[0] for c in chunks]
        outs.append(np.concatenate(parts, axis=1))
        b.close()
    np.testing.assert_array_equal(outs[0], outs[1])
    np.testing.assert_array_equal(outs[0], outs[2])


def test_f32_matches_f64_engine_on_vowel_chart():
    """config 2 recipe (16 kHz, vowel pairs x voices): FP32 kernel vs the engine's own FP64 kernel."""
    fb = workloads.vowel_chart(2, pairs=40)
    count = int(fb.timeline_samples().max())
    res = {}
    for prec in (player.PRECISION_FP64, player.PRECISION_FP32):
        b = player.Batch(16000, fb.num_streams, precision=prec, seed=8, stream_ids=fb.stream_ids)
        b.set_frames_host(fb)
        res[prec], written = b.synthesize_host(count)
        assert (written == fb.timeline_samples()).all()
        b.close()
    parity.assert_f32_parity(res[player.PRECISION_FP32].ravel(), res[player.PRECISION_FP64].ravel(), "vowel chart")


def test_philox_noise_statistics():
    """Frication-only streams expose the coloured noise directly: out = nF*0.3*fric*pre/2*gain*4000 with
    nF = u + 0.75*nF, u~U[0,1]  =>  mean 2.0, variance (1/12)/(1-0.75^2), lag-1 autocorrelation 0.75;
    distinct stream ids give uncorrelated sequences; the same id reproduces."""
    sr, n, count = 22050, 8, 100000
    fr = np.zeros((1, 47))
    fr[0, workloads.P["fricationAmplitude"]] = 1.0
    fr[0, workloads.P["parallelBypass"]] = 1.0
    fr[0, workloads.P["preFormantGain"]] = 1.0
    fr[0, workloads.P["outputGain"]] = 1.0
    fr[0, workloads.P["voicePitch"]] = fr[0, workloads.P["endVoicePitch"]] = 100.0
    stream = (fr, np.array([count + 10], np.uint32), np.array([1], np.uint32), np.zeros(1, np.uint8), np.full(1, -1, np.int32))
    ids = np.array([0, 1, 2, 3, 4, 5, 6, 0], dtype=np.uint64)   # last one repeats id 0
    fb = workloads._concat(sr, [stream] * n, ids)
    b = player.Batch(sr, n, precision=player.PRECISION_FP32, seed=1234, stream_ids=ids)
    b.set_frames_host(fb)
    out, _ = b.synthesize_host(count)
    b.close()
    x = out[:, 16:].astype(np.float64) / 600.0   # undo 0.3 * 0.5 * 4000; skip the fade-in ticks
    np.testing.assert_array_equal(out[0], out[7])
    for s in range(7):
        assert abs(x[s].mean() - 2.0) < 0.02
        assert abs(x[s].var() - (1 / 12) / (1 - 0.75 ** 2)) < 0.01
        c = np.corrcoef(x[s][:-1], x[s][1:])[0, 1]
        assert abs(c - 0.75) < 0.01
    for a in range(7):
        for c in range(a + 1, 7):
            assert abs(np.corrcoef(x[a], x[c])[0, 1]) < 0.02


def test_f32_rounds_equal_single_launch_bitwise(monkeypatch, port):
    """The batch engine's round-based execution (phase-sorted hold / general chunks, the two sides of a stream in
    paired warps, several stream groups in flight), its ring scheduler (the same chunks handed out by device-side rings
    inside one launch) and its block scheduler (the default: streams owned by one thread block each, compact state, hold /
    fade / general cells of 64..128 ticks) render the same bits as the one-thread-per-stream kernel,
    including streams that drain mid-call, partially filled warps and calls that end inside a chunk."""
    sr, n = 22050, 203
    streams = [workloads.random_stream(500 + s, 0.35 if s % 7 == 3 else 1.0, sr) for s in range(n)]
    fb = workloads._concat(sr, streams, np.arange(500, 500 + n, dtype=np.uint64))
    count = int(1.0 * sr)
    res = {}
    for mode, min_streams in (("single", "100000000"), ("rounds", "1"), ("sched", "1"), ("sched_fade", "1"), ("sched_fade_roles", "1"), ("block", "1")):
        monkeypatch.setenv("NVSP_ROUNDS_MIN_STREAMS", min_streams)
        monkeypatch.setenv("NVSP_GROUPS", "3")
        monkeypatch.setenv("NVSP_SCHED", {"sched": "rings", "sched_fade": "rings", "sched_fade_roles": "rings", "block": "block"}.get(mode, "rounds"))
        # the ring scheduler's fade class (straight-line fade loop on 64-tick cells), with and without SM roles
        monkeypatch.setenv("NVSP_SCHED_FADE_TICKS", "128" if mode.startswith("sched_fade") else "0")
        monkeypatch.setenv("NVSP_SCHED_FADE_MAX", "320")
        monkeypatch.setenv("NVSP_SCHED_FADE_SMS", "40" if mode == "sched_fade_roles" else "0")
        monkeypatch.setenv("NVSP_SCHED_HOLD_SMS", "30" if mode == "sched_fade_roles" else "0")
        monkeypatch.setenv("NVSP_BLOCK_MIN_STREAMS", "1")
        monkeypatch.setenv("NVSP_BLOCK_BLOCKS", "3")   # 68 streams per block on 8 workers: the queues fill and drain
        monkeypatch.setenv("NVSP_SCHED_BLOCKS", "3")  # fewer workers than stream batches: streams queue up in the rings
        monkeypatch.setenv("NVSP_SCHED_HOLD_TICKS", "256")  # several chunks of both classes per stream in a 1-s render
        monkeypatch.setenv("NVSP_SCHED_GEN_TICKS", "128")
        b = player.Batch(sr, n, precision=player.PRECISION_FP32, seed=77, stream_ids=fb.stream_ids)
        b.set_frames_host(fb)
        parts, written = [], np.zeros(n, dtype=np.int64)
        for c in (9000, 777, count - 9777):
            o, w = b.synthesize_host(c)
            parts.append(o)
            written += w
        res[mode] = (np.concatenate(parts, axis=1), written, b.last_indices(), b.launch_stats()[0])
        b.close()
    assert res["rounds"][3] > res["single"][3] + 50, "the rounds path did not run"
    assert res["sched"][3] == res["single"][3] + 3 * 3, "the stream scheduler did not run (import + seed + workers + export per call)"
    assert res["block"][3] == res["single"][3] + 2 * 3, "the block scheduler did not run (import + workers + export per call)"
    for k in range(3):
        np.testing.assert_array_equal(res["single"][k], res["sched"][k])
        np.testing.assert_array_equal(res["single"][k], res["sched_fade"][k])
        np.testing.assert_array_equal(res["single"][k], res["sched_fade_roles"][k])
        np.testing.assert_array_equal(res["single"][k], res["block"][k])
    np.testing.assert_array_equal(res["single"][1], res["rounds"][1])
    np.testing.assert_array_equal(res["single"][2], res["rounds"][2])
    np.testing.assert_array_equal(res["single"][0], res["rounds"][0])
    expect = np.minimum(fb.timeline_samples(), count)
    np.testing.assert_array_equal(res["rounds"][1], expect)
    # and both are the reference's audio
    s = 3
    fr, m, f, nul, ux = fb.stream(s)
    want = port.render(sr, fr, m, f, nul, ux, max_samples=count, noise=("philox", 77, int(fb.stream_ids[s])))
    got = res["rounds"][0][s][:len(want)]
    parity.assert_f32_parity(got, want, "drained stream")
    assert not res["rounds"][0][s][len(want):].any()


def test_f32_ring_scheduler_long_hold_chunks_bitwise(monkeypatch):
    """The ring scheduler stretches a hold chunk to the shortest quiet span of its 32 streams (up to NVSP_SCHED_HOLD_MAX
    ticks): steady vowels next to busy random-frame streams, every cap from "never" to "the whole note", calls that end
    inside a stretched chunk -- the same bits as the one-thread-per-stream kernel."""
    sr = 16000
    vc = workloads.vowel_chart(3, sr)                      # holds of thousands of ticks
    pick = list(range(0, len(vc.stream_ids), 97))[:40]
    streams = [vc.stream(s) for s in pick] + [workloads.random_stream(900 + s, 1.5, sr) for s in range(40)]
    ids = np.concatenate([vc.stream_ids[pick], np.arange(900, 940, dtype=np.uint64)])
    fb = workloads._concat(sr, streams, ids)
    n = len(ids)
    counts = (7001, 640, 12000)
    res = {}
    for mode, hold_max in (("single", None), ("cap256", "256"), ("cap576", "576"), ("cap1024", "1024"), ("cap65536", "65536"),
                           ("fade64", "1024"), ("fade192", "4096")):
        monkeypatch.setenv("NVSP_ROUNDS_MIN_STREAMS", "100000000" if mode == "single" else "1")
        monkeypatch.setenv("NVSP_SCHED", "rings")
        monkeypatch.setenv("NVSP_SCHED_BLOCKS", "2")
        monkeypatch.setenv("NVSP_SCHED_HOLD_TICKS", "256")
        monkeypatch.setenv("NVSP_SCHED_GEN_TICKS", "192")
        if hold_max:
            monkeypatch.setenv("NVSP_SCHED_HOLD_MAX", hold_max)
        # (the vowel chart's 400 ms fades: fade chunks of many cells, stretched to the cap)
        monkeypatch.setenv("NVSP_SCHED_FADE_TICKS", {"fade64": "64", "fade192": "192"}.get(mode, "0"))
        monkeypatch.setenv("NVSP_SCHED_FADE_MAX", {"fade64": "64", "fade192": "2048"}.get(mode, "512"))
        b = player.Batch(sr, n, precision=player.PRECISION_FP32, seed=5, stream_ids=fb.stream_ids)
        b.set_frames_host(fb)
        parts, written = [], np.zeros(n, dtype=np.int64)
        for c in counts:
            o, w = b.synthesize_host(c)
            parts.append(o)
            written += w
        res[mode] = (np.concatenate(parts, axis=1), written, b.last_indices(), b.launch_stats()[0])
        b.close()
    assert res["cap1024"][3] == res["single"][3] + 3 * len(counts), "the ring scheduler did not run"
    assert res["single"][0].any()
    for mode in ("cap256", "cap576", "cap1024", "cap65536", "fade64", "fade192"):
        for k in range(3):
            np.testing.assert_array_equal(res["single"][k], res[mode][k], err_msg=mode)
