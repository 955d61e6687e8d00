"""GPU: the BASELINE.json workload families that are not the bench line -- config 2 (vowel-chart pairs x voices at
16 kHz, reference test_playVowelchart.py:27-43) and config 5 (midi-sing style pitch sweeps with vibrato, reference
test_midiSing.py:23-60,116-134) -- through the batch C-ABI in both precisions, against the plain-C oracle on seeded
subsets, plus size-independent properties on larger batches (the oracle would need minutes there):

  * a batch is the concatenation of its streams: rendering streams [a, b) alone gives the same bits as rows a..b of the
    full batch (sharding by stream);
  * chunked pulls equal one pull (the state block carries everything, reference src/speechWaveGenerator.cpp:184-192);
  * every stream produces exactly its timeline length (frame.cpp:41-80 occupancy law) and silence after it.

Bars: tests/parity.py (FP64: >= 99.99 % exact, rest <= 1 LSB; FP32: >= 99.9 % within 1 LSB and >= 60 dB SNR).
"""
import numpy as np
import pytest

from nvspeechplayer_b200 import player, workloads
from tests import parity
from tests.test_gpu_parity_f64 import assert_f64_parity

pytestmark = pytest.mark.gpu

SEED = 0xB200


def _render(fb, sr, count, precision, chunks=None, streams=None):
    n = fb.num_streams if streams is None else len(streams)
    if streams is not None:
        fb = workloads._concat(sr, [fb.stream(s) for s in streams], fb.stream_ids[list(streams)])
    b = player.Batch(sr, n, precision=precision, noise=player.NOISE_PHILOX, seed=SEED, stream_ids=fb.stream_ids)
    b.set_frames_host(fb)
    parts, written = [], np.zeros(n, dtype=np.int64)
    for c in (chunks or [count]):
        o, w = b.synthesize_host(c)
        parts.append(o)
        written += w
    b.close()
    return np.concatenate(parts, axis=1), written


def _oracle(port, fb, sr, s, count):
    fr, m, f, nul, ux = fb.stream(s)
    return port.render(sr, fr, m, f, nul, ux, max_samples=count, noise=("philox", SEED, int(fb.stream_ids[s])))


@pytest.mark.parametrize("precision", [player.PRECISION_FP64, player.PRECISION_FP32])
def test_config2_vowel_chart_vs_oracle(port, precision):
    sr = 16000
    fb = workloads.vowel_chart(3, sr, pairs=24)  # 72 (voice, pair) streams, 13 926 ticks each incl. the trailing silence
    count = int(fb.timeline_samples().max())
    out, written = _render(fb, sr, count, precision)
    np.testing.assert_array_equal(written, np.minimum(fb.timeline_samples(), count))
    for s in range(0, fb.num_streams, 5):
        want = _oracle(port, fb, sr, s, count)
        got = out[s, :written[s]]
        if precision == player.PRECISION_FP64:
            assert_f64_parity(got, want, "vowel chart stream %d" % s)
        else:
            parity.assert_f32_parity(got, want, "vowel chart stream %d" % s)
        assert not out[s, written[s]:].any()


@pytest.mark.parametrize("precision", [player.PRECISION_FP64, player.PRECISION_FP32])
def test_config5_midi_sing_vs_oracle(port, precision):
    sr, secs = 22050, 2.0
    fb = workloads.midi_sing(40, secs, sr)
    count = int(secs * sr)
    out, written = _render(fb, sr, count, precision)
    for s in range(0, 40, 4):
        want = _oracle(port, fb, sr, s, count)
        got = out[s, :len(want)]
        assert written[s] == len(want)
        if precision == player.PRECISION_FP64:
            assert_f64_parity(got, want, "midi stream %d" % s)
        else:
            parity.assert_f32_parity(got, want, "midi stream %d" % s)


def test_config5_whole_sample_pitch_periods_wrap_with_the_reference(port):
    """Notes whose period is a whole number of samples (150 Hz at 22 050 Hz = 147 samples) put the sawtooth wrap on the
    last bit of the phase sum: the FP32 kernels must wrap on the reference's sample (FP64 phase, exact division)."""
    sr, count = 22050, 22050
    fr = np.zeros((2, 47))
    for j, hz in enumerate((150.0, 225.0)):
        fr[j, workloads.P["voicePitch"]] = fr[j, workloads.P["endVoicePitch"]] = hz
        fr[j, workloads.P["voiceAmplitude"]] = 1.0
        fr[j, workloads.P["preFormantGain"]] = 1.0
        fr[j, workloads.P["outputGain"]] = 1.0
        workloads.set_frame(fr[j], "a")
    m = np.array([11025, 11025], dtype=np.uint32)
    f = np.array([441, 441], dtype=np.uint32)
    nul = np.zeros(2, dtype=np.uint8)
    ux = np.array([0, 1], dtype=np.int32)
    fb = workloads._concat(sr, [(fr, m, f, nul, ux)] * 4, np.arange(4, dtype=np.uint64))
    out, _ = _render(fb, sr, count, player.PRECISION_FP32)
    for s in range(4):
        want = _oracle(port, fb, sr, s, count)
        parity.assert_f32_parity(out[s, :len(want)], want, "whole-sample periods, stream %d" % s)


@pytest.mark.parametrize("family", ["vowel", "midi"])
def test_batch_is_the_concatenation_of_its_streams(family):
    """Sharding property at a size the oracle is not asked to render: 2 368 vowel-chart streams / 1 500 midi streams."""
    if family == "vowel":
        sr = 16000
        fb = workloads.vowel_chart(2, sr, pairs=1184)
        count = 6000
    else:
        sr = 22050
        fb = workloads.midi_sing(1500, 1.0, sr)
        count = 11025
    full, wfull = _render(fb, sr, count, player.PRECISION_FP32)
    lo, hi = fb.num_streams // 3, fb.num_streams // 3 + 257  # a shard that starts and ends inside warps
    part, wpart = _render(fb, sr, count, player.PRECISION_FP32, streams=range(lo, hi))
    np.testing.assert_array_equal(part, full[lo:hi])
    np.testing.assert_array_equal(wpart, wfull[lo:hi])
    # chunked pulls (ending inside chunks of the round scheduler) equal one pull
    chunked, wch = _render(fb, sr, count, player.PRECISION_FP32, chunks=[1000, 333, count - 1333], streams=range(lo, hi))
    np.testing.assert_array_equal(chunked, part)
    np.testing.assert_array_equal(wch, wpart)
    np.testing.assert_array_equal(wfull, np.minimum(fb.timeline_samples(), count))


@pytest.mark.parametrize("precision", [player.PRECISION_FP64, player.PRECISION_FP32])
def test_voices_applied_on_device_equal_host_rewritten_frames(port, precision):
    """speechPlayer_batchApplyVoices rewrites the queued frames like the NVDA driver's
    applyVoiceToFrame (absolute override, then _mul factor, per parameter); the render equals uploading frames rewritten
    on the host with the same rule, bit for bit, and the oracle's render of those frames within the precision's bar."""
    sr, secs, n = 22050, 0.5, 70
    fb = workloads.random_frames(n, secs, sr, first_stream=4000)
    count = int(secs * sr)
    va, vm = player.nvda_voice_tables([
        {"cb1_mul": 1.3, "pa6_mul": 1.3, "fricationAmplitude_mul": 0.85},
        {"cf1_mul": 1.01, "cf2_mul": 1.02, "cf4": 3770, "cf5": 4100, "cf6": 5000, "cfNP_mul": 0.9, "cb1_mul": 1.3,
         "fricationAmplitude_mul": 0.7, "pa6_mul": 1.3},
        {"aspirationAmplitude": 1, "voiceAmplitude": 0},
        {"voicePitch_mul": 0.75, "endVoicePitch_mul": 0.75, "cf1_mul": 0.75, "cf2_mul": 0.85, "cf3_mul": 0.85}])
    rng = np.random.default_rng(5)
    vos = rng.integers(0, 4, size=n).astype(np.uint32)
    # host rule
    host = workloads._concat(sr, [fb.stream(s) for s in range(n)], fb.stream_ids)
    for s in range(n):
        a, b_ = int(host.offsets[s]), int(host.offsets[s + 1])
        rows = host.frames[a:b_]
        keep = host.is_null[a:b_] == 0
        new = np.where(np.isnan(va[vos[s]]), rows, va[vos[s]]) * vm[vos[s]]
        rows[keep] = new[keep]
    b = player.Batch(sr, n, precision=precision, noise=player.NOISE_PHILOX, seed=SEED, stream_ids=fb.stream_ids)
    b.set_frames_host(fb)
    b.apply_voices(va, vm, vos)
    dev_out, dev_w = b.synthesize_host(count)
    b.set_frames_host(host)
    host_out, host_w = b.synthesize_host(count)
    b.close()
    np.testing.assert_array_equal(dev_out, host_out)
    np.testing.assert_array_equal(dev_w, host_w)
    assert not np.array_equal(dev_out, _render(fb, sr, count, precision)[0])  # the voices did change the audio
    for s in (0, 17, 42):
        want = _oracle(port, host, sr, s, count)
        if precision == player.PRECISION_FP64:
            assert_f64_parity(dev_out[s, :len(want)], want, "voiced stream %d" % s)
        else:
            parity.assert_f32_parity(dev_out[s, :len(want)], want, "voiced stream %d" % s)


def test_ipa_text_to_pcm_end_to_end(golden_config1):
    """Config 1 without the reference's Python in the loop: sampleIpa.txt -> native bulk frame producer
    (include/speechPlayer_ipa.h) -> one FP64 player -> the int16 stream the compiled reference renders from its own
    ipa.py + speechPlayer.py (tests/golden/config1.npz), bit for bit."""
    import os
    from nvspeechplayer_b200 import ipa
    g = golden_config1
    lines = [str(t) for t in np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ipa_frames.npz"))["texts"][:8]]
    sr = int(g["sample_rate"])
    fb = ipa.frames_for_texts(lines, speed=0.6, sample_rate=sr, trailing_silence_ms=150.0)
    p = player.SpeechPlayer(sr, precision=player.PRECISION_FP64, noise=player.NOISE_PHILOX, seed=int(g["philox_seed"]),
                            streamId=int(g["philox_stream"]))
    p.queue_frames(fb.frames, fb.min_dur, fb.fade_dur, None, fb.is_null)  # all eight lines into ONE player, in order
    pcm = p.synthesize_np(300000)
    p.close()
    assert_f64_parity(pcm, g["pcm_philox"], "sampleIpa.txt end to end")
