"""GPU: the long-utterance path (kernel b: chunks rendered in parallel, phase and resonator state recovered by scans)
against the reference restatement, through the C-ABI.

Bar (BASELINE.json north_star, FP32 variant): <= 1 LSB on >= 99.9 % of samples and >= 60 dB SNR; the sample count
follows the timeline law.  Config 4's recipe (config-1 frames looped, 44.1 kHz) at a length the oracle renders in
seconds; the full hour is exercised by bench.py --workload long."""
import numpy as np
import pytest

from nvspeechplayer_b200 import player, workloads
from tests import parity

pytestmark = pytest.mark.gpu


def _looped_config1(g, sr, repeats):
    scale = sr / 22050.0
    m = np.tile((g["min_dur"] * scale).astype(np.uint32), repeats)
    f = np.tile((g["fade_dur"] * scale).astype(np.uint32), repeats)
    return np.tile(g["frames"], (repeats, 1)), m, f, np.tile(g["is_null"], repeats)


@pytest.mark.parametrize("sr,repeats,chunk", [(44100, 2, 1024), (22050, 1, 256), (16000, 1, 4096)])
def test_long_config1_looped(port, golden_config1, sr, repeats, chunk):
    fr, m, f, nul = _looped_config1(golden_config1, sr, repeats)
    want = port.render(sr, fr, m, f, nul, noise=("philox", 5, 9))
    got, ms, launches = player.synthesize_long(sr, fr, m, f, nul, seed=5, stream_id=9, chunk_ticks=chunk)
    assert launches >= 20
    assert len(got) == len(want) == workloads._concat(sr, [(fr, m, f, nul, np.zeros(len(m), np.int32))], [9]).timeline_samples()[0]
    w1, exact, snr, mx = parity.assert_f32_parity(got, want, "long config1 @%d" % sr)
    assert snr >= 80.0


@pytest.mark.parametrize("sr,chunk", [(22050, 1024), (44100, 512)])
def test_long_random_frames(port, sr, chunk):
    """all 47 params randomised, 5 % NULL requests, fades from 5 ms to 0.6 M: every stage's scan is exercised"""
    fr, m, f, nul, ux = workloads.random_stream(31337, 6.0, sr)
    want = port.render(sr, fr, m, f, nul, ux, noise=("philox", 0xB200, 31337))
    got, ms, _ = player.synthesize_long(sr, fr, m, f, nul, seed=0xB200, stream_id=31337, chunk_ticks=chunk)
    assert len(got) == len(want)
    parity.assert_f32_parity(got, want, "long random @%d" % sr)


def test_long_matches_batch_kernel_and_chunk_sizes():
    """same stream through kernel (a) (FP32 batch path) and kernel (b) at two chunk sizes: within 1 LSB of each other"""
    sr = 22050
    fr, m, f, nul, ux = workloads.random_stream(77, 3.0, sr)
    fb = workloads._concat(sr, [(fr, m, f, nul, ux)], [77])
    n = int(fb.timeline_samples()[0])
    b = player.Batch(sr, 1, precision=player.PRECISION_FP32, seed=3, stream_ids=fb.stream_ids)
    b.set_frames_host(fb)
    a_out, written = b.synthesize_host(n)
    b.close()
    assert written[0] == n
    outs = [player.synthesize_long(sr, fr, m, f, nul, seed=3, stream_id=77, chunk_ticks=c)[0] for c in (128, 2048)]
    for o in outs:
        assert len(o) == n
        parity.assert_f32_parity(o, a_out[0], "kernel b vs kernel a")
    parity.assert_f32_parity(outs[0], outs[1], "chunk 128 vs 2048")


def test_long_truncated_and_null_only():
    sr = 22050
    fr, m, f, nul, ux = workloads.random_stream(5, 1.0, sr)
    full, _, _ = player.synthesize_long(sr, fr, m, f, nul, seed=1, stream_id=5)
    part, _, _ = player.synthesize_long(sr, fr, m, f, nul, seed=1, stream_id=5, max_samples=10000)
    assert len(part) == 10000
    parity.assert_f32_parity(part, full[:10000], "truncated")
    # a queue of NULL requests only renders silence of the lawful length (reference src/frame.cpp:59-63)
    got, _, _ = player.synthesize_long(sr, np.zeros((2, 47)), [100, 0], [50, 30], is_null=[1, 1], seed=1, stream_id=5)
    assert len(got) == 101 + 32 and not got.any()
