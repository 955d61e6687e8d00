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


def _random_script(seed):
    """Random interleavings of queueFrame (real / NULL frames, short and zero durations, purges at any moment -- on the pop
    tick, inside a fade, inside a hold, while idle, twice in a row) and pulls of random sizes."""
    rng = np.random.default_rng(seed)
    sr = int(rng.choice([16000, 22050, 44100]))
    ops, ux = [], 0
    for _ in range(int(rng.integers(8, 30))):
        r = rng.random()
        if r < 0.6:
            for _ in range(int(rng.integers(1, 4))):
                null = rng.random() < 0.2
                fr = None if null else scenarios._rand_frame(rng)
                m = int(rng.choice([0, 1, 2, int(rng.integers(3, 700))])) if not null else int(rng.integers(0, 300))
                if fr is not None and m == 0:
                    m = 1   # a real frame with M = 0 divides by zero in the reference (frame.cpp:98): not a use case
                f = int(rng.choice([0, 1, 2, int(rng.integers(3, 700))]))
                purge = rng.random() < 0.15
                ux += 1
                ops.append(("q", fr, m, f, ux if rng.random() < 0.7 else -1, purge))
        else:
            ops.append(("s", int(rng.choice([1, 2, 7, int(rng.integers(8, 3000))]))))
    ops.append(("s", 4000))
    return dict(sr=sr, ops=ops)


@pytest.mark.parametrize("seed", range(40))
def test_pull_random_scripts_vs_oracle(port, seed):
    """The request-level host frame manager against the reference's per-sample one (plain-C restatement, pinned bit-exact
    to the compiled reference) on scripts nobody wrote by hand."""
    sc = _random_script(1000 + seed)

    def oracle_player(sr):
        p = port.player(sr)
        p.noise_philox(scenarios.SEED, scenarios.STREAM)
        return p
    want, wcounts, widx = scenarios.run_script(oracle_player, sc)
    got, counts, idx = scenarios.run_script(_player, sc)
    assert counts == wcounts
    assert idx == widx
    parity.assert_f32_parity(got, want, "random script %d" % seed, short_ok=3)


def test_pull_phase_runs_bit_identical(golden_config1):
    """The run decomposition of the glottal-phase recurrence (klatt_pull_core.cuh pullRuns*, the device's default) renders
    the same bits as the plain serial FP64 loop -- on every scenario (whole-sample pitch periods, negative pitches that
    hand the pull back to the serial loop, purges), on config 1 and on random frames -- while leaving only a fraction of
    the ticks to the serial thread."""
    def render_all(mode):
        outs = []
        for name, sc in sorted(scenarios.all_scenarios().items()):
            pcm, _, _ = scenarios.run_script(lambda sr: _player(sr, phase_mode=mode), sc)
            outs.append(pcm)
        g = golden_config1
        p = sim.PullPlayer(int(g["sample_rate"]), seed=int(g["philox_seed"]), stream=int(g["philox_stream"]), phase_mode=mode)
        for j in range(len(g["min_dur"])):
            p.queue_frame(None if g["is_null"][j] else g["frames"][j], int(g["min_dur"][j]), int(g["fade_dur"][j]))
        while True:
            c = p.synthesize(8192)
            if c.size == 0:
                break
            outs.append(c.copy())
        p.close()
        for sr, pull in ((16000, 8192), (22050, 2048), (44100, 333)):
            fr, m, f, nul, ux = workloads.random_stream(99, 1.5, sr)
            p = sim.PullPlayer(sr, seed=3, stream=99, phase_mode=mode)
            for j in range(len(m)):
                p.queue_frame(None if nul[j] else fr[j], int(m[j]), int(f[j]), int(ux[j]))
            for _ in range(int(1.5 * sr) // pull):
                outs.append(p.synthesize(pull).copy())
            p.close()
        return np.concatenate(outs)

    l0, s0 = sim.pull_runs_stats()
    plain = render_all(0)
    assert sim.pull_runs_stats() == (l0, s0)
    runs = render_all(1)
    l1, s1 = sim.pull_runs_stats()
    np.testing.assert_array_equal(runs, plain)
    assert l1 - l0 > 100                                  # the decomposition did render most launches ...
    assert (s1 - s0) < 0.25 * plain.size                  # ... and left the serial thread a fraction of the ticks


def _pull_stream(sr, fr, m, f, nul, ux, count, seed, stream, pull=8192):
    p = sim.PullPlayer(sr, seed=seed, stream=stream)
    for j in range(len(m)):
        p.queue_frame(None if nul[j] else fr[j], int(m[j]), int(f[j]), int(ux[j]))
    got = []
    while sum(len(c) for c in got) < count:
        c = p.synthesize(min(pull, count - sum(len(c) for c in got)))
        if c.size == 0:
            break
        got.append(c.copy())
    p.close()
    return np.concatenate(got) if got else np.zeros(0, np.int16)


def test_pull_config2_vowel_chart_streams(port):
    """BASELINE config 2 recipe (16 kHz, vowel pair x voice, pitch sweeps 40 -> 300 -> 40 Hz) through pulls."""
    fb = workloads.vowel_chart(2, pairs=6)
    for s in range(fb.num_streams):
        fr, m, f, nul, ux = fb.stream(s)
        count = int(fb.timeline_samples()[s])
        want = port.render(16000, fr, m, f, nul, ux, max_samples=count, noise=("philox", 8, int(fb.stream_ids[s])))
        got = _pull_stream(16000, fr, m, f, nul, ux, count, 8, int(fb.stream_ids[s]))
        parity.assert_f32_parity(got, want, "vowel chart stream %d via pulls" % s)


def test_pull_config5_midi_sing_streams(port):
    """BASELINE config 5 recipe (midi-sing note lists with pitch sweeps and vibrato) through pulls of 2048."""
    sr, secs = 22050, 2.0
    fb = workloads.midi_sing(4, seconds=secs, sample_rate=sr, first_stream=40)
    count = int(secs * sr)
    for s in range(fb.num_streams):
        fr, m, f, nul, ux = fb.stream(s)
        want = port.render(sr, fr, m, f, nul, ux, max_samples=count, noise=("philox", 5, int(fb.stream_ids[s])))
        got = _pull_stream(sr, fr, m, f, nul, ux, len(want), 5, int(fb.stream_ids[s]), pull=2048)
        parity.assert_f32_parity(got, want, "midi stream %d via pulls" % s)


def test_pull_whole_sample_pitch_periods(port):
    """150 Hz and 225 Hz at 22 050 Hz (periods of exactly 147 and 98 samples): the wrap instant hangs on the last bit of the
    FP64 phase sum, and the run decomposition must reproduce it."""
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
    want = port.render(sr, fr, m, f, nul, ux, max_samples=count, noise=("philox", 2, 3))
    got = _pull_stream(sr, fr, m, f, nul, ux, len(want), 2, 3)
    parity.assert_f32_parity(got, want, "whole-sample periods via pulls")


def test_pull_nvda_speak_cancel_speak(port):
    """The NVDA driver's speak / cancel / speak with its audio thread pulling 8192 samples at a time (tests/nvda_flow.py):
    same samples per pull, same getLastIndex after every pull, same audio as the reference's frame manager and generator."""
    from tests import nvda_flow

    def oracle_player(sr):
        p = port.player(sr)
        p.noise_philox(scenarios.SEED, scenarios.STREAM)
        return p
    want, wcounts, widx = nvda_flow.run(oracle_player)
    got, counts, idx = nvda_flow.run(_player)
    assert counts == wcounts and len(counts) > 8
    assert idx == widx and max(idx) == 299
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
    got, counts = run(_player)
    assert counts == wcounts
    parity.assert_f32_parity(got, want, "live control: " + flow)


def _adversarial_increments(rng, n):
    sr = float(rng.choice([16000, 22050, 44100]))
    kind = int(rng.integers(0, 12))
    t = np.arange(n)
    if kind == 0: return np.full(n, rng.uniform(40, 500) / sr)
    if kind == 1: return np.full(n, 2.0 ** -float(rng.integers(2, 12)))                       # dyadic: rounding ties everywhere
    if kind == 2: return np.full(n, float(rng.integers(1, 400)) / float(rng.integers(401, 4000)))
    if kind == 3: return rng.uniform(0, 0.02, n)
    if kind == 4: return rng.uniform(0, 0.49999, n)                                            # absurd pitches
    if kind == 5: return np.where(rng.random(n) < 0.3, 0.0, rng.uniform(40, 500) / sr)         # zeros interspersed
    if kind == 6: return (200 + 100 * np.sin(2 * np.pi * 5.5 * t / sr)) / sr                   # deep vibrato
    if kind == 7: return np.full(n, rng.uniform(1e-9, 1e-5))                                   # almost no pitch
    if kind == 8: return 2.0 ** -rng.integers(2, 30, n).astype(float)                          # random dyadics
    if kind == 9: return np.full(n, 1.0 / float(rng.integers(2, 400)))                         # whole-sample periods
    if kind == 10: return np.abs(rng.normal(0.01, 0.01, n))
    return np.linspace(rng.uniform(40, 400), rng.uniform(40, 400), n) / sr                     # glide


def test_pull_phase_runs_adversarial_increments():
    """The run decomposition against the plain FP64 recurrence on increment sequences chosen to hurt: exact dyadics (every
    addition a rounding tie), whole-sample periods, zeros, near-Nyquist pitches, starts on and next to binade boundaries.
    Bit for bit, carry included; and the plain loop itself against Python's float arithmetic."""
    import math
    rng = np.random.default_rng(11)
    ticks = visited = 0
    for it in range(400):
        n = int(rng.choice([1, 7, 333, 512, 2048, 4000, 8192]))
        inc = _adversarial_increments(rng, n)
        pos0 = float(rng.choice([0.0, rng.random(), 1 - 2.0 ** -53, 2.0 ** -45, 0.5, 0.25 - 2.0 ** -55]))
        a, ca, _ = sim.phase(inc, pos0, 0)
        b, cb, v = sim.phase(inc, pos0, 1)
        assert np.array_equal(a.view(np.uint64), b.view(np.uint64)) and ca == cb, (it, n, pos0)
        if v >= 0:
            ticks += n
            visited += v
        if it % 40 == 0:
            pos = pos0
            for k in range(n):
                s = pos + float(inc[k])
                pos = s - math.trunc(s)
                assert pos == a[k]
            assert pos == ca
    assert visited < 0.2 * ticks
