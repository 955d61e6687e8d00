"""Workload generators for the BASELINE.json configs (frame queues, not audio).

A workload is a ``FrameBatch``: every stream's queueFrame calls, flattened.  Durations
are in SAMPLES (the C side's unit, reference ``src/speechPlayer.h:28``); the ms->samples
conversion of the reference wrapper is ``int(ms*(sampleRate/1000.0))``
(reference ``speechPlayer.py:53``) and is applied here where a recipe is given in ms.

  config 2  vowel_chart()    reference test_playVowelchart.py:24-45 x "voices" in the style of
                             nvdaAddon/synthDrivers/nvSpeechPlayer/__init__.py:117-125
  config 3  random_frames()  SURVEY.md section 8(d): all frame params randomised
  config 5  midi_sing()      reference test_midiSing.py:23-60,116-134 style note lists

Generation is counter-based per stream (numpy Philox keyed by (seed, stream id)), so
any shard of streams can be generated independently on any rank.
"""
import os
from dataclasses import dataclass

import numpy as np

PARAM_NAMES = [
    "voicePitch", "vibratoPitchOffset", "vibratoSpeed", "voiceTurbulenceAmplitude", "glottalOpenQuotient",
    "voiceAmplitude", "aspirationAmplitude",
    "cf1", "cf2", "cf3", "cf4", "cf5", "cf6", "cfN0", "cfNP",
    "cb1", "cb2", "cb3", "cb4", "cb5", "cb6", "cbN0", "cbNP",
    "caNP", "fricationAmplitude",
    "pf1", "pf2", "pf3", "pf4", "pf5", "pf6",
    "pb1", "pb2", "pb3", "pb4", "pb5", "pb6",
    "pa1", "pa2", "pa3", "pa4", "pa5", "pa6",
    "parallelBypass", "preFormantGain", "outputGain", "endVoicePitch",
]  # ABI order, reference src/frame.h:22-46
NUM_PARAMS = len(PARAM_NAMES)
assert NUM_PARAMS == 47
P = {n: i for i, n in enumerate(PARAM_NAMES)}


@dataclass
class FrameBatch:
    sample_rate: int
    offsets: np.ndarray     # int64 [numStreams+1] into the flat arrays
    frames: np.ndarray      # float64 [total, 47]   (rows of NULL frames are ignored)
    min_dur: np.ndarray     # uint32 [total]  minFrameDuration in samples
    fade_dur: np.ndarray    # uint32 [total]  fadeDuration in samples (0 is clamped to 1 by the ABI)
    is_null: np.ndarray     # uint8  [total]  1 = queueFrame(NULL, ...)
    user_index: np.ndarray  # int32  [total]
    stream_ids: np.ndarray  # uint64 [numStreams] noise-stream id of each stream

    @property
    def num_streams(self):
        return len(self.offsets) - 1

    def stream(self, s):
        a, b = int(self.offsets[s]), int(self.offsets[s + 1])
        return (self.frames[a:b], self.min_dur[a:b], self.fade_dur[a:b], self.is_null[a:b], self.user_index[a:b])

    def timeline_samples(self):
        """Samples each stream yields when fully pre-queued: sum max(M+1, max(F,1)+2)."""
        m = self.min_dur.astype(np.int64)
        f = np.maximum(self.fade_dur.astype(np.int64), 1)
        occ = np.maximum(m + 1, f + 2)
        c = np.concatenate([[0], np.cumsum(occ)])
        return c[self.offsets[1:]] - c[self.offsets[:-1]]

    def fade_fraction(self, sample_count=None):
        """phi = fade ticks / rendered ticks (exact from the timeline law; if sample_count is given the
        per-stream tail beyond it is cut)."""
        m = self.min_dur.astype(np.int64)
        f = np.maximum(self.fade_dur.astype(np.int64), 1)
        occ = np.maximum(m + 1, f + 2)
        fade = 0
        total = 0
        for s in range(self.num_streams):
            a, b = int(self.offsets[s]), int(self.offsets[s + 1])
            start = np.concatenate([[0], np.cumsum(occ[a:b])])[:-1] + 1  # fade ticks are start .. start+F-1
            limit = int(occ[a:b].sum()) if sample_count is None else min(int(sample_count), int(occ[a:b].sum()))
            fade += int(np.clip(limit - start, 0, f[a:b]).sum())
            total += limit
        return fade / max(total, 1)


def _concat(sample_rate, per_stream, stream_ids):
    offsets = np.zeros(len(per_stream) + 1, dtype=np.int64)
    for i, st in enumerate(per_stream):
        offsets[i + 1] = offsets[i] + len(st[1])
    cat = lambda k, dt: (np.concatenate([st[k] for st in per_stream]).astype(dt) if per_stream else np.zeros(0, dt))
    frames = np.concatenate([st[0] for st in per_stream]).reshape(-1, NUM_PARAMS) if per_stream else np.zeros((0, 47))
    return FrameBatch(sample_rate, offsets, np.ascontiguousarray(frames, dtype=np.float64), cat(1, np.uint32),
                      cat(2, np.uint32), cat(3, np.uint8), cat(4, np.int32), np.asarray(stream_ids, dtype=np.uint64))


def ms_to_samples(ms, sample_rate):
    return int(ms * (sample_rate / 1000.0))  # reference speechPlayer.py:53


# ----------------------------------------------------------------------------------------------
# config 3: synthetic random frames
# ----------------------------------------------------------------------------------------------
RANDOM_RANGES = {
    "voicePitch": (60, 400), "endVoicePitch": (60, 400), "vibratoPitchOffset": (0, 0.3), "vibratoSpeed": (0, 8),
    "voiceTurbulenceAmplitude": (0, 0.5), "glottalOpenQuotient": (0, 0.6), "voiceAmplitude": (0, 1),
    "aspirationAmplitude": (0, 1), "fricationAmplitude": (0, 1), "caNP": (0, 1),
    "cf1": (150, 900), "cf2": (500, 2500), "cf3": (1300, 3500), "cf4": (3000, 4000), "cf5": (3500, 4500),
    "cf6": (4500, 5500), "cfN0": (200, 500), "cfNP": (180, 300),
    "cb1": (40, 400), "cb2": (40, 400), "cb3": (40, 400), "cb4": (150, 1000), "cb5": (150, 1000),
    "cb6": (150, 1000), "cbN0": (50, 200), "cbNP": (50, 200),
    "pf1": (150, 900), "pf2": (500, 2500), "pf3": (1300, 3500), "pf4": (3000, 4000), "pf5": (3500, 4500),
    "pf6": (4500, 5500),
    "pb1": (40, 1000), "pb2": (40, 1000), "pb3": (40, 1000), "pb4": (40, 1000), "pb5": (40, 1000), "pb6": (40, 1000),
    "pa1": (0, 1), "pa2": (0, 1), "pa3": (0, 1), "pa4": (0, 1), "pa5": (0, 1), "pa6": (0, 1),
    "parallelBypass": (0, 1.1), "preFormantGain": (0.5, 1.5), "outputGain": (0.5, 2),
}
_LO = np.array([RANDOM_RANGES[n][0] for n in PARAM_NAMES], dtype=np.float64)
_HI = np.array([RANDOM_RANGES[n][1] for n in PARAM_NAMES], dtype=np.float64)
DEFAULT_SEED = 0xB200


def random_stream(stream_id, seconds, sample_rate, seed=DEFAULT_SEED, null_fraction=0.05,
                  min_ms=(30.0, 150.0), fade_lo_ms=5.0, fade_hi_frac=0.6):
    """One config-3 stream: frames until >= seconds*sample_rate ticks; M_ms~U(30,150), F_ms~U(5,0.6*M_ms),
    5% NULL frames, all params uniform in RANDOM_RANGES; no NaN, M>=1."""
    rng = np.random.Generator(np.random.Philox(key=[int(seed), int(stream_id)]))
    target = int(round(seconds * sample_rate))
    n_max = target // max(ms_to_samples(min_ms[0], sample_rate) + 1, 2) + 2
    d = rng.random((n_max, 3))
    m_ms = min_ms[0] + (min_ms[1] - min_ms[0]) * d[:, 0]
    f_ms = fade_lo_ms + (fade_hi_frac * m_ms - fade_lo_ms) * d[:, 1]
    scale = sample_rate / 1000.0
    m = np.maximum((m_ms * scale).astype(np.int64), 1)
    f = (f_ms * scale).astype(np.int64)
    occ = np.maximum(m + 1, np.maximum(f, 1) + 2)
    n = int(np.searchsorted(np.cumsum(occ), target, side="left")) + 1
    n = min(n, n_max)
    frames = _LO + (_HI - _LO) * rng.random((n, NUM_PARAMS))
    is_null = (d[:n, 2] < null_fraction).astype(np.uint8)
    user_index = np.arange(n, dtype=np.int32)
    return frames, m[:n].astype(np.uint32), f[:n].astype(np.uint32), is_null, user_index


def random_frames(num_streams, seconds, sample_rate=22050, seed=DEFAULT_SEED, first_stream=0, **kw):
    ids = np.arange(first_stream, first_stream + num_streams, dtype=np.uint64)
    return _concat(sample_rate, [random_stream(int(s), seconds, sample_rate, seed, **kw) for s in ids], ids)


# ----------------------------------------------------------------------------------------------
# phoneme table (numeric content of the reference's data.py, extracted by tests/golden/make_golden.py)
# ----------------------------------------------------------------------------------------------
_TABLE = None


def phoneme_table():
    global _TABLE
    if _TABLE is None:
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "phoneme_table.npz")
        z = np.load(path)
        _TABLE = {"names": [str(n) for n in z["names"]], "table": z["table"], "present": z["present"].astype(bool),
                  "flags": {k[4:]: z[k] for k in z.files if k.startswith("flag")}}
    return _TABLE


def set_frame(frame, phoneme):
    """reference ipa.py:29-32 setFrame: copy the phoneme's numeric params into the frame (others untouched)."""
    t = phoneme_table()
    i = t["names"].index(phoneme) if isinstance(phoneme, str) else int(phoneme)
    present = t["present"][i]  # only the keys the phoneme's dict holds are written; the rest keep their value
    frame[present] = t["table"][i][present]
    return frame


# ----------------------------------------------------------------------------------------------
# config 2: vowel chart pairs x voices
# ----------------------------------------------------------------------------------------------
VOICE_PARAMS = ["cf1", "cf2", "cf3", "cb1", "pa6", "fricationAmplitude", "voicePitch", "endVoicePitch"]


def voice_multipliers(voice, seed=DEFAULT_SEED):
    """Per-parameter multipliers in the style of the NVDA driver's voice table
    (nvdaAddon/synthDrivers/nvSpeechPlayer/__init__.py:86-125): seeded U(0.75,1.3); pitch start/end share one."""
    rng = np.random.Generator(np.random.Philox(key=[int(seed) ^ 0x766F6963, int(voice)]))
    u = 0.75 + 0.55 * rng.random(len(VOICE_PARAMS) - 1)
    mult = np.ones(NUM_PARAMS)
    for name, v in zip(VOICE_PARAMS[:-1], u):
        mult[P[name]] = v
    mult[P["endVoicePitch"]] = mult[P["voicePitch"]]
    return mult


def vowel_chart_stream(first, last, voice, sample_rate=16000, seed=DEFAULT_SEED):
    """One (voice, ordered pair) stream of test_playVowelchart.py:27-43: NULL(0 ms, 20 ms, purge), A(300,50) with
    pitch 40->300, B(500,400) with pitch 300->40, NULL(50,50)."""
    mult = voice_multipliers(voice, seed)
    base = np.zeros(NUM_PARAMS)
    base[P["preFormantGain"]] = 1.0
    base[P["voiceAmplitude"]] = 1.0
    base[P["outputGain"]] = 1.0
    a = base.copy()
    a[P["voicePitch"]], a[P["endVoicePitch"]] = 40, 300
    set_frame(a, first)
    b = a.copy()
    b[P["voicePitch"]], b[P["endVoicePitch"]] = 300, 40
    set_frame(b, last)
    frames = np.stack([np.zeros(NUM_PARAMS), a * mult, b * mult, np.zeros(NUM_PARAMS)])
    ms = [(0, 20), (300, 50), (500, 400), (50, 50)]
    m = np.array([ms_to_samples(x, sample_rate) for x, _ in ms], dtype=np.uint32)
    f = np.array([ms_to_samples(y, sample_rate) for _, y in ms], dtype=np.uint32)
    return frames, m, f, np.array([1, 0, 0, 1], dtype=np.uint8), np.full(4, -1, dtype=np.int32)


def vowel_chart(num_voices, sample_rate=16000, seed=DEFAULT_SEED, pairs=None, first_stream=0):
    """config 2, 'fresh player per (voice, pair)' reading: stream id = voice*npairs + pair."""
    t = phoneme_table()
    voiced = [i for i, v in enumerate(t["flags"]["_isVoiced"]) if v]
    all_pairs = [(x, y) for x in voiced for y in voiced]
    if pairs is not None:
        all_pairs = all_pairs[:pairs]
    streams, ids = [], []
    for v in range(num_voices):
        for k, (x, y) in enumerate(all_pairs):
            streams.append(vowel_chart_stream(x, y, v, sample_rate, seed))
            ids.append(first_stream + v * len(all_pairs) + k)
    return _concat(sample_rate, streams, ids)


# ----------------------------------------------------------------------------------------------
# config 5: midi-sing style pitch sweeps
# ----------------------------------------------------------------------------------------------
def midi_sing_stream(stream_id, seconds, sample_rate=22050, seed=DEFAULT_SEED):
    """test_midiSing.py-style: base outputGain=1, voiceAmplitude=1, vibrato 0.125/5.5 Hz; seeded notes (MIDI 36-84,
    hz=440*2^((n-69)/12), velocity -> preFormantGain=vel/32), each note = 'i' start (50 ms / 30 ms fade) + 'a' held
    for the note length (fade 30 ms) sweeping to the next note's pitch, then NULL(0, 20 ms) at the end."""
    rng = np.random.Generator(np.random.Philox(key=[int(seed) ^ 0x6D696469, int(stream_id)]))
    target = int(round(seconds * sample_rate))
    base = np.zeros(NUM_PARAMS)
    base[P["outputGain"]] = 1.0
    base[P["voiceAmplitude"]] = 1.0
    base[P["vibratoPitchOffset"]] = 0.125
    base[P["vibratoSpeed"]] = 5.5
    frames, m, f, nul = [], [], [], []
    ticks = 0
    n_notes = max(2, int(seconds / 0.2) + 2)
    notes = rng.integers(36, 85, size=n_notes)
    vel = rng.integers(32, 128, size=n_notes)
    length_ms = 150.0 + 450.0 * rng.random(n_notes)
    hz = 440.0 * 2.0 ** ((notes - 69) / 12.0)
    k = 0
    while ticks < target and k < n_notes - 1:
        start = base.copy()
        set_frame(start, "i")
        start[P["preFormantGain"]] = vel[k] / 32.0
        start[P["voicePitch"]] = start[P["endVoicePitch"]] = hz[k]
        mid = base.copy()
        set_frame(mid, "a")
        mid[P["preFormantGain"]] = vel[k] / 32.0
        mid[P["voicePitch"]], mid[P["endVoicePitch"]] = hz[k], hz[k + 1]
        for fr, dur, fade in ((start, 50.0, 30.0), (mid, float(length_ms[k]), 30.0)):
            frames.append(fr)
            m.append(max(ms_to_samples(dur, sample_rate), 1))
            f.append(ms_to_samples(fade, sample_rate))
            nul.append(0)
            ticks += max(m[-1] + 1, max(f[-1], 1) + 2)
        k += 1
    frames.append(np.zeros(NUM_PARAMS))
    m.append(0)
    f.append(ms_to_samples(20.0, sample_rate))
    nul.append(1)
    n = len(m)
    return (np.stack(frames), np.array(m, dtype=np.uint32), np.array(f, dtype=np.uint32),
            np.array(nul, dtype=np.uint8), np.arange(n, dtype=np.int32))


def midi_sing(num_streams, seconds=5.0, sample_rate=22050, seed=DEFAULT_SEED, first_stream=0):
    ids = np.arange(first_stream, first_stream + num_streams, dtype=np.uint64)
    return _concat(sample_rate, [midi_sing_stream(int(s), seconds, sample_rate, seed) for s in ids], ids)
