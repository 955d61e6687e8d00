"""The reference's real consumer, as a script: the NVDA driver's speak / cancel and its audio thread
(reference nvdaAddon/synthDrivers/nvSpeechPlayer/__init__.py:168-241, :56-82).

    speak   : queueFrame(frame, duration, fade, userIndex=...) per phoneme, then queueFrame(None, endPause, fade)
    thread  : loop { data = synthesize(8192); if not data: break; idx = getLastIndex() }
    cancel  : queueFrame(None, 20 ms, 5 ms, purgeQueue=True)      (:237-238)

`run(make_player, sr)` drives any player with queue_frame / synthesize / last_index / close (durations in samples) through
two utterances, the first cancelled after three pulls, and returns (pcm, samples per pull, getLastIndex after each pull)."""
import numpy as np

from nvspeechplayer_b200 import workloads

PULL = 8192


def utterance(seed, sr, n_phonemes, first_index):
    rng = np.random.default_rng(seed)
    names = sorted(workloads.phoneme_table()["names"])
    ops = []
    for j in range(n_phonemes):
        fr = np.zeros(workloads.NUM_PARAMS)
        fr[workloads.P["preFormantGain"]] = 1.0
        fr[workloads.P["outputGain"]] = 1.0
        fr[workloads.P["voiceAmplitude"]] = 1.0
        fr[workloads.P["vibratoPitchOffset"]] = 0.1
        fr[workloads.P["vibratoSpeed"]] = 5.5
        workloads.set_frame(fr, names[int(rng.integers(len(names)))])
        pitch = float(rng.uniform(90, 180))
        fr[workloads.P["voicePitch"]] = pitch
        fr[workloads.P["endVoicePitch"]] = pitch * float(rng.uniform(0.9, 1.1))
        dur = int(sr * rng.uniform(0.05, 0.16))
        fade = int(sr * rng.uniform(0.01, 0.04))
        ops.append((fr, dur, fade, first_index + j if j % 3 == 0 else -1))
    return ops


def run(make_player, sr=16000):
    p = make_player(sr)
    pcm, counts, idx = [], [], []

    def pull():
        c = p.synthesize(PULL)
        pcm.append(np.asarray(c, dtype=np.int16).copy())
        counts.append(len(c))
        idx.append(p.last_index())
        return len(c)

    # speak #1, cancelled after three pulls (the audio thread is mid-utterance, mid-fade or mid-hold as it falls)
    for fr, dur, fade, ux in utterance(1, sr, 24, 100):
        p.queue_frame(fr, dur, fade, ux, False)
    p.queue_frame(None, int(sr * 0.1), int(sr * 0.01), -1, False)
    for _ in range(3):
        assert pull() == PULL
    p.queue_frame(None, int(sr * 0.020), int(sr * 0.005), -1, True)   # cancel()
    while pull() == PULL:
        pass
    assert pull() == 0                                                  # idle: the thread would block on its event here
    # speak #2, played to the end
    for fr, dur, fade, ux in utterance(2, sr, 12, 200):
        p.queue_frame(fr, dur, fade, ux, False)
    p.queue_frame(None, int(sr * 0.1), int(sr * 0.01), 299, False)
    while pull() == PULL:
        pass
    assert pull() == 0
    p.close()
    return np.concatenate(pcm), counts, idx


# ---------------------------------------------------------------------------------------------------------------
# The reference's two live-control demos as scripts: every control event is a queueFrame with purgeQueue=True while the
# audio thread keeps pulling (durations in ms there: int(ms * (sampleRate / 1000.0)), reference speechPlayer.py:53).
# ---------------------------------------------------------------------------------------------------------------
def _ms(ms, sr):
    return int(ms * (sr / 1000.0))


def _base_frame(phoneme="a"):
    fr = np.zeros(workloads.NUM_PARAMS)
    fr[workloads.P["preFormantGain"]] = 1.0
    fr[workloads.P["outputGain"]] = 1.0
    fr[workloads.P["voiceAmplitude"]] = 1.0
    workloads.set_frame(fr, phoneme)
    return fr


def run_leap(make_player, sr=22050, pull=512, events=150):
    """reference test_leap.py:26-39: a hand steers pitch / cf1 / cf2 and every tracking frame re-queues the frame with
    purgeQueue=True (20 000 ms hold, 50 ms fade); no hand: queueFrame(None, 0, 50, purgeQueue=True)."""
    rng = np.random.default_rng(7)
    p = make_player(sr)
    fr = _base_frame("a")
    pcm, counts = [], []
    x = y = z = 0.5
    for e in range(events):
        if 60 <= e < 75 or e >= events - 10:          # the hand leaves the box
            p.queue_frame(None, 0, _ms(50, sr), -1, True)
        else:
            x, y, z = (float(np.clip(v + rng.normal(0, 0.05), 0, 1)) for v in (x, y, z))
            fr[workloads.P["voicePitch"]] = fr[workloads.P["endVoicePitch"]] = 100 * (8 ** y)
            fr[workloads.P["cf1"]] = 200 + 600 * x
            fr[workloads.P["cf2"]] = 500 + 1500 * z
            fr[workloads.P["cf3"]] = 3200
            p.queue_frame(fr, _ms(20000, sr), _ms(50, sr), -1, True)
        c = p.synthesize(pull)
        pcm.append(np.asarray(c, dtype=np.int16).copy())
        counts.append(len(c))
    for _ in range(8):                                  # ring out and go idle
        c = p.synthesize(pull)
        pcm.append(np.asarray(c, dtype=np.int16).copy())
        counts.append(len(c))
    p.close()
    return np.concatenate(pcm), counts


def run_midi(make_player, sr=22050, pull=2048):
    """reference test_midiSing.py:96-134: note on = patch 'start' frames (the first with purgeQueue) + a 'mid' frame held
    for 10 000 000 ms; controller and pitch-bend messages re-queue the frame with purgeQueue=True; note off = patch
    'end' frames (the first with purgeQueue) + queueFrame(None, 0, 20)."""
    p = make_player(sr)
    fr = _base_frame("a")
    fr[workloads.P["vibratoPitchOffset"]] = 0.125
    fr[workloads.P["vibratoSpeed"]] = 5.5
    pcm, counts = [], []

    def pulls(k):
        for _ in range(k):
            c = p.synthesize(pull)
            pcm.append(np.asarray(c, dtype=np.int16).copy())
            counts.append(len(c))

    def note_on(note, vel):
        fr[workloads.P["voicePitch"]] = fr[workloads.P["endVoicePitch"]] = 440 * (2 ** ((note - 69) / 12.0))
        fr[workloads.P["preFormantGain"]] = vel / 32.0
        workloads.set_frame(fr, "i")
        p.queue_frame(fr, _ms(50, sr), _ms(30, sr), -1, True)
        workloads.set_frame(fr, "a")
        p.queue_frame(fr, _ms(10000000, sr), _ms(30, sr), -1, False)

    def note_off():
        workloads.set_frame(fr, "a")
        p.queue_frame(fr, _ms(30, sr), _ms(30, sr), -1, True)
        p.queue_frame(None, 0, _ms(20, sr), -1, False)

    note_on(57, 96); pulls(5)
    for bend in (70, 90, 120, 64, 20):                  # pitch-bend wheel: vibrato depth / speed, open quotient
        if bend < 64:
            fr[workloads.P["glottalOpenQuotient"]] = 0.1 * ((64 - bend) / 64.0)
        else:
            fr[workloads.P["voiceTurbulenceAmplitude"]] = 0
        fr[workloads.P["vibratoSpeed"]] = (5.5 + ((bend - 64) / 64.0)) if bend >= 64 else 5.5
        fr[workloads.P["vibratoPitchOffset"]] = (0.125 + (((bend - 64) / 64.0) * 0.875)) if bend >= 64 else (0.125 * (bend / 64.0))
        p.queue_frame(fr, _ms(10000000, sr), _ms(100, sr), -1, True)
        pulls(2)
    workloads.set_frame(fr, "u")                        # controller: another vowel while the note sounds
    p.queue_frame(fr, _ms(10000000, sr), _ms(50, sr), -1, True); pulls(3)
    note_on(64, 120); pulls(4)                          # legato: a new note purges the held one
    note_off(); pulls(3)
    note_on(45, 60); pulls(2)
    note_off(); pulls(4)
    p.close()
    return np.concatenate(pcm), counts
