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
