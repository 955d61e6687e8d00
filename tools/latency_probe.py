#!/usr/bin/env python
"""Per-handle pull latency of the five-symbol drop-in (what the NVDA audio thread sees): sampleIpa.txt frames queued on
one player, speechPlayer_synthesize pulls of 8192 samples (the NVDA pull size) and shorter, wall clock per pull.
"stream" is SPEECHPLAYER_PRECISION_STREAM: the time-parallel block kernel with the frame manager on the host
(SURVEY 8f rank 3).  "reference" is the unmodified reference C++ (oracle/_ref, test infrastructure) on one host core of
the same box, timed the same way -- the latency a pull has today.  NVSP_PROBE_ONLY=<name> runs one of them."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
g = np.load(os.path.join(ROOT, "tests", "golden", "config1.npz"))
sr = int(g["sample_rate"])
only = os.environ.get("NVSP_PROBE_ONLY")


def report(name, pull, ts):
    ts = np.array(ts[:-1]) * 1e3
    print("%-9s pull %5d samples (%.0f ms of audio): median %.3f ms, p95 %.3f ms, max %.3f ms over %d pulls  -> %.0fx real time"
          % (name, pull, 1e3 * pull / sr, np.median(ts), np.percentile(ts, 95), ts.max(), len(ts), (1e3 * pull / sr) / np.median(ts)))


def run(make, queue, pull):
    p = make()
    queue(p)
    p_syn = p.synthesize_np if hasattr(p, "synthesize_np") else p.synthesize
    p_syn(pull)  # warm-up (first launch, allocations)
    ts = []
    while True:
        t0 = time.perf_counter()
        c = p_syn(pull)
        ts.append(time.perf_counter() - t0)
        if len(c) < pull:
            break
    p.close()
    return ts


if not only or only == "reference":
    from oracle import oracle
    if oracle.have_ref():
        ref = oracle.RefLib()

        def queue_ref(p):
            for j in range(len(g["min_dur"])):
                p.queue_frame(None if g["is_null"][j] else g["frames"][j], int(g["min_dur"][j]), int(g["fade_dur"][j]))
        for pull in (8192, 2048, 512):
            report("reference", pull, run(lambda: ref.player(sr), queue_ref, pull))
    else:
        print("reference: oracle/_ref not built here")

from nvspeechplayer_b200 import player
for prec, name in ((player.PRECISION_FP64, "fp64"), (player.PRECISION_FP32, "fp32"), (player.PRECISION_STREAM, "stream")):
    if only and name != only:
        continue
    for pull in (8192, 2048) + ((512,) if prec == player.PRECISION_STREAM else ()):
        ts = run(lambda: player.SpeechPlayer(sr, precision=prec, noise=player.NOISE_PHILOX, seed=1, streamId=0),
                 lambda p: p.queue_frames(g["frames"], g["min_dur"], g["fade_dur"], None, g["is_null"]), pull)
        report(name, pull, ts)
