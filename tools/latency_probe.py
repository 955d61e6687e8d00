#!/usr/bin/env python
"""Per-handle pull latency of the five-symbol drop-in (what the NVDA audio thread sees): sampleIpa.txt frames queued on
one player, 8192-sample speechPlayer_synthesize pulls, wall clock per pull.  "stream" is SPEECHPLAYER_PRECISION_STREAM: the
time-parallel block kernel with the frame manager on the host."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nvspeechplayer_b200 import player
g = np.load(os.path.join(ROOT, "tests", "golden", "config1.npz"))
sr = int(g["sample_rate"])
only = os.environ.get("NVSP_PROBE_ONLY")
for prec, name in ((player.PRECISION_FP64, "fp64"), (player.PRECISION_FP32, "fp32"), (player.PRECISION_STREAM, "stream")):
    if only and name != only:
        continue
    for pull in (8192, 2048) + ((512,) if prec == player.PRECISION_STREAM else ()):
        p = player.SpeechPlayer(sr, precision=prec, noise=player.NOISE_PHILOX, seed=1, streamId=0)
        p.queue_frames(g["frames"], g["min_dur"], g["fade_dur"], None, g["is_null"])
        p.synthesize_np(pull)  # warm-up (first launch, allocations)
        ts = []
        while True:
            t0 = time.perf_counter()
            c = p.synthesize_np(pull)
            ts.append(time.perf_counter() - t0)
            if c.size < pull:
                break
        p.close()
        ts = np.array(ts[:-1]) * 1e3
        print("%s pull %5d samples (%.0f ms of audio): median %.2f ms, p95 %.2f ms, max %.2f ms over %d pulls  -> %.0fx real time"
              % (name, pull, 1e3 * pull / sr, np.median(ts), np.percentile(ts, 95), ts.max(), len(ts), (1e3 * pull / sr) / np.median(ts)))
