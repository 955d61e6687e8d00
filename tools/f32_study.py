"""Numerics study (CPU): FP32 production arithmetic (host build) vs the reference, per workload and sample rate."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle
from tests.hostsim import sim
from nvspeechplayer_b200 import workloads

port = oracle.PortLib()
g = np.load("tests/golden/config1.npz")
for sr in (16000, 22050, 44100):
    scale = sr / 22050.0
    m = (g["min_dur"] * scale).astype(np.uint32); f = (g["fade_dur"] * scale).astype(np.uint32)
    want = port.render(sr, g["frames"], m, f, g["is_null"], noise=("philox", 1, 2))
    got, _ = sim.render_f32(sr, g["frames"], m, f, g["is_null"], seed=1, stream=2)
    print("config1 @%5d: n=%7d  <=1LSB %.5f exact %.4f SNR %.1f dB max %g" % ((sr, len(want)) + sim.parity(got, want)))
for sr in (16000, 22050, 44100):
    res = []
    for sid in range(8):
        fr, m, f, nul, ux = workloads.random_stream(sid, 10.0, sr)
        n = int(10.0 * sr)
        want = port.render(sr, fr, m, f, nul, ux, max_samples=n, noise=("philox", 3, sid))
        got, _ = sim.render_f32(sr, fr, m, f, nul, ux, max_samples=n, seed=3, stream=sid)
        res.append(sim.parity(got, want))
    r = np.array(res)
    print("random  @%5d: 8x10s   <=1LSB min %.5f mean %.5f  exact %.4f  SNR min %.1f dB  max|d| %g" %
          (sr, r[:, 0].min(), r[:, 0].mean(), r[:, 1].mean(), r[:, 2].min(), r[:, 3].max()))
