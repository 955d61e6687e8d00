"""Shared parity metrics and bars (BASELINE.json north_star).

FP64 kernel : int16 bit-exact on >= 99.99 % of samples, <= 1 LSB on the rest.
FP32 kernel : <= 1 LSB on >= 99.9 % of samples and >= 60 dB SNR against the reference.
"""
import numpy as np

F32_WITHIN_1LSB = 0.999
F32_SNR_DB = 60.0


def metrics(got, want):
    """(fraction within 1 LSB, fraction exact, SNR dB, max |diff|) over the common prefix."""
    n = min(len(got), len(want))
    g = np.asarray(got[:n], dtype=np.float64)
    w = np.asarray(want[:n], dtype=np.float64)
    d = g - w
    sig, err = float((w * w).sum()), float((d * d).sum())
    snr = float("inf") if err == 0 else 10 * np.log10(max(sig, 1e-30) / err)
    if n == 0:
        return 1.0, 1.0, float("inf"), 0.0
    return float((np.abs(d) <= 1).mean()), float((d == 0).mean()), snr, float(np.abs(d).max())


def assert_f32_parity(got, want, what="", within=F32_WITHIN_1LSB, snr_db=F32_SNR_DB, short_ok=2):
    assert len(got) == len(want), "%s: %d vs %d samples" % (what, len(got), len(want))
    w1, exact, snr, mx = metrics(got, want)
    bad = int(round((1 - w1) * len(want)))
    # short renders: a couple of 2-LSB samples would break a percentage bar that is meant for long audio
    assert w1 >= within or bad <= short_ok, "%s: only %.5f within 1 LSB (max %g, snr %.1f dB)" % (what, w1, mx, snr)
    quiet = float(np.abs(np.asarray(want, dtype=np.float64)).max()) < 50 if len(want) else True
    assert snr >= snr_db or quiet or mx <= 1, "%s: SNR %.1f dB (max %g)" % (what, snr, mx)
    return w1, exact, snr, mx
