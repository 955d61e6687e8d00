#!/usr/bin/env python
"""Study for the next step of the pull path: how much of the serial FP64 glottal-phase recurrence
    pos[t] = fmod(pos[t-1] + inc[t], 1)            (reference src/speechWaveGenerator.cpp:55)
can be taken off the single thread WITHOUT changing a bit.

Observation: while pos stays inside one binade [2^k, 2^(k+1)), every value it takes is a multiple of that binade's ulp
u = 2^(k-52), so RN(pos + inc) = pos + RN_u(inc) unless inc/u is exactly half-way (round-half-even then looks at pos).  The
rounded increments are integers on a fixed grid: inside such a RUN the recurrence is an integer prefix sum, associative,
scannable.  Only the SPECIAL ticks need the real FP64 addition in order: wraps, binade crossings, ties, and ticks whose
approximate phase (exact 2^-64 fixed-point prefix sum, off by < 1e-12) is too close to a binade boundary to classify.

The script classifies every tick from the APPROXIMATE phase only (what all threads can compute in parallel), rebuilds
the sequence with exact arithmetic on normal ticks and the FP64 recurrence on special ticks, and checks it bit for bit
against the plain serial loop.  Output: fraction of special ticks and of 8-tick groups that contain one."""
import math
import sys

import numpy as np


def serial(inc, pos0):
    out = np.empty_like(inc)
    pos = pos0
    for t, x in enumerate(inc):
        s = pos + x
        pos = s - math.trunc(s)
        out[t] = pos
    return out


def binade(x):
    return math.frexp(x)[1] - 1   # x in [2^k, 2^(k+1))


def classify(inc, pos0, margin=2.0 ** -30):
    """normal[t], I[t] (integer increment on the grid of tick t's binade), k[t]; uses only the approximate phase"""
    n = len(inc)
    q = [int(round(float(x) * 2.0 ** 64)) for x in inc]
    A = int(round(pos0 * 2.0 ** 64))
    normal = np.zeros(n, bool)
    I = np.zeros(n, object)
    K = np.zeros(n, int)
    for t in range(n):
        a_prev = (A % (1 << 64)) / 2.0 ** 64
        A += q[t]
        x = float(inc[t])
        s = a_prev + x
        if not (0.0 < x < 0.5) or a_prev < 2.0 ** -40 or s >= 1.0 - margin:
            continue
        k = binade(s)
        lo, hi = 2.0 ** k, 2.0 ** (k + 1)
        if binade(a_prev) != k or a_prev < lo * (1 + margin) or s > hi * (1 - margin):
            continue
        u = 2.0 ** (k - 52)
        r = x / u                      # exact: division by a power of two
        fl = math.floor(r)
        if r - fl == 0.5:
            continue                   # tie: round-half-even depends on the parity of pos/u
        normal[t] = True
        I[t] = int(fl) + (1 if r - fl > 0.5 else 0)
        K[t] = k
    return normal, I, K


def rebuild(inc, pos0, normal, I, K):
    out = np.empty_like(inc)
    pos = pos0
    for t in range(len(inc)):
        if normal[t]:
            pos = pos + float(I[t]) * 2.0 ** (int(K[t]) - 52)   # exact: multiples of u inside one binade
        else:
            s = pos + float(inc[t])
            pos = s - math.trunc(s)
        out[t] = pos
    return out


def cases(n, rng):
    sr = 22050.0
    yield "150 Hz @ 22050 (period of exactly 147 samples)", np.full(n, 150.0 / sr)
    yield "100 Hz @ 16000 (160 samples)", np.full(n, 100.0 / 16000.0)
    yield "261.6256 Hz constant", np.full(n, 261.6256 / sr)
    t = np.arange(n)
    yield "glide 110 -> 180 Hz", (110.0 + 70.0 * t / n) / sr
    yield "vibrato 5.5 Hz, 6 % * 0.125 around 220 Hz", 220.0 * (1.0 + 0.06 * 0.125 * np.sin(2 * np.pi * 5.5 * t / sr)) / sr
    yield "random walk 60..400 Hz", np.clip(200.0 + np.cumsum(rng.normal(0, 0.5, n)), 60.0, 400.0) / sr
    yield "60 Hz @ 44100", np.full(n, 60.0 / 44100.0)
    yield "400 Hz @ 16000", np.full(n, 400.0 / 16000.0)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    rng = np.random.default_rng(1)
    print("%-46s %8s %10s %12s %10s" % ("increments", "special", "runs", "groups w/", "bit-exact"))
    for name, inc in cases(n, rng):
        inc = np.asarray(inc, dtype=np.float64)
        pos0 = float(rng.random())
        want = serial(inc, pos0)
        normal, I, K = classify(inc, pos0)
        got = rebuild(inc, pos0, normal, I, K)
        special = int((~normal).sum())
        runs = int(np.count_nonzero(np.diff(np.concatenate([[0], normal.astype(int)])) == 1))
        groups = int(sum((~normal[g:g + 8]).any() for g in range(0, n, 8)))
        print("%-46s %7.1f%% %10d %11.1f%% %10s" % (name, 100.0 * special / n, runs, 100.0 * groups / ((n + 7) // 8),
                                                      "yes" if np.array_equal(got, want) else "NO"))


if __name__ == "__main__":
    main()
