"""TEST/BENCH INFRASTRUCTURE ONLY: the reference's CPU path timed on this box's host cores.

One OS PROCESS per core (the reference's noise is the process-global, internally locked libc rand(), SURVEY.md
7.3-8), each rendering a bounded share of the SAME synthetic workload the GPU arm renders (config 3 random
frames, same generator, same stream ids), timing only speechPlayer_synthesize.

kind "reference": oracle/_ref/libspeechPlayer_ref.so, the unmodified reference C++ compiled by oracle/Makefile.
kind "port":      oracle/libklatt_oracle.so, the plain-C restatement (fallback when _ref is absent).
"""
import multiprocessing as mp
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(args):
    kind, stream_ids, seconds, sample_rate, seed, workload = args
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    import numpy as np
    from nvspeechplayer_b200 import workloads
    from oracle import oracle
    n = int(round(seconds * sample_rate))
    lib = oracle.RefLib(philox=False) if kind == "reference" else oracle.PortLib()
    samples, busy = 0, 0.0
    buf = np.zeros(n, dtype=np.int16)
    import ctypes
    for sid in stream_ids:
        if workload == "random_frames":
            fr, m, f, nul, ux = workloads.random_stream(int(sid), seconds, sample_rate, seed)
        else:
            raise ValueError(workload)
        if kind == "reference":
            lib.seed(int(sid) + 1)
        p = lib.player(sample_rate)
        if kind == "port":
            p.noise_libc()
        for j in range(len(m)):
            p.queue_frame(None if nul[j] else fr[j], int(m[j]), int(f[j]), int(ux[j]))
        t0 = time.perf_counter()
        if kind == "reference":
            got = lib.lib.speechPlayer_synthesize(p.h, n, buf.ctypes.data_as(ctypes.c_void_p))
        else:
            got = lib.lib.klatt_oracle_synthesize(p.h, n, buf.ctypes.data_as(ctypes.c_void_p))
        busy += time.perf_counter() - t0
        samples += max(got, 0)
        p.close()
    return samples, busy


def available_kind():
    sys.path.insert(0, ROOT)
    from oracle import oracle
    return "reference" if os.path.exists(oracle.REF_SO) else "port"


def run(streams_per_core=4, seconds=10.0, sample_rate=22050, cores=None, seed=0xB200, first_stream=0, kind=None,
        workload="random_frames"):
    """Returns {"value": audio-seconds per wall-second (aggregate over cores), "cores", "kind", "sample",
    "ns_per_sample_core"}."""
    cores = cores or os.cpu_count() or 1
    kind = kind or available_kind()
    if kind == "port":
        from oracle import oracle
        if not oracle.have_port():
            oracle.build(quiet=True)
    ids = [list(range(first_stream + c * streams_per_core, first_stream + (c + 1) * streams_per_core)) for c in range(cores)]
    ctx = mp.get_context("spawn")  # never fork a process that may hold a CUDA context
    t0 = time.perf_counter()
    with ctx.Pool(cores) as pool:
        res = pool.map(_worker, [(kind, ids[c], seconds, sample_rate, seed, workload) for c in range(cores)])
    wall = time.perf_counter() - t0
    total = sum(r[0] for r in res)
    slowest = max(r[1] for r in res)  # the processes run concurrently: the job ends with the slowest core
    busy = sum(r[1] for r in res)
    return {
        "value": (total / sample_rate) / slowest if slowest > 0 else 0.0,
        "unit": "audio-seconds/s",
        "cores": cores,
        "kind": kind,
        "sample": "%d streams x %.1f s of the %s workload @%d Hz, one process per core, timing speechPlayer_synthesize only"
                  % (cores * streams_per_core, seconds, workload, sample_rate),
        "ns_per_sample_core": 1e9 * busy / max(total, 1),
        "synth_seconds_slowest_core": slowest,
        "wall_seconds_incl_spawn": wall,
    }


if __name__ == "__main__":
    import json
    print(json.dumps(run()))
