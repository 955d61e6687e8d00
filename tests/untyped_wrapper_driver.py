"""Drives the engine the way the reference's UNCHANGED ctypes wrapper does (reference speechPlayer.py:42-65), in a fresh
process: the library is found under the literal file name speechPlayer.dll, loaded with a bare cdll.LoadLibrary, and NO
restype / argtypes are declared -- so the handle comes back through a C int, the frame goes in as byref(Structure) or None,
purgeQueue is a Python bool, and the sample buffer is a (c_short * n)() array.  Renders the queueFrame sequence recorded for
BASELINE config 1 (tests/golden/config1.npz, durations already in samples) and writes the PCM to argv[1].

usage: python tests/untyped_wrapper_driver.py out.npy     (test infrastructure; run by tests/test_gpu_reference_wrapper.py)"""
import os
import sys
from ctypes import Structure, byref, c_double, c_short, cdll

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nvspeechplayer_b200.workloads import PARAM_NAMES  # noqa: E402  (the 47 field names, ABI order)


class Frame(Structure):
    _fields_ = [(name, c_double) for name in PARAM_NAMES]


def main(out_path):
    g = np.load(os.path.join(ROOT, "tests", "golden", "config1.npz"))
    dll = cdll.LoadLibrary(os.path.join(ROOT, "nvspeechplayer_b200", "speechPlayer.dll"))
    handle = dll.speechPlayer_initialize(int(g["sample_rate"]))   # restype unset: a C int
    assert isinstance(handle, int) and 0 < handle < 2 ** 31, "the handle must survive the trip through a C int"
    assert dll.speechPlayer_getLastIndex(handle) == -1
    for j in range(len(g["min_dur"])):
        frame = None
        if not g["is_null"][j]:
            frame = Frame(*[float(x) for x in g["frames"][j]])
        dll.speechPlayer_queueFrame(handle, byref(frame) if frame else None, int(g["min_dur"][j]), int(g["fade_dur"][j]), -1, False)
    chunks = []
    while True:
        buf = (c_short * 8192)()
        res = dll.speechPlayer_synthesize(handle, 8192, buf)
        if res <= 0:
            break
        chunks.append(np.frombuffer(buf, dtype=np.int16, count=min(res, 8192)).copy())
    # a purge with a Python bool, then silence: the cancel path of the NVDA driver
    dll.speechPlayer_queueFrame(handle, None, 100, 50, 7, True)
    tail = (c_short * 512)()
    got = dll.speechPlayer_synthesize(handle, 512, tail)
    assert got == 101 and dll.speechPlayer_getLastIndex(handle) == 7, (got, dll.speechPlayer_getLastIndex(handle))
    dll.speechPlayer_terminate(handle)
    np.save(out_path, np.concatenate(chunks))


if __name__ == "__main__":
    main(sys.argv[1])
