"""CPU: the C-ABI library loads and exports every symbol the public headers declare; host-only helpers work;
without a CUDA device the engine fails loudly instead of falling back."""
import ctypes
import os
import re

import numpy as np
import pytest

from nvspeechplayer_b200 import player
from oracle import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    names = set()
    for h in ("speechPlayer.h", "speechPlayer_batch.h", "speechPlayer_ipa.h"):
        src = open(os.path.join(ROOT, "include", h)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names |= set(re.findall(r"\b(speechPlayer_[A-Za-z]+)\s*\(", src))
    return sorted(names)


def test_every_declared_symbol_is_exported():
    lib = player.load_library()
    syms = _declared_symbols()
    assert {"speechPlayer_initialize", "speechPlayer_queueFrame", "speechPlayer_synthesize", "speechPlayer_getLastIndex",
            "speechPlayer_terminate", "speechPlayer_synthesizeBatch"} <= set(syms)
    for name in syms:
        assert hasattr(lib, name), name
    assert b"sm_100a" in lib.speechPlayer_version()


def test_reference_wrapper_file_name_is_present():
    # the unchanged reference wrapper loads "speechPlayer.dll" next to itself (reference speechPlayer.py:42)
    assert os.path.exists(player.dllPath)
    assert ctypes.sizeof(player.Frame) == 376


def test_timeline_law_helper():
    lib = player.load_library()
    m = np.array([10, 10, 4, 0, 1, 10, 10, 10], dtype=np.uint32)
    f = np.array([4, 1, 10, 1, 1, 9, 10, 0], dtype=np.uint32)
    want = [11, 11, 12, 3, 3, 11, 12, 11]  # SURVEY.md 3.3.1, probe-verified on the compiled reference
    for i in range(len(m)):
        assert lib.speechPlayer_timelineSamples(m[i:].ctypes.data, f[i:].ctypes.data, 1) == want[i]
    assert lib.speechPlayer_timelineSamples(m.ctypes.data, f.ctypes.data, len(m)) == sum(want)
    assert oracle.timeline_samples(m, f) == sum(want)


@pytest.mark.parametrize("seed", [1, 2, 12345, 0])
def test_glibc_rand_replica_matches_libc(seed):
    lib = player.load_library()
    n = 100000
    got = np.zeros(n, dtype=np.int32)
    lib.speechPlayer_debugGlibcRand(seed, n, got.ctypes.data_as(ctypes.c_void_p))
    want = oracle.libc_rand_sequence(seed, n)
    np.testing.assert_array_equal(got, want)
    if seed == 1:
        assert got[0] == 1804289383 and got[1] == 846930886


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    lib = player.load_library()
    assert not lib.speechPlayer_initialize(22050)
    assert b"CUDA" in lib.speechPlayer_lastError()
    with pytest.raises(player.EngineError):
        player.SpeechPlayer(22050)
    with pytest.raises(player.EngineError):
        player.Batch(22050, 4)


def test_output_sinks_and_voice_tables(tmp_path):
    """Host-side sinks of SURVEY.md 8f rank 4 (reference lavPlayer.py:14-17: int16 / 32767.0) and the NVDA-style voice
    tables of rank 2 (reference nvdaAddon/synthDrivers/nvSpeechPlayer/__init__.py:86-125)."""
    import wave
    pcm = np.array([0, 1, -1, 32767, -32767, -32768, 12345], dtype=np.int16)
    f = player.to_float32(pcm)
    assert f.dtype == np.float32 and f[3] == 1.0 and f[4] == -1.0 and abs(f[6] - 12345 / 32767.0) < 1e-7
    path = str(tmp_path / "x.wav")
    player.write_wav(path, pcm, 22050)
    with wave.open(path, "rb") as w:
        assert (w.getnchannels(), w.getsampwidth(), w.getframerate(), w.getnframes()) == (1, 2, 22050, len(pcm))
        np.testing.assert_array_equal(np.frombuffer(w.readframes(len(pcm)), dtype="<i2"), pcm)
    va, vm = player.nvda_voice_tables([{"cb1_mul": 1.3, "pa6_mul": 1.3, "fricationAmplitude_mul": 0.85},
                                       {"cf4": 3770, "cf1_mul": 1.01, "voiceAmplitude": 0, "aspirationAmplitude": 1}])
    names = list(player.PARAM_NAMES)
    assert vm[0, names.index("cb1")] == 1.3 and np.isnan(va[0]).all()
    assert va[1, names.index("cf4")] == 3770 and vm[1, names.index("cf1")] == 1.01 and va[1, names.index("voiceAmplitude")] == 0
