"""IPA text -> Klatt frame queues through the native bulk producer (include/speechPlayer_ipa.h).

Mirrors the caller-facing names of the reference's ipa.py (generateFramesAndTiming, reference ipa.py:336-353) so that a
script written against it runs unchanged, and adds the bulk form the batch engine wants: many utterances in one native
call, straight into the flat arrays of speechPlayer_batchSetFramesHost (no per-frame ctypes call, no Frame objects;
SURVEY.md section 8f rank 1).  The phoneme table is the numeric content of the reference's data.py
(nvspeechplayer_b200/data/phoneme_table.npz, written by tests/golden/make_golden.py)."""
import ctypes

import numpy as np

from . import player, workloads

_FLAG_BITS = {"_isNasal": 1, "_isStop": 2, "_isLiquid": 4, "_isVowel": 8, "_isVoiced": 16, "_isAfricate": 32,
              "_copyAdjacent": 64, "_isSemivowel": 128}
_table = None


def _lib():
    L = player.load_library()
    if not getattr(L, "_ipa_ready", False):
        vp = ctypes.c_void_p
        L.speechPlayer_ipaTableCreate.restype = vp
        L.speechPlayer_ipaTableCreate.argtypes = [vp, vp, vp, vp, ctypes.c_uint]
        L.speechPlayer_ipaTableDestroy.restype = None
        L.speechPlayer_ipaTableDestroy.argtypes = [vp]
        L.speechPlayer_ipaFrames.restype = ctypes.c_longlong
        L.speechPlayer_ipaFrames.argtypes = [vp, vp, ctypes.c_uint, vp, vp, vp, vp, ctypes.c_int, ctypes.c_double, vp, vp, vp, vp,
                                             vp, vp, vp, ctypes.c_ulonglong, ctypes.c_uint]
        L._ipa_ready = True
    return L


def table():
    """The native phoneme table handle (created once from the shipped .npz)."""
    global _table
    if _table is None:
        t = workloads.phoneme_table()
        n = len(t["names"])
        keys = np.zeros((n, 3), dtype=np.uint32)
        for i, name in enumerate(t["names"]):
            cps = [ord(c) for c in name]
            keys[i, :len(cps)] = cps
        flags = np.zeros(n, dtype=np.uint32)
        for fname, bit in _FLAG_BITS.items():
            flags |= (np.asarray(t["flags"][fname]) != 0).astype(np.uint32) * np.uint32(bit)
        values = np.ascontiguousarray(t["table"], dtype=np.float64)
        present = np.ascontiguousarray(t["present"], dtype=np.uint8)
        h = _lib().speechPlayer_ipaTableCreate(keys.ctypes.data, values.ctypes.data, present.ctypes.data, flags.ctypes.data, n)
        if not h:
            raise player.EngineError(player.last_error())
        _table = h
    return _table


def _per_text(x, n, dtype=np.float64):
    if x is None:
        return None
    a = np.ascontiguousarray(np.broadcast_to(np.asarray(x, dtype=dtype), (n,)))
    return a


def frames_for_texts(texts, speed=1, basePitch=100, inflection=0.5, clauseType=None, sample_rate=22050,
                     trailing_silence_ms=-1.0, stream_ids=None, threads=0, with_ms=False):
    """Many IPA clauses -> one workloads.FrameBatch (durations in samples at sample_rate), one stream per clause.
    speed / basePitch / inflection / clauseType: scalars or per-text sequences.  trailing_silence_ms >= 0 appends the
    NULL frame reference test_speakIpa.py:26 queues after each line."""
    L, t = _lib(), table()
    n = len(texts)
    enc = [s.encode("utf8") for s in texts]
    arr = (ctypes.c_char_p * max(n, 1))(*enc)
    sp, bp, inf = _per_text(speed, n), _per_text(basePitch, n), _per_text(inflection, n)
    ct = None
    if clauseType is not None:
        cl = [clauseType] * n if isinstance(clauseType, str) or clauseType is None else list(clauseType)
        ct = np.frombuffer(b"".join((c or ".").encode("ascii")[:1] for c in cl), dtype=np.uint8).copy()
    ptr = lambda a: None if a is None else a.ctypes.data
    offsets = np.zeros(n + 1, dtype=np.int64)
    # one native pass: a character yields at most three entries (aspiration, gap, phoneme), plus the trailing silence
    cap = max(3 * sum(len(s) for s in texts) + n, 1)
    frames = np.empty((cap, 47))
    m = np.empty(cap, dtype=np.uint32)
    f = np.empty(cap, dtype=np.uint32)
    nul = np.empty(cap, dtype=np.uint8)
    dms = np.empty(cap) if with_ms else None
    fms = np.empty(cap) if with_ms else None
    total = L.speechPlayer_ipaFrames(t, arr, n, ptr(sp), ptr(bp), ptr(inf), ptr(ct), sample_rate, float(trailing_silence_ms),
                                     offsets.ctypes.data, frames.ctypes.data, m.ctypes.data, f.ctypes.data, nul.ctypes.data,
                                     ptr(dms), ptr(fms), cap, threads)
    if total < 0 or total > cap:
        raise player.EngineError(player.last_error() or "frame capacity estimate too small")
    k = int(total)
    ids = np.arange(n, dtype=np.uint64) if stream_ids is None else np.asarray(stream_ids, dtype=np.uint64)
    fb = workloads.FrameBatch(sample_rate=sample_rate, offsets=offsets, frames=frames[:k], min_dur=m[:k], fade_dur=f[:k],
                              is_null=nul[:k], user_index=np.full(k, -1, dtype=np.int32), stream_ids=ids)
    return (fb, dms[:k], fms[:k]) if with_ms else fb


def generateFramesAndTiming(ipaText, speed=1, basePitch=100, inflection=0.5, clauseType=None):
    """Same generator as reference ipa.py:336-353: yields (Frame or None, durationMs, fadeMs)."""
    fb, dms, fms = frames_for_texts([ipaText], speed, basePitch, inflection, clauseType, with_ms=True)
    for j in range(len(dms)):
        if fb.is_null[j]:
            yield None, float(dms[j]), float(fms[j])
        else:
            fr = player.Frame()
            for name, v in zip(player.PARAM_NAMES, fb.frames[j]):
                setattr(fr, name, float(v))
            yield fr, float(dms[j]), float(fms[j])
