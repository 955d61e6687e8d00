"""GPU: the in-library multi-GPU path (speechPlayer_multiBatch*): contiguous stream ranges balanced by ticks,
one host thread + per-device batch per shard, gather into disjoint rows of one host buffer.  On a one-GPU box the shards all
sit on device 0 (an ordinal may repeat) -- the partition, the threading, the re-based queue offsets and the gather are the
same code; with more GPUs visible the shards spread over them."""
import ctypes

import numpy as np
import pytest

from nvspeechplayer_b200 import player, sharding, workloads

pytestmark = pytest.mark.gpu


def _device_count():
    cu = ctypes.CDLL("libcuda.so.1")
    n = ctypes.c_int(0)
    cu.cuInit(0)
    cu.cuDeviceGetCount(ctypes.byref(n))
    return n.value


@pytest.mark.parametrize("prec", [player.PRECISION_FP32, player.PRECISION_FP64], ids=["fp32", "fp64"])
def test_multibatch_equals_one_batch_and_balances_by_ticks(prec):
    sr, n, shards = 22050, 301, 3
    rng = np.random.default_rng(8)
    secs = rng.uniform(0.05, 0.6, n)
    secs[:40] = 0.6            # the first streams are long: equal COUNTS would be unbalanced
    streams = [workloads.random_stream(900 + s, float(secs[s]), sr) for s in range(n)]
    fb = workloads._concat(sr, streams, np.arange(900, 900 + n, dtype=np.uint64))
    count = int(0.5 * sr)
    devices = [d % max(_device_count(), 1) for d in range(shards)]
    mb = player.MultiBatch(sr, n, devices, precision=prec, seed=21, stream_ids=fb.stream_ids)
    mb.set_frames_host(fb)
    first = mb.shards()
    assert list(first) == sharding.balanced_ranges(fb.timeline_samples(), shards)
    ticks = fb.timeline_samples().astype(np.int64)
    per = [int(ticks[first[d]:first[d + 1]].sum()) for d in range(shards)]
    assert max(per) - min(per) <= 2 * int(ticks.max()) + 2, per   # a cut lands within one stream of the ideal point
    assert first[1] - first[0] < first[2] - first[1]   # fewer of the long streams in the first shard
    parts = [mb.synthesize_host(c) for c in (3000, count - 3000)]
    out = np.concatenate([p[0] for p in parts], axis=1)
    written = parts[0][1].astype(np.int64) + parts[1][1]
    idx = mb.last_indices()
    mb.close()
    b = player.Batch(sr, n, precision=prec, seed=21, stream_ids=fb.stream_ids)
    b.set_frames_host(fb)
    ref = [b.synthesize_host(c) for c in (3000, count - 3000)]
    np.testing.assert_array_equal(out, np.concatenate([p[0] for p in ref], axis=1))
    np.testing.assert_array_equal(written, ref[0][1].astype(np.int64) + ref[1][1])
    np.testing.assert_array_equal(idx, b.last_indices())
    np.testing.assert_array_equal(written, np.minimum(ticks, count))
    b.close()
