"""CPU, world_size 2 over gloo: the host logic of the multi-GPU path.  Streams shard by id with no data-path
collective; the only communication is combining per-rank scalars.  Checks that the shards tile the job exactly, that a
shard generated on its own rank equals the same streams of the unsharded workload (counter-based generator), and that
the combine step yields max(time) / sum(samples)."""
import hashlib
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nvspeechplayer_b200 import sharding, workloads

TOTAL, SECS, SR = 37, 0.2, 22050


def _digest(fb):
    h = hashlib.sha256()
    for a in (fb.offsets, fb.frames, fb.min_dur, fb.fade_dur, fb.is_null, fb.user_index, fb.stream_ids):
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, last = sharding.shard_range(TOTAL, rank, world)
    fb = workloads.random_frames(last - first, SECS, SR, first_stream=first)
    samples = int(fb.timeline_samples().sum())
    elapsed, total = sharding.combine(dist, torch.device("cpu"), 10.0 + rank, samples)
    gathered = [None] * world
    dist.all_gather_object(gathered, (first, last, _digest(fb), samples))
    if rank == 0:
        q.put((elapsed, total, gathered))
    dist.barrier()
    dist.destroy_process_group()


def test_shards_tile_the_job_and_match_the_unsharded_workload():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    world = 2
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    elapsed, total, gathered = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert elapsed == 11.0  # max over ranks
    assert [g[0] for g in gathered] == [0, 19] and [g[1] for g in gathered] == [19, 37]
    whole = workloads.random_frames(TOTAL, SECS, SR)
    assert total == float(whole.timeline_samples().sum()) == float(sum(g[3] for g in gathered))
    for first, last, digest, _ in gathered:
        assert digest == _digest(workloads.random_frames(last - first, SECS, SR, first_stream=first))
    # the union of the shards is the unsharded queue set, stream for stream
    for first, last, _, _ in gathered:
        part = workloads.random_frames(last - first, SECS, SR, first_stream=first)
        for k in range(last - first):
            for a, b in zip(part.stream(k), whole.stream(first + k)):
                np.testing.assert_array_equal(a, b)


def test_shard_range_properties():
    for total in (0, 1, 7, 64, 65536, 1_000_003):
        for world in (1, 2, 3, 8):
            r = [sharding.shard_range(total, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == total
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1
    assert sharding.weak_shard(65536, 3) == (196608, 262144)


def test_balanced_ranges_properties():
    """The in-library multi-GPU partition (speechPlayer_multiBatchSetFramesHost applies the same rule; the GPU test compares
    the two): contiguous, monotone, covers everything, tick sums within one stream of each other."""
    rng = np.random.default_rng(3)
    for n, shards in ((1, 1), (5, 8), (301, 3), (65536, 8), (1000, 2)):
        ticks = rng.integers(0, 20000, n)
        first = sharding.balanced_ranges(ticks, shards)
        assert first[0] == 0 and first[-1] == n and len(first) == shards + 1
        assert all(first[k] <= first[k + 1] for k in range(shards))
        if n >= 4 * shards:
            per = [int(ticks[first[k]:first[k + 1]].sum()) for k in range(shards)]
            assert max(per) - min(per) <= 2 * int(ticks.max()) + 2
    assert sharding.balanced_ranges([10, 10, 10, 10], 2) == [0, 2, 4]
    assert sharding.balanced_ranges([100, 1, 1, 1, 1], 2)[1] == 1
