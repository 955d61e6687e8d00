#!/usr/bin/env python
"""Single-process multi-GPU end-to-end line (in-library sharding, speechPlayer_multiBatch*): N GPUs, one host thread + one
per-device batch per GPU, host frames in, int16 out into ONE pinned host buffer.  The torchrun mode of bench.py (one process
per GPU) is the contract's N>1 launch; this is the same work driven through the library's own multi-GPU entry.

    python tools/multibatch_bench.py --gpus 8 [--streams-per-gpu 16384] [--seconds 10] [--slice-seconds 2] [--steps 2]"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--streams-per-gpu", type=int, default=16384)
    ap.add_argument("--seconds", type=float, default=10.0)
    ap.add_argument("--slice-seconds", type=float, default=2.0)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--sample-rate", type=int, default=22050)
    args = ap.parse_args()
    import torch
    from nvspeechplayer_b200 import player, workloads
    sr, n = args.sample_rate, args.gpus * args.streams_per_gpu
    count = int(round(args.seconds * sr))
    slice_ticks = min(count, max(int(args.slice_seconds * sr) // 64 * 64, 64))
    fb = workloads.random_frames(n, args.seconds, sr)
    out = torch.empty((n, slice_ticks), dtype=torch.int16).pin_memory().numpy()
    mb = player.MultiBatch(sr, n, list(range(args.gpus)), precision=player.PRECISION_FP32, seed=0xB200, stream_ids=fb.stream_ids)

    def step():
        mb.set_frames_host(fb)
        done, total = 0, 0
        while done < count:
            k = min(slice_ticks, count - done)
            o, w = mb.synthesize_host(k, out[:, :k] if k == slice_ticks else None)
            total += int(w.sum())
            done += k
        return total

    step()
    t0 = time.perf_counter()
    total = 0
    for _ in range(args.steps):
        total = step()
    dt = (time.perf_counter() - t0) / args.steps
    print(json.dumps({"tool": "multibatch_bench", "n_gpus": args.gpus, "streams": n, "seconds_per_stream": args.seconds,
                      "shards": [int(x) for x in mb.shards()], "ms_per_step": dt * 1e3, "e2e_audio_seconds_per_s": total / sr / dt,
                      "d2h_bytes_per_step": int(total * 2), "d2h_GBps": total * 2 / dt / 1e9,
                      "note": "one process, in-library sharding: host frames in, int16 out into one pinned host buffer"}))
    mb.close()


if __name__ == "__main__":
    main()
