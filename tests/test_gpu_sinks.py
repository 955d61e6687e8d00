"""GPU: the output sinks on the device against the int16 path and the reference's own formulas:
float32 = int16 / 32767.0 (reference lavPlayer.py:17), and the ragged rows of a batch packed back to back (the NVDA driver
appends every buffer it pulls to one audio stream, nvdaAddon/synthDrivers/nvSpeechPlayer/__init__.py:70-73)."""
import numpy as np
import pytest

from nvspeechplayer_b200 import player, workloads

pytestmark = pytest.mark.gpu


def _rendered_batch():
    import torch
    sr, n = 22050, 257
    # ragged: every stream drains at its own length (0.05 .. 0.5 s), some of them odd
    rng = np.random.default_rng(5)
    streams = [workloads.random_stream(300 + s, float(rng.uniform(0.05, 0.5)), sr) for s in range(n)]
    fb = workloads._concat(sr, streams, np.arange(300, 300 + n, dtype=np.uint64))
    count = int(0.4 * sr) + 3
    stride = (count + 7) // 8 * 8
    dev = torch.device("cuda", 0)
    d_out = torch.zeros((n, stride), dtype=torch.int16, device=dev)
    d_written = torch.zeros(n, dtype=torch.int32, device=dev)
    b = player.Batch(sr, n, precision=player.PRECISION_FP32, seed=9, stream_ids=fb.stream_ids)
    b.set_frames_host(fb)
    b.synthesize_device(count, d_out.data_ptr(), stride, d_written.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    b.close()
    return torch, dev, n, count, stride, d_out, d_written


def test_float32_on_the_device_is_int16_over_32767():
    torch, dev, n, count, stride, d_out, d_written = _rendered_batch()
    for out_stride in (stride, count + 1):  # the 16-byte path and the scalar path
        d_f = torch.full((n, out_stride), 7.0, dtype=torch.float32, device=dev)
        player.to_float32_device(d_out.data_ptr(), stride, n, count, d_f.data_ptr(), out_stride, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        got = d_f.cpu().numpy()
        want = player.to_float32(d_out[:, :count].cpu().numpy())          # (int16 / 32767.0) -> float32, as lavPlayer.py:17
        assert got[:, :count].tobytes() == want.tobytes()
        assert (got[:, count:] == 7.0).all()                              # nothing beyond sampleCount is touched
    assert np.abs(want).max() > 0.05


def test_concatenation_on_the_device():
    torch, dev, n, count, stride, d_out, d_written = _rendered_batch()
    written = d_written.cpu().numpy().astype(np.int64)
    assert (written < count).any() and (written == count).any() and (written % 2 == 1).any()
    host = d_out.cpu().numpy()
    d_off = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    d_packed = torch.full((n * count + 8,), 12345, dtype=torch.int16, device=dev)
    player.concatenate_device(d_out.data_ptr(), stride, n, count, d_written.data_ptr(), d_off.data_ptr(), d_packed.data_ptr(),
                              torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    off = d_off.cpu().numpy()
    np.testing.assert_array_equal(off, np.concatenate([[0], np.cumsum(written)]))
    want = np.concatenate([host[s, :written[s]] for s in range(n)])
    got = d_packed.cpu().numpy()
    np.testing.assert_array_equal(got[:len(want)], want)
    assert (got[len(want):] == 12345).all()
    # without counts every row contributes sampleCount samples
    player.concatenate_device(d_out.data_ptr(), stride, n, count, None, d_off.data_ptr(), d_packed.data_ptr(),
                              torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(d_packed.cpu().numpy()[:n * count], host[:, :count].reshape(-1))
    assert int(d_off[-1]) == n * count
