"""Event encodings on the GPU: same free-function signatures as dataloader/encodings.py:30-85.

Inputs are CUDA fp32 tensors (``xs, ys, ts, ps`` of shape [N]); outputs are CUDA fp32 tensors with
the reference's shapes.  Counts are bit-exact; the voxel grid is accumulated in 64-bit fixed point
and is run-to-run deterministic.  ``events_to_channels_batched`` is the B-window variant the training
loop uses ([B,N] -> [B,2,H,W]) in a single launch.
"""
import torch

from . import _lib


def _chk(*ts):
    out = []
    for t in ts:
        if not t.is_cuda:
            raise _lib.SnnflowError("snnflow encodings run on CUDA tensors only (no CPU fallback)")
        out.append(t.float().contiguous())
    return out


def events_to_image(xs, ys, ps, sensor_size=(180, 240), accumulate=True):
    """dataloader/encodings.py:30-45."""
    xs, ys, ps = _chk(xs, ys, ps)
    H, W = sensor_size
    out = torch.empty((H, W), dtype=torch.float32, device=xs.device)
    scratch = None if accumulate else torch.empty((H, W), dtype=torch.int32, device=xs.device)
    _lib.check(_lib.lib().snnflow_encode_image(_lib.ptr(xs), _lib.ptr(ys), _lib.ptr(ps), _lib.ptr(out),
                                               _lib.ptr(scratch), xs.numel(), H, W, int(bool(accumulate)),
                                               _lib.stream()), "snnflow_encode_image")
    return out


def events_to_channels(xs, ys, ps, sensor_size=(180, 240)):
    """dataloader/encodings.py:70-85 -> [2,H,W] per-polarity event counts."""
    assert len(xs) == len(ys) and len(ys) == len(ps)
    xs, ys, ps = _chk(xs, ys, ps)
    H, W = sensor_size
    out = torch.empty((2, H, W), dtype=torch.float32, device=xs.device)
    _lib.check(_lib.lib().snnflow_encode_cnt(_lib.ptr(xs), _lib.ptr(ys), _lib.ptr(ps), _lib.ptr(out), xs.numel(), 1,
                                             H, W, _lib.stream()), "snnflow_encode_cnt")
    return out


def events_to_channels_batched(xs, ys, ps, sensor_size):
    """[B,N] event arrays -> [B,2,H,W] counts, one launch."""
    xs, ys, ps = _chk(xs, ys, ps)
    B, N = xs.shape
    H, W = sensor_size
    out = torch.empty((B, 2, H, W), dtype=torch.float32, device=xs.device)
    _lib.check(_lib.lib().snnflow_encode_cnt(_lib.ptr(xs), _lib.ptr(ys), _lib.ptr(ps), _lib.ptr(out), N, B, H, W,
                                             _lib.stream()), "snnflow_encode_cnt")
    return out


def events_to_voxel(xs, ys, ts, ps, num_bins, sensor_size=(180, 240), round_ts=False):
    """dataloader/encodings.py:48-67 -> [num_bins,H,W]."""
    assert len(xs) == len(ys) and len(ys) == len(ts) and len(ts) == len(ps)
    xs, ys, ts, ps = _chk(xs, ys, ts, ps)
    H, W = sensor_size
    out = torch.empty((num_bins, H, W), dtype=torch.float32, device=xs.device)
    scratch = torch.empty((num_bins, H, W), dtype=torch.int64, device=xs.device)
    _lib.check(_lib.lib().snnflow_encode_voxel(_lib.ptr(xs), _lib.ptr(ys), _lib.ptr(ts), _lib.ptr(ps), _lib.ptr(out),
                                               _lib.ptr(scratch), xs.numel(), num_bins, H, W, int(bool(round_ts)),
                                               _lib.stream()), "snnflow_encode_voxel")
    return out
