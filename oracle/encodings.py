"""Oracle (TEST INFRASTRUCTURE): event encodings on CPU.

Restates dataloader/encodings.py:30-85 (events_to_image / events_to_voxel /
events_to_channels).  Counts are integers (exact in fp32), so the CUDA kernels must match
these bit-exactly; the voxel grid has fractional weights and is a tolerance-tier check.
"""
import numpy as np
import torch


def events_to_image(xs, ys, ps, sensor_size, accumulate=True):
    """encodings.py:30-45: img[ys, xs] (+)= ps.  Non-accumulating writes keep the LAST event
    per pixel (index_put_ on CPU applies indices in order)."""
    H, W = sensor_size
    img = torch.zeros(H * W, dtype=torch.float32)
    lin = ys.long() * W + xs.long()
    if accumulate:
        img.index_add_(0, lin, ps.float())
    else:
        img[lin] = ps.float()
    return img.view(H, W)


def events_to_channels(xs, ys, ps, sensor_size):
    """encodings.py:70-85: two per-polarity count images; a negative event adds (-1)*(-1)=+1."""
    pos = ps * (ps * (ps >= 0))
    neg = ps * (ps * (ps <= 0))
    return torch.stack([events_to_image(xs, ys, pos, sensor_size), events_to_image(xs, ys, neg, sensor_size)])


def events_to_voxel(xs, ys, ts, ps, num_bins, sensor_size, round_ts=False):
    """encodings.py:48-67: temporal bilinear voxel grid, ts in [0,1]."""
    t = ts * (num_bins - 1)
    if round_ts:
        t = torch.round(t)
    out = []
    for b in range(num_bins):
        w = torch.clamp_min(1.0 - torch.abs(t - b), 0.0)
        out.append(events_to_image(xs, ys, ps * w, sensor_size))
    return torch.stack(out)


def events_to_channels_np(xs, ys, ps, sensor_size):
    """Pure-numpy integer restatement (np.add.at) used to cross-check the torch one."""
    H, W = sensor_size
    out = np.zeros((2, H, W), dtype=np.int64)
    xi, yi = xs.astype(np.int64), ys.astype(np.int64)
    np.add.at(out[0], (yi[ps > 0], xi[ps > 0]), 1)
    np.add.at(out[1], (yi[ps < 0], xi[ps < 0]), 1)
    return out.astype(np.float32)


def synth_events(n, sensor_size, gen):
    """Synthetic window of events in the loader's layout (base.py:71-99):
    integer-valued float coords, ts sorted and min-max normalised to [0,1], ps in {-1,+1}."""
    H, W = sensor_size
    xs = torch.randint(0, W, (n,), generator=gen).float()
    ys = torch.randint(0, H, (n,), generator=gen).float()
    ts = torch.sort(torch.rand(n, generator=gen)).values
    if n > 1 and float(ts.max() - ts.min()) > 0:
        ts = (ts - ts.min()) / (ts.max() - ts.min())
    else:
        ts = torch.zeros_like(ts)
    ps = torch.randint(0, 2, (n,), generator=gen).float() * 2 - 1
    return xs, ys, ts, ps
