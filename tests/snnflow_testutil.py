"""Shared helpers for the test-suite (golden fixture loading, comparison utilities)."""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False) as f:
        return {k: f[k] for k in f.files}


def spike_mismatch_outside_band(z, z_ref, v_ref, theta, band=1e-5):
    """Number of spike mismatches at neurons whose reference membrane is further than ``band`` from
    threshold, and the number of neurons inside the band."""
    near = np.abs(v_ref - theta) <= band
    bad = (z != z_ref) & ~near
    return int(bad.sum()), int(near.sum())


LOADER_FIXTURES = ["loader_events_hot", "loader_events_round", "loader_gtflow_pool"]
LOADER_KEYS = ("event_cnt", "event_voxel", "event_mask", "event_list", "event_list_pol_mask")


def loader_windows(g):
    """Replay plan of a loader fixture: for every item, the raw event window of every batch slot exactly as
    ``H5Loader.get_events`` hands it to ``event_formatting`` (h5.py:124-132: float64 seconds minus t0)."""
    B, n_win, n_items = int(g["B"]), int(g["n_win"]), int(g["n_items"])
    plan = []
    for it in range(n_items):
        wins = []
        for b in range(B):
            if str(g["mode"]) == "events":
                i0, i1 = it * n_win, (it + 1) * n_win
            else:
                i0, i1 = [int(v) for v in g[f"item{it}.range"]]
            wins.append(tuple(g[f"raw{b}.{k}"][i0:i1] for k in ("xs", "ys", "ts", "ps")))
        plan.append(wins)
    return plan


def synth_window(T, B, N, H, W, seed):
    """A synthetic training window in the loader's layout (dataloader/base.py:261-278), CPU tensors."""
    import torch
    g = torch.Generator().manual_seed(seed)
    xs = torch.randint(0, W, (T, B, N), generator=g).float()
    ys = torch.randint(0, H, (T, B, N), generator=g).float()
    ts = torch.sort(torch.rand(T, B, N, generator=g), dim=2).values
    ts = (ts - ts.amin(2, keepdim=True)) / (ts.amax(2, keepdim=True) - ts.amin(2, keepdim=True))
    ps = torch.randint(0, 2, (T, B, N), generator=g).float() * 2 - 1
    lin = ys.long() * W + xs.long()
    cnt = torch.zeros(T, B, 2, H * W)
    cnt[:, :, 0].scatter_add_(2, lin, (ps > 0).float())
    cnt[:, :, 1].scatter_add_(2, lin, (ps < 0).float())
    cnt = cnt.view(T, B, 2, H, W)
    return {"event_cnt": cnt.contiguous(), "event_list": torch.stack([ts, ys, xs, ps], dim=3).contiguous(),
            "event_list_pol_mask": torch.stack([(ps > 0).float(), (ps < 0).float()], dim=3).contiguous(),
            "event_mask": (cnt.sum(2, keepdim=True) > 0).float().contiguous()}


def grad_report(got, ref, rtol=1e-4, atol_rel=1e-6):
    """Element-wise comparison of a gradient tensor with its reference: returns (fraction of elements with
    |got - ref| <= rtol * |ref| + atol_rel * max|ref|,  norm-wise relative error).  Both numpy float arrays."""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    scale = float(np.abs(ref).max()) if ref.size else 0.0
    ok = np.abs(got - ref) <= rtol * np.abs(ref) + atol_rel * scale
    nrm = float(np.linalg.norm(ref))
    return float(ok.mean()) if ok.size else 1.0, (float(np.linalg.norm(got - ref)) / nrm if nrm > 0 else float(np.linalg.norm(got)))
