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
