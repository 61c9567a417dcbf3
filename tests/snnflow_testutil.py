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
