"""GPU parity: the loader window formatter (raw event windows -> batch tensors) against fixtures
written by the reference's own ``H5Loader.__getitem__`` and against the CPU oracle on seeded streams."""
import numpy as np
import pytest
import torch

from snnflow_testutil import LOADER_FIXTURES, LOADER_KEYS, load_golden, loader_windows

pytestmark = pytest.mark.gpu


def _config(mode, B, H, W, target, hot):
    return {"data": {"mode": mode}, "loader": {"resolution": list(target if mode != "events" else (H, W)),
                                               "std_resolution": [H, W], "batch_size": B,
                                               "augment": ["Horizontal", "Vertical", "Polarity"],
                                               "augment_prob": [0.5, 0.5, 0.5]},
            "hot_filter": dict(enabled=hot is not None, **(hot or dict(max_px=100, min_obvs=5, max_rate=0.8)))}


def _check(batch, want, what, voxel_tol=2e-6):
    for k in LOADER_KEYS:
        got = batch[k].cpu().numpy()
        assert got.shape == want[k].shape, (what, k, got.shape, want[k].shape)
        if k == "event_voxel":   # fractional weights: fixed-point accumulation vs the reference's sequential fp32 sums
            np.testing.assert_allclose(got, want[k], rtol=0, atol=voxel_tol, err_msg=f"{what} {k}")
        else:                    # counts, masks, lists: bit-exact
            assert np.array_equal(got, want[k]), f"{what} {k}: {np.abs(got - want[k]).max()}"


@pytest.mark.parametrize("name", LOADER_FIXTURES)
@pytest.mark.parametrize("ts64", [True, False])
def test_formatter_matches_reference_loader(name, ts64):
    import snnflow_b200 as snnflow
    g = load_golden(name)
    B, H, W, nb = int(g["B"]), int(g["H"]), int(g["W"]), int(g["num_bins"])
    hot = dict(max_px=int(g["hot"][0]), min_obvs=int(g["hot"][1]), max_rate=float(g["hot"][2])) if g["hot"].size else None
    fmt = snnflow.EventWindowFormatter(_config(str(g["mode"]), B, H, W, tuple(int(v) for v in g["target"]), hot), nb,
                                       round_encoding=bool(g["round_enc"]))
    for b in range(B):
        for j, m in enumerate(("Horizontal", "Vertical", "Polarity")):
            fmt.batch_augmentation[m][b] = bool(g["flips"][b][j])
    fmt._flips = None
    for it, wins in enumerate(loader_windows(g)):
        xs = torch.from_numpy(np.stack([w[0] for w in wins])).cuda()          # int16, as stored in the HDF5 files
        ys = torch.from_numpy(np.stack([w[1] for w in wins])).cuda()
        ps = torch.from_numpy(np.stack([w[3] for w in wins])).cuda()          # int8 in {0,1}
        ts = np.stack([w[2] for w in wins])                                   # float64 absolute seconds, t0 = 10
        if ts64:
            batch = fmt.format_batch(xs, ys, torch.from_numpy(ts).cuda(), ps, t0=[10.0] * B)
        else:
            batch = fmt.format_batch(xs, ys, torch.from_numpy((ts - 10.0).astype(np.float32)).cuda(), ps)
        _check(batch, {k: g[f"item{it}.{k}"] for k in LOADER_KEYS}, f"{name} item {it}")
    if hot is not None:
        assert int(fmt.hot_idx[0]) == int(g["n_items"])


def test_formatter_vs_oracle_seeded_and_hot_tiebreak():
    """Larger seeded stream at 128x128 (the training resolution) against the CPU oracle, with more hot pixels than
    max_px so the literal argmax order (highest rate, then lowest index) decides which ones go."""
    import snnflow_b200 as snnflow
    from oracle import loader as oload
    B, H, W, N, nb = 2, 128, 128, 20000, 5
    hotcfg = dict(max_px=3, min_obvs=1, max_rate=0.6)
    np.random.seed(5)
    fmt = snnflow.EventWindowFormatter(_config("events", B, H, W, (H, W), hotcfg), nb)
    ohot = oload.HotFilter(B, (H, W), **hotcfg)
    gen = torch.Generator().manual_seed(11)
    for it in range(4):
        xs = torch.randint(0, W, (B, N), generator=gen)
        ys = torch.randint(0, H, (B, N), generator=gen)
        # six pixels that fire in (almost) every window, max_px = 3: two of them in all four windows, four of them in
        # three (rate 0.75 > max_rate) - the cut falls between counts, and ties at the cut go by lowest flat index
        for j in range(6 if it > 0 else 2):
            xs[:, j::997], ys[:, j::997] = 5 + 9 * j, 100 - 7 * j
        ts = torch.sort(torch.rand(B, N, generator=gen, dtype=torch.float64), dim=1).values + 3.0
        ps = torch.randint(0, 2, (B, N), generator=gen)
        batch = fmt.format_batch(xs.cuda(), ys.cuda(), ts.cuda(), ps.cuda(), t0=[3.0] * B)
        items = []
        for b in range(B):
            flips = tuple(fmt.batch_augmentation[m][b] for m in ("Horizontal", "Vertical", "Polarity"))
            items.append(oload.format_item(xs[b].float(), ys[b].float(), (ts[b] - 3.0).float(), ps[b].float(),
                                           resolution=(H, W), num_bins=nb, flips=flips, hot=ohot, batch=b))
        want = {k: v.numpy() for k, v in oload.collate(items).items()}
        _check(batch, want, f"seeded item {it}", voxel_tol=2e-5)
        if it >= 1:
            assert int((want["event_mask"][0, 0] == 0).sum()) > 0
    # exactly max_px pixels are removed per slot once the filter is active
    act = oload.hot_event_mask(ohot.events[0] / ohot.idx[0], ohot.idx[0], **hotcfg)
    assert int((act == 0).sum()) == 3


def test_formatter_empty_window_and_errors():
    import snnflow_b200 as snnflow
    from snnflow_b200 import _lib
    fmt = snnflow.EventWindowFormatter(_config("events", 2, 16, 16, (16, 16), None), 3)
    z = torch.zeros((2, 0), device="cuda")
    batch = fmt.format_batch(z, z, z, z)
    assert batch["event_list"].shape == (2, 0, 4) and batch["event_list_pol_mask"].shape == (2, 0, 2)
    for k in ("event_cnt", "event_voxel", "event_mask"):
        assert float(batch[k].abs().sum()) == 0.0
    with pytest.raises(_lib.SnnflowError):
        fmt.format_batch(torch.zeros(2, 4), torch.zeros(2, 4), torch.zeros(2, 4), torch.zeros(2, 4))
    with pytest.raises(ValueError):
        fmt.format_batch(torch.zeros(3, 4).cuda(), torch.zeros(3, 4).cuda(), torch.zeros(3, 4).cuda(), torch.zeros(3, 4).cuda())
    # a window whose timestamps are all equal normalises to zeros (base.py:97-98)
    one = torch.ones((2, 5), device="cuda")
    for m in fmt.batch_augmentation:
        fmt.batch_augmentation[m] = [False, False]
    fmt._flips = None
    b2 = fmt.format_batch(one, one, one * 7.0, one)
    assert float(b2["event_list"][..., 0].abs().sum()) == 0.0
    assert float(b2["event_cnt"][:, 0, 1, 1].sum()) == 10.0
