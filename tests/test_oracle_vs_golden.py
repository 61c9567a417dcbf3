"""CPU: the oracle restatement reproduces the reference-generated golden fixtures.

Bit-exact for spikes, counts and (dyadic-weight) membranes; tolerance (stated per assert) for
gradients, fractional-weight splats and loss values.
"""
import numpy as np
import pytest
import torch

from oracle import encodings as oenc
from oracle import firenet as ofn
from oracle import iwe as oiwe
from oracle import lif as olif
from oracle import loss as oloss
from snnflow_testutil import load_golden

T = torch.from_numpy

LAYER_FIXTURES = ["layer_ff_hard_arctan", "layer_ff_soft_super_res", "layer_rec_hard_arctan",
                  "layer_rec_soft_triangle", "layer_rec_hard_mgspike", "layer_rec_nodetach", "layer_head_counts", "layer_ff_c32",
                  "layer_rec_c32", "layer_rec_c32_rand"]


def run_layer(g, manual_backward=False):
    rec, hard, detach, has_res = [bool(v) for v in g["meta"]]
    act = str(g["activation"])
    x = T(g["x"]).clone().requires_grad_(True)
    w_ff = T(g["w_ff"]).clone().requires_grad_(True)
    w_rec = T(g["w_rec"]).clone().requires_grad_(True) if rec else None
    leak = T(g["leak"]).clone().requires_grad_(True)
    thresh = T(g["thresh"]).clone().requires_grad_(True)
    v = z = None
    vs, zs, outs, loss = [], [], [], 0
    for t in range(x.shape[0]):
        out, v, z, _ = olif.lif_step(x[t], w_ff, leak, thresh, v, z, w_rec,
                                     T(g["residual"][t]) if has_res else None,
                                     hard_reset=hard, detach=detach, activation=act)
        vs.append(v); zs.append(z); outs.append(out)
        loss = loss + (out * T(g["gout"][t])).sum()
    loss = loss + (v * T(g["gv_last"])).sum()
    loss.backward()
    return dict(v=torch.stack(vs), z=torch.stack(zs), out=torch.stack(outs), g_x=x.grad, dw_ff=w_ff.grad,
                dw_rec=None if w_rec is None else w_rec.grad, dleak=leak.grad, dthresh=thresh.grad)


@pytest.mark.parametrize("name", LAYER_FIXTURES)
def test_layer_forward_backward(name):
    g = load_golden(name)
    r = run_layer(g)
    exact = not name.endswith("_rand")
    if exact:
        assert np.array_equal(r["z"].detach().numpy(), g["z"])
        assert np.array_equal(r["v"].detach().numpy(), g["v"])
        assert np.array_equal(r["out"].detach().numpy(), g["out"])
    else:
        np.testing.assert_allclose(r["v"].detach().numpy(), g["v"], rtol=1e-5, atol=1e-6)
    for k in ("g_x", "dw_ff", "dw_rec", "dleak", "dthresh"):
        if r[k] is None:
            continue
        np.testing.assert_allclose(r[k].numpy(), g[k], rtol=1e-4, atol=1e-5, err_msg=k)


@pytest.mark.parametrize("name", LAYER_FIXTURES)
def test_layer_manual_backward_recurrences(name):
    """The hand-derived BPTT recurrences (what the CUDA kernel implements) match reference autograd."""
    g = load_golden(name)
    rec, hard, detach, has_res = [bool(v) for v in g["meta"]]
    act = str(g["activation"])
    x, w_ff = T(g["x"]), T(g["w_ff"])
    w_rec = T(g["w_rec"]) if rec else None
    lam, theta = T(g["lam"]), T(g["theta"])
    v_all, z_all = T(g["v"]), T(g["z"])
    nT = x.shape[0]
    zero = torch.zeros_like(v_all[0])
    g_v, g_z_next = T(g["gv_last"]).clone(), zero.clone()
    acc = dict(dw_ff=0, dw_rec=0, dlam=0, dtheta=0)
    gxs = [None] * nT
    for t in reversed(range(nT)):
        v_in = v_all[t - 1] if t > 0 else zero
        z_in = z_all[t - 1] if t > 0 else zero
        import torch.nn.functional as F
        cur = F.conv2d(x[t], w_ff, padding=1) + (F.conv2d(z_in, w_rec, padding=1) if rec else 0)
        b = olif.lif_step_backward(x[t], w_ff, w_rec, lam, theta, v_in, z_in, v_all[t], cur,
                                   T(g["gout"][t]) + g_z_next, g_v, hard_reset=hard, detach=detach, activation=act)
        gxs[t] = b["g_x"]
        g_v, g_z_next = b["g_v_in"], b["g_z_in"]
        for k in acc:
            if b[k] is not None:
                acc[k] = acc[k] + b[k]
    np.testing.assert_allclose(torch.stack(gxs).numpy(), g["g_x"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(acc["dw_ff"].numpy(), g["dw_ff"], rtol=1e-4, atol=1e-5)
    if rec:
        np.testing.assert_allclose(acc["dw_rec"].numpy(), g["dw_rec"], rtol=1e-4, atol=1e-5)
    lamv = lam.reshape(-1)
    dleak = acc["dlam"] * lamv * (1 - lamv)
    dthresh = acc["dtheta"] * (T(g["thresh"]).reshape(-1) >= 0.01).float()
    np.testing.assert_allclose(dleak.numpy(), g["dleak"].reshape(-1), rtol=2e-4, atol=1e-5)
    np.testing.assert_allclose(dthresh.numpy(), g["dthresh"].reshape(-1), rtol=2e-4, atol=1e-5)


@pytest.mark.parametrize("name", ["net_firenet_c8", "net_fireflownet_c8", "net_firenet_c32"])
def test_network_forward(name):
    g = load_golden(name)
    params = {k[len("param."):]: T(v) for k, v in g.items() if k.startswith("param.")}
    cnt = T(g["cnt"])
    states = [None] * 7
    with torch.no_grad():
        for t in range(cnt.shape[0]):
            flow, states, spikes = ofn.forward(params, cnt[t], states)
            assert np.array_equal(flow.numpy(), g["flow"][t])
            act = [float((cnt[t] != 0).float().mean())] + [float((s != 0).float().mean()) for s in spikes]
            np.testing.assert_allclose(act, g["activity"][t][:8], rtol=0, atol=1e-12)
    for i, (v, z) in enumerate(states):
        assert np.array_equal(v.numpy(), g[f"state{i}"][0])
        assert np.array_equal(z.numpy(), g[f"state{i}"][1])
    # the fixture must exercise every layer (default init goes silent)
    assert (g["activity"][-1][1:8] > 0.01).all(), g["activity"][-1]


def test_encodings():
    g = load_golden("encode_small")
    xs, ys, ts, ps = (T(g[k]) for k in ("xs", "ys", "ts", "ps"))
    H, W = g["cnt"].shape[1:]
    cnt = oenc.events_to_channels(xs, ys, ps, (H, W)).numpy()
    assert np.array_equal(cnt, g["cnt"])
    assert np.array_equal(oenc.events_to_channels_np(g["xs"], g["ys"], g["ps"], (H, W)), g["cnt"])
    assert cnt.sum() == len(g["xs"])
    assert np.array_equal(oenc.events_to_image(xs, ys, ps.abs(), (H, W), accumulate=False).numpy(), g["mask"])
    assert np.array_equal(oenc.events_to_image(xs, ys, ps, (H, W)).numpy(), g["img_acc"])
    for key, nb, rnd in (("voxel5", 5, False), ("voxel2", 2, False), ("voxel5_round", 5, True)):
        np.testing.assert_allclose(oenc.events_to_voxel(xs, ys, ts, ps, nb, (H, W), rnd).numpy(), g[key],
                                   rtol=1e-6, atol=1e-6)


def test_encodings_empty():
    g = load_golden("encode_empty")
    e = torch.zeros(0)
    assert np.array_equal(oenc.events_to_channels(e, e, e, (8, 8)).numpy(), g["cnt"])
    assert np.array_equal(oenc.events_to_voxel(e, e, e, e, 5, (8, 8)).numpy(), g["voxel5"])


@pytest.mark.parametrize("name", ["iwe_rand", "iwe_zero_flow", "iwe_fractional"])
def test_iwe(name):
    g = load_golden(name)
    H, W, S = [int(v) for v in g["params"]]
    ev, pm = T(g["events"]), T(g["pol_mask"])
    flow = T(g["flow"]).clone().requires_grad_(True)
    ef = oiwe.gather_event_flow(flow, ev, (H, W))
    assert np.array_equal(ef.detach().numpy(), g["ev_flow"])
    for tag, tref, tsw in (("fw", 3, ev[:, :, 0:1]), ("bw", 0, 3 - ev[:, :, 0:1])):
        idx, w = oiwe.interpolation(ev, ef, tref, (H, W), S)
        assert np.array_equal(idx.detach().numpy(), g[f"{tag}_idx"])
        np.testing.assert_allclose(w.detach().numpy(), g[f"{tag}_w"], rtol=0, atol=0)
        img = oiwe.warp_images(ev, ef, pm, tref, (H, W), S, ts_weight=tsw)
        np.testing.assert_allclose(img.detach().numpy(), g[f"{tag}_img"], rtol=1e-6, atol=1e-6)
        flow.grad = None
        (img * T(g[f"{tag}_gimg"])).sum().backward(retain_graph=True)
        np.testing.assert_allclose(flow.grad.numpy(), g[f"{tag}_gflow"], rtol=1e-5, atol=1e-5)
    ev1 = ev.clone()
    ev1[:, :, 0] /= 3.0
    with torch.no_grad():
        for key, rnd in (("pol_iwe_round", True), ("pol_iwe_bilinear", False)):
            out = oiwe.pol_iwe(flow, ev1, (H, W), pm[:, :, 0:1], pm[:, :, 1:2], S, rnd)
            np.testing.assert_allclose(out.numpy(), g[key], rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("name", ["train_firenet_c8", "train_fireflownet_c8_mask"])
def test_train_window(name):
    g = load_golden(name)
    C, B, H, W, nT, n = [int(v) for v in g["dims"]]
    params = {k[len("param."):]: T(v).clone().requires_grad_(True) for k, v in g.items() if k.startswith("param.")
              if v.dtype == np.float32 and v.ndim > 0}
    lossf = oloss.EventWarpingOracle((H, W), 0.001, mask_output=bool(g["mask_output"]))
    states, flows = [None] * 7, []
    for t in range(nT):
        flow, states, _ = ofn.forward(params, T(g[f"cnt{t}"]), states)
        flow.retain_grad()
        flows.append(flow)
        lossf.associate(flow, T(g[f"events{t}"]).clone(), T(g[f"pol{t}"]), T(g[f"mask{t}"]))
    loss = lossf()
    loss.backward()
    np.testing.assert_allclose(float(loss.detach()), float(g["loss"]), rtol=1e-6)
    np.testing.assert_allclose(torch.stack(flows).detach().numpy(), g["flow"], rtol=0, atol=0)
    np.testing.assert_allclose(torch.stack([f.grad for f in flows]).numpy(), g["gflow"], rtol=1e-5, atol=1e-7)
    for k, p in params.items():
        if "grad." + k in g:
            np.testing.assert_allclose(p.grad.numpy(), g["grad." + k], rtol=1e-4, atol=1e-6, err_msg=k)


# ------------------------------------------------------------------------------------------------
# loader: raw event window -> batch tensors (fixtures written by the reference's own H5Loader.__getitem__)
# ------------------------------------------------------------------------------------------------
from oracle import loader as oload  # noqa: E402
from snnflow_testutil import LOADER_FIXTURES, LOADER_KEYS, loader_windows  # noqa: E402


@pytest.mark.parametrize("name", LOADER_FIXTURES)
def test_loader_oracle_matches_reference(name):
    g = load_golden(name)
    B, H, W, nb = int(g["B"]), int(g["H"]), int(g["W"]), int(g["num_bins"])
    target = tuple(int(v) for v in g["target"])
    hot = None
    if g["hot"].size:
        hot = oload.HotFilter(B, (H, W), max_px=int(g["hot"][0]), min_obvs=int(g["hot"][1]), max_rate=float(g["hot"][2]))
    removed = 0
    for it, wins in enumerate(loader_windows(g)):
        items = []
        for b, (xs, ys, ts, ps) in enumerate(wins):
            ts32 = (ts - 10.0).astype(np.float32)                     # get_events: ts -= t0, then astype(float32)
            items.append(oload.format_item(T(xs.astype(np.float32)), T(ys.astype(np.float32)), T(ts32),
                                           T(ps.astype(np.float32)), resolution=(H, W), num_bins=nb,
                                           round_ts=bool(g["round_enc"]), flips=tuple(bool(v) for v in g["flips"][b]),
                                           hot=hot, batch=b, target=target))
        batch = oload.collate(items)
        for k in LOADER_KEYS:
            ref = g[f"item{it}.{k}"]
            assert batch[k].shape == ref.shape, (k, batch[k].shape, ref.shape)
            assert np.array_equal(batch[k].numpy(), ref), f"{name} item {it} {k}"      # same ATen ops: bit-exact
        if hot is not None and target == (H, W):
            for b, (xs, ys, ts, ps) in enumerate(wins):
                seen = np.zeros((H, W), bool)
                fx = (W - 1 - xs) if g["flips"][b][0] else xs
                fy = (H - 1 - ys) if g["flips"][b][1] else ys
                seen[fy, fx] = True
                removed += int((seen & (g[f"item{it}.event_mask"][b, 0] == 0)).sum())
    if name == "loader_events_hot":
        assert removed > 0, "the hot-pixel filter never fired: the fixture would be vacuous"


def _hot_mask_by_cut(hot_events, idx, max_px, min_obvs, max_rate):
    """The selection rule the CUDA kernel implements (csrc/loader.cu, ld_hot_kernel), restated in numpy: candidates are
    pixels with hot_events / idx > max_rate; if there are more than max_px of them, a cut c* on the integer hit count is
    found (largest c with >= max_px candidates of count >= c), everything above the cut goes and the remaining places are
    filled with the lowest flat indices AT the cut."""
    he = hot_events.reshape(-1)
    mask = np.ones(he.shape, dtype=np.float32)
    if not idx > min_obvs or max_px <= 0:
        return mask.reshape(hot_events.shape)
    cand = (he / np.float32(idx)).astype(np.float32) > np.float32(max_rate)
    if cand.sum() <= max_px:
        mask[cand] = 0
        return mask.reshape(hot_events.shape)
    lo, hi = 0, idx
    while lo < hi:
        mid = (lo + hi + 1) >> 1
        if (cand & (he >= mid)).sum() >= max_px:
            lo = mid
        else:
            hi = mid - 1
    above = cand & (he > lo)
    at = np.flatnonzero(cand & (he == lo))
    mask[above] = 0
    mask[at[:max_px - int(above.sum())]] = 0
    return mask.reshape(hot_events.shape)


@pytest.mark.parametrize("seed", range(12))
def test_hot_pixel_cut_rule_equals_the_reference_argmax_loop(seed):
    """get_hot_event_mask (dataloader/encodings.py:88-103) removes one argmax at a time; the kernel uses a cut on the hit
    count instead.  Both must pick the same pixels, ties included (many equal rates, more candidates than max_px)."""
    rng = np.random.default_rng(seed)
    H, W = int(rng.integers(3, 12)), int(rng.integers(3, 12))
    idx = int(rng.integers(1, 40))
    he = rng.integers(0, idx + 1, size=(H, W)).astype(np.float32)          # integer hit counts <= idx
    if seed % 3 == 0:
        he[rng.random((H, W)) < 0.5] = idx                                   # heavy ties at rate 1.0
    max_px = int(rng.integers(1, 8))
    min_obvs = int(rng.integers(0, 6))
    max_rate = float(rng.choice([0.3, 0.5, 0.8, 0.95]))
    want = oload.hot_event_mask(T(he) / idx, idx, max_px=max_px, min_obvs=min_obvs, max_rate=max_rate).numpy()
    got = _hot_mask_by_cut(he, idx, max_px, min_obvs, max_rate)
    assert np.array_equal(got, want), (H, W, idx, max_px, min_obvs, max_rate)
    if ref_shim_available():
        from oracle import ref_shim
        ref = ref_shim.load().encodings.get_hot_event_mask(T(he) / idx, idx, max_px=max_px, min_obvs=min_obvs,
                                                           max_rate=max_rate).numpy()
        assert np.array_equal(want, ref)


def ref_shim_available():
    from oracle import ref_shim
    return ref_shim.reference_available()


@pytest.mark.parametrize("name", LOADER_FIXTURES)
def test_loader_numpy_restatement_matches_reference(name):
    """A second, numpy-only restatement of the exact loader outputs (counts, mask, event list, polarity mask) against the
    same reference-written fixtures - independent of oracle/loader.py and of torch's scatter / pooling ops."""
    from oracle.loader_np import format_item_np
    g = load_golden(name)
    B, H, W = int(g["B"]), int(g["H"]), int(g["W"])
    target = tuple(int(v) for v in g["target"])
    hot_cfg = dict(max_px=int(g["hot"][0]), min_obvs=int(g["hot"][1]), max_rate=float(g["hot"][2])) if g["hot"].size else None
    state = [dict(events=np.zeros((H, W), np.float32), idx=0) for _ in range(B)]
    for it, wins in enumerate(loader_windows(g)):
        for b, (xs, ys, ts, ps) in enumerate(wins):
            out = format_item_np(xs, ys, (ts - 10.0).astype(np.float32), ps, resolution=(H, W),
                                 flips=tuple(bool(v) for v in g["flips"][b]), hot_state=state[b], hot_cfg=hot_cfg, target=target)
            assert np.array_equal(out["event_cnt"], g[f"item{it}.event_cnt"][b]), (name, it, b, "cnt")
            assert np.array_equal(out["event_mask"], g[f"item{it}.event_mask"][b]), (name, it, b, "mask")
            assert np.array_equal(out["event_list"].T, g[f"item{it}.event_list"][b]), (name, it, b, "list")
            assert np.array_equal(out["event_list_pol_mask"].T, g[f"item{it}.event_list_pol_mask"][b]), (name, it, b, "pol")


def test_snntorch_leaky_restatement_basics():
    """oracle/snntorch_lif.py (PARITY UNPINNED restatement of snntorch 0.9.4's Leaky.forward): the properties its
    docstring promises - clamp of beta, immediate reset without double reset, ATan(alpha=2) surrogate."""
    import math
    import torch
    from oracle import snntorch_lif as osl
    beta, thr = torch.tensor([[[1.7]], [[0.5]]]), torch.tensor([[[1.0]], [[1.0]]])
    cur = torch.tensor([[[[0.4, 1.5]], [[0.4, 1.5]]]], requires_grad=True)       # [1,2,1,2]
    mem = torch.tensor([[[[0.8, 0.8]], [[0.8, 0.8]]]])
    spk, out = osl.leaky_step(cur, mem, beta, thr, "zero")
    assert spk.tolist() == [[[[1.0, 1.0]], [[0.0, 1.0]]]]                         # channel 0: beta clamped to 1 -> 1.2 > 1
    assert out[0, 0].tolist() == [[0.0, 0.0]] and abs(float(out[0, 1, 0, 0]) - 0.8) < 1e-6 and float(out[0, 1, 0, 1]) == 0.0
    spk.sum().backward()
    m = 0.5 * 0.8 + 0.4
    assert abs(float(cur.grad[0, 1, 0, 0]) - 1 / (1 + (math.pi * (m - 1.0)) ** 2)) < 1e-6
    # membrane entering above threshold, no new spike: the state function subtracts theta (reset = 1) and the "no double
    # reset" correction do_reset = spk - reset = -1 adds it back (the published code, literally)
    spk2, out2 = osl.leaky_step(cur.detach(), torch.full_like(mem, 1.5), beta, thr, "subtract")
    assert abs(float(out2[0, 1, 0, 0]) - (0.5 * 1.5 + 0.4)) < 1e-6 and float(spk2[0, 1, 0, 0]) == 0.0
