"""GPU parity: event encodings (bit-exact counts) and image-of-warped-events kernels."""
import numpy as np
import pytest
import torch

from snnflow_testutil import load_golden

pytestmark = pytest.mark.gpu


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_encode_golden():
    from snnflow_b200 import encodings as enc
    g = load_golden("encode_small")
    xs, ys, ts, ps = (dev(g[k]) for k in ("xs", "ys", "ts", "ps"))
    H, W = g["cnt"].shape[1:]
    assert np.array_equal(enc.events_to_channels(xs, ys, ps, (H, W)).cpu().numpy(), g["cnt"])
    assert np.array_equal(enc.events_to_image(xs, ys, ps.abs(), (H, W), accumulate=False).cpu().numpy(), g["mask"])
    assert np.array_equal(enc.events_to_image(xs, ys, ps, (H, W)).cpu().numpy(), g["img_acc"])
    for key, nb, rnd in (("voxel5", 5, False), ("voxel2", 2, False), ("voxel5_round", 5, True)):
        a = enc.events_to_voxel(xs, ys, ts, ps, nb, (H, W), rnd)
        b = enc.events_to_voxel(xs, ys, ts, ps, nb, (H, W), rnd)
        assert torch.equal(a, b), "voxel encoding must be run-to-run deterministic"
        np.testing.assert_allclose(a.cpu().numpy(), g[key], rtol=1e-6, atol=1e-6)
    # "last event wins" for accumulate=False with distinct values
    vals = torch.arange(1, len(g["xs"]) + 1, dtype=torch.float32)
    from oracle import encodings as oenc
    want = oenc.events_to_image(torch.from_numpy(g["xs"]), torch.from_numpy(g["ys"]), vals, (H, W), accumulate=False)
    got = enc.events_to_image(xs, ys, vals.cuda(), (H, W), accumulate=False)
    assert torch.equal(got.cpu(), want)


def test_encode_empty_and_batched():
    from snnflow_b200 import encodings as enc
    e = torch.zeros(0, device="cuda")
    assert float(enc.events_to_channels(e, e, e, (8, 8)).abs().sum()) == 0
    assert float(enc.events_to_voxel(e, e, e, e, 5, (8, 8)).abs().sum()) == 0
    from oracle import encodings as oenc
    gen = torch.Generator().manual_seed(7)
    B, N, H, W = 3, 1001, 32, 48    # N not a multiple of 4: exercises the scalar tail / unaligned rows
    evs = [oenc.synth_events(N, (H, W), gen) for _ in range(B)]
    xs, ys, ps = (torch.stack([e[i] for e in evs]) for i in (0, 1, 3))
    got = enc.events_to_channels_batched(xs.cuda(), ys.cuda(), ps.cuda(), (H, W)).cpu()
    want = torch.stack([oenc.events_to_channels(*[e[i] for i in (0, 1, 3)], (H, W)) for e in evs])
    assert torch.equal(got, want)


def test_encode_full_size_properties():
    """BASELINE config 5 size: 10M events into 256x256.  Size-independent checks: the counts sum to N, the
    batched and the flat call agree, and the result equals a torch.bincount of the same data."""
    from snnflow_b200 import encodings as enc
    N, H, W = 10_000_000, 256, 256
    g = torch.Generator(device="cuda").manual_seed(3)
    xs = torch.randint(0, W, (N,), generator=g, device="cuda").float()
    ys = torch.randint(0, H, (N,), generator=g, device="cuda").float()
    ps = torch.randint(0, 2, (N,), generator=g, device="cuda").float() * 2 - 1
    cnt = enc.events_to_channels(xs, ys, ps, (H, W))
    assert float(cnt.sum()) == N
    lin = (ys.long() * W + xs.long())
    want = torch.stack([torch.bincount(lin[ps > 0], minlength=H * W), torch.bincount(lin[ps < 0], minlength=H * W)])
    assert torch.equal(cnt.reshape(2, -1).long(), want)


@pytest.mark.parametrize("name", ["iwe_rand", "iwe_zero_flow", "iwe_fractional"])
def test_iwe_golden(name):
    from snnflow_b200 import iwe
    g = load_golden(name)
    H, W, S = [int(v) for v in g["params"]]
    ev, pm = dev(g["events"]), dev(g["pol_mask"])
    flow = dev(g["flow"]).requires_grad_(True)
    ef = iwe.gather_event_flow(flow, ev, (H, W))
    assert np.array_equal(ef.detach().cpu().numpy(), g["ev_flow"])
    for tag, tref, mode in (("fw", 3, 1), ("bw", 0, 2)):
        img = iwe.warp_images(ev, ef, pm, tref, (H, W), S, ts_mode=mode, ts_ref=3.0)
        img2 = iwe.warp_images(ev, ef, pm, tref, (H, W), S, ts_mode=mode, ts_ref=3.0)
        assert torch.equal(img, img2), "splat must be run-to-run deterministic"
        np.testing.assert_allclose(img.detach().cpu().numpy(), g[f"{tag}_img"], rtol=1e-5, atol=1e-5)
        flow.grad = None
        (img * dev(g[f"{tag}_gimg"])).sum().backward(retain_graph=True)
        ref = g[f"{tag}_gflow"]
        np.testing.assert_allclose(flow.grad.cpu().numpy(), ref, rtol=1e-4, atol=1e-5 * max(1.0, np.abs(ref).max()))
    ev1 = ev.clone()
    ev1[:, :, 0] /= 3.0
    with torch.no_grad():
        for key, rnd in (("pol_iwe_round", True), ("pol_iwe_bilinear", False)):
            out = iwe.compute_pol_iwe(flow, ev1, (H, W), pm[:, :, 0:1].contiguous(), pm[:, :, 1:2].contiguous(), S, rnd)
            np.testing.assert_allclose(out.cpu().numpy(), g[key], rtol=1e-5, atol=1e-5)


def test_iwe_full_size_properties():
    """10M events, 256x256 (BASELINE config 5): mass conservation - every in-range bilinear splat deposits
    total weight 1, so sum(img) == number of events whose four corners are all inside; zero flow keeps every
    event on its own pixel and reproduces the count encoding exactly."""
    from snnflow_b200 import encodings as enc
    from snnflow_b200 import iwe
    N, H, W = 10_000_000, 256, 256
    g = torch.Generator(device="cuda").manual_seed(5)
    xs = torch.randint(0, W, (N,), generator=g, device="cuda").float()
    ys = torch.randint(0, H, (N,), generator=g, device="cuda").float()
    ts = torch.rand(N, generator=g, device="cuda")
    ps = torch.randint(0, 2, (N,), generator=g, device="cuda").float() * 2 - 1
    ev = torch.stack([ts, ys, xs, ps], dim=1)[None].contiguous()
    pos, neg = (ps > 0).float()[None, :, None], (ps < 0).float()[None, :, None]
    zero = torch.zeros(1, 2, H, W, device="cuda")
    out = iwe.compute_pol_iwe(zero, ev, (H, W), pos, neg, 128, round_idx=False)
    assert torch.equal(out[0], enc.events_to_channels(xs, ys, ps, (H, W)))
    flow = torch.tanh(0.5 * torch.randn(1, 2, H, W, generator=g, device="cuda")) * 0.02
    out = iwe.compute_pol_iwe(flow, ev, (H, W), pos, neg, 128, round_idx=False)
    total = float(out.double().sum())
    assert 0.9 * N < total <= N + 1e-3 * N
