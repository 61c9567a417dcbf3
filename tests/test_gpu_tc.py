"""GPU parity: the tcgen05 (tensor-core) ConvLIF forward against the golden fixtures, the CPU oracle and the
exact-fp32 CUDA-core kernel.  Spike inputs are exact in fp16 and dyadic (2^-12) weights split exactly into the
two fp16 terms, so v, z and the input current must be IDENTICAL; random fp32 weights: rel 1e-4 (north star)."""
import numpy as np
import pytest
import torch

from snnflow_testutil import load_golden, spike_mismatch_outside_band

pytestmark = pytest.mark.gpu


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def pack(w_ff, w_rec, Cin, C):
    from snnflow_b200 import _lib
    L = _lib.lib()
    n = L.snnflow_convlif_packed_bytes(Cin, C, int(w_rec is not None))
    assert n > 0
    blob = torch.empty(n, dtype=torch.uint8, device="cuda")
    _lib.check(L.snnflow_convlif_pack(_lib.ptr(w_ff), _lib.ptr(w_rec), blob.data_ptr(), Cin, C, _lib.stream()), "pack")
    return blob


def run_tc(x, blob, rec, lam, theta, hard=True, residual=None, want_cur=True):
    """x [T,B,Cin,H,W] on the GPU -> v, z, out, cur stacked over T."""
    from snnflow_b200 import _lib
    L = _lib.lib()
    T, B, Cin, H, W = x.shape
    C = lam.numel()
    v = torch.empty((T, B, C, H, W), device="cuda")
    z, cur = torch.empty_like(v), torch.empty_like(v)
    out = torch.empty_like(v) if residual is not None else None
    flags = (_lib.HARD_RESET if hard else 0) | _lib.DETACH_RESET
    for t in range(T):
        _lib.check(L.snnflow_convlif_fwd_tc(
            _lib.ptr(x[t]), blob.data_ptr(), int(rec), _lib.ptr(v[t - 1]) if t else None, _lib.ptr(z[t - 1]) if t else None,
            _lib.ptr(lam), _lib.ptr(theta), _lib.ptr(residual[t]) if residual is not None else None, _lib.ptr(v[t]),
            _lib.ptr(z[t]), _lib.ptr(out[t]) if out is not None else None, _lib.ptr(cur[t]) if want_cur else None,
            B, Cin, C, H, W, flags, _lib.stream()), "fwd_tc")
    torch.cuda.synchronize()
    return v, z, (out if out is not None else z), cur


@pytest.mark.parametrize("name", ["layer_ff_c32", "layer_rec_c32", "layer_rec_c32_rand"])
def test_tc_forward_golden(name):
    from snnflow_b200 import _lib
    g = load_golden(name)
    rec = bool(g["meta"][0])
    x, w_ff = dev(g["x"]), dev(g["w_ff"])
    w_rec = dev(g["w_rec"]) if rec else None
    blob = pack(w_ff, w_rec, x.shape[2], w_ff.shape[0])
    _lib.lib().snnflow_tc_inexact_count(1)
    v, z, out, cur = run_tc(x, blob, rec, dev(g["lam"].reshape(-1)), dev(g["theta"].reshape(-1)))
    assert _lib.lib().snnflow_tc_inexact_count(0) == 0
    v, z = v.cpu().numpy(), z.cpu().numpy()
    if name.endswith("_rand"):
        np.testing.assert_allclose(v, g["v"], rtol=1e-4, atol=1e-5)
        bad, near = spike_mismatch_outside_band(z, g["z"], g["v"], g["theta"][None, None], band=1e-5)
        assert bad == 0, (bad, near)
    else:
        assert np.array_equal(z, g["z"]), f"spike mismatches: {(z != g['z']).sum()}"
        assert np.array_equal(v, g["v"]), f"max |dv| = {np.abs(v - g['v']).max()}"


@pytest.mark.parametrize("shape", [(2, 32, 32, 40, 200, True), (1, 32, 32, 128, 128, False), (2, 16, 32, 17, 256, False),
                                   (1, 64, 64, 9, 130, False), (1, 48, 32, 9, 130, True), (3, 32, 16, 33, 47, True)])
def test_tc_matches_simt_and_oracle(shape):
    """Odd widths (partial tiles, W > 128), every supported channel count, 3 steps with recurrence."""
    from oracle import lif as olif
    from snnflow_b200 import _lib
    B, Cin, C, H, W, rec = shape
    gen = torch.Generator().manual_seed(99 + W)
    w_ff = olif.dyadic((torch.rand(C, Cin, 3, 3, generator=gen) * 2 - 1) * (1 / Cin) ** 0.5)
    w_rec = olif.dyadic((torch.rand(C, C, 3, 3, generator=gen) * 2 - 1) * (1 / C) ** 0.5) if rec else None
    leak = torch.randn(C, 1, 1, generator=gen)
    thresh = torch.randn(C, 1, 1, generator=gen) * 0.1 + 0.3
    lam, theta = torch.sigmoid(leak).reshape(-1).cuda(), thresh.clamp_min(0.01).reshape(-1).cuda()
    T = 3
    x = (torch.rand(T, B, Cin, H, W, generator=gen) < 0.2).float()
    res = (torch.rand(T, B, C, H, W, generator=gen) < 0.2).float()
    w_ff_d, w_rec_d = w_ff.cuda(), (w_rec.cuda() if rec else None)
    blob = pack(w_ff_d, w_rec_d, Cin, C)
    v, z, out, cur = run_tc(x.cuda(), blob, rec, lam, theta, residual=res.cuda())
    vr = zr = None
    for t in range(T):
        o, vr, zr, cr = olif.lif_step(x[t], w_ff, leak, thresh, vr, zr, w_rec, res[t])
        assert torch.equal(cur[t].cpu(), cr), f"t={t}: conv differs, max {(cur[t].cpu() - cr).abs().max()}"
        assert torch.equal(z[t].cpu(), zr), f"t={t}: {(z[t].cpu() != zr).sum()} spike mismatches"
        assert torch.equal(v[t].cpu(), vr)
        assert torch.equal(out[t].cpu(), o)
    assert 0.02 < float(zr.mean()) < 0.9


def test_tc_random_weights_close_to_fp32_conv():
    """Raw fp32 weights: the hi+lo fp16 split keeps 22 mantissa bits -> conv within 1e-6 of the fp64 conv."""
    B, C, H, W = 2, 32, 24, 128
    gen = torch.Generator().manual_seed(5)
    w_ff = ((torch.rand(C, C, 3, 3, generator=gen) * 2 - 1) * (1 / C) ** 0.5)
    x = (torch.rand(1, B, C, H, W, generator=gen) < 0.3).float()
    lam, theta = torch.full((C,), 0.5).cuda(), torch.full((C,), 0.3).cuda()
    w_d = w_ff.cuda()
    blob = pack(w_d, None, C, C)
    _, _, _, cur = run_tc(x.cuda(), blob, False, lam, theta)
    ref = torch.nn.functional.conv2d(x[0].double(), w_ff.double(), padding=1)
    err = float((cur[0].cpu().double() - ref).abs().max())
    err32 = float((torch.nn.functional.conv2d(x[0], w_ff, padding=1).double() - ref).abs().max())
    # no worse than a few times the rounding error of an fp32 convolution of the same data (|I| up to ~5)
    assert err < max(4 * err32, 5e-6), (err, err32)


def test_tc_flags_inexact_input():
    from snnflow_b200 import _lib
    C, H, W = 32, 4, 128
    w_d = torch.zeros(C, C, 3, 3, device="cuda")
    blob = pack(w_d, None, C, C)
    x = torch.full((1, 1, C, H, W), 0.1, device="cuda")     # 0.1 is not representable in fp16
    _lib.lib().snnflow_tc_inexact_count(1)
    run_tc(x, blob, False, torch.full((C,), 0.5).cuda(), torch.full((C,), 0.3).cuda())
    assert _lib.lib().snnflow_tc_inexact_count(1) > 0


def test_module_uses_tensor_cores_and_matches_simt():
    import snnflow_b200 as snnflow
    from snnflow_b200 import _lib
    torch.manual_seed(0)
    net = snnflow.LIFFireNet(dict(num_bins=2, encoding="cnt", base_num_channels=32, kernel_size=3,
                                  neuron_kwargs=dict(leak=(0.0, 1.0), thresh=(0.3, 0.1)))).cuda()
    from oracle.lif import dyadic
    with torch.no_grad():
        for n, p in net.named_parameters():
            if n.endswith("weight"):
                p.copy_(dyadic(p))
    g = torch.Generator().manual_seed(3)
    cnt = torch.poisson(torch.full((4, 2, 2, 32, 160), 0.25), generator=g).cuda()
    net.stream_forward = False      # this test is about the per-bin cells (the streamed path uses the window engine)

    def run(tc):
        snnflow.ConvLIF.use_tensor_cores = tc
        snnflow.ConvLIFRecurrent.use_tensor_cores = tc
        net.reset_states()
        _lib.profile(True)
        with torch.no_grad():
            for t in range(4):
                net(None, cnt[t])
        prof = _lib.profile_summary()
        _lib.profile(False)
        return [s.clone() for s in net._states], prof

    try:
        s_tc, prof_tc = run(True)
        s_simt, prof_simt = run(False)
    finally:
        snnflow.ConvLIF.use_tensor_cores = True
        snnflow.ConvLIFRecurrent.use_tensor_cores = True
    assert prof_tc["convlif_fwd_tc"]["launches"] == 4 * 6 and prof_tc["convlif_fwd_simt"]["launches"] == 4
    assert "convlif_fwd_tc" not in prof_simt
    for a, b in zip(s_tc, s_simt):
        assert torch.equal(a, b)
    assert float(s_tc[-1][1].mean()) > 0.01


@pytest.mark.parametrize("name", ["layer_ff_c32", "layer_rec_c32", "layer_rec_c32_rand"])
def test_tc_backward_golden(name):
    """Tensor-core weight gradient (bf16 hi/lo split of g_I, spikes exact) vs reference autograd: rel 1e-4."""
    from snnflow_b200 import _lib
    from test_gpu_layers import capi_backward, capi_forward
    g = load_golden(name)
    _, _, _, cur = capi_forward(g)
    _lib.profile(True)
    r = capi_backward(g, dev(g["v"]), cur, flags_extra=_lib.INPUT_EXACT16)
    prof = _lib.profile_summary()
    _lib.profile(False)
    assert "wgrad_tc" in prof and "wgrad_simt" not in prof
    assert "dgrad_tc" in prof and "dgrad_simt" not in prof
    for k in ("g_x", "dw_ff", "dw_rec", "dleak", "dthresh"):
        if r[k] is None:
            continue
        ref = g[k].reshape(r[k].shape)
        scale = max(1.0, float(np.abs(ref).max()))
        np.testing.assert_allclose(r[k].cpu().numpy(), ref, rtol=1e-4, atol=1e-5 * scale, err_msg=k)


def test_tc_wgrad_matches_simt_large():
    """128x128, batch 4, recurrent: tensor-core partial sums vs the exact-fp32 CUDA-core kernel, rel 1e-4 of max."""
    from snnflow_b200 import _lib
    L = _lib.lib()
    B, C, H, W = 4, 32, 128, 128
    gen = torch.Generator().manual_seed(11)
    x = (torch.rand(B, C, H, W, generator=gen) < 0.2).float().cuda()
    z = (torch.rand(B, C, H, W, generator=gen) < 0.15).float().cuda()
    v_in = torch.randn(B, C, H, W, generator=gen).cuda()
    v_out = (torch.randn(B, C, H, W, generator=gen) * 0.3 + 0.2).cuda()
    cur = torch.randn(B, C, H, W, generator=gen).cuda()
    g_out = torch.randn(B, C, H, W, generator=gen).cuda()
    w_ff = ((torch.rand(C, C, 3, 3, generator=gen) - 0.5) * 0.3).cuda()
    w_rec = ((torch.rand(C, C, 3, 3, generator=gen) - 0.5) * 0.3).cuda()
    lam, theta = torch.full((C,), 0.6).cuda(), torch.full((C,), 0.3).cuda()
    ws = torch.empty(L.snnflow_convlif_bwd_workspace_bytes(B, C, C, H, W, 1), dtype=torch.uint8, device="cuda")
    outs = []
    for extra in (_lib.NO_TENSOR_CORES, _lib.INPUT_EXACT16):
        g_x, g_v, g_z = torch.empty_like(x), torch.empty_like(x), torch.empty_like(x)
        dwf, dwr = torch.zeros_like(w_ff), torch.zeros_like(w_rec)
        dl, dt = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
        _lib.check(L.snnflow_convlif_bwd(
            _lib.ptr(x), _lib.ptr(w_ff), _lib.ptr(w_rec), _lib.ptr(v_in), _lib.ptr(z), _lib.ptr(v_out), _lib.ptr(cur),
            _lib.ptr(lam), _lib.ptr(theta), _lib.ptr(g_out), None, None, _lib.ptr(g_x), _lib.ptr(g_v), _lib.ptr(g_z),
            _lib.ptr(dwf), _lib.ptr(dwr), _lib.ptr(dl), _lib.ptr(dt), ws.data_ptr(), ws.numel(), B, C, C, H, W,
            _lib.HARD_RESET | _lib.DETACH_RESET | extra, 0, 10.0, _lib.stream()), "bwd")
        torch.cuda.synchronize()
        outs.append((dwf.cpu().numpy(), dwr.cpu().numpy(), g_x.cpu().numpy(), g_z.cpu().numpy()))
    for a, b in zip(outs[0], outs[1]):
        np.testing.assert_allclose(b, a, rtol=1e-4, atol=1e-4 * np.abs(a).max())


@pytest.mark.parametrize("shape", [(2, 32, 32, 20, 200, True, True), (1, 16, 32, 9, 130, False, True), (2, 32, 32, 7, 64, True, False)])
def test_tc_dgrad_matches_simt(shape):
    """Odd widths / channel mixes / non-detached reset (g_z_in accumulates onto the reset term)."""
    from snnflow_b200 import _lib
    L = _lib.lib()
    B, Cin, C, H, W, rec, detach = shape
    gen = torch.Generator().manual_seed(21 + W)
    x = (torch.rand(B, Cin, H, W, generator=gen) < 0.2).float().cuda()
    z = (torch.rand(B, C, H, W, generator=gen) < 0.15).float().cuda()
    v_in = torch.randn(B, C, H, W, generator=gen).cuda()
    v_out = (torch.randn(B, C, H, W, generator=gen) * 0.3 + 0.2).cuda()
    cur = torch.randn(B, C, H, W, generator=gen).cuda()
    g_out = (torch.randn(B, C, H, W, generator=gen) * torch.rand(B, C, H, W, generator=gen) ** 8).cuda()   # wide dynamic range
    w_ff = ((torch.rand(C, Cin, 3, 3, generator=gen) - 0.5) * 0.3).cuda()
    w_rec = ((torch.rand(C, C, 3, 3, generator=gen) - 0.5) * 0.3).cuda() if rec else None
    lam, theta = torch.full((C,), 0.6).cuda(), torch.full((C,), 0.3).cuda()
    ws = torch.empty(L.snnflow_convlif_bwd_workspace_bytes(B, Cin, C, H, W, int(rec)), dtype=torch.uint8, device="cuda")
    base = _lib.HARD_RESET | (_lib.DETACH_RESET if detach else 0)
    outs = []
    for extra in (_lib.NO_TENSOR_CORES, 0):
        g_x, g_v, g_z = torch.empty_like(x), torch.empty_like(z), torch.empty_like(z)
        dwf = torch.zeros_like(w_ff)
        dwr = torch.zeros_like(w_rec) if rec else None
        dl, dt = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
        _lib.check(L.snnflow_convlif_bwd(
            _lib.ptr(x), _lib.ptr(w_ff), _lib.ptr(w_rec), _lib.ptr(v_in), _lib.ptr(z), _lib.ptr(v_out), _lib.ptr(cur),
            _lib.ptr(lam), _lib.ptr(theta), _lib.ptr(g_out), None, None, _lib.ptr(g_x), _lib.ptr(g_v), _lib.ptr(g_z),
            _lib.ptr(dwf), _lib.ptr(dwr), _lib.ptr(dl), _lib.ptr(dt), ws.data_ptr(), ws.numel(), B, Cin, C, H, W,
            base | extra, 0, 10.0, _lib.stream()), "bwd")
        torch.cuda.synchronize()
        outs.append((g_x.cpu().numpy(), g_z.cpu().numpy()))
    np.testing.assert_allclose(outs[1][0], outs[0][0], rtol=1e-4, atol=1e-5 * np.abs(outs[0][0]).max())
    if rec or not detach:
        np.testing.assert_allclose(outs[1][1], outs[0][1], rtol=1e-4, atol=1e-5 * np.abs(outs[0][1]).max())
