"""GPU parity: ConvLIF / ConvLIFRecurrent kernels vs the reference-generated golden fixtures and the
CPU oracle.  Bit-exact tier: dyadic weights + spike/count inputs -> v, z, out must be IDENTICAL.
Tolerance tier (random fp32 weights, gradients): rel 1e-4 as stated by the north star."""
import numpy as np
import pytest
import torch

from snnflow_testutil import load_golden, spike_mismatch_outside_band

pytestmark = pytest.mark.gpu

LAYER_FIXTURES = ["layer_ff_hard_arctan", "layer_ff_soft_super_res", "layer_rec_hard_arctan",
                  "layer_rec_soft_triangle", "layer_rec_hard_mgspike", "layer_rec_nodetach", "layer_head_counts", "layer_ff_c32",
                  "layer_rec_c32", "layer_rec_c32_rand"]


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def capi_forward(g, flags_extra=0):
    """Run the fixture's T steps straight through the C ABI with the reference's lam/theta injected."""
    import snnflow_b200 as snnflow
    from snnflow_b200 import _lib
    L = _lib.lib()
    rec, hard, detach, has_res = [bool(v) for v in g["meta"]]
    x, w_ff = dev(g["x"]), dev(g["w_ff"])
    w_rec = dev(g["w_rec"]) if rec else None
    lam, theta = dev(g["lam"].reshape(-1)), dev(g["theta"].reshape(-1))
    res = dev(g["residual"]) if has_res else None
    T, B, Cin, H, W = x.shape
    C = w_ff.shape[0]
    v = torch.empty((T, B, C, H, W), device="cuda")
    z = torch.empty_like(v)
    out = torch.empty_like(v) if has_res else None
    cur = torch.empty_like(v)
    flags = (_lib.HARD_RESET if hard else 0) | (_lib.DETACH_RESET if detach else 0) | flags_extra
    for t in range(T):
        _lib.check(L.snnflow_convlif_fwd(
            _lib.ptr(x[t]), _lib.ptr(w_ff), _lib.ptr(w_rec), _lib.ptr(v[t - 1]) if t else None,
            _lib.ptr(z[t - 1]) if t else None, _lib.ptr(lam), _lib.ptr(theta), _lib.ptr(res[t]) if has_res else None,
            _lib.ptr(v[t]), _lib.ptr(z[t]), _lib.ptr(out[t]) if has_res else None, _lib.ptr(cur[t]), B, Cin, C, H, W,
            flags, _lib.stream()), "fwd")
    torch.cuda.synchronize()
    return v, z, (out if has_res else z), cur


@pytest.mark.parametrize("name", LAYER_FIXTURES)
def test_forward_capi(name):
    g = load_golden(name)
    v, z, out, _ = capi_forward(g)
    v, z, out = v.cpu().numpy(), z.cpu().numpy(), out.cpu().numpy()
    if name.endswith("_rand"):
        # tolerance tier: fp32 weights, summation order differs from oneDNN's
        np.testing.assert_allclose(v, g["v"], rtol=1e-4, atol=1e-5)
        bad, near = spike_mismatch_outside_band(z, g["z"], g["v"], g["theta"][None, None], band=1e-5)
        assert bad == 0, (bad, near)
    else:
        assert np.array_equal(z, g["z"]), f"spike mismatches: {(z != g['z']).sum()}"
        assert np.array_equal(v, g["v"]), f"max |dv| = {np.abs(v - g['v']).max()}"
        assert np.array_equal(out, g["out"])


def capi_backward(g, v, cur, flags_extra=0):
    from snnflow_b200 import _lib
    L = _lib.lib()
    rec, hard, detach, has_res = [bool(v_) for v_ in g["meta"]]
    sg = _lib.SURROGATE_ID[str(g["activation"])]
    x, w_ff = dev(g["x"]), dev(g["w_ff"])
    w_rec = dev(g["w_rec"]) if rec else None
    lam, theta = dev(g["lam"].reshape(-1)), dev(g["theta"].reshape(-1))
    z = dev(g["z"])
    gout = dev(g["gout"])
    T, B, Cin, H, W = x.shape
    C = w_ff.shape[0]
    flags = (_lib.HARD_RESET if hard else 0) | (_lib.DETACH_RESET if detach else 0) | flags_extra
    nbytes = L.snnflow_convlif_bwd_workspace_bytes(B, Cin, C, H, W, int(rec))
    ws = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    g_x = torch.empty_like(x)
    dw_ff = torch.zeros_like(w_ff)
    dw_rec = torch.zeros_like(w_rec) if rec else None
    dlam = torch.zeros(C, device="cuda")
    dtheta = torch.zeros(C, device="cuda")
    g_v = dev(g["gv_last"])
    g_z = torch.zeros_like(g_v)
    for t in reversed(range(T)):
        g_v_in, g_z_in = torch.empty_like(g_v), torch.empty_like(g_v)
        _lib.check(L.snnflow_convlif_bwd(
            _lib.ptr(x[t]), _lib.ptr(w_ff), _lib.ptr(w_rec), _lib.ptr(v[t - 1]) if t else None,
            _lib.ptr(z[t - 1]) if t else None, _lib.ptr(v[t]), _lib.ptr(cur[t]), _lib.ptr(lam), _lib.ptr(theta),
            _lib.ptr(gout[t]), _lib.ptr(g_v), _lib.ptr(g_z), _lib.ptr(g_x[t]), _lib.ptr(g_v_in), _lib.ptr(g_z_in),
            _lib.ptr(dw_ff), _lib.ptr(dw_rec), _lib.ptr(dlam), _lib.ptr(dtheta), ws.data_ptr(), ws.numel(), B, Cin, C,
            H, W, flags, sg, 10.0, _lib.stream()), "bwd")
        g_v, g_z = g_v_in, g_z_in
    torch.cuda.synchronize()
    lamv = lam
    dleak = dlam * lamv * (1 - lamv)
    dthresh = dtheta * (dev(g["thresh"].reshape(-1)) >= 0.01).float()
    return dict(g_x=g_x, dw_ff=dw_ff, dw_rec=dw_rec, dleak=dleak, dthresh=dthresh)


@pytest.mark.parametrize("name", LAYER_FIXTURES)
def test_backward_capi(name):
    g = load_golden(name)
    v, z, out, cur = capi_forward(g)
    if not name.endswith("_rand"):
        assert np.array_equal(z.cpu().numpy(), g["z"])
    from snnflow_b200 import _lib
    # teacher-forced membranes from the reference; CUDA-core (exact fp32) data/weight gradients
    r = capi_backward(g, dev(g["v"]), cur, flags_extra=_lib.NO_TENSOR_CORES)
    for k in ("g_x", "dw_ff", "dw_rec", "dleak", "dthresh"):
        if r[k] is None:
            continue
        ref = g[k].reshape(r[k].shape)
        scale = max(1.0, float(np.abs(ref).max()))
        np.testing.assert_allclose(r[k].cpu().numpy(), ref, rtol=1e-4, atol=1e-5 * scale, err_msg=k)


@pytest.mark.parametrize("name", ["layer_ff_hard_arctan", "layer_rec_hard_arctan", "layer_ff_soft_super_res",
                                  "layer_rec_nodetach"])
def test_module_autograd(name):
    """The nn.Module face + torch.autograd.Function: same loss as the fixture generator, grads vs reference."""
    import snnflow_b200 as snnflow
    g = load_golden(name)
    rec, hard, detach, has_res = [bool(v) for v in g["meta"]]
    act = str(g["activation"])
    x = dev(g["x"]).requires_grad_(True)
    T, B, Cin, H, W = x.shape
    C = g["w_ff"].shape[0]
    cls = snnflow.ConvLIFRecurrent if rec else snnflow.ConvLIF
    layer = cls(Cin, C, 3, activation=act, hard_reset=hard, detach=detach).cuda()
    sd = {"ff.weight": dev(g["w_ff"]), "leak": dev(g["leak"]), "thresh": dev(g["thresh"]),
          "act_width": torch.tensor(10.0)}
    if rec:
        sd["rec.weight"] = dev(g["w_rec"])
    layer.load_state_dict(sd)
    state, loss = None, 0
    gout, res = dev(g["gout"]), dev(g["residual"]) if has_res else None
    zs = []
    for t in range(T):
        out, state = layer(x[t], state, residual=res[t] if has_res else 0)
        assert state.shape == (2, B, C, H, W)
        zs.append(state[1])
        loss = loss + (out * gout[t]).sum()
    loss = loss + (state[0] * dev(g["gv_last"])).sum()
    loss.backward()
    assert np.array_equal(torch.stack(zs).detach().cpu().numpy(), g["z"])
    pairs = [(x.grad, "g_x"), (layer.ff.weight.grad, "dw_ff"), (layer.leak.grad, "dleak"), (layer.thresh.grad, "dthresh")]
    if rec:
        pairs.append((layer.rec.weight.grad, "dw_rec"))
    for got, k in pairs:
        ref = g[k]
        scale = max(1.0, float(np.abs(ref).max()))
        np.testing.assert_allclose(got.cpu().numpy(), ref, rtol=1e-4, atol=1e-5 * scale, err_msg=k)


@pytest.mark.parametrize("shape", [(2, 32, 32, 64, 64, True), (3, 2, 32, 33, 47, False), (1, 8, 8, 128, 128, True),
                                   (2, 32, 32, 128, 128, False)])
def test_forward_vs_oracle_seeded(shape):
    """Sizes beyond the fixtures: seeded inputs, dyadic weights -> bit-exact against the CPU oracle."""
    from oracle import lif as olif
    from snnflow_b200 import _lib
    B, Cin, C, H, W, rec = shape
    gen = torch.Generator().manual_seed(1234 + H)
    w_ff = olif.dyadic((torch.rand(C, Cin, 3, 3, generator=gen) * 2 - 1) * (1 / Cin) ** 0.5)
    w_rec = olif.dyadic((torch.rand(C, C, 3, 3, generator=gen) * 2 - 1) * (1 / C) ** 0.5) if rec else None
    leak = torch.randn(C, 1, 1, generator=gen)
    thresh = torch.randn(C, 1, 1, generator=gen) * 0.1 + 0.3
    lam, theta = torch.sigmoid(leak), thresh.clamp_min(0.01)
    L = _lib.lib()
    v = z = None
    vg = zg = None
    # device copies are kept alive in named variables: a temporary's memory could be recycled by the caching
    # allocator before the asynchronous kernel has read it
    w_ff_d, w_rec_d = w_ff.cuda(), (w_rec.cuda() if rec else None)
    lam_d, theta_d = lam.reshape(-1).cuda(), theta.reshape(-1).cuda()
    for t in range(3):
        x = (torch.rand(B, Cin, H, W, generator=gen) < 0.2).float()
        _, v, z, _ = olif.lif_step(x, w_ff, leak, thresh, v, z, w_rec)
        vo = torch.empty((B, C, H, W), device="cuda")
        zo = torch.empty_like(vo)
        xd = x.cuda()
        _lib.check(L.snnflow_convlif_fwd(
            _lib.ptr(xd), _lib.ptr(w_ff_d), _lib.ptr(w_rec_d), _lib.ptr(vg), _lib.ptr(zg),
            _lib.ptr(lam_d), _lib.ptr(theta_d), None, _lib.ptr(vo), _lib.ptr(zo),
            None, None, B, Cin, C, H, W, _lib.HARD_RESET | _lib.DETACH_RESET, _lib.stream()), "fwd")
        vg, zg = vo, zo
        assert torch.equal(zo.cpu(), z), f"t={t}: {(zo.cpu() != z).sum()} spike mismatches"
        assert torch.equal(vo.cpu(), v), f"t={t}: max |dv| {(vo.cpu() - v).abs().max()}"
    assert 0.02 < float(z.mean()) < 0.9


def test_error_paths():
    import snnflow_b200 as snnflow
    from snnflow_b200 import _lib
    L = _lib.lib()
    assert L.snnflow_convlif_fwd(None, None, None, None, None, None, None, None, None, None, None, None, 1, 1, 1, 1, 1,
                                 0, None) == -1
    assert b"null" in L.snnflow_last_error()
    with pytest.raises(NotImplementedError):
        snnflow.ConvLIF(2, 8, 3, stride=2)
    with pytest.raises(NotImplementedError):
        snnflow.ConvLIF(2, 8, 5)
    with pytest.raises(NotImplementedError):
        snnflow.ConvLIF(2, 8, 3, activation="no_such_spike")
    with pytest.raises(_lib.SnnflowError):
        snnflow.ConvLIF(2, 8, 3)(torch.zeros(1, 2, 8, 8), None)
