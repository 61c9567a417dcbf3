"""GPU parity: LIFFireNet / LIFFireFlowNet (7 fused layers + flow head) against the reference-generated
network fixtures, and one full training window (forward, contrast loss, BPTT) against reference autograd."""
import numpy as np
import pytest
import torch

from snnflow_testutil import load_golden

pytestmark = pytest.mark.gpu


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def build_net(g, kind, C):
    import snnflow_b200 as snnflow
    cls = getattr(snnflow, kind)
    net = cls(dict(num_bins=2, encoding="cnt", base_num_channels=C, kernel_size=3, mask_output=False)).cuda()
    sd = {k[len("param."):]: dev(v) for k, v in g.items() if k.startswith("param.")}
    net.load_state_dict(sd, strict=True)
    return net


@pytest.mark.parametrize("name,C", [("net_firenet_c8", 8), ("net_fireflownet_c8", 8), ("net_firenet_c32", 32)])
def test_network_forward(name, C):
    g = load_golden(name)
    net = build_net(g, str(g["kind"]), C)
    cnt = dev(g["cnt"])
    with torch.no_grad():
        for t in range(cnt.shape[0]):
            o = net(None, cnt[t], log=True)
            # flow goes through tanhf on the GPU vs the CPU's tanh: tolerance 1e-6 abs
            np.testing.assert_allclose(o["flow"][0].cpu().numpy(), g["flow"][t], rtol=1e-5, atol=1e-6)
            got = [o["activity"][k] for k in sorted(o["activity"])][:8]
            np.testing.assert_allclose(got, g["activity"][t][:8], rtol=0, atol=1e-7)
    for i, st in enumerate(net._states):
        st = st.cpu().numpy()
        assert np.array_equal(st[1], g[f"state{i}"][1]), f"layer {i}: spikes differ"
        # lam = sigmoid(leak) is evaluated by torch on the GPU here (CPU in the fixture): allow 1-ulp-of-lam drift
        np.testing.assert_allclose(st[0], g[f"state{i}"][0], rtol=2e-6, atol=2e-6)


@pytest.mark.parametrize("name", ["train_firenet_c8", "train_fireflownet_c8_mask"])
def test_training_window(name):
    import snnflow_b200 as snnflow
    g = load_golden(name)
    C, B, H, W, nT, n = [int(v) for v in g["dims"]]
    net = build_net(g, str(g["kind"]), C)
    mask_output = bool(g["mask_output"])
    cfg = {"loader": {"resolution": [H, W]}, "loss": {"flow_regul_weight": 0.001}, "model": {"mask_output": mask_output}}
    lossf = snnflow.EventWarping(cfg, torch.device("cuda"))
    flows = []
    for t in range(nT):
        out = net(None, dev(g[f"cnt{t}"]))
        out["flow"][0].retain_grad()
        flows.append(out["flow"][0])
        lossf.event_flow_association(out["flow"], dev(g[f"events{t}"]), dev(g[f"pol{t}"]), dev(g[f"mask{t}"]))
    loss = lossf()
    loss.backward()
    np.testing.assert_allclose(float(loss.detach()), float(g["loss"]), rtol=1e-5)
    np.testing.assert_allclose(torch.stack(flows).detach().cpu().numpy(), g["flow"], rtol=1e-5, atol=1e-6)
    # Gradient tolerance: the loss divides by (count + 1e-9) (loss/flow.py:214-215), so pixels that only receive
    # ~1e-7 of bilinear weight amplify fp32 summation-order noise of the splat by up to 1e9; a handful of elements
    # of d loss / d flow are therefore ill-conditioned in the reference itself.  Check rel 1e-4 on >= 99% of the
    # elements and 3e-3 norm-wise on everything (observed: 1 element in 1536 off by 1.6e-3 relative).
    # Parameter gradients inherit that perturbation through the few ill-conditioned pixels (every element moves by
    # the same ~1e-3 relative amount), so they are checked norm-wise; the per-kernel gradient tests in
    # test_gpu_layers.py hold the 1e-4 element-wise bar on a well-conditioned loss.
    def close(got, ref, name, elementwise):
        got, ref = np.asarray(got, dtype=np.float64), np.asarray(ref, dtype=np.float64)
        scale = max(np.abs(ref).max(), 1e-12)
        ok = np.abs(got - ref) <= 1e-4 * np.abs(ref) + 1e-5 * scale
        if elementwise:
            assert ok.mean() >= 0.99, (name, ok.mean())
        err, nrm = np.linalg.norm(got - ref), np.linalg.norm(ref)
        assert err <= 3e-3 * nrm + 1e-7, (name, err, nrm)

    close(torch.stack([f.grad for f in flows]).cpu().numpy(), g["gflow"], "gflow", True)
    for k, p in net.named_parameters():
        close(p.grad.cpu().numpy(), g["grad." + k], k, False)


def test_state_plumbing():
    import snnflow_b200 as snnflow
    net = snnflow.LIFFireNet(dict(num_bins=2, encoding="cnt", base_num_channels=8, kernel_size=3)).cuda()
    assert net.states == [None] * 7
    x = torch.ones(1, 2, 16, 16, device="cuda")
    net(None, x)
    st = net.states
    assert all(s.shape == (2, 1, 8, 16, 16) for s in st)
    assert st[0].data_ptr() != net._states[0].data_ptr()      # clones, like model_util.copy_states
    net.detach_states()
    assert not any(s.requires_grad for s in net._states)
    net.reset_states()
    assert net._states == [None] * 7


def test_graphed_per_bin_forward_equals_eager():
    """model.graph_forward(): the per-bin forward() replayed as a CUDA graph under no_grad - same flows, same states, state
    plumbing (reset_states, externally assigned states, .states clones) intact."""
    import snnflow_b200 as snnflow
    torch.manual_seed(0)
    net = snnflow.LIFFireNet(dict(num_bins=2, encoding="cnt", base_num_channels=32, kernel_size=3,
                                  neuron_kwargs=dict(leak=(0.0, 1.0), thresh=(0.3, 0.1)))).cuda()
    with torch.no_grad():
        net.pred.conv2d.weight.mul_(20)
    g = torch.Generator().manual_seed(3)
    cnt = torch.poisson(torch.full((6, 1, 2, 32, 48), 0.3), generator=g).cuda()
    net.stream_forward = False      # eager reference = the per-bin cells (what the graph captures)
    with torch.no_grad():
        ref = [net(None, cnt[t])["flow"][0].clone() for t in range(6)]
        s_ref = net.states
        net.reset_states()
        net.graph_forward()
        got = []
        for t in range(3):
            got.append(net(None, cnt[t])["flow"][0].clone())
        kept = net.states                                  # deep clones survive the next replays
        for t in range(3, 6):
            got.append(net(None, cnt[t])["flow"][0].clone())
        for a, b in zip(ref, got):
            assert torch.equal(a, b)
        for a, b in zip(s_ref, net._states):
            assert torch.equal(a, b)
        # rewind to the state after bin 2 (assigned from outside) and replay bins 3..5
        net._states = kept
        again = [net(None, cnt[t])["flow"][0].clone() for t in range(3, 6)]
        for a, b in zip(ref[3:], again):
            assert torch.equal(a, b)
        net.reset_states()
        assert torch.equal(net(None, cnt[0])["flow"][0], ref[0])
    # with autograd on, the call goes through the eager cells
    net.reset_states()
    out = net(None, cnt[0])["flow"][0]
    assert out.requires_grad


@pytest.mark.parametrize("graph", [False, True])
@pytest.mark.parametrize("kind,C,B,H,W", [("LIFFireNet", 32, 2, 24, 136), ("LIFFireFlowNet", 32, 3, 16, 128), ("LIFFireNet", 16, 1, 130, 64),
                                          ("LIFFireNet", 32, 1, 8, 256)])
def test_streamed_per_bin_forward_equals_cells(kind, C, B, H, W, graph):
    """Under no_grad the per-bin forward() runs on the window engine with the state kept in the engine's layout between calls
    (SNNFLOW_STATE_INTERNAL) and `_states` materialised lazily: flows, spikes and membranes must equal the per-bin cells bit for
    bit (2^-12-grid weights), through reads of .states, reset_states(), externally assigned states, a forward_window() call in
    between and a switch back to autograd mode.  graph=True: the same with model.graph_forward() - every streamed bin is a
    CUDA-graph replay (one graph per phase of the recurrent layers' ping-pong slots).  W = 256: 128-pixel column tiles."""
    import snnflow_b200 as snnflow
    from oracle.lif import dyadic as snap
    torch.manual_seed(1)
    net = getattr(snnflow, kind)(dict(num_bins=2, encoding="cnt", base_num_channels=C, kernel_size=3,
                                      neuron_kwargs=dict(leak=(0.0, 1.0), thresh=(0.3, 0.1))))
    with torch.no_grad():
        net.pred.conv2d.weight.mul_(20)
        for n, p in net.named_parameters():
            if n.endswith("weight"):
                p.copy_(snap(p))
    net = net.cuda()
    g = torch.Generator().manual_seed(8)
    cnt = torch.poisson(torch.full((9, B, 2, H, W), 0.25), generator=g).cuda()
    with torch.no_grad():
        net.stream_forward = False
        ref = [net(None, cnt[t])["flow"][0].clone() for t in range(9)]
        s_ref_3 = None
        net.reset_states()
        for t in range(3):
            net(None, cnt[t])
        s_ref_3 = net.states
        for t in range(3, 9):
            net(None, cnt[t])
        s_ref = net.states
        # streamed
        net.stream_forward = True
        if graph:
            net.graph_forward()
        net.reset_states()
        got = [net(None, cnt[t])["flow"][0].clone() for t in range(3)]
        assert net._window_runner.stream_live, "the streamed path was not taken"
        if graph:
            assert net._window_runner._stream["graphs"] is not None, "the streamed bins were not replayed as graphs"
        mid = net.states                                   # lazily materialised, deep-cloned
        for a, b in zip(s_ref_3, mid):
            assert torch.equal(a, b)
        got += [net(None, cnt[t])["flow"][0].clone() for t in range(3, 5)]      # continues from the arena (list only read)
        # a T = 2 window in between (other arena), then streaming again
        got += list(net.forward_window(cnt[5:7]).clone())
        got += [net(None, cnt[t])["flow"][0].clone() for t in range(7, 9)]
        for t, (a, b) in enumerate(zip(ref, got)):
            assert torch.equal(a, b), f"bin {t}: flows differ ({float((a - b).abs().max())})"
        for i, (a, b) in enumerate(zip(s_ref, net.states)):
            assert torch.equal(a[1], b[1]), f"layer {i}: spikes differ"
            assert torch.equal(a[0], b[0]), f"layer {i}: membranes differ"
        assert float(s_ref[-1][1].mean()) > 0.01
        # externally assigned state: rewind to bin 3
        net._states = [s.clone() for s in s_ref_3]
        again = [net(None, cnt[t])["flow"][0].clone() for t in range(3, 6)]
        for a, b in zip(ref[3:6], again):
            assert torch.equal(a, b)
    # autograd on: the cells, continuing from the streamed state
    out = net(None, cnt[6])["flow"][0]
    assert out.requires_grad and torch.equal(out.detach(), ref[6])
