"""GPU: the whole-window engine (one C call per direction) equals T per-bin module calls - identical states and
flows, gradients equal up to fp32 accumulation order - and reproduces the reference training-window fixture."""
import numpy as np
import pytest
import torch

from snnflow_testutil import load_golden

pytestmark = pytest.mark.gpu


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def make_net(kind, C, seed=0):
    import snnflow_b200 as snnflow
    torch.manual_seed(seed)
    net = getattr(snnflow, kind)(dict(num_bins=2, encoding="cnt", base_num_channels=C, kernel_size=3,
                                      neuron_kwargs=dict(leak=(0.0, 1.0), thresh=(0.3, 0.1)))).cuda()
    with torch.no_grad():
        net.pred.conv2d.weight.mul_(20)
    import importlib
    WindowRunner = importlib.import_module("snn_event-based_optical_flow_b200.engine").WindowRunner
    runner = WindowRunner(net)
    runner.engine = "per_step"   # this file pins the per-step window engine; test_gpu_window.py covers the layer-major one
    object.__setattr__(net, "_window_runner", runner)
    return net


@pytest.mark.parametrize("kind,C,H,W", [("LIFFireNet", 32, 24, 136), ("LIFFireFlowNet", 32, 16, 128), ("LIFFireNet", 8, 20, 36)])
def test_window_equals_per_bin(kind, C, H, W):
    net = make_net(kind, C)
    g = torch.Generator().manual_seed(5)
    T, B = 4, 2
    cnt = torch.poisson(torch.full((2, T, B, 2, H, W), 0.25), generator=g).cuda()
    gout = torch.randn(2, T, B, 2, H, W, generator=g).cuda()

    def run(window):
        net.reset_states()
        net.zero_grad(set_to_none=True)
        flows = []
        for k in range(2):   # two consecutive windows: exercises the state hand-over between arenas
            if window:
                f = net.forward_window(cnt[k])   # per-step engine: same kernels as the modules => identical results
            else:
                f = torch.stack([net(None, cnt[k, t])["flow"][0] for t in range(T)])
            (f * gout[k]).sum().backward()
            net.detach_states()
            flows.append(f.detach().clone())
        return flows, [s.clone() for s in net._states], {n: p.grad.clone() for n, p in net.named_parameters()}

    f_a, s_a, g_a = run(False)
    f_b, s_b, g_b = run(True)
    for a, b in zip(f_a, f_b):
        assert torch.equal(a, b)
    for a, b in zip(s_a, s_b):
        assert torch.equal(a, b)
    assert float(s_a[-1][1].mean()) > 0.01
    for n in g_a:
        scale = float(g_a[n].abs().max()) + 1e-12
        assert torch.allclose(g_a[n], g_b[n], rtol=1e-4, atol=1e-5 * scale), (n, float((g_a[n] - g_b[n]).abs().max()), scale)


def test_window_eval_mode_matches():
    net = make_net("LIFFireFlowNet", 32)
    g = torch.Generator().manual_seed(6)
    cnt = torch.poisson(torch.full((5, 2, 2, 16, 130), 0.25), generator=g).cuda()
    with torch.no_grad():
        net.reset_states()
        ref = torch.stack([net(None, cnt[t])["flow"][0] for t in range(5)])
        s_ref = [s.clone() for s in net._states]
        net.reset_states()
        got = torch.cat([net.forward_window(cnt[:3]), net.forward_window(cnt[3:])])
    assert torch.equal(ref, got)
    for a, b in zip(s_ref, net._states):
        assert torch.equal(a, b)


def test_training_window_fixture_through_engine():
    import snnflow_b200 as snnflow
    g = load_golden("train_firenet_c8")
    C, B, H, W, nT, n = [int(v) for v in g["dims"]]
    net = snnflow.LIFFireNet(dict(num_bins=2, encoding="cnt", base_num_channels=C, kernel_size=3)).cuda()
    net.load_state_dict({k[len("param."):]: dev(v) for k, v in g.items() if k.startswith("param.")})
    cfg = {"loader": {"resolution": [H, W]}, "loss": {"flow_regul_weight": 0.001}, "model": {"mask_output": False}}
    lossf = snnflow.EventWarping(cfg, torch.device("cuda"))
    flows = net.forward_window(torch.stack([dev(g[f"cnt{t}"]) for t in range(nT)]))
    for t in range(nT):
        lossf.event_flow_association([flows[t]], dev(g[f"events{t}"]), dev(g[f"pol{t}"]), dev(g[f"mask{t}"]))
    loss = lossf()
    loss.backward()
    np.testing.assert_allclose(float(loss.detach()), float(g["loss"]), rtol=1e-5)
    np.testing.assert_allclose(flows.detach().cpu().numpy(), g["flow"], rtol=1e-5, atol=1e-6)
    for k, p in net.named_parameters():
        ref = g["grad." + k].astype(np.float64)
        err = np.linalg.norm(p.grad.cpu().numpy().astype(np.float64) - ref)
        assert err <= 3e-3 * np.linalg.norm(ref) + 1e-7, (k, err, np.linalg.norm(ref))
