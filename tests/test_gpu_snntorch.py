"""GPU: the SNNtorch_ConvLIF / SNNtorch_ConvLIFRecurrent cells (what the reference's LIFFireNet
instantiates by default, models/model.py:37-39) against oracle/snntorch_lif.py.  PARITY UNPINNED: the oracle restates
snntorch 0.9.4's published Leaky.forward from memory (snntorch is not installable offline); these tests pin the CUDA cells to
that restatement, not to snntorch itself."""
import numpy as np
import pytest
import torch

from oracle import snntorch_lif as osl
from snnflow_testutil import grad_report

pytestmark = pytest.mark.gpu


def _params(cell):
    return {k: v.detach().cpu().clone() for k, v in cell.state_dict().items()}


@pytest.mark.parametrize("recurrent,hard_reset,training", [(False, True, True), (True, True, True), (True, False, True),
                                                           (False, False, False), (True, True, False)])
def test_cell_steps_match_oracle(recurrent, hard_reset, training):
    import snnflow_b200 as snnflow
    torch.manual_seed(4)
    Cin, C, B, H, W, T = 8, 16, 2, 12, 20, 4
    cls = snnflow.SNNtorch_ConvLIFRecurrent if recurrent else snnflow.SNNtorch_ConvLIF
    cell = cls(Cin, C, 3, leak=(0.3, 1.2), thresh=(0.05, 0.5), hard_reset=hard_reset)
    with torch.no_grad():
        cell.bn.weight.uniform_(0.5, 1.5)
        cell.bn.bias.uniform_(-0.2, 0.4)
        cell.bn.running_mean.uniform_(-0.1, 0.1)
        cell.bn.running_var.uniform_(0.5, 1.5)
    cell.train(training)
    p = _params(cell)
    for k in ("ff.weight", "rec.weight", "bn.weight", "bn.bias", "lif.beta", "lif.threshold"):
        if k in p:
            p[k].requires_grad_(True)
    cell = cell.cuda()
    g = torch.Generator().manual_seed(5)
    xs = (torch.rand(T, B, Cin, H, W, generator=g) < 0.3).float()
    gout = torch.randn(T, B, C, H, W, generator=g)
    xs_r = xs.clone().requires_grad_(True)
    xs_g = xs.clone().cuda().requires_grad_(True)
    st_r = st_g = None
    loss_r = loss_g = 0
    near = mism = total = 0
    for t in range(T):
        spk_r, st_r = osl.cell_step(xs_r[t], st_r, p, recurrent=recurrent, hard_reset=hard_reset, training=training)
        spk_g, st_g = cell(xs_g[t], st_g)
        assert tuple(st_g.shape) == (2, B, C, H, W)
        diff = (spk_g.detach().cpu() != spk_r.detach())
        mism += int(diff.sum())
        total += diff.numel()
        if int(diff.sum()):
            # a flipped spike is only acceptable within fp32 round-off of threshold; afterwards the trajectories differ,
            # so the comparison is teacher-forced: the oracle continues from the GPU state
            st_r = st_g.detach().cpu().clone()
            near += int(diff.sum())
        else:
            np.testing.assert_allclose(st_g[0].detach().cpu().numpy(), st_r[0].detach().numpy(), rtol=1e-4, atol=2e-5)
        loss_r = loss_r + (spk_r * gout[t]).sum()
        loss_g = loss_g + (spk_g * gout[t].cuda()).sum()
    assert mism <= 2e-4 * total, (mism, total)
    assert float(st_r[1].mean()) > 0.01, "silent cell: vacuous"
    if mism:
        return   # gradients of diverged trajectories are not comparable
    loss_r.backward()
    loss_g.backward()
    named = dict(cell.named_parameters())
    for k in ("ff.weight", "rec.weight", "bn.weight", "bn.bias", "lif.beta", "lif.threshold"):
        if k in p:
            frac, rel = grad_report(named[k].grad.cpu().numpy(), p[k].grad.numpy(), rtol=1e-4, atol_rel=1e-5)
            assert rel <= 2e-4, (k, frac, rel)
    frac, rel = grad_report(xs_g.grad.cpu().numpy(), xs_r.grad.numpy(), rtol=1e-4, atol_rel=1e-5)
    assert rel <= 2e-4, ("g_x", frac, rel)
    if training:   # BatchNorm's running statistics were updated like the oracle's
        np.testing.assert_allclose(cell.bn.running_mean.cpu().numpy(), p["bn.running_mean"].numpy(), rtol=1e-4, atol=1e-6)
        np.testing.assert_allclose(cell.bn.running_var.cpu().numpy(), p["bn.running_var"].numpy(), rtol=1e-4, atol=1e-6)


def test_shipped_network_runs_and_keeps_reference_keys():
    """SNNtorchLIFFireNet = the reference's LIFFireNet as shipped: state_dict keys of its cells, state plumbing, a training
    step's worth of forward + backward through all seven cells."""
    import snnflow_b200 as snnflow
    torch.manual_seed(0)
    net = snnflow.SNNtorchLIFFireNet(dict(num_bins=2, encoding="cnt", base_num_channels=16, kernel_size=3)).cuda()
    keys = list(net.state_dict())
    assert keys[:10] == ["head.ff.weight", "head.lif.beta", "head.lif.threshold", "head.lif.graded_spikes_factor",
                         "head.lif.reset_mechanism_val", "head.bn.weight", "head.bn.bias", "head.bn.running_mean",
                         "head.bn.running_var", "head.bn.num_batches_tracked"]
    assert "G1.rec.weight" in keys and "pred.conv2d.weight" in keys
    g = torch.Generator().manual_seed(1)
    cnt = torch.poisson(torch.full((3, 2, 2, 16, 24), 0.3), generator=g).cuda()
    loss = 0
    for t in range(3):
        loss = loss + net(None, cnt[t])["flow"][0].square().sum()
    assert len(net._states) == 7 and all(tuple(s.shape) == (2, 2, 16, 16, 24) for s in net._states)
    assert max(float(s[1].mean()) for s in net._states) > 0.01
    loss.backward()
    for n, p in net.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all(), n
    assert float(net.G1.rec.weight.grad.abs().sum()) > 0 and float(net.head.lif.beta.grad.abs().sum()) > 0
    net.detach_states()
    # a checkpoint of the ConvLIF-cell network is refused with an explicit message, and vice versa
    other = snnflow.LIFFireNet(dict(num_bins=2, encoding="cnt", base_num_channels=16, kernel_size=3))
    with pytest.raises(RuntimeError, match="SNNtorch"):
        other.load_state_dict(net.state_dict())
