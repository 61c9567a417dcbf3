"""Drive the UNMODIFIED reference (TEST / BASELINE INFRASTRUCTURE): its own ``LIFFireNet`` / ``LIFFireFlowNet``,
``EventWarping`` and the loop bodies of ``train_flow.py:232-279`` / ``eval_flow.py:220-237``.

The reference is resolved by ``oracle/ref_shim.py`` (``/root/reference`` in the build container, the staged copy
under the git-ignored ``baseline/_ref`` on the GPU box).  Two users:

  * ``bench.py --impl reference`` and the ``cpu_baseline`` leg: the reference's own cells on the host cores;
  * ``tests/test_gpu_reference_seam.py``: the same classes with ``head_neuron / ff_neuron / rec_neuron``
    (models/model.py:37-39) pointed at the CUDA cells - the drop-in seam, executed on the GPU.

The shipped ``LIFFireNet`` wires the ``SNNtorch_*`` cells, which need snntorch (absent offline); the hot path of this
repo is the ``ConvLIF`` / ``ConvLIFRecurrent`` pair of models/spiking_submodules.py, installed through the class
attributes exactly as a user of the reference would.
"""
import torch

from . import ref_shim


def available():
    return ref_shim.reference_available()


def model_config(channels, mask_output=False):
    """The `model` section LIFFireNet reads (configs/train_SNN.yml:14-24; model.py:43-71)."""
    return dict(name="LIFFireNet", encoding="cnt", num_bins=2, round_encoding=False, norm_input=False, mask_output=mask_output,
                spiking_neuron=None, base_num_channels=channels, kernel_size=3, activations=["arctanspike", "arctanspike"],
                quantization={"enabled": False})


def build_net(kind="LIFFireNet", channels=32, cells=None, leak=(0.0, 1.0), thresh=(0.3, 0.1), seed=0, dyadic=False,
              pred_gain=1.0, mask_output=False):
    """The reference's network class with its cells chosen through the class-attribute seam.
    cells: None -> the reference's own ConvLIF / ConvLIFRecurrent (adapted for the kwargs LIFFireNet passes),
           or a (ff_cell, rec_cell) pair, e.g. the CUDA cells.
    model.py never forwards leak / thresh statistics to the cells: they are re-drawn here, in a
    fixed order, so that two nets built with the same seed hold identical parameters whatever cells they use."""
    ref = ref_shim.load()
    base = getattr(ref.model, kind)
    if cells is None:
        ff, rec = ref.adapt(ref.ConvLIF), ref.adapt(ref.ConvLIFRecurrent)
    else:
        ff, rec = cells
    recurrent = kind == "LIFFireNet"

    class Net(base):
        head_neuron = ff
        ff_neuron = ff
        rec_neuron = rec if recurrent else ff

    torch.manual_seed(seed)
    net = Net(model_config(channels, mask_output))
    g = torch.Generator().manual_seed(seed + 12345)
    with torch.no_grad():
        for name in ("head", "G1", "R1a", "R1b", "G2", "R2a", "R2b"):
            cell = getattr(net, name)
            cell.leak.copy_(torch.randn(cell.leak.shape, generator=g) * leak[1] + leak[0])
            cell.thresh.copy_(torch.randn(cell.thresh.shape, generator=g) * thresh[1] + thresh[0])
        net.pred.conv2d.weight.mul_(pred_gain)
        if dyadic:
            for n, p in net.named_parameters():
                if n.endswith("weight"):
                    p.copy_(torch.round(p * 4096.0) / 4096.0)
    return net


def loss_config(res, mask_output=False, weight=0.001):
    return {"loader": {"resolution": list(res)}, "loss": {"flow_regul_weight": weight}, "model": {"mask_output": mask_output}}


def make_loss(res, device, mask_output=False, weight=0.001):
    ref = ref_shim.load()
    return ref.flow.EventWarping(loss_config(res, mask_output, weight), device)


def train_step(net, loss_fn, optimizer, window, device, clip_grad=1.0):
    """One optimizer step over one loss window: the loop body of train_flow.py:232-279, bin by bin.
    window: event_cnt [T,B,2,H,W], event_list [T,B,N,4], event_list_pol_mask [T,B,N,2], event_mask [T,B,1,H,W]
    (host tensors, moved to `device` per bin like the reference does)."""
    T = window["event_cnt"].shape[0]
    x = None
    for t in range(T):
        x = net(None, window["event_cnt"][t].to(device))
        loss_fn.event_flow_association(x["flow"], window["event_list"][t].clone().to(device),
                                       window["event_list_pol_mask"][t].to(device), window["event_mask"][t].to(device))
    loss = loss_fn()
    loss.backward()
    grads = {n: p.grad.detach().clone() for n, p in net.named_parameters() if p.grad is not None}
    if clip_grad is not None:
        torch.nn.utils.clip_grad.clip_grad_norm_(net.parameters(), clip_grad)
    optimizer.step()
    optimizer.zero_grad()
    net.detach_states()
    loss_fn.reset()
    return loss.detach(), grads, x["flow"][-1].detach()


def eval_frames(net, cnt_window, device, events=None, pol_masks=None, res=None, flow_scaling=None):
    """The evaluation loop body of eval_flow.py:220-237 under no_grad: one model() call per frame and, when the frame's
    event list is given, compute_pol_iwe(round_idx=True) on its flow.  Returns the flows (and IWEs)."""
    ref = ref_shim.load()
    flows, iwes = [], []
    with torch.no_grad():
        for t in range(cnt_window.shape[0]):
            x = net(None, cnt_window[t].to(device))
            flows.append(x["flow"][-1])
            if events is not None:
                pm = pol_masks[t].to(device)
                iwes.append(ref.iwe.compute_pol_iwe(x["flow"][-1], events[t].to(device), res, pm[:, :, 0:1], pm[:, :, 1:2],
                                                    flow_scaling=flow_scaling if flow_scaling is not None else max(res),
                                                    round_idx=True))
    return flows, iwes
