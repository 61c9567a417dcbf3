"""Oracle (TEST INFRASTRUCTURE, **PARITY UNPINNED**): the SNNtorch_ConvLIF / SNNtorch_ConvLIFRecurrent cell on CPU.

Restates models/SNNtorch_spiking_submodules.py:283-322 (SNNtorch_ConvLIF.forward) and :492-567
(SNNtorch_ConvLIFRecurrent.forward), whose neuron is the third-party ``snn.Leaky`` of snntorch 0.9.4
(requirements.txt:8).  snntorch is not installed in this image and its source is not on disk, so ``leaky_step`` restates
the PUBLISHED algorithm of ``snntorch.Leaky.forward`` (reset_delay=False, default ATan(alpha=2) surrogate) from memory:

    reset = H(mem - threshold).detach()
    "zero":     mem = beta.clamp(0, 1) * ((1 - reset) * mem) + input
    "subtract": mem = beta.clamp(0, 1) * mem + input - reset * threshold
    spk = H(mem - threshold)           backward: grad / (1 + (pi * (mem - threshold))**2)     (ATan, alpha = 2)
    do_reset = spk - reset             (no double reset)
    "zero": mem = mem - do_reset * mem        "subtract": mem = mem - do_reset * threshold

Neither the reference's tests nor a runnable snntorch pin this restatement: the GPU cells are checked against it, and both
are labelled parity-unpinned.
"""
import math

import torch
import torch.nn.functional as F


class _ATan(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return (x > 0).float()

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        return g / (1 + (math.pi * x) ** 2)      # alpha / 2 / (1 + (pi / 2 * alpha * x)^2) with alpha = 2


def leaky_step(cur, mem, beta, threshold, reset_mechanism="zero"):
    if mem is None:
        mem = torch.zeros_like(cur)
    reset = _ATan.apply(mem - threshold).detach()
    b = beta.clamp(0, 1)
    if reset_mechanism == "zero":
        m = b * ((1 - reset) * mem) + cur
    else:
        m = b * mem + cur - reset * threshold
    spk = _ATan.apply(m - threshold)
    do_reset = spk - reset
    mem_out = m - do_reset * m if reset_mechanism == "zero" else m - do_reset * threshold
    return spk, mem_out


def cell_step(x, prev_state, p, *, recurrent, hard_reset=True, training=True, bn_momentum=0.1, bn_eps=1e-5):
    """One SNNtorch_ConvLIF(/Recurrent) step.  p: dict with ff.weight[, rec.weight], bn.weight, bn.bias, bn.running_mean,
    bn.running_var, lif.beta [C,1,1], lif.threshold [C,1,1] (running stats are updated in place when training)."""
    thr = p["lif.threshold"].clamp_min(0.01)                                     # :284 (in-place clamp of the parameter data)
    cur = F.conv2d(x, p["ff.weight"], padding=1)
    mem = None if prev_state is None else prev_state[0]
    if recurrent:
        prev_spk = torch.zeros_like(cur) if prev_state is None else prev_state[1]
        cur = cur + F.conv2d(prev_spk, p["rec.weight"], padding=1)               # :497-521
    cur = F.batch_norm(cur, p["bn.running_mean"], p["bn.running_var"], p["bn.weight"], p["bn.bias"], training, bn_momentum, bn_eps)
    spk, mem_out = leaky_step(cur, mem, p["lif.beta"], thr, "zero" if hard_reset else "subtract")
    mem_out = mem_out.detach()                                                   # :309-311
    return spk, torch.stack([mem_out, spk], dim=0)
