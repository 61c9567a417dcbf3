"""Oracle (TEST INFRASTRUCTURE): LIFFireNet / LIFFireFlowNet on CPU, torch fp32.

Restates models/model.py:172-182 (LIFFireNet.forward: head -> G1 -> R1a -> R1b -> G2 -> R2a -> R2b
-> 1x1 tanh pred) and :393-395 (LIFFireFlowNet: the same chain with feed-forward cells at G1/G2)
on top of ``oracle.lif.lif_step``; the prediction head is models/submodules.py:96-113
(conv 1x1 + bias, tanh).  Parameters live in a flat dict keyed like the reference state_dict
(``head.ff.weight``, ``G1.rec.weight``, ``R1a.leak`` ... ``pred.conv2d.weight``).
Also the CPU baseline that ``bench.py`` times.
"""
import math

import torch
import torch.nn.functional as F

from .lif import lif_step

LAYERS = ("head", "G1", "R1a", "R1b", "G2", "R2a", "R2b")
RECURRENT = ("G1", "G2")


def init_params(base_channels=32, num_bins=2, recurrent=True, leak=(-4.0, 0.1), thresh=(0.8, 0.0),
                w_scale_pred=0.01, gen=None):
    """Parameter init following spiking_submodules.py:88-100,234-239 and model.py:43,105-107."""
    C = base_channels
    p = {}

    def unif(shape, scale):
        return (torch.rand(shape, generator=gen) * 2 - 1) * scale

    for name in LAYERS:
        cin = num_bins if name == "head" else C
        p[f"{name}.ff.weight"] = unif((C, cin, 3, 3), math.sqrt(1 / cin))
        if recurrent and name in RECURRENT:
            p[f"{name}.rec.weight"] = unif((C, C, 3, 3), math.sqrt(1 / C))
        p[f"{name}.leak"] = torch.randn(C, 1, 1, generator=gen) * leak[1] + leak[0]
        p[f"{name}.thresh"] = torch.randn(C, 1, 1, generator=gen) * thresh[1] + thresh[0]
    p["pred.conv2d.weight"] = unif((2, C, 1, 1), w_scale_pred)
    p["pred.conv2d.bias"] = torch.zeros(2)
    return p


def forward(params, x, states, *, residual=False, hard_reset=True, detach=True, activation="arctanspike",
            act_width=10.0):
    """One time bin.  states: list of 7 entries, each None or (v, z).  Returns (flow, new_states, spikes)."""
    new_states, spikes = [], []
    h = x
    skip = None
    for i, name in enumerate(LAYERS):
        st = states[i]
        v, z = (None, None) if st is None else st
        res = None
        if residual and name in ("R1b", "R2b"):
            res = skip                                          # model.py:176,180
        out, v2, z2, _ = lif_step(h, params[f"{name}.ff.weight"], params[f"{name}.leak"], params[f"{name}.thresh"],
                                  v, z, params.get(f"{name}.rec.weight"), res, hard_reset=hard_reset,
                                  detach=detach, activation=activation, act_width=act_width)
        new_states.append((v2, z2))
        spikes.append(out)
        if name in RECURRENT:
            skip = out
        h = out
    flow = torch.tanh(F.conv2d(h, params["pred.conv2d.weight"], params["pred.conv2d.bias"]))
    return flow, new_states, spikes
