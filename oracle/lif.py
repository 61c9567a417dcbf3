"""Oracle (TEST INFRASTRUCTURE): ConvLIF / ConvLIFRecurrent layer-step on CPU, torch fp32.

Restates models/spiking_submodules.py:121-151 (ConvLIF.forward) and :265-300
(ConvLIFRecurrent.forward) plus the surrogate gradients of models/spiking_util.py:28-93
as free functions.  The op sequence and rounding order follow the reference exactly
(every Python operator is one separately-rounded fp32 op):

    hard reset:  v' = ((v * lam) * (1 - z)) + ((1 - lam) * I)
    soft reset:  v' = ((v * lam) + ((1 - lam) * I)) - (z * theta)
    spike:       z' = float((v' - theta) > 0)

``lif_step`` is differentiable through torch autograd with the same surrogate
Functions as the reference, so it doubles as the gradient oracle;
``lif_step_backward`` restates the hand-derived BPTT recurrences of SURVEY.md
section 8(a3) that the CUDA backward kernel implements.
"""
import torch
import torch.nn.functional as F

SURROGATES = ("arctanspike", "superspike", "trianglespike", "mgspike")


def _gaussian(x, mu, sigma):
    """spiking_util.py:6-10."""
    import math
    return torch.exp(-((x - mu) * (x - mu)) / (2 * sigma * sigma)) / (sigma * math.sqrt(2 * math.pi))


def surrogate_grad(u, width, kind):
    """d spike / d u for u = v' - theta.  spiking_util.py:42 (superspike), :60-64 (multi-Gauss), :78 (triangle), :92 (arctan)."""
    if kind == "arctanspike":
        return 1 / (1 + width * u * u)
    if kind == "superspike":
        return 1 / (1 + width * u.abs()) ** 2
    if kind == "trianglespike":
        return F.relu(1 - width * u.abs())
    if kind == "mgspike":                                    # spiking_util.py:56-64 (MultiGaussSpike)
        return 1.15 * _gaussian(u, 0.0, width) - 0.15 * _gaussian(u, width, 6 * width) - 0.15 * _gaussian(u, -width, 6 * width)
    raise ValueError(kind)


class _Spike(torch.autograd.Function):
    """spiking_util.py:13-21 forward (x.gt(0).float()), backward = grad * surrogate."""

    @staticmethod
    def forward(ctx, u, width, kind):
        ctx.save_for_backward(u)
        ctx.width, ctx.kind = width, kind
        return u.gt(0).float()

    @staticmethod
    def backward(ctx, g):
        (u,) = ctx.saved_tensors
        return g * surrogate_grad(u, ctx.width, ctx.kind), None, None


def lif_step(x, w_ff, leak, thresh, v=None, z=None, w_rec=None, residual=None, *, hard_reset=True,
             detach=True, activation="arctanspike", act_width=10.0):
    """One layer-step.  Returns (out, v', z', I) with out = z' + residual.

    x [B,Cin,H,W]; w_ff [C,Cin,k,k]; w_rec [C,C,k,k] or None; leak/thresh [C,1,1] raw parameters;
    v, z [B,C,H,W] previous state (None = zeros, spiking_submodules.py:128-129).
    """
    pad = w_ff.shape[-1] // 2
    cur = F.conv2d(x, w_ff, padding=pad)                       # :125 / :269
    if v is None:
        v = torch.zeros_like(cur)
        z = torch.zeros_like(cur)
    if w_rec is not None:
        cur = cur + F.conv2d(z, w_rec, padding=pad)            # :279, :293 (non-detached z)
    theta = thresh.clamp_min(0.01)                             # :133
    lam = torch.sigmoid(leak)                                  # :136
    if detach:
        z = z.detach()                                         # :139-140
    if hard_reset:
        v_out = v * lam * (1 - z) + (1 - lam) * cur            # :144
    else:
        v_out = v * lam + (1 - lam) * cur - z * theta          # :146
    z_out = _Spike.apply(v_out - theta, act_width, activation)  # :149, spiking_util.py:108-109
    out = z_out if residual is None else z_out + residual      # :151
    return out, v_out, z_out, cur


def lif_step_backward(x, w_ff, w_rec, lam, theta, v_in, z_in, v_out, cur, g_z, g_v, *, hard_reset=True,
                      detach=True, activation="arctanspike", act_width=10.0, need_gx=True):
    """Hand-derived backward of one layer-step.

    g_z = total gradient w.r.t. z' (output spikes + next step's use of state[1]),
    g_v = gradient w.r.t. v' coming from the next step's use of state[0].
    lam/theta are the *effective* per-channel values [C,1,1] (sigmoid / clamp already applied).
    Returns dict(g_x, g_v_in, g_z_in, dw_ff, dw_rec, dlam, dtheta).
    """
    pad = w_ff.shape[-1] // 2
    sg = surrogate_grad(v_out - theta, act_width, activation)
    gs = g_z * sg                                   # through the spike
    gv = g_v + gs                                   # total d/d v'
    g_cur = gv * (1 - lam)
    if hard_reset:
        g_v_in = gv * lam * (1 - z_in)
        dlam = (gv * (v_in * (1 - z_in) - cur)).sum(dim=(0, 2, 3))
        dtheta = -gs.sum(dim=(0, 2, 3))
        g_z_reset = gv * (-(v_in * lam))
    else:
        g_v_in = gv * lam
        dlam = (gv * (v_in - cur)).sum(dim=(0, 2, 3))
        dtheta = -gs.sum(dim=(0, 2, 3)) - (gv * z_in).sum(dim=(0, 2, 3))
        g_z_reset = gv * (-theta)
    g_z_in = torch.zeros_like(z_in) if detach else g_z_reset
    dw_rec = None
    if w_rec is not None:
        g_z_in = g_z_in + F.conv_transpose2d(g_cur, w_rec, padding=pad)
        dw_rec = torch.nn.grad.conv2d_weight(z_in, w_rec.shape, g_cur, padding=pad)
    g_x = F.conv_transpose2d(g_cur, w_ff, padding=pad) if need_gx else None
    dw_ff = torch.nn.grad.conv2d_weight(x, w_ff.shape, g_cur, padding=pad)
    return dict(g_x=g_x, g_v_in=g_v_in, g_z_in=g_z_in, dw_ff=dw_ff, dw_rec=dw_rec, dlam=dlam, dtheta=dtheta)


def dyadic(w, bits=12):
    """Snap to a 2^-bits grid: makes every conv partial sum exact in fp32 for spike / small-integer
    inputs, hence independent of summation order."""
    s = float(2 ** bits)
    return torch.round(w * s) / s
