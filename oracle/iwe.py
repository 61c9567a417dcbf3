"""Oracle (TEST INFRASTRUCTURE): image of warped events (IWE) on CPU, torch fp32.

Restates utils/iwe.py:4-93 (purge_unfeasible, get_interpolation, interpolate), the per-event
flow gather of loss/flow.py:66-81 / utils/iwe.py:110-120 and compute_pol_iwe (utils/iwe.py:133-154).
Differentiable through torch autograd, which is the gradient oracle for the CUDA backward
(including the reference's tie semantics: abs'(0)=0 and max(0,0) passing half the gradient,
SURVEY.md section 8 a9).
"""
import torch


def gather_event_flow(flow, events, res):
    """flow [B,2,H,W] (channel 0 = x, 1 = y); events [B,N,4] = (ts,y,x,p) -> [B,N,2] = (fy,fx).
    The flat index is formed in fp32 as y*W + x and then truncated (loss/flow.py:67-69,77)."""
    idx = (events[:, :, 1] * res[1] + events[:, :, 2]).long()
    f = flow.reshape(flow.shape[0], 2, -1)
    fy = torch.gather(f[:, 1, :], 1, idx)
    fx = torch.gather(f[:, 0, :], 1, idx)
    return torch.stack([fy, fx], dim=2)


def interpolation(events, ev_flow, tref, res, flow_scaling, round_idx=False):
    """utils/iwe.py:20-71.  Returns (idx [B,K,1] float, weights [B,K,1]) with K = N (round) or 4N
    (bilinear; corner order top-left, top-right, bottom-left, bottom-right, concatenated along N)."""
    warped = events[:, :, 1:3] + (tref - events[:, :, 0:1]) * ev_flow * flow_scaling   # :37
    if round_idx:
        idx = torch.round(warped)                                                      # :41 half-to-even
        w = torch.ones_like(idx)
    else:
        ty, by = torch.floor(warped[:, :, 0:1]), torch.floor(warped[:, :, 0:1] + 1)    # :45-48
        lx, rx = torch.floor(warped[:, :, 1:2]), torch.floor(warped[:, :, 1:2] + 1)
        idx = torch.cat([torch.cat([ty, lx], 2), torch.cat([ty, rx], 2),
                         torch.cat([by, lx], 2), torch.cat([by, rx], 2)], dim=1)       # :50-54
        rep = torch.cat([warped] * 4, dim=1)
        w = torch.max(torch.zeros_like(rep), 1 - torch.abs(rep - idx))                 # :57-59
    oob = (idx[:, :, 0:1] < 0) | (idx[:, :, 0:1] >= res[0]) | (idx[:, :, 1:2] < 0) | (idx[:, :, 1:2] >= res[1])
    mask = (~oob).float()                                                              # :13-17
    idx = idx * mask
    w = torch.prod(w, dim=-1, keepdim=True) * mask                                     # :65
    flat = idx[:, :, 0:1] * res[1] + idx[:, :, 1:2]                                    # :68-69
    return flat, w


def scatter_image(idx, weights, res, polarity_mask=None):
    """utils/iwe.py:74-93: zeros(B, H*W).scatter_add_(idx, weights [* mask]) -> [B,1,H,W]."""
    if polarity_mask is not None:
        weights = weights * polarity_mask
    B = idx.shape[0]
    img = torch.zeros(B, res[0] * res[1], 1, dtype=weights.dtype)
    img = img.scatter_add(1, idx.long(), weights)
    return img.view(B, 1, res[0], res[1])


def warp_images(events, ev_flow, pol_mask, tref, res, flow_scaling, ts_weight=None, round_idx=False):
    """The four images one direction of the contrast loss needs (loss/flow.py:199-213):
    [B,4,H,W] = (count+, count-, sum_ts+, sum_ts-); ts_weight [B,N,1] multiplies the weights of
    the last two (ts for the forward warp, max_ts - ts for the backward warp)."""
    idx, w = interpolation(events, ev_flow, tref, res, flow_scaling, round_idx)
    rep = 1 if round_idx else 4
    pm = torch.cat([pol_mask] * rep, dim=1)
    outs = [scatter_image(idx, w, res, pm[:, :, 0:1]), scatter_image(idx, w, res, pm[:, :, 1:2])]
    if ts_weight is not None:
        tw = torch.cat([ts_weight] * rep, dim=1)
        outs += [scatter_image(idx, w * tw, res, pm[:, :, 0:1]), scatter_image(idx, w * tw, res, pm[:, :, 1:2])]
    return torch.cat(outs, dim=1)


def pol_iwe(flow, events, res, pos_mask, neg_mask, flow_scaling=128, round_idx=True):
    """utils/iwe.py:133-154 compute_pol_iwe (tref = 1): [B,2,H,W]."""
    ev_flow = gather_event_flow(flow, events, res)
    idx, w = interpolation(events, ev_flow, 1, res, flow_scaling, round_idx)
    rep = 1 if round_idx else 4
    pm = torch.cat([pos_mask] * rep, dim=1)
    nm = torch.cat([neg_mask] * rep, dim=1)
    return torch.cat([scatter_image(idx, w, res, pm), scatter_image(idx, w, res, nm)], dim=1)
