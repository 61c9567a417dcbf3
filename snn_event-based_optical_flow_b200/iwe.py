"""Image of warped events on the GPU: host-side mirror of utils/iwe.py.

``gather_event_flow`` is the per-event flow lookup of loss/flow.py:66-81 / utils/iwe.py:110-120;
``warp_images`` fuses get_interpolation (utils/iwe.py:20-71) with the 2 or 4 ``interpolate`` scatters
(:74-93) one direction of the contrast loss needs; ``compute_pol_iwe`` / ``deblur_events`` keep the
reference signatures (:96-154).  All are differentiable w.r.t. the flow with the reference's autograd
semantics (tie rules included).  The reference's host-synchronising bounds check (utils/iwe.py:87-89)
is unnecessary here: out-of-range corners are purged inside the kernel.
"""
import torch

from . import _lib
from .spiking_submodules import _f32c


class _GatherFlow(torch.autograd.Function):
    @staticmethod
    def forward(ctx, flow, events, H, W):
        flow, events = _f32c(flow), _f32c(events)
        B, N = events.shape[0], events.shape[1]
        out = torch.empty((B, N, 2), dtype=torch.float32, device=flow.device)
        _lib.check(_lib.lib().snnflow_flow_gather_fwd(_lib.ptr(flow), _lib.ptr(events), _lib.ptr(out), B, N, H, W,
                                                      _lib.stream()), "snnflow_flow_gather_fwd")
        # Not save_for_backward: EventWarping.event_flow_association later shifts the timestamps of this very
        # tensor in place (loss/flow.py:91), which would trip autograd's version check although the gather only
        # reads the (unmodified) y/x columns.
        ctx.events = events
        ctx.dims = (B, N, H, W)
        return out

    @staticmethod
    def backward(ctx, g):
        events = ctx.events
        B, N, H, W = ctx.dims
        g_flow = torch.zeros((B, 2, H, W), dtype=torch.float32, device=g.device)
        _lib.check(_lib.lib().snnflow_flow_gather_bwd(_lib.ptr(_f32c(g)), _lib.ptr(events), _lib.ptr(g_flow), B, N, H,
                                                      W, _lib.stream()), "snnflow_flow_gather_bwd")
        return g_flow, None, None, None


class _WarpImages(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ev_flow, events, pol_mask, H, W, tref, flow_scaling, n_img, ts_mode, ts_ref, round_idx):
        ev_flow, events, pol_mask = _f32c(ev_flow), _f32c(events), _f32c(pol_mask)
        B, N = events.shape[0], events.shape[1]
        out = torch.empty((B, n_img, H, W), dtype=torch.float32, device=events.device)
        scratch = torch.empty((B, n_img, H, W), dtype=torch.int64, device=events.device)
        _lib.check(_lib.lib().snnflow_iwe_splat_fwd(
            _lib.ptr(events), _lib.ptr(ev_flow), _lib.ptr(pol_mask), _lib.ptr(out), _lib.ptr(scratch), B, N, H, W,
            float(tref), float(flow_scaling), n_img, ts_mode, float(ts_ref), int(bool(round_idx)), _lib.stream()),
            "snnflow_iwe_splat_fwd")
        ctx.save_for_backward(ev_flow, events, pol_mask)
        ctx.cfg = (B, N, H, W, float(tref), float(flow_scaling), n_img, ts_mode, float(ts_ref), bool(round_idx))
        return out

    @staticmethod
    def backward(ctx, g_img):
        ev_flow, events, pol_mask = ctx.saved_tensors
        B, N, H, W, tref, S, n_img, ts_mode, ts_ref, round_idx = ctx.cfg
        if round_idx:   # rounding has zero gradient (utils/iwe.py:39-42)
            return (torch.zeros_like(ev_flow),) + (None,) * 10
        g = torch.empty_like(ev_flow)
        _lib.check(_lib.lib().snnflow_iwe_splat_bwd(
            _lib.ptr(events), _lib.ptr(ev_flow), _lib.ptr(pol_mask), _lib.ptr(_f32c(g_img)), _lib.ptr(g), B, N, H, W,
            tref, S, n_img, ts_mode, ts_ref, _lib.stream()), "snnflow_iwe_splat_bwd")
        return (g,) + (None,) * 10


def gather_event_flow(flow, events, res):
    """flow [B,2,H,W] (x, y); events [B,N,4] (ts,y,x,p) -> per-event flow [B,N,2] (fy, fx)."""
    return _GatherFlow.apply(flow, events, int(res[0]), int(res[1]))


def warp_images(events, ev_flow, pol_mask, tref, res, flow_scaling, ts_mode=0, ts_ref=0.0, round_idx=False):
    """[B,2,H,W] (count+, count-) for ts_mode 0, else [B,4,H,W] (+ ts-weighted sums; ts_mode 1: weight = ts,
    ts_mode 2: weight = ts_ref - ts)."""
    n_img = 2 if ts_mode == 0 else 4
    return _WarpImages.apply(ev_flow, events, pol_mask, int(res[0]), int(res[1]), tref, flow_scaling, n_img, ts_mode,
                             ts_ref, round_idx)


def deblur_events(flow, event_list, res, flow_scaling=128, round_idx=True, polarity_mask=None):
    """utils/iwe.py:96-130 -> [B,1,H,W]."""
    B, N = event_list.shape[0], event_list.shape[1]
    if polarity_mask is None:
        polarity_mask = torch.ones((B, N, 1), dtype=torch.float32, device=event_list.device)
    pm = torch.cat([polarity_mask, torch.zeros_like(polarity_mask)], dim=2)
    ev_flow = gather_event_flow(flow, event_list, res)
    return warp_images(event_list, ev_flow, pm, 1, res, flow_scaling, round_idx=round_idx)[:, 0:1]


def compute_pol_iwe(flow, event_list, res, pos_mask, neg_mask, flow_scaling=128, round_idx=True):
    """utils/iwe.py:133-154 -> [B,2,H,W] per-polarity image of warped events (one fused launch)."""
    pm = torch.cat([pos_mask, neg_mask], dim=2)
    ev_flow = gather_event_flow(flow, event_list, res)
    return warp_images(event_list, ev_flow, pm, 1, res, flow_scaling, round_idx=round_idx)
