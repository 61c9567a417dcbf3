"""Contrast-maximisation loss: host-side mirror of loss/flow.py:28-303 (EventWarping).

Same constructor, ``event_flow_association`` / ``reset`` / ``num_events`` / ``forward`` behaviour as the
reference (including the in-place timestamp shift of the event list, :91).  The eight
``interpolate`` scatters of the reference forward (:199-213, :232-246) are two fused
``snnflow_iwe_splat_fwd`` launches; their backward is ``snnflow_iwe_splat_bwd``.  The arithmetic after
the splat (:214-301) is a handful of small torch ops on [B,4,H,W] images.
"""
import torch

from . import _lib
from .iwe import gather_event_flow, warp_images
from .spiking_submodules import _f32c


class _WindowLoss(torch.autograd.Function):
    """loss, d loss / d flow of a whole window in one C call (snnflow_window_loss, include/snnflow.h)."""

    @staticmethod
    def forward(ctx, flow, events, pol_mask, event_mask, owner):
        L = _lib.lib()
        flow, events, pol_mask = _f32c(flow), _f32c(events), _f32c(pol_mask)
        T, B, N = events.shape[0], events.shape[1], events.shape[2]
        H, W = flow.shape[-2], flow.shape[-1]
        mask = _f32c(event_mask) if (owner.smoothing_mask and event_mask is not None) else None
        nbytes = L.snnflow_window_loss_workspace_bytes(T, B, N, H, W)
        ws = owner._workspace(nbytes, flow.device)
        loss = torch.empty(1, dtype=torch.float32, device=flow.device)
        g_flow = torch.empty_like(flow)
        _lib.check(L.snnflow_window_loss(_lib.ptr(flow), _lib.ptr(events), _lib.ptr(pol_mask), _lib.ptr(mask), _lib.ptr(loss),
                                         _lib.ptr(g_flow), ws.data_ptr(), ws.numel(), T, B, N, H, W, float(owner.flow_scaling),
                                         float(owner.weight), int(bool(owner.loss_scaling)), _lib.stream()),
                   "snnflow_window_loss")
        ctx.save_for_backward(g_flow)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        (g_flow,) = ctx.saved_tensors
        return g_flow * g, None, None, None, None


class EventWarping(torch.nn.Module):
    def __init__(self, config, device, flow_scaling=None, loss_scaling=True):
        super().__init__()
        self.loss_scaling = loss_scaling
        self.res = config["loader"]["resolution"]
        self.flow_scaling = flow_scaling if flow_scaling is not None else max(config["loader"]["resolution"])
        self.weight = config["loss"]["flow_regul_weight"]
        self.smoothing_mask = config["model"].get("mask_output", False)
        self.overwrite_intermediate = config["loss"].get("overwrite_intermediate", False)
        self.device = device
        self.reset()

    def _workspace(self, nbytes, dev):
        ws = getattr(self, "_ws", None)
        if ws is None or ws.numel() < nbytes or ws.device != dev:
            ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            self._ws = ws
        return ws

    @property
    def fused_window_loss_ok(self):
        """The one-call window loss covers the training configuration of the reference (per-bin flow association)."""
        return not self.overwrite_intermediate

    def window_loss(self, flows, event_list, pol_mask, event_mask=None):
        """The loss of one window from its T flow maps and event lists at once - what T calls of
        event_flow_association followed by forward() return (loss/flow.py:58-121, :178-303) - as one fused call.
        flows [T,B,2,H,W]; event_list [T,B,N,4] with per-bin timestamps in [0,1] (not modified); pol_mask [T,B,N,2];
        event_mask [T,B,1,H,W] (only read when config model.mask_output is set)."""
        if not flows.is_cuda:
            raise _lib.SnnflowError("snnflow EventWarping.window_loss runs on CUDA tensors only (no CPU fallback)")
        return _WindowLoss.apply(flows, event_list, pol_mask, event_mask, self)

    def window_loss_and_grad(self, flows, event_list, pol_mask, event_mask=None):
        """window_loss() without autograd: returns (loss [scalar tensor], d loss / d flows [T,B,2,H,W]) - one C call.
        Used by the direct training step (train.TrainWindow.step_direct)."""
        if not flows.is_cuda:
            raise _lib.SnnflowError("snnflow EventWarping.window_loss runs on CUDA tensors only (no CPU fallback)")
        L = _lib.lib()
        flow, events, pol = _f32c(flows.detach()), _f32c(event_list), _f32c(pol_mask)
        T, B, N = events.shape[0], events.shape[1], events.shape[2]
        H, W = flow.shape[-2], flow.shape[-1]
        mask = _f32c(event_mask) if (self.smoothing_mask and event_mask is not None) else None
        ws = self._workspace(L.snnflow_window_loss_workspace_bytes(T, B, N, H, W), flow.device)
        loss = torch.empty(1, dtype=torch.float32, device=flow.device)
        g_flow = torch.empty_like(flow)
        _lib.check(L.snnflow_window_loss(_lib.ptr(flow), _lib.ptr(events), _lib.ptr(pol), _lib.ptr(mask), _lib.ptr(loss),
                                         _lib.ptr(g_flow), ws.data_ptr(), ws.numel(), T, B, N, H, W, float(self.flow_scaling),
                                         float(self.weight), int(bool(self.loss_scaling)), _lib.stream()), "snnflow_window_loss")
        return loss.reshape(()), g_flow

    def reset(self):
        self._passes = 0
        self._event_list = None
        self._flow_list = None
        self._flow_maps_x = None
        self._flow_maps_y = None
        self._pol_mask_list = None
        self._event_mask = None

    @property
    def num_events(self):
        return 0 if self._event_list is None else self._event_list.shape[1]

    @property
    def event_mask(self):
        if self.overwrite_intermediate:
            return self._event_mask                 # mask of the training window (loss/flow.py:170-175)
        return self._event_mask[:, -1:, :, :]       # mask of the last forward pass

    def overwrite_intermediate_flow(self, flow_list):
        """loss/flow.py:123-153: every event of the window is re-associated with the FINAL flow estimate (one flow map per
        scale instead of one per bin), and the event mask becomes the union over the window."""
        self._flow_list, self._flow_maps_x, self._flow_maps_y = [], [], []
        for flow in flow_list:
            self._flow_maps_x.append(flow[:, 0:1])
            self._flow_maps_y.append(flow[:, 1:2])
            self._flow_list.append(gather_event_flow(flow, self._event_list, self.res))
        self._event_mask = torch.sum(self._event_mask, dim=1, keepdim=True).clamp_(max=1)

    def event_flow_association(self, flow_list, event_list, pol_mask, event_mask):
        """loss/flow.py:58-121."""
        if self._flow_list is None:
            self._flow_list, self._flow_maps_x, self._flow_maps_y = [], [], []
        for i, flow in enumerate(flow_list):
            ev_flow = gather_event_flow(flow, event_list, self.res)
            if i == len(self._flow_list):
                self._flow_list.append(ev_flow)
                self._flow_maps_x.append(flow[:, 0:1])
                self._flow_maps_y.append(flow[:, 1:2])
            else:
                self._flow_list[i] = torch.cat([self._flow_list[i], ev_flow], dim=1)
                self._flow_maps_x[i] = torch.cat([self._flow_maps_x[i], flow[:, 0:1]], dim=1)
                self._flow_maps_y[i] = torch.cat([self._flow_maps_y[i], flow[:, 1:2]], dim=1)
        if self._event_list is None:
            self._event_list, self._pol_mask_list, self._event_mask = event_list, pol_mask, event_mask
        else:
            event_list[:, :, 0:1] += self._passes
            self._event_list = torch.cat([self._event_list, event_list], dim=1)
            self._pol_mask_list = torch.cat([self._pol_mask_list, pol_mask], dim=1)
            self._event_mask = torch.cat([self._event_mask, event_mask], dim=1)
        self._passes += 1

    def _direction(self, ev_flow, tref, ts_mode, max_ts):
        img = warp_images(self._event_list, ev_flow, self._pol_mask_list, tref, self.res, self.flow_scaling,
                          ts_mode=ts_mode, ts_ref=max_ts)
        cnt_p, cnt_n = img[:, 0:1], img[:, 1:2]
        ts_p = img[:, 2:3] / (cnt_p + 1e-9) / max_ts                       # loss/flow.py:214-217
        ts_n = img[:, 3:4] / (cnt_n + 1e-9) / max_ts
        B = img.shape[0]
        loss = (ts_p.reshape(B, -1) ** 2).sum(1) + (ts_n.reshape(B, -1) ** 2).sum(1)
        if self.loss_scaling:
            tot = cnt_p + cnt_n                                            # :224-227 (zero entries keep their grad path)
            loss = loss / torch.where(tot > 0, torch.ones_like(tot), tot).reshape(B, -1).sum(1)
        return loss.sum()

    def _smoothness(self, fx, fy):
        def charb(a, b):
            return torch.sqrt((a + b) ** 2 + 1e-6)

        terms = [
            charb(fx[:, :, :, :-1] - fx[:, :, :, 1:], fy[:, :, :, :-1] - fy[:, :, :, 1:]),
            charb(fx[:, :, :-1, :] - fx[:, :, 1:, :], fy[:, :, :-1, :] - fy[:, :, 1:, :]),
            charb(fx[:, :, :-1, :-1] - fx[:, :, 1:, 1:], fy[:, :, :-1, :-1] - fy[:, :, 1:, 1:]),
            charb(fx[:, :, 1:, :-1] - fx[:, :, :-1, 1:], fy[:, :, 1:, :-1] - fy[:, :, :-1, 1:]),
            charb(fx[:, :-1] - fx[:, 1:], fy[:, :-1] - fy[:, 1:]),
        ]
        if self.overwrite_intermediate:
            terms = terms[:4]
        if self.smoothing_mask:
            m = self._event_mask
            masks = [m[:, :, :, :-1] * m[:, :, :, 1:], m[:, :, :-1, :] * m[:, :, 1:, :],
                     m[:, :, :-1, :-1] * m[:, :, 1:, 1:], m[:, :, 1:, :-1] * m[:, :, :-1, 1:], m[:, :-1] * m[:, 1:]]
            terms = [a * b for a, b in zip(masks, terms)]
        total = terms[0].sum() + terms[1].sum() + terms[2].sum() + terms[3].sum()
        if self.overwrite_intermediate:                       # loss/flow.py:289-295: no temporal term, four components
            return total / 4 / fx.shape[1]
        return (total + terms[4].sum()) / 5 / fx.shape[1]

    def forward(self):
        max_ts = self._passes
        loss = 0
        for i in range(len(self._flow_list)):
            fw = self._direction(self._flow_list[i], max_ts, 1, max_ts)    # loss/flow.py:197-228
            bw = self._direction(self._flow_list[i], 0, 2, max_ts)         # :230-261
            loss = loss + fw + bw + self.weight * self._smoothness(self._flow_maps_x[i], self._flow_maps_y[i])
        return loss / len(self._flow_list)
