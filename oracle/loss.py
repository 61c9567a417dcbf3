"""Oracle (TEST INFRASTRUCTURE): contrast-maximisation loss on CPU, torch fp32.

Restates loss/flow.py:28-303 (EventWarping: event_flow_association bookkeeping and forward)
for the configuration the training script uses (overwrite_intermediate=False), as a small
functional accumulator.  Differentiable through autograd.
"""
import torch

from . import iwe as _iwe


class EventWarpingOracle:
    def __init__(self, res, flow_regul_weight=0.001, flow_scaling=None, mask_output=False, loss_scaling=True):
        self.res = tuple(res)
        self.flow_scaling = flow_scaling if flow_scaling is not None else max(res)   # flow.py:42
        self.weight = flow_regul_weight
        self.smoothing_mask = mask_output                                            # flow.py:44
        self.loss_scaling = loss_scaling
        self.reset()

    def reset(self):                                                                 # flow.py:155-162
        self.passes = 0
        self.events = None
        self.ev_flow = None
        self.flow_x = None
        self.flow_y = None
        self.pol_mask = None
        self.event_mask = None

    @property
    def num_events(self):
        return 0 if self.events is None else self.events.shape[1]

    def associate(self, flow, event_list, pol_mask, event_mask):
        """flow.py:58-121 for a single flow scale.  NOTE: like the reference (:91), this shifts the
        timestamps of ``event_list`` IN PLACE by the number of passes so far."""
        ef = _iwe.gather_event_flow(flow, event_list, self.res)
        if self.events is None:
            self.events, self.ev_flow, self.pol_mask, self.event_mask = event_list, ef, pol_mask, event_mask
            self.flow_x, self.flow_y = flow[:, 0:1], flow[:, 1:2]
        else:
            event_list[:, :, 0:1] += self.passes
            self.events = torch.cat([self.events, event_list], dim=1)
            self.ev_flow = torch.cat([self.ev_flow, ef], dim=1)
            self.pol_mask = torch.cat([self.pol_mask, pol_mask], dim=1)
            self.event_mask = torch.cat([self.event_mask, event_mask], dim=1)
            self.flow_x = torch.cat([self.flow_x, flow[:, 0:1]], dim=1)
            self.flow_y = torch.cat([self.flow_y, flow[:, 1:2]], dim=1)
        self.passes += 1

    def _direction(self, tref, ts_weight, max_ts):
        img = _iwe.warp_images(self.events, self.ev_flow, self.pol_mask, tref, self.res, self.flow_scaling,
                               ts_weight=ts_weight)
        cnt_p, cnt_n, ts_p, ts_n = img[:, 0:1], img[:, 1:2], img[:, 2:3], img[:, 3:4]
        ts_p = ts_p / (cnt_p + 1e-9) / max_ts                                        # flow.py:214-217
        ts_n = ts_n / (cnt_n + 1e-9) / max_ts
        B = img.shape[0]
        loss = (ts_p.reshape(B, -1) ** 2).sum(1) + (ts_n.reshape(B, -1) ** 2).sum(1)  # :222
        if self.loss_scaling:
            # :224-227 sets the positive entries to 1 IN PLACE; zero entries stay `cnt_p + cnt_n` and keep
            # their gradient path (weight-0 corner contributions leak d loss / d nz through them).
            tot = cnt_p + cnt_n
            nz = torch.where(tot > 0, torch.ones_like(tot), tot).reshape(B, -1).sum(1)
            loss = loss / nz
        return loss.sum()

    def smoothness(self):
        fx, fy = self.flow_x, self.flow_y                                            # [B,T,H,W]

        def charb(a, b):
            return torch.sqrt((a + b) ** 2 + 1e-6)

        terms = [
            charb(fx[:, :, :, :-1] - fx[:, :, :, 1:], fy[:, :, :, :-1] - fy[:, :, :, 1:]),              # dx
            charb(fx[:, :, :-1, :] - fx[:, :, 1:, :], fy[:, :, :-1, :] - fy[:, :, 1:, :]),              # dy
            charb(fx[:, :, :-1, :-1] - fx[:, :, 1:, 1:], fy[:, :, :-1, :-1] - fy[:, :, 1:, 1:]),        # dr
            charb(fx[:, :, 1:, :-1] - fx[:, :, :-1, 1:], fy[:, :, 1:, :-1] - fy[:, :, :-1, 1:]),        # ur
            charb(fx[:, :-1] - fx[:, 1:], fy[:, :-1] - fy[:, 1:]),                                      # dt
        ]
        if self.smoothing_mask:
            m = self.event_mask
            masks = [m[:, :, :, :-1] * m[:, :, :, 1:], m[:, :, :-1, :] * m[:, :, 1:, :],
                     m[:, :, :-1, :-1] * m[:, :, 1:, 1:], m[:, :, 1:, :-1] * m[:, :, :-1, 1:],
                     m[:, :-1] * m[:, 1:]]
            terms = [a * b for a, b in zip(masks, terms)]
        s = terms[0].sum() + terms[1].sum() + terms[2].sum() + terms[3].sum() + terms[4].sum()   # :289-292
        return s / 5 / fx.shape[1]                                                               # :294-295

    def __call__(self):
        max_ts = self.passes                                                          # flow.py:179
        ts = self.events[:, :, 0:1]
        fw = self._direction(max_ts, ts, max_ts)                                      # :197-228
        bw = self._direction(0, max_ts - ts, max_ts)                                  # :230-261
        return fw + bw + self.weight * self.smoothness()                              # :298
