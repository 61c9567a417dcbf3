"""Training window of train_flow.py:232-279 on the GPU, plus the data-parallel gradient exchange.

One optimizer step = T time bins of (network forward + event/flow association), the contrast loss, BPTT,
(SUM all-reduce of the flat gradient across ranks), gradient-norm clipping, Adam, state detach and loss
reset - the exact order of the reference loop.  The loss is a SUM over the batch (loss/flow.py:228,261,291),
so data-parallel ranks all-reduce with SUM and clip AFTER the reduction; that reproduces the single-process
global-batch step.
"""
import torch
import torch.distributed as dist


class FlatGradAllReduce:
    """One NCCL (or gloo) SUM all-reduce of all gradients as a single flat fp32 buffer (~300 KB at C=32)."""

    def __init__(self, params, group=None):
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        n = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(n, dtype=torch.float32, device=self.params[0].device)
        self.views, o = [], 0
        for p in self.params:
            self.views.append(self.flat[o:o + p.numel()].view_as(p))
            o += p.numel()

    def __call__(self):
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(self.group) == 1:
            return
        grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in self.params]
        torch._foreach_copy_(self.views, grads)
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
        for p, v in zip(self.params, self.views):
            if p.grad is None:
                p.grad = v.clone()
            else:
                p.grad.copy_(v)


class TrainWindow:
    """Runs optimizer steps on windows of T bins.  `batch` is the loader dict of the reference
    (dataloader/base.py:261-278) stacked over the T bins of the window:
        event_cnt [T,B,2,H,W], event_list [T,B,N,4], event_list_pol_mask [T,B,N,2], event_mask [T,B,1,H,W]."""

    def __init__(self, model, loss_fn, optimizer, clip_grad=1.0, group=None):
        self.model, self.loss_fn, self.opt, self.clip = model, loss_fn, optimizer, clip_grad
        self.reducer = FlatGradAllReduce(model.parameters(), group)

    def step(self, batch, use_window=True):
        T = batch["event_cnt"].shape[0]
        flows = self.model.forward_window(batch["event_cnt"]) if (use_window and hasattr(self.model, "forward_window")) else None
        for t in range(T):
            flow = flows[t] if flows is not None else self.model(None, batch["event_cnt"][t])["flow"][0]
            self.loss_fn.event_flow_association([flow], batch["event_list"][t], batch["event_list_pol_mask"][t],
                                                batch["event_mask"][t])
        loss = self.loss_fn()
        loss.backward()
        self.reducer()
        if self.clip is not None:
            torch.nn.utils.clip_grad_norm_(self.model.parameters(), self.clip)
        self.opt.step()
        self.opt.zero_grad(set_to_none=True)
        self.model.detach_states()
        self.loss_fn.reset()
        return loss.detach()
