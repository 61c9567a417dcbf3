"""Gradient clipping + Adam of the training step (train_flow.py:264-271) as one C call over a flat parameter buffer.

``FusedClipAdam(params, lr, betas, eps, max_norm)`` does what
``torch.nn.utils.clip_grad.clip_grad_norm_(params, max_norm); torch.optim.Adam(params, lr).step()`` does (Adam without
amsgrad / weight decay - the reference builds ``eval(config["optimizer"]["name"])(model.parameters(), lr=...)``,
train_flow.py:82), in two kernel launches (``snnflow_clip_adam``) instead of ~20.  The parameters are re-pointed at
slices of ONE flat fp32 buffer (``p.data`` becomes a view, values preserved), so the update is a single streaming pass;
the step counter and the hyper-parameters live on the device, which makes the step replayable inside a CUDA graph.
"""
import torch

from . import _lib


class FusedClipAdam:
    fused_clip = True   # train.TrainWindow: this optimizer clips by itself

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, max_norm=None):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("FusedClipAdam: no trainable parameters")
        dev = self.params[0].device
        if dev.type != "cuda" or any(p.device != dev or p.dtype != torch.float32 for p in self.params):
            raise _lib.SnnflowError("FusedClipAdam needs fp32 parameters on one CUDA device (no CPU fallback)")
        self.offsets, n = [], 0
        for p in self.params:
            self.offsets.append(n)
            n += (p.numel() + 63) // 64 * 64          # 256-byte aligned slices
        self.n = n
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=dev)
        self.grad_views = []
        with torch.no_grad():
            for p, o in zip(self.params, self.offsets):
                view = self.flat[o:o + p.numel()].view_as(p)
                view.copy_(p)
                p.data = view                          # the module's parameter now lives in the flat buffer
                self.grad_views.append(self.grad[o:o + p.numel()].view_as(p))
        self.hyper = torch.tensor([lr, betas[0], betas[1], eps, max_norm if max_norm else 0.0], dtype=torch.float32, device=dev)
        self.step_count = torch.zeros(1, dtype=torch.int64, device=dev)
        self.bias_state = torch.zeros(4, dtype=torch.float64, device=dev)   # beta^t running products, step size (kernel-owned)
        self.partials = torch.zeros(_lib.lib().snnflow_clip_adam_partials(n), dtype=torch.float32, device=dev)
        self.grad_norm = torch.zeros(1, dtype=torch.float32, device=dev)   # total norm of the last step (before clipping)
        self.param_groups = [{"params": self.params, "lr": lr, "betas": betas, "eps": eps}]

    def set_lr(self, lr):
        self.hyper[0] = lr
        self.param_groups[0]["lr"] = lr

    @torch.no_grad()
    def step(self, grads=None):
        """Clip (if ``max_norm``) and update.  ``grads``: tensors to use instead of ``p.grad`` (missing gradients count
        as zeros, like parameters torch's Adam skips while their moments are still zero)."""
        if grads is None:
            grads = [p.grad for p in self.params]
        have = [(v, g) for v, g in zip(self.grad_views, grads) if g is not None]
        if len(have) != len(grads):
            self.grad.zero_()
        torch._foreach_copy_([v for v, _ in have], [g for _, g in have])
        L = _lib
        L.check(L.lib().snnflow_clip_adam(self.flat.data_ptr(), self.grad.data_ptr(), self.exp_avg.data_ptr(),
                                          self.exp_avg_sq.data_ptr(), self.n, self.hyper.data_ptr(),
                                          self.step_count.data_ptr(), self.bias_state.data_ptr(), self.partials.data_ptr(),
                                          self.grad_norm.data_ptr(),
                                          L.stream()), "snnflow_clip_adam")
        # the kernel wrote the parameters behind autograd's back: bump their version counters so that anything keyed on
        # them (the cells' packed-weight cache, spiking_submodules.py) sees the update
        torch.autograd.graph.increment_version(self.params)

    def zero_grad(self, set_to_none=True):
        for p in self.params:
            if set_to_none:
                p.grad = None
            elif p.grad is not None:
                p.grad.zero_()

    def state_dict(self):
        return {"step": self.step_count.clone(), "exp_avg": self.exp_avg.clone(), "exp_avg_sq": self.exp_avg_sq.clone(),
                "hyper": self.hyper.clone(), "bias_state": self.bias_state.clone()}

    def load_state_dict(self, sd):
        self.step_count.copy_(sd["step"]); self.exp_avg.copy_(sd["exp_avg"]); self.exp_avg_sq.copy_(sd["exp_avg_sq"])
        self.hyper.copy_(sd["hyper"]); self.bias_state.copy_(sd["bias_state"])
