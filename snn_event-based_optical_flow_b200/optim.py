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
        # one-launch update (snnflow_dp_clip_adam): per-CTA launch counters, norm partials, grid-barrier counter
        L = _lib.lib()
        self._one_launch_ok = n <= int(L.snnflow_dp_clip_adam_max_n())
        ctas = int(L.snnflow_dp_clip_adam_ctas())
        self._dp_counter = torch.zeros(ctas, dtype=torch.int32, device=dev)
        self._dp_partials = torch.zeros(ctas, dtype=torch.float32, device=dev)
        self._dp_gridcnt = torch.zeros(1, dtype=torch.int32, device=dev)
        self._peers = None

    def set_lr(self, lr):
        self.hyper[0] = lr
        self.param_groups[0]["lr"] = lr

    @torch.no_grad()
    def step(self, grads=None, gate=None):
        """Clip (if ``max_norm``) and update.  ``grads``: tensors to use instead of ``p.grad`` (missing gradients count
        as zeros, like parameters torch's Adam skips while their moments are still zero).  ``gate``: optional int32
        device tensor; when it is non-zero at execution time the kernels skip the whole update (see snnflow.h)."""
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
                                          self.grad_norm.data_ptr(), None if gate is None else gate.data_ptr(),
                                          L.stream()), "snnflow_clip_adam")
        # the kernel wrote the parameters behind autograd's back: bump their version counters so that anything keyed on
        # them (the cells' packed-weight cache, spiking_submodules.py) sees the update
        torch.autograd.graph.increment_version(self.params)

    # ---- the whole update as ONE launch, optionally with the data-parallel gradient SUM in front -----------------------
    def attach_peers(self, group=None):
        """Data parallel: allocate this rank's symmetric gradient buffer (two slots, torch symmetric memory over NVLink)
        and exchange the peers' addresses.  Afterwards step_flat() sums the flat gradient over the ranks of `group` inside
        the same launch that clips and applies Adam."""
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        group = group if group is not None else dist.group.WORLD
        ctas = int(_lib.lib().snnflow_dp_clip_adam_ctas())
        slot = (self.n + ctas + 63) // 64 * 64
        dev = self.flat.device
        sym = symm_mem.empty(2 * slot, dtype=torch.float32, device=dev)
        hdl = symm_mem.rendezvous(sym, group)
        sym.zero_()
        self._peers = {"sym": sym, "hdl": hdl, "slot": slot, "rank": dist.get_rank(group), "world": dist.get_world_size(group),
                       "bufs": torch.tensor([int(x) for x in hdl.buffer_ptrs], dtype=torch.int64, device=dev),
                       "pads": torch.tensor([int(x) for x in hdl.signal_pad_ptrs], dtype=torch.int64, device=dev)}
        hdl.barrier()
        return self

    @torch.no_grad()
    def step_flat(self, gate=None, reduced=None):
        """Clip + Adam on the flat gradient buffer ``self.grad`` (filled directly by the caller, e.g. the window engine's
        backward) in ONE kernel launch (snnflow_dp_clip_adam); with attach_peers() the gradient is first summed over the
        ranks inside the same launch.  ``reduced``: optional flat tensor that receives the (summed) gradient."""
        L = _lib
        pr = self._peers
        L.check(L.lib().snnflow_dp_clip_adam(
            self.grad.data_ptr(), None if pr is None else pr["bufs"].data_ptr(), None if pr is None else pr["pads"].data_ptr(),
            self._dp_counter.data_ptr(), 0 if pr is None else pr["rank"], 1 if pr is None else pr["world"], self.n,
            0 if pr is None else pr["slot"], self.flat.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(),
            self.hyper.data_ptr(), self.step_count.data_ptr(), self.bias_state.data_ptr(), self._dp_partials.data_ptr(),
            self._dp_gridcnt.data_ptr(), self.grad_norm.data_ptr(), None if gate is None else gate.data_ptr(),
            None if reduced is None else reduced.data_ptr(), L.stream()), "snnflow_dp_clip_adam")
        torch.autograd.graph.increment_version(self.params)

    def grad_view(self, param):
        """The slice of the flat gradient buffer that belongs to `param` (shaped like it)."""
        for p, v in zip(self.params, self.grad_views):
            if p is param:
                return v
        raise KeyError("parameter is not managed by this optimizer")

    def zero_grad(self, set_to_none=True):
        for p in self.params:
            if set_to_none:
                p.grad = None
            elif p.grad is not None:
                p.grad.zero_()

    def state_dict(self):
        """The layout of ``torch.optim.Adam.state_dict()`` - what the reference's checkpoints store as
        ``optimizer_state_dict`` (train_flow.py save_data) - sliced out of the flat moment buffers, so a run can be resumed
        from / handed back to the reference's optimizer."""
        t = self.step_count.to(torch.float32).reshape(()).cpu()
        state = {}
        if int(t) > 0:
            for i, (p, o) in enumerate(zip(self.params, self.offsets)):
                n = p.numel()
                state[i] = {"step": t.clone(), "exp_avg": self.exp_avg[o:o + n].view_as(p).clone(),
                            "exp_avg_sq": self.exp_avg_sq[o:o + n].view_as(p).clone()}
        lr, b1, b2, eps, max_norm = [float(v) for v in self.hyper.tolist()]
        group = {"lr": lr, "betas": (b1, b2), "eps": eps, "weight_decay": 0, "amsgrad": False, "maximize": False,
                 "foreach": None, "capturable": False, "differentiable": False, "fused": None, "decoupled_weight_decay": False,
                 "params": list(range(len(self.params))), "max_norm": max_norm}
        return {"state": state, "param_groups": [group]}

    @torch.no_grad()
    def load_state_dict(self, sd):
        """Accepts ``torch.optim.Adam.state_dict()`` (and this class's own output)."""
        group = sd["param_groups"][0]
        if group.get("amsgrad") or group.get("weight_decay"):
            raise ValueError("FusedClipAdam implements Adam without amsgrad / weight decay")
        if len(group["params"]) != len(self.params):
            raise ValueError(f"optimizer state has {len(group['params'])} parameters, this optimizer {len(self.params)}")
        b1, b2 = group["betas"]
        self.hyper[:4] = torch.tensor([group["lr"], b1, b2, group["eps"]], dtype=torch.float32)
        if "max_norm" in group:
            self.hyper[4] = float(group["max_norm"])
        self.param_groups[0].update(lr=group["lr"], betas=(b1, b2), eps=group["eps"])
        self.exp_avg.zero_()
        self.exp_avg_sq.zero_()
        steps = set()
        for i, (p, o) in enumerate(zip(self.params, self.offsets)):
            st = sd["state"].get(i, sd["state"].get(str(i)))
            if st is None:
                continue
            n = p.numel()
            self.exp_avg[o:o + n].copy_(st["exp_avg"].reshape(-1))
            self.exp_avg_sq[o:o + n].copy_(st["exp_avg_sq"].reshape(-1))
            steps.add(int(st["step"]))
        if len(steps) > 1:
            raise ValueError(f"FusedClipAdam keeps ONE step counter; the state holds {sorted(steps)}")
        t = steps.pop() if steps else 0
        self.step_count.fill_(t)
        # the kernel carries beta^t as running products in double precision (re-initialised by itself at step 0)
        p1, p2 = float(b1) ** t, float(b2) ** t
        self.bias_state.copy_(torch.tensor([p1, p2, group["lr"] / (1.0 - p1) if t else 0.0,
                                            1.0 / (1.0 - p2) ** 0.5 if t else 0.0], dtype=torch.float64))
