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

    def active(self):
        return dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1

    def __call__(self, grads=None):
        """SUM-reduce the gradients in place (`grads`: tensors to use instead of p.grad, e.g. graph-static ones)."""
        if not self.active():
            return
        if grads is None:
            for p in self.params:
                if p.grad is None:
                    p.grad = torch.zeros_like(p)
            grads = [p.grad for p in self.params]
        torch._foreach_copy_(self.views, grads)
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
        torch._foreach_copy_(grads, self.views)


class PeerGradAllReduce:
    """The same SUM all-reduce as one peer-memory kernel per rank (snnflow_dp_allreduce_sum): every rank reads every
    peer's gradient buffer over NVLink between two flag barriers.  No NCCL call, no stream hand-over: it is captured
    inside the step's CUDA graph.  Needs torch symmetric memory (one NVLink domain); `available()` tells."""

    graph_safe = True

    def __init__(self, params, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        from . import _lib
        self._lib = _lib
        self.params = [p for p in params if p.requires_grad]
        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        dev = self.params[0].device
        n = sum(p.numel() for p in self.params)
        self.n = n
        self.sym = symm_mem.empty(n, dtype=torch.float32, device=dev)     # this rank's contribution, readable by peers
        self.hdl = symm_mem.rendezvous(self.sym, self.group)
        self.out = torch.zeros(n, dtype=torch.float32, device=dev)
        self.views_in, self.views_out, o = [], [], 0
        for p in self.params:
            self.views_in.append(self.sym[o:o + p.numel()].view_as(p))
            self.views_out.append(self.out[o:o + p.numel()].view_as(p))
            o += p.numel()
        self.bufs = torch.tensor([int(x) for x in self.hdl.buffer_ptrs], dtype=torch.int64, device=dev)
        self.pads = torch.tensor([int(x) for x in self.hdl.signal_pad_ptrs], dtype=torch.int64, device=dev)
        self.counter = torch.zeros(_lib.lib().snnflow_dp_allreduce_ctas(), dtype=torch.int32, device=dev)
        self.hdl.barrier()

    @staticmethod
    def available():
        try:
            import torch.distributed._symmetric_memory  # noqa: F401
            return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1 and dist.get_backend() == "nccl"
        except Exception:  # noqa: BLE001
            return False

    def active(self):
        return True

    def __call__(self, grads=None):
        if grads is None:
            for p in self.params:
                if p.grad is None:
                    p.grad = torch.zeros_like(p)
            grads = [p.grad for p in self.params]
        L = self._lib
        torch._foreach_copy_(self.views_in, grads)
        L.check(L.lib().snnflow_dp_allreduce_sum(self.bufs.data_ptr(), self.pads.data_ptr(), self.out.data_ptr(),
                                                 self.counter.data_ptr(), self.rank, self.world, self.n, L.stream()),
                "snnflow_dp_allreduce_sum")
        torch._foreach_copy_(grads, self.views_out)


class TrainWindow:
    """Runs optimizer steps on windows of T bins.  `batch` is the loader dict of the reference
    (dataloader/base.py:261-278) stacked over the T bins of the window:
        event_cnt [T,B,2,H,W], event_list [T,B,N,4], event_list_pol_mask [T,B,N,2], event_mask [T,B,1,H,W]."""

    def __init__(self, model, loss_fn, optimizer, clip_grad=1.0, group=None, peer_allreduce=False):
        self.model, self.loss_fn, self.opt, self.clip = model, loss_fn, optimizer, clip_grad
        if getattr(optimizer, "fused_clip", False):
            have = float(optimizer.hyper[4])
            if (clip_grad or 0.0) != have:
                raise ValueError(f"FusedClipAdam clips at max_norm={have}, TrainWindow was given clip_grad={clip_grad}")
        self.reducer = FlatGradAllReduce(model.parameters(), group)
        if peer_allreduce and PeerGradAllReduce.available():
            try:   # NVLink peer memory: the exchange becomes one kernel inside the step graph
                self.reducer = PeerGradAllReduce(model.parameters(), group)
            except Exception as e:  # noqa: BLE001 - no symmetric memory on this system: keep the NCCL all-reduce
                import warnings
                warnings.warn(f"peer-memory all-reduce unavailable ({e}); using the NCCL all-reduce")
        self._graph = None
        self._direct_cache = {}
        self._peer_direct = False
        if peer_allreduce and isinstance(self.reducer, PeerGradAllReduce) and hasattr(optimizer, "attach_peers"):
            try:   # the direct step sums the gradient inside the optimizer's own launch (optim.FusedClipAdam.step_flat)
                optimizer.attach_peers(group)
                self._peer_direct = True
            except Exception as e:  # noqa: BLE001
                import warnings
                warnings.warn(f"one-launch data-parallel update unavailable ({e}); using the all-reduce kernel + fused Adam")
        self._params = [p for p in model.parameters()]
        self.fused_loss = True   # False: per-bin event_flow_association + EventWarping.forward (the reference's call pattern)

    # ---- the step in two halves (train_flow.py:232-262 and :265-279) -------------------------------------
    def _forward_backward(self, batch, use_window=True):
        T = batch["event_cnt"].shape[0]
        flows = self.model.forward_window(batch["event_cnt"]) if (use_window and hasattr(self.model, "forward_window")) else None
        if flows is not None and self.fused_loss and hasattr(self.loss_fn, "window_loss") and getattr(self.loss_fn, "fused_window_loss_ok", True):
            # the whole window's association + contrast loss + its gradient as one fused call (flow_loss.window_loss)
            loss = self.loss_fn.window_loss(flows, batch["event_list"], batch["event_list_pol_mask"], batch["event_mask"])
            loss.backward()
            return loss.detach()
        for t in range(T):
            flow = flows[t] if flows is not None else self.model(None, batch["event_cnt"][t])["flow"][0]
            self.loss_fn.event_flow_association([flow], batch["event_list"][t], batch["event_list_pol_mask"][t],
                                                batch["event_mask"][t])
        if getattr(self.loss_fn, "overwrite_intermediate", False):       # train_flow.py:244-246
            self.loss_fn.overwrite_intermediate_flow([flow])
        loss = self.loss_fn()
        loss.backward()
        return loss.detach()

    def _runner(self):
        return getattr(self.model, "_window_runner", None)

    def _update(self):
        if getattr(self.opt, "fused_clip", False):
            # optim.FusedClipAdam: clipping and Adam in one C call (max_norm given to the optimizer), gated by the window
            # engine's sticky "input was not bf16-exact" flag: an update computed from rounded inputs is never applied
            r = self._runner()
            self.opt.step(gate=None if r is None else r.input_flag)
        else:
            if self.clip is not None:
                torch.nn.utils.clip_grad_norm_(self.model.parameters(), self.clip)
            self.opt.step()
        self.opt.zero_grad(set_to_none=True)
        self.model.detach_states()
        self.loss_fn.reset()

    def step(self, batch, use_window=True):
        if use_window and self.direct_ok(batch):
            return self.step_direct(batch)
        loss = self._forward_backward(batch, use_window)
        self.reducer()
        self._update()
        return loss

    # ---- the direct step: no autograd, no gradient copies, one launch for all-reduce + clip + Adam ---------------------
    direct = True   # False: always go through torch.autograd (p.grad is populated, any optimizer works)

    def direct_ok(self, batch):
        """The direct step needs: the layer-major window engine at the network's own width, the fused window loss, a
        FusedClipAdam over exactly the network's parameters and - data parallel - the peer-memory exchange."""
        if not self.direct or not self.fused_loss or not getattr(self.loss_fn, "fused_window_loss_ok", True) or not getattr(self.opt, "fused_clip", False) or not getattr(self.opt, "_one_launch_ok", False):
            return False
        if not hasattr(self.model, "forward_window") or not hasattr(self.loss_fn, "window_loss_and_grad") or not torch.is_grad_enabled():
            return False
        if self.reducer.active() and not self._peer_direct:
            return False
        from .engine import WindowRunner
        r = self._runner()
        if r is None:
            r = WindowRunner(self.model)
            object.__setattr__(self.model, "_window_runner", r)
        key = tuple(batch["event_cnt"].shape)
        ok = self._direct_cache.get(key)
        if ok is None:
            ok = self._direct_cache[key] = bool(r.direct_ok(batch["event_cnt"], self.opt))
        return ok

    def step_direct(self, batch):
        """train_flow.py:232-279 as a fixed sequence of C calls: window forward (T bins), fused contrast loss with its flow
        gradient, window BPTT writing every parameter gradient into the optimizer's flat buffer, then ONE kernel that sums
        the gradient over the ranks (peer memory), clips and applies Adam.  p.grad is not populated."""
        r = self._runner()
        flows = r.direct_forward(batch["event_cnt"])
        loss, g_flow = self.loss_fn.window_loss_and_grad(flows, batch["event_list"], batch["event_list_pol_mask"], batch["event_mask"])
        r.direct_backward(g_flow, self.opt)
        self.opt.step_flat(gate=r.input_flag)
        self.model.detach_states()
        self.loss_fn.reset()
        return loss

    # ---- whole-step CUDA graph -----------------------------------------------------------------------------
    def capture(self, example_batch, warmup=3):
        """Capture one optimizer step in CUDA graphs.  Every buffer the step touches has a fixed address (window
        arena, workspace, static input copies), so ``step_graphed`` only copies the new window into the static inputs
        and replays.  Single process: ONE graph (window forward, association, loss, BPTT, clip, Adam: ~600 kernels, one
        launch).  Data parallel: two graphs around the one eager NCCL all-reduce of the flat gradient - [forward ..
        BPTT] | all-reduce | [clip, Adam] - so no collective is ever captured.  The optimizer must be capturable
        (torch.optim.Adam(..., capturable=True)).  Returns the snnflow kernel launches inside one step."""
        from . import _lib
        self._static = {k: torch.empty_like(v) for k, v in example_batch.items()}
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._load(example_batch)
                self.step(self._static)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self._load(example_batch)
        n0 = _lib.launch_count()
        self._graph = torch.cuda.CUDAGraph()
        if not self.reducer.active() or getattr(self.reducer, "graph_safe", False):
            self._graph2 = None
            with torch.cuda.graph(self._graph):
                self._static_loss = self.step(self._static)
        else:
            with torch.cuda.graph(self._graph):
                self._static_loss = self._forward_backward(self._static)
            self._grads = [p.grad for p in self.reducer.params]   # static addresses inside the graph's memory pool
            self.reducer(self._grads)
            self._graph2 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._graph2, pool=self._graph.pool()):
                self._update()
        self.launches_per_step = _lib.launch_count() - n0
        return self.launches_per_step

    def _load(self, batch):
        for k, v in self._static.items():
            v.copy_(batch[k], non_blocking=True)

    def prefetch(self, batch):
        """Start copying the NEXT window (pinned host memory or device) into a staging buffer on a side stream; the copy
        overlaps the step that is running.  `step_graphed(batch)` with the same object then only moves it device to device."""
        if getattr(self, "_staging", None) is None:
            self._staging = {k: torch.empty_like(v) for k, v in self._static.items()}
            self._copy_stream = torch.cuda.Stream()
            self._staging_free = None
        cs = self._copy_stream
        if self._staging_free is not None:
            cs.wait_event(self._staging_free)          # the previous window has left the staging buffer
        with torch.cuda.stream(cs):
            for k, v in self._staging.items():
                v.copy_(batch[k], non_blocking=True)
            self._staged = (batch, cs.record_event())

    def step_graphed(self, batch):
        """Same as step() after capture(): `batch` may live on the device or in pinned host memory."""
        staged = getattr(self, "_staged", None)
        if staged is not None and staged[0] is batch:
            torch.cuda.current_stream().wait_event(staged[1])
            for k, v in self._static.items():
                v.copy_(self._staging[k], non_blocking=True)
            self._staging_free = torch.cuda.current_stream().record_event()
            self._staged = None
        else:
            self._load(batch)
        self._graph.replay()
        if self._graph2 is not None:
            self.reducer(self._grads)
            self._graph2.replay()
        # The replay rewrote the parameters in place without passing through autograd: bump their version counters so
        # that everything keyed on them - the cells' packed tensor-core weights (spiking_submodules._packed_weights) -
        # is rebuilt by the next eager forward() instead of running the weights of an earlier step.
        torch.autograd.graph.increment_version(self._params)
        self._replays = getattr(self, "_replays", 0) + 1
        r = self._runner()
        if r is not None and self._replays % r.validate_every == 0:
            # graph replays bypass the engine's per-window input check: poll this runner's sticky flag (the fused
            # optimizer skipped every update since the flag was raised)
            r.check_input_flag()
        return self._static_loss
