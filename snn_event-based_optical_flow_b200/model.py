"""LIFFireNet / LIFFireFlowNet: host-side mirror of models/model.py:29-207 and :387-554.

On the GPU box the reference tree does not exist, so the benchmark and the GPU parity tests need
the network container itself; this mirror keeps the reference's constructor (a ``unet_kwargs`` dict),
the class-attribute seam ``head_neuron / ff_neuron / rec_neuron`` (models/model.py:37-39), the
``states`` / ``detach_states`` / ``reset_states`` plumbing (:109-130), the layer order and the
``{"flow": [flow], "activity": ...}`` return, and the state_dict keys the reference's networks have WHEN THEY ARE
BUILT ON THE ``ConvLIF`` / ``ConvLIFRecurrent`` CELLS of models/spiking_submodules.py (``head.ff.weight``,
``G1.rec.weight``, ``R1a.leak``, ``R1a.thresh`` ...).  In a checkout that has the reference, the same cells drop into
the reference's own classes (tests/test_gpu_reference_seam.py runs exactly this on the GPU):

    class Net(models.model.LIFFireNet):
        head_neuron = ff_neuron = snnflow.ConvLIF; rec_neuron = snnflow.ConvLIFRecurrent

The reference's ``LIFFireNet`` as shipped wires the ``SNNtorch_ConvLIF`` / ``SNNtorch_ConvLIFRecurrent`` cells
(models/model.py:37-39: snntorch ``Leaky`` + ``BatchNorm2d``, keys ``lif.beta``, ``lif.threshold``, ``bn.*``), a
different neuron equation: that network is ``SNNtorchLIFFireNet`` below (cells in snntorch_submodules.py, parity unpinned
because snntorch is absent offline).  A checkpoint of one kind does not load into the other: ``load_state_dict`` says so
explicitly instead of reporting a wall of missing keys.
"""
import torch
import torch.nn as nn

from .spiking_submodules import ConvLIF, ConvLIFRecurrent
from .submodules import ConvLayer


class LIFFireNet(nn.Module):
    head_neuron = ConvLIF
    ff_neuron = ConvLIF
    rec_neuron = ConvLIFRecurrent
    residual = False
    num_recurrent_units = 7
    w_scale_pred = 0.01

    def __init__(self, unet_kwargs):
        super().__init__()
        self.num_bins = unet_kwargs["num_bins"]
        self.encoding = unet_kwargs["encoding"]
        self.norm_input = unet_kwargs.get("norm_input", False)
        self.mask = unet_kwargs.get("mask_output", False)
        C = unet_kwargs["base_num_channels"]
        k = unet_kwargs["kernel_size"]
        q = unet_kwargs.get("quantization", {})
        # model.py never forwards `spiking_neuron` / `activations` to the cells;
        # `neuron_kwargs` is this mirror's explicit way to set them (leak/thresh statistics etc.).
        nk = dict(unet_kwargs.get("neuron_kwargs", {}))
        mk = lambda cls, cin: cls(cin, C, k, quantization_config=q, **nk)
        self.head = mk(self.head_neuron, self.num_bins)
        self.G1 = mk(self.rec_neuron, C)
        self.R1a = mk(self.ff_neuron, C)
        self.R1b = mk(self.ff_neuron, C)
        self.G2 = mk(self.rec_neuron, C)
        self.R2a = mk(self.ff_neuron, C)
        self.R2b = mk(self.ff_neuron, C)
        self.pred = ConvLayer(C, out_channels=2, kernel_size=1, activation="tanh", w_scale=self.w_scale_pred,
                              quantization_config=q)
        self.reset_states()

    # --- state plumbing (models/model.py:109-130) ---
    # `_states` is the reference's list of 7 tensors [2,B,C,H,W].  Under no_grad the per-bin forward() keeps the state inside
    # the window engine's arena in its own layout (engine.WindowRunner.stream_forward) and this list is materialised only when
    # somebody reads it; assigning it (reset_states, detach_states, a caller-provided state) makes the list the truth again.
    stream_forward = True    # False: every per-bin forward() goes through the cells (fp32 NCHW state each call)

    @property
    def _states(self):
        r = self.__dict__.get("_window_runner")
        if r is not None and r.stream_live:
            self.__dict__["_states_list"] = r.stream_export()
            r.stream_live = False          # the list is current (the arena stays valid until the list is assigned again)
            self.__dict__["_states_from_stream"] = True
        return self.__dict__.get("_states_list")

    @_states.setter
    def _states(self, value):
        self.__dict__["_states_list"] = value
        self.__dict__["_states_from_stream"] = False
        r = self.__dict__.get("_window_runner")
        if r is not None:
            r.stream_live = False

    @property
    def states(self):
        return [None if s is None else s.clone() for s in self._states]   # model_util.py:95-101

    @states.setter
    def states(self, states):
        self._states = states

    def detach_states(self):
        self._states = [None if s is None else s.detach() for s in self._states]

    def reset_states(self):
        self._states = [None] * self.num_recurrent_units

    def init_cropping(self, width, height):
        pass

    def load_state_dict(self, state_dict, *args, **kwargs):
        foreign = [k for k in state_dict if ".lif." in k or ".bn." in k or ".tebn." in k or ".mpbn." in k]
        if foreign and not hasattr(self.head, "lif"):
            raise RuntimeError(
                "snnflow LIFFireNet: this state_dict comes from the reference's SNNtorch_ConvLIF cells (keys such as "
                f"{foreign[0]!r}: snntorch Leaky + BatchNorm2d, models/SNNtorch_spiking_submodules.py).  This network "
                "implements the ConvLIF / ConvLIFRecurrent cells of models/spiking_submodules.py (keys ff.weight, "
                "rec.weight, leak, thresh); the two neuron models are not weight-compatible.  SNNtorchLIFFireNet is the "
                "network built on the SNNtorch_* cells.")
        return super().load_state_dict(state_dict, *args, **kwargs)

    def forward_window(self, event_cnt_window):
        """T bins at once: [T,B,num_bins,H,W] -> flow [T,B,2,H,W], equivalent to T calls of forward() (same states,
        same gradients) but executed as one C call per direction (engine.WindowRunner)."""
        from .engine import WindowRunner
        runner = getattr(self, "_window_runner", None)
        if runner is None:
            runner = WindowRunner(self)
            object.__setattr__(self, "_window_runner", runner)
        if not runner.supported():
            return torch.stack([self.forward(None, event_cnt_window[t])["flow"][0] for t in range(event_cnt_window.shape[0])])
        return runner(event_cnt_window)

    # ---- CUDA-graphed per-bin inference (opt-in) ------------------------------------------------------------------------
    def graph_forward(self, enabled=True):
        """Opt in to replaying the per-bin ``forward()`` as a CUDA graph whenever autograd is off (the reference's streaming
        eval loop, eval_flow.py:220: one ``model()`` call per frame).  The call sequence, the arguments, the returned dict
        and the state plumbing are unchanged; what changes is ownership: the flow map and the tensors in ``_states`` returned
        by a graphed call are static buffers that the next graphed call overwrites (``.states``, which deep-clones, is safe
        to keep).  At batch 1 a frame is launch-latency bound: the graph removes the host overhead.  Where the streaming
        mode applies (``stream_forward``: counts encoding, no residual) the graph replays the streamed bin - state inside the
        engine, nothing copied; elsewhere it replays the 7 cells + flow head and ends with a copy of the new state into the
        static buffers (16 B per neuron), which pays at small batches, not at batch 16 / 256x256."""
        object.__setattr__(self, "_graphs", {} if enabled else None)
        return self

    def _forward_graphed(self, x):
        if self.stream_forward and self.encoding == "cnt" and self.num_bins == 2 and not self.residual:
            flow = self._forward_streamed(x, graph=True)      # the streamed bin as a graph replay (engine.stream_forward)
            if flow is not None:
                return flow
        key = (tuple(x.shape), x.dtype, x.device)
        g = self._graphs.get(key)
        if g is None:
            side = torch.cuda.Stream()
            saved = self._states
            sx = torch.zeros_like(x)
            with torch.no_grad():
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):          # warm-up outside the capture (module loading, weight packing)
                    self._states = [None] * self.num_recurrent_units
                    self._forward_eager(sx)
                    shapes = [tuple(s.shape) for s in self._states]
                torch.cuda.current_stream().wait_stream(side)
                static = [torch.zeros(sh, dtype=torch.float32, device=x.device) for sh in shapes]
                self._states = list(static)
                cg = torch.cuda.CUDAGraph()
                with torch.cuda.graph(cg):
                    flow = self._forward_eager(sx)
                    for dst, src in zip(static, self._states):   # the new state goes back into the static set
                        dst.copy_(src)
            self._states = saved
            g = self._graphs[key] = dict(x=sx, states=static, graph=cg, flow=flow)
        st = self._states
        for i, buf in enumerate(g["states"]):          # states that are not the graph's own buffers are copied in
            if st[i] is None:
                buf.zero_()
            elif st[i].data_ptr() != buf.data_ptr():
                buf.copy_(st[i])
        g["x"].copy_(x)
        g["graph"].replay()
        self._states = list(g["states"])
        return g["flow"]

    def _forward_streamed(self, x, graph=False):
        """Per-bin inference on the window engine with the state kept in the engine's layout between calls."""
        from .engine import WindowRunner
        r = self.__dict__.get("_window_runner")
        if r is None:
            r = WindowRunner(self)
            object.__setattr__(self, "_window_runner", r)
        if not r.stream_ok(x):
            return None
        if self.__dict__.get("_states_from_stream") and not r.stream_live and r._stream is not None and r._stream["shape"] == tuple(x.shape):
            # the list was only READ since the last streamed call (states getter): the arena is still current
            r.stream_live = True
        if r.stream_live and r._stream is not None and r._stream["shape"] != tuple(x.shape):
            self._states      # another input shape: the streamed state becomes the list (and fails the shape check below, like
                              # the reference does when batch size / resolution change without reset_states())
        states = self.__dict__.get("_states_list")
        flow = r.stream_forward(x, states, graph=graph)
        self.__dict__["_states_from_stream"] = False
        return flow

    def forward(self, event_voxel=None, event_cnt=None, log=False, return_dict=True):
        if (self.stream_forward and getattr(self, "_graphs", None) is None and not torch.is_grad_enabled()
                and not (isinstance(log, bool) and log) and not self.norm_input and self.encoding == "cnt" and self.num_bins == 2
                and event_cnt is not None and event_cnt.is_cuda and not self.residual):
            flow = self._forward_streamed(event_cnt)
            if flow is not None:
                return {"flow": [flow], "activity": None} if return_dict else flow
        if getattr(self, "_graphs", None) is not None and not torch.is_grad_enabled() and not (isinstance(log, bool) and log):
            x = event_voxel if self.encoding == "voxel" else event_cnt
            if x is not None and x.is_cuda and not self.norm_input and (self.encoding == "voxel" or self.num_bins == 2):
                flow = self._forward_graphed(x)
                return {"flow": [flow], "activity": None} if return_dict else flow
        return self._forward_impl(event_voxel, event_cnt, log, return_dict)

    def _forward_eager(self, x):
        return self._forward_impl(x if self.encoding == "voxel" else None, x if self.encoding != "voxel" else None, False, False)

    def _forward_impl(self, event_voxel=None, event_cnt=None, log=False, return_dict=True):
        if self.encoding == "voxel":
            x = event_voxel
        elif self.encoding == "cnt" and self.num_bins == 2:
            x = event_cnt
        else:
            raise AttributeError("Model error: Incorrect input encoding.")
        if self.norm_input:                                               # model.py:165-170
            nz = x != 0
            mean, std = x[nz].mean(), x[nz].std()
            x = x.clone()
            x[nz] = (x[nz] - mean) / std
        s = self._states
        x1, s[0] = self.head(x, s[0])
        x2, s[1] = self.G1(x1, s[1])
        x3, s[2] = self.R1a(x2, s[2])
        x4, s[3] = self.R1b(x3, s[3], residual=x2 if self.residual else 0)
        x5, s[4] = self.G2(x4, s[4])
        x6, s[5] = self.R2a(x5, s[5])
        x7, s[6] = self.R2b(x6, s[6], residual=x5 if self.residual else 0)
        flow = self.pred(x7)
        self._states = s          # (through the setter: the list is the truth, a streamed state in the arena is stale)
        if not return_dict:
            return flow
        activity = None
        if isinstance(log, bool) and log:
            names = ["0:input", "1:head", "2:G1", "3:R1a", "4:R1b", "5:G2", "6:R2a", "7:R2b", "8:pred"]
            acts = torch.stack([t.detach().ne(0).float().mean() for t in (x, x1, x2, x3, x4, x5, x6, x7, flow)])
            activity = dict(zip(names, acts.tolist()))                    # one host sync instead of nine
        return {"flow": [flow], "activity": activity}


class SNNtorchLIFFireNet(LIFFireNet):
    """The reference's LIFFireNet AS SHIPPED (models/model.py:37-39): SNNtorch_ConvLIF / SNNtorch_ConvLIFRecurrent cells
    (snntorch Leaky + BatchNorm2d).  Runs bin by bin through the cells (snntorch_submodules.py); the layer-major window engine
    covers the ConvLIF cells only.  PARITY UNPINNED (snntorch is absent offline)."""

    from .snntorch_submodules import SNNtorch_ConvLIF as head_neuron, SNNtorch_ConvLIF as ff_neuron, \
        SNNtorch_ConvLIFRecurrent as rec_neuron

    def forward_window(self, event_cnt_window):
        return torch.stack([self.forward(None, event_cnt_window[t])["flow"][0] for t in range(event_cnt_window.shape[0])])


class LIFFireFlowNet(LIFFireNet):
    """Feed-forward variant: G1/G2 are plain ConvLIF cells (models/model.py:387-395)."""

    rec_neuron = ConvLIF
