"""ConvLIF / ConvLIFRecurrent on the B200: host-side mirror of models/spiking_submodules.py.

Same constructor arguments, parameter / buffer names (``ff.weight``, ``rec.weight``, ``leak``,
``thresh``, ``act_width``), state layout (``stack([v, z])``, shape [2,B,C,H,W]) and return values as
the reference cells (models/spiking_submodules.py:29-151, :154-300), so ``LIFFireNet`` /
``LIFFireFlowNet`` (models/model.py:37-39, :393-395) run on them by assigning the class attributes
``head_neuron / ff_neuron / rec_neuron``.  The arithmetic is one fused CUDA kernel per layer-step
(``snnflow_convlif_fwd``) and one fused BPTT step (``snnflow_convlif_bwd``) behind the C ABI of
``include/snnflow.h``; there is no PyTorch/CPU fallback - CPU tensors raise.
"""
import math

import torch
import torch.nn as nn

from . import _lib

_SUPPORTED_ACTIVATIONS = ("arctanspike", "superspike", "trianglespike", "mgspike")   # mgspike: per-step engine only
_workspaces = {}


def _workspace(device, nbytes):
    """Grow-only per-device scratch buffer (kernels on one stream run in order, so it is shared)."""
    buf = _workspaces.get(device)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        _workspaces[device] = buf
    return buf


def _f32c(t):
    if t.dtype != torch.float32:
        t = t.float()
    return t if t.is_contiguous() else t.contiguous()


def _grad_arg(t):
    """A parameter as an autograd.Function argument: itself with autograd on, detached under no_grad (so that
    ctx.needs_input_grad, which ignores the grad mode, is False and the forward kernels skip what only a backward needs)."""
    return t if (t is None or torch.is_grad_enabled()) else t.detach()


class _ConvLIFStep(torch.autograd.Function):
    """(x, prev_state, w_ff, w_rec, leak, thresh, residual) -> (state [2,B,C,H,W], out or None)."""

    @staticmethod
    def forward(ctx, x, prev_state, w_ff, w_rec, leak, thresh, residual, hard_reset, detach, surrogate, act_width,
                packed=None):
        L = _lib.lib()
        x = _f32c(x)
        w_ff = _f32c(w_ff)
        B, Cin, H, W = x.shape
        C = w_ff.shape[0]
        if w_rec is not None:
            w_rec = _f32c(w_rec)
        if prev_state is not None:
            # the kernels read B*C*H*W floats per half through the raw pointer: a state left over from another batch
            # size / resolution is a shape error here as it is in the reference (spiking_submodules.py:144)
            if tuple(prev_state.shape) != (2, B, C, H, W):
                raise _lib.SnnflowError(f"snnflow ConvLIF: prev_state is {tuple(prev_state.shape)}, this input needs "
                                        f"{(2, B, C, H, W)} (reset the states when batch size or resolution change)")
            prev_state = _f32c(prev_state)
        if residual is not None:
            residual = _f32c(residual)
        lam = torch.sigmoid(leak.detach()).reshape(-1).contiguous()          # spiking_submodules.py:136
        theta = thresh.detach().clamp_min(0.01).reshape(-1).contiguous()     # :133
        state = torch.empty((2, B, C, H, W), dtype=torch.float32, device=x.device)
        out = torch.empty((B, C, H, W), dtype=torch.float32, device=x.device) if residual is not None else None
        need_bwd = any(ctx.needs_input_grad)
        cur = torch.empty((B, C, H, W), dtype=torch.float32, device=x.device) if need_bwd else None
        flags = (_lib.HARD_RESET if hard_reset else 0) | (_lib.DETACH_RESET if detach else 0)
        v_in = prev_state[0] if prev_state is not None else None
        z_in = prev_state[1] if prev_state is not None else None
        if packed is not None:   # tensor-core path: fp16-exact input (spikes), pre-packed weights
            _lib.check(L.snnflow_convlif_fwd_tc(
                _lib.ptr(x), _lib.ptr(packed), int(w_rec is not None), _lib.ptr(v_in), _lib.ptr(z_in), _lib.ptr(lam),
                _lib.ptr(theta), _lib.ptr(residual), _lib.ptr(state[0]), _lib.ptr(state[1]), _lib.ptr(out),
                _lib.ptr(cur), B, Cin, C, H, W, flags, _lib.stream()), "snnflow_convlif_fwd_tc")
        else:
            _lib.check(L.snnflow_convlif_fwd(
                _lib.ptr(x), _lib.ptr(w_ff), _lib.ptr(w_rec), _lib.ptr(v_in), _lib.ptr(z_in), _lib.ptr(lam),
                _lib.ptr(theta), _lib.ptr(residual), _lib.ptr(state[0]), _lib.ptr(state[1]), _lib.ptr(out),
                _lib.ptr(cur), B, Cin, C, H, W, flags, _lib.stream()), "snnflow_convlif_fwd")
        if need_bwd:
            ctx.save_for_backward(x, prev_state, w_ff, w_rec, lam, theta, state, cur, thresh)
            # packed is only passed for inputs tagged fp16/bf16-exact (spikes): the backward may use tensor cores too
            bwd_flags = flags | (_lib.INPUT_EXACT16 if packed is not None else 0)
            ctx.cfg = (bwd_flags, surrogate, float(act_width), leak.shape, thresh.shape)
            ctx.has_residual = residual is not None
        return state, out

    @staticmethod
    def backward(ctx, g_state, g_out):
        L = _lib.lib()
        x, prev_state, w_ff, w_rec, lam, theta, state, cur, thresh = ctx.saved_tensors
        flags, surrogate, width, leak_shape, thresh_shape = ctx.cfg
        B, Cin, H, W = x.shape
        C = w_ff.shape[0]
        dev = x.device
        recurrent = w_rec is not None
        g_state = _f32c(g_state) if g_state is not None else None
        g_out = _f32c(g_out) if g_out is not None else None
        need_gx = ctx.needs_input_grad[0]
        g_x = torch.empty_like(x) if need_gx else None
        g_prev = torch.empty((2, B, C, H, W), dtype=torch.float32, device=dev)
        dw_ff = torch.zeros_like(w_ff)
        dw_rec = torch.zeros_like(w_rec) if recurrent else None
        dlam = torch.zeros(C, dtype=torch.float32, device=dev)
        dtheta = torch.zeros(C, dtype=torch.float32, device=dev)
        nbytes = L.snnflow_convlif_bwd_workspace_bytes(B, Cin, C, H, W, int(recurrent))
        ws = _workspace(dev, nbytes)
        v_in = prev_state[0] if prev_state is not None else None
        z_in = prev_state[1] if prev_state is not None else None
        _lib.check(L.snnflow_convlif_bwd(
            _lib.ptr(x), _lib.ptr(w_ff), _lib.ptr(w_rec), _lib.ptr(v_in), _lib.ptr(z_in), _lib.ptr(state[0]),
            _lib.ptr(cur), _lib.ptr(lam), _lib.ptr(theta), _lib.ptr(g_out),
            _lib.ptr(g_state[0]) if g_state is not None else None,
            _lib.ptr(g_state[1]) if g_state is not None else None,
            _lib.ptr(g_x), _lib.ptr(g_prev[0]), _lib.ptr(g_prev[1]), _lib.ptr(dw_ff), _lib.ptr(dw_rec),
            _lib.ptr(dlam), _lib.ptr(dtheta), ws.data_ptr(), ws.numel(), B, Cin, C, H, W, flags, surrogate, width,
            _lib.stream()), "snnflow_convlif_bwd")
        d_leak = (dlam * lam * (1.0 - lam)).reshape(leak_shape)                       # sigmoid'
        d_thresh = (dtheta * (thresh.reshape(-1) >= 0.01).float()).reshape(thresh_shape)  # clamp_min'
        g_prev_ret = g_prev if (prev_state is not None and ctx.needs_input_grad[1]) else None
        g_res = g_out if ctx.has_residual else None
        return (g_x, g_prev_ret, dw_ff, dw_rec, d_leak, d_thresh, g_res, None, None, None, None, None)


class _LIFBase(nn.Module):
    recurrent = False

    def _setup(self, input_size, hidden_size, kernel_size, stride, activation, act_width, leak, thresh, learn_leak,
               learn_thresh, hard_reset, detach, norm, quantization_config):
        if kernel_size != 3 or stride != 1:
            raise NotImplementedError("snnflow ConvLIF: the fused sm_100a kernel covers kernel_size=3, stride=1")
        if norm is not None:
            raise NotImplementedError("snnflow ConvLIF: norm='weight'/'group' is outside the fused-kernel envelope")
        if quantization_config and quantization_config.get("enabled", False):
            raise NotImplementedError("snnflow ConvLIF: brevitas-quantised convolutions are out of scope")
        if not isinstance(activation, str) or activation not in _SUPPORTED_ACTIVATIONS:
            raise NotImplementedError(f"snnflow ConvLIF: activation {activation!r} not in {_SUPPORTED_ACTIVATIONS}")
        self.input_size = input_size
        self.hidden_size = hidden_size
        padding = kernel_size // 2
        # construction order (and therefore RNG consumption) follows spiking_submodules.py:86-100 / :222-239
        self.ff = nn.Conv2d(input_size, hidden_size, kernel_size, stride=stride, padding=padding, bias=False)
        if self.recurrent:
            self.rec = nn.Conv2d(hidden_size, hidden_size, kernel_size, padding=padding, bias=False)
        leak_init = torch.randn(hidden_size, 1, 1) * leak[1] + leak[0]
        if learn_leak:
            self.leak = nn.Parameter(leak_init)
        else:
            self.register_buffer("leak", leak_init)
        thresh_init = torch.randn(hidden_size, 1, 1) * thresh[1] + thresh[0]
        if learn_thresh:
            self.thresh = nn.Parameter(thresh_init)
        else:
            self.register_buffer("thresh", thresh_init)
        nn.init.uniform_(self.ff.weight, -math.sqrt(1 / input_size), math.sqrt(1 / input_size))
        if self.recurrent:
            nn.init.uniform_(self.rec.weight, -math.sqrt(1 / hidden_size), math.sqrt(1 / hidden_size))
        self.activation = activation
        self.register_buffer("act_width", torch.tensor(act_width))
        self._act_width = float(act_width)
        self.hard_reset = hard_reset
        self.detach = detach
        self.norm = None

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)
        self._act_width = float(self.act_width)

    use_tensor_cores = True   # class-level switch (tests flip it to compare both paths)

    def _packed_weights(self):
        """fp16 hi/lo split of the weights in the UMMA smem layout; rebuilt only when the weights change."""
        w_ff = self.ff.weight
        w_rec = self.rec.weight if self.recurrent else None
        key = (w_ff.data_ptr(), w_ff._version, None if w_rec is None else (w_rec.data_ptr(), w_rec._version))
        cached = getattr(self, "_packed_cache", None)
        if cached is not None and cached[0] == key:
            return cached[1]
        L = _lib.lib()
        nbytes = L.snnflow_convlif_packed_bytes(self.input_size, self.hidden_size, int(self.recurrent))
        if nbytes == 0:
            return None
        blob = torch.empty(nbytes, dtype=torch.uint8, device=w_ff.device)
        wf = _f32c(w_ff.detach())
        wr = _f32c(w_rec.detach()) if w_rec is not None else None
        _lib.check(L.snnflow_convlif_pack(_lib.ptr(wf), _lib.ptr(wr), blob.data_ptr(), self.input_size,
                                          self.hidden_size, _lib.stream()), "snnflow_convlif_pack")
        self._packed_cache = (key, blob)
        return blob

    def _step(self, input_, prev_state, residual):
        if not input_.is_cuda:
            raise _lib.SnnflowError("snnflow ConvLIF runs on CUDA tensors only (no CPU fallback)")
        res = residual if torch.is_tensor(residual) else None
        # The tensor-core kernel needs fp16-exact inputs.  Tensors produced by these cells (spikes, spikes +
        # tagged residual) carry the tag `_snnflow_exact16`; anything else takes the exact-fp32 CUDA-core path.
        exact16 = getattr(input_, "_snnflow_exact16", False)
        packed = self._packed_weights() if (self.use_tensor_cores and exact16) else None
        # (_grad_arg: under no_grad the kernels must not save anything for a backward pass - Function.forward only sees
        # the tensors' requires_grad flags, not the caller's grad mode)
        state, out = _ConvLIFStep.apply(
            input_, prev_state, _grad_arg(self.ff.weight), _grad_arg(self.rec.weight) if self.recurrent else None,
            _grad_arg(self.leak), _grad_arg(self.thresh),
            res, self.hard_reset, self.detach, _lib.SURROGATE_ID[self.activation], self._act_width, packed)
        if out is None:
            out = state[1]
            if not torch.is_tensor(residual) and residual != 0:
                return out + residual, state
            out._snnflow_exact16 = True
        elif getattr(res, "_snnflow_exact16", False):
            out._snnflow_exact16 = True
        return out, state


class ConvLIF(_LIFBase):
    """Convolutional spiking LIF cell (models/spiking_submodules.py:29-151)."""

    def __init__(self, input_size, hidden_size, kernel_size, stride=1, activation="arctanspike", act_width=10.0,
                 leak=(-4.0, 0.1), thresh=(0.8, 0.0), learn_leak=True, learn_thresh=True, hard_reset=True, detach=True,
                 norm=None, quantization_config=None, exporting=False, tebn=False, num_timesteps=4, mpbn=False):
        super().__init__()
        self._setup(input_size, hidden_size, kernel_size, stride, activation, act_width, leak, thresh, learn_leak,
                    learn_thresh, hard_reset, detach, norm, quantization_config)

    def forward(self, input_, prev_state, residual=0, timestep=None):
        return self._step(input_, prev_state, residual)


class ConvLIFRecurrent(_LIFBase):
    """Convolutional recurrent spiking LIF cell (models/spiking_submodules.py:154-300)."""

    recurrent = True

    def __init__(self, input_size, hidden_size, kernel_size, activation="arctanspike", act_width=10.0,
                 leak=(-4.0, 0.1), thresh=(0.8, 0.0), learn_leak=True, learn_thresh=True, hard_reset=True, detach=True,
                 norm=None, quantization_config=None, exporting=False, tebn=False, num_timesteps=4, mpbn=False):
        super().__init__()
        self._setup(input_size, hidden_size, kernel_size, 1, activation, act_width, leak, thresh, learn_leak,
                    learn_thresh, hard_reset, detach, norm, quantization_config)

    def forward(self, input_, prev_state, residual=0, timestep=None):
        # the reference cell takes no residual (spiking_submodules.py:265); LIFFireNet never passes one to it
        return self._step(input_, prev_state, 0)
