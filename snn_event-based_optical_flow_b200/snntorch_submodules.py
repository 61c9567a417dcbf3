"""SNNtorch_ConvLIF / SNNtorch_ConvLIFRecurrent on the B200: host-side mirror of models/SNNtorch_spiking_submodules.py
(:124-322, :324-567) - the cells the reference's LIFFireNet wires in by default (models/model.py:37-39).

Same constructor arguments, same sub-module / parameter names (``ff.weight``, ``rec.weight``, ``bn.*``, ``lif.beta``,
``lif.threshold`` and snn.Leaky's two buffers), same state layout ``stack([mem, spk])`` and return values.  A step is

    I   = ff(x) [+ rec(spk_prev)]          snnflow_convlif_fwd / _bwd  (the fused 3x3 conv kernels of the ConvLIF cells with
                                            lam = 0: the layer-step degenerates to "membrane = input current")
    I   = BatchNorm2d(I)                   torch (training mode couples the whole batch through its statistics)
    spk, mem = Leaky(I, mem_prev)          snnflow_leaky_fwd / _bwd    (csrc/leaky.cu: clamp(beta), reset-to-zero / subtract
                                            without reset delay, ATan(alpha = 2) surrogate, detached membrane)

PARITY UNPINNED: snntorch 0.9.4 is not installable offline, so the neuron arithmetic restates the
published ``Leaky.forward`` (the tests compare with a CPU restatement of the same text).  TEBN / MPBN, the
brevitas-quantised branches and ``detach=False`` raise ``NotImplementedError``.  There is no CPU fallback.
"""
import math

import torch
import torch.nn as nn

from . import _lib
from .spiking_submodules import _ConvLIFStep, _f32c, _grad_arg, _workspace


class _LeakyStep(torch.autograd.Function):
    """(cur, mem_in or None, beta [C,1,1], threshold [C,1,1]) -> (spk, mem_out); mem_out carries no gradient."""

    @staticmethod
    def forward(ctx, cur, mem_in, beta, threshold, subtract):
        L = _lib.lib()
        cur = _f32c(cur)
        B, C, H, W = cur.shape
        if mem_in is not None:
            if tuple(mem_in.shape) != (B, C, H, W):
                raise _lib.SnnflowError(f"snnflow Leaky: membrane is {tuple(mem_in.shape)}, the input current {tuple(cur.shape)} "
                                        "(reset the states when batch size or resolution change)")
            mem_in = _f32c(mem_in.detach())
        b = _f32c(beta.detach().reshape(-1))
        th = _f32c(threshold.detach().reshape(-1))
        need_bwd = any(ctx.needs_input_grad)
        mem_out, spk = torch.empty_like(cur), torch.empty_like(cur)
        m_pre = torch.empty_like(cur) if need_bwd else None
        _lib.check(L.snnflow_leaky_fwd(_lib.ptr(cur), _lib.ptr(mem_in), _lib.ptr(b), _lib.ptr(th), _lib.ptr(mem_out), _lib.ptr(spk),
                                       _lib.ptr(m_pre), B, C, H, W, int(subtract), _lib.stream()), "snnflow_leaky_fwd")
        if need_bwd:
            ctx.save_for_backward(m_pre, mem_in, b, th)
            ctx.cfg = (int(subtract), beta.shape, threshold.shape)
        ctx.mark_non_differentiable(mem_out)
        return spk, mem_out

    @staticmethod
    def backward(ctx, g_spk, _g_mem):
        L = _lib.lib()
        m_pre, mem_in, b, th = ctx.saved_tensors
        subtract, bshape, tshape = ctx.cfg
        B, C, H, W = m_pre.shape
        g_spk = _f32c(g_spk)
        g_cur = torch.empty_like(m_pre)
        d_beta = torch.empty(C, dtype=torch.float32, device=m_pre.device)
        d_theta = torch.empty(C, dtype=torch.float32, device=m_pre.device)
        ws = _workspace(m_pre.device, L.snnflow_leaky_bwd_workspace_bytes(B, C, H, W))
        _lib.check(L.snnflow_leaky_bwd(_lib.ptr(g_spk), _lib.ptr(m_pre), _lib.ptr(mem_in), _lib.ptr(b), _lib.ptr(th), _lib.ptr(g_cur),
                                       _lib.ptr(d_beta), _lib.ptr(d_theta), ws.data_ptr(), ws.numel(), B, C, H, W, subtract,
                                       _lib.stream()), "snnflow_leaky_bwd")
        return g_cur, None, d_beta.reshape(bshape), d_theta.reshape(tshape), None


class Leaky(nn.Module):
    """Parameter container with snn.Leaky's names (``beta``, ``threshold``; buffers ``graded_spikes_factor``,
    ``reset_mechanism_val``), so that ``lif.*`` state_dict keys of the reference's checkpoints load."""

    def __init__(self, beta, threshold, learn_beta, learn_threshold, reset_mechanism):
        super().__init__()
        if learn_beta:
            self.beta = nn.Parameter(beta)
        else:
            self.register_buffer("beta", beta)
        if learn_threshold:
            self.threshold = nn.Parameter(threshold)
        else:
            self.register_buffer("threshold", threshold)
        self.reset_mechanism = reset_mechanism
        self.register_buffer("graded_spikes_factor", torch.tensor(1.0))
        self.register_buffer("reset_mechanism_val", torch.tensor({"subtract": 0, "zero": 1}[reset_mechanism]))

    def forward(self, cur, mem):
        return _LeakyStep.apply(cur, mem, _grad_arg(self.beta), _grad_arg(self.threshold), self.reset_mechanism == "subtract")


class _SNNtorchBase(nn.Module):
    recurrent = False
    use_tensor_cores = True

    def _setup(self, input_size, hidden_size, kernel_size, stride, leak, thresh, learn_leak, learn_thresh, hard_reset, detach,
               norm, quantization_config, tebn, mpbn):
        if kernel_size != 3 or stride != 1:
            raise NotImplementedError("snnflow SNNtorch_ConvLIF: the fused sm_100a kernels cover kernel_size=3, stride=1")
        if norm is not None:
            raise NotImplementedError("snnflow SNNtorch_ConvLIF: norm='weight'/'group' is outside the fused-kernel envelope")
        if quantization_config and quantization_config.get("enabled", False):
            raise NotImplementedError("snnflow SNNtorch_ConvLIF: brevitas-quantised convolutions are out of scope")
        if tebn or mpbn:
            raise NotImplementedError("snnflow SNNtorch_ConvLIF: TEBN / MPBN are not covered")
        if not detach:
            raise NotImplementedError("snnflow SNNtorch_ConvLIF: detach=False (gradient through the membrane) is not covered")
        self.input_size, self.hidden_size = input_size, hidden_size
        # construction / RNG order of SNNtorch_spiking_submodules.py:165-166, :237-250 (:365-366, :440-452)
        beta_init = torch.empty(hidden_size, 1, 1).uniform_(leak[0], leak[1])
        threshold_init = torch.empty(hidden_size, 1, 1).uniform_(thresh[0], thresh[1])
        padding = kernel_size // 2
        self.ff = nn.Conv2d(input_size, hidden_size, kernel_size, stride=stride, padding=padding, bias=False)
        if self.recurrent:
            self.rec = nn.Conv2d(hidden_size, hidden_size, kernel_size, padding=padding, bias=False)
        self.lif = Leaky(beta_init, threshold_init, learn_leak, learn_thresh, "zero" if hard_reset else "subtract")
        nn.init.uniform_(self.ff.weight, -math.sqrt(1 / input_size), math.sqrt(1 / input_size))
        if self.recurrent:
            nn.init.uniform_(self.rec.weight, -math.sqrt(1 / hidden_size), math.sqrt(1 / hidden_size))
        self.bn = nn.BatchNorm2d(hidden_size, momentum=0.1, eps=1e-5)
        self.tebn_enabled, self.mpbn_enabled, self.mpbn = False, False, None
        self.detach = detach
        self.norm = None
        # constants that turn the ConvLIF layer-step kernel into a plain conv: lam = sigmoid(-inf) = 0, no spike ever
        self.register_buffer("_conv_leak", torch.full((hidden_size, 1, 1), -1.0e4), persistent=False)
        self.register_buffer("_conv_thresh", torch.full((hidden_size, 1, 1), 3.0e38), persistent=False)

    _packed_weights = None   # bound below (shares the cache logic of the ConvLIF cells)

    def _current(self, input_, prev_spk):
        """ff(x) [+ rec(spk_prev)] through the fused conv kernels (fp32 CUDA cores, or tcgen05 when the input is tagged
        fp16-exact - spikes / event counts)."""
        if not input_.is_cuda:
            raise _lib.SnnflowError("snnflow SNNtorch_ConvLIF runs on CUDA tensors only (no CPU fallback)")
        exact16 = getattr(input_, "_snnflow_exact16", False)
        packed = self._packed_weights() if (self.use_tensor_cores and exact16) else None
        state_in = None
        if self.recurrent and prev_spk is not None:
            state_in = torch.stack([torch.zeros_like(prev_spk), prev_spk])    # (v, z): v is multiplied by lam = 0
        state, _ = _ConvLIFStep.apply(input_, state_in, _grad_arg(self.ff.weight), _grad_arg(self.rec.weight) if self.recurrent else None,
                                      self._conv_leak, self._conv_thresh, None, True, True, 0, 10.0, packed)
        return state[0]

    def _step(self, input_, prev_state):
        with torch.no_grad():
            self.lif.threshold.data.clamp_(min=0.01)                          # SNNtorch_spiking_submodules.py:284 / :493
        mem = None if prev_state is None else prev_state[0]
        prev_spk = None if prev_state is None else prev_state[1]
        cur = self.bn(self._current(input_, prev_spk))                        # :287-296 / :497-529
        spk, mem_out = self.lif(cur, mem)                                     # :305 / :550 ; mem_out is detached (:309-311)
        spk._snnflow_exact16 = True
        return spk, torch.stack([mem_out, spk], dim=0)                        # :320-322


from .spiking_submodules import _LIFBase  # noqa: E402

_SNNtorchBase._packed_weights = _LIFBase._packed_weights


class SNNtorch_ConvLIF(_SNNtorchBase):
    """models/SNNtorch_spiking_submodules.py:124-322."""

    def __init__(self, input_size, hidden_size, kernel_size, stride=1, activation="arctanspike", act_width=10.0,
                 leak=(0.0, 1.0), thresh=(0.0, 0.8), learn_leak=True, learn_thresh=True, hard_reset=True, detach=True,
                 norm=None, quantization_config=None, exporting=False, tebn=False, num_timesteps=4, mpbn=False):
        super().__init__()
        self._setup(input_size, hidden_size, kernel_size, stride, leak, thresh, learn_leak, learn_thresh, hard_reset, detach,
                    norm, quantization_config, tebn, mpbn)

    def forward(self, input_, prev_state, residual=0, timestep=None):
        return self._step(input_, prev_state)


class SNNtorch_ConvLIFRecurrent(_SNNtorchBase):
    """models/SNNtorch_spiking_submodules.py:324-567."""

    recurrent = True

    def __init__(self, input_size, hidden_size, kernel_size, activation="arctanspike", act_width=10.0, leak=(0.0, 1.0),
                 thresh=(0.0, 0.8), learn_leak=True, learn_thresh=True, hard_reset=True, detach=True, norm=None,
                 quantization_config=None, exporting=False, tebn=False, num_timesteps=4, mpbn=False):
        super().__init__()
        self._setup(input_size, hidden_size, kernel_size, 1, leak, thresh, learn_leak, learn_thresh, hard_reset, detach, norm,
                    quantization_config, tebn, mpbn)

    def forward(self, input_, prev_state, residual=0, timestep=None):
        return self._step(input_, prev_state)
