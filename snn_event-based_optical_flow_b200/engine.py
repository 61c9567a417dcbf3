"""Whole-window execution of LIFFireNet / LIFFireFlowNet: T time bins forward (and the matching BPTT) as one
C call each (``snnflow_net_forward`` / ``snnflow_net_backward``, include/snnflow.h), wrapped in a single
``torch.autograd.Function``.

The per-layer modules stay the drop-in face (models/model.py calls them once per bin); this is the fast path for
a caller that has the T bins of a loss window at hand (train_flow.py:232-279 processes them back to back anyway).
Saved activations live in a preallocated arena ([v | z | I] per layer and bin), so nothing is allocated, stacked
or cloned per layer-step, and gradients accumulate straight into per-parameter buffers.
"""
import ctypes
import math

import torch

from . import _lib

N_LAYERS = 7


class NetDesc(ctypes.Structure):
    _fields_ = [("B", ctypes.c_int), ("C", ctypes.c_int), ("H", ctypes.c_int), ("W", ctypes.c_int), ("T", ctypes.c_int),
                ("num_bins", ctypes.c_int), ("recurrent_mask", ctypes.c_uint), ("flags", ctypes.c_uint),
                ("surrogate", ctypes.c_int), ("act_width", ctypes.c_float)]


class LayerPtrs(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in ("w_ff", "w_rec", "lam", "theta", "packed", "dw_ff", "dw_rec", "dlam",
                                               "dtheta", "thresh_raw", "d_leak", "d_thresh")]


def _bind(L):
    if getattr(L, "_net_bound", False):
        return
    P = ctypes.c_void_p
    L.snnflow_net_acts_floats.restype = ctypes.c_size_t
    L.snnflow_net_acts_floats.argtypes = [P, ctypes.c_int]
    L.snnflow_net_bwd_workspace_bytes.restype = ctypes.c_size_t
    L.snnflow_net_bwd_workspace_bytes.argtypes = [P]
    L.snnflow_net_forward.restype = ctypes.c_int
    L.snnflow_net_forward.argtypes = [P, P, P, P, P, ctypes.POINTER(P), P, P,
                                      ctypes.c_int, P]
    L.snnflow_net_backward.restype = ctypes.c_int
    L.snnflow_net_backward.argtypes = [P, P, P, P, ctypes.POINTER(P), P, P, P,
                                       P, P, P, ctypes.c_size_t, P]
    L.snnflow_window_supported.restype = ctypes.c_int
    L.snnflow_window_supported.argtypes = [P, ctypes.c_int]
    L.snnflow_window_arena_bytes.restype = ctypes.c_size_t
    L.snnflow_window_arena_bytes.argtypes = [P, ctypes.c_int]
    L.snnflow_window_workspace_bytes.restype = ctypes.c_size_t
    L.snnflow_window_workspace_bytes.argtypes = [P]
    L.snnflow_window_state_offsets.restype = ctypes.c_int
    L.snnflow_window_state_offsets.argtypes = [P, ctypes.c_int, ctypes.POINTER(ctypes.c_size_t)]
    L.snnflow_window_flags_offset.restype = ctypes.c_size_t
    L.snnflow_window_flags_offset.argtypes = [P, ctypes.c_int]
    L.snnflow_window_forward.restype = ctypes.c_int
    L.snnflow_window_forward.argtypes = [P, P, P, P, P, ctypes.POINTER(P), P, P,
                                         ctypes.c_int, P]
    L.snnflow_window_import_state.restype = ctypes.c_int
    L.snnflow_window_import_state.argtypes = [P, ctypes.POINTER(P), P, P]
    L.snnflow_window_export_state.restype = ctypes.c_int
    L.snnflow_window_export_state.argtypes = [P, P, ctypes.c_int, ctypes.POINTER(P), P]
    L.snnflow_window_backward.restype = ctypes.c_int
    L.snnflow_window_backward.argtypes = [P, P, P, ctypes.POINTER(P), P, P, P,
                                          P, P, P, ctypes.c_size_t, P]
    L._net_bound = True


def _state_ptrs(states, shape=None, contiguous=True):
    """The states as an array of raw pointers.  `shape` = (B, C, H, W) of the window: every state must be a contiguous
    fp32 CUDA tensor [2,B,C,H,W] - the kernels read B*C*H*W floats per half through the raw pointer, so a state left
    over from another batch size / resolution (no reset_states() in between) is an error here, as it is a shape error
    in the reference (spiking_submodules.py:144: v * leak * (1 - z) + (1 - leak) * ff)."""
    arr = (ctypes.c_void_p * N_LAYERS)()
    if states is None or all(s is None for s in states):
        return None, arr
    for i, s in enumerate(states):
        if s is None:
            arr[i] = None
            continue
        if shape is not None:
            want = (2,) + tuple(shape)
            if tuple(s.shape) != want or s.dtype != torch.float32 or not s.is_cuda or (contiguous and not s.is_contiguous()):
                raise _lib.SnnflowError(
                    f"layer {i}: state is {tuple(s.shape)} {s.dtype} on {s.device}"
                    f"{'' if s.is_contiguous() else ' (non-contiguous)'}, the window needs a contiguous float32 CUDA "
                    f"tensor {want}: call reset_states() when the batch size or the resolution changes")
        arr[i] = s.data_ptr()
    return arr, arr


def _effective_params(runner, layers):
    """lam = sigmoid(leak) (spiking_submodules.py:136) and theta = clamp_min(thresh, 0.01) (:133) of all layers, [7,C]
    each.  `runner.param_override = (lam, theta)` injects values evaluated elsewhere (the parity tests pass the
    reference's CPU values: torch's CUDA and CPU sigmoid differ in the last bit)."""
    ov = getattr(runner, "param_override", None)
    if ov is not None:
        dev = layers[0].leak.device
        return ov[0].to(dev, torch.float32).contiguous(), ov[1].to(dev, torch.float32).contiguous()
    lam = torch.sigmoid(torch.stack([l.leak.detach().reshape(-1) for l in layers]))
    theta = torch.stack([l.thresh.detach().reshape(-1) for l in layers]).clamp_min(0.01)
    return lam, theta


class _WindowFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, runner, cnt, *params):
        L = _lib.lib()
        _bind(L)
        net = runner.net
        layers = runner.layers
        T, B, nb, H, W = cnt.shape
        C = layers[0].hidden_size
        dev = cnt.device
        cnt = cnt.float().contiguous()
        need_bwd = any(ctx.needs_input_grad)   # (grad mode is off inside Function.forward)
        # effective leak / threshold of all layers in two launches (spiking_submodules.py:133,136)
        lam, theta = _effective_params(runner, layers)
        first = layers[0]
        flags = (_lib.HARD_RESET if first.hard_reset else 0) | (_lib.DETACH_RESET if first.detach else 0)
        if not first.use_tensor_cores:
            flags |= _lib.NO_TENSOR_CORES
        mask = sum(1 << i for i, l in enumerate(layers) if l.recurrent)
        desc = NetDesc(B, C, H, W, T, nb, mask, flags, _lib.SURROGATE_ID[first.activation], first._act_width)
        packed = [None if (i == 0 or not l.use_tensor_cores) else l._packed_weights() for i, l in enumerate(layers)]
        lp = (LayerPtrs * N_LAYERS)()
        for i, l in enumerate(layers):
            lp[i].w_ff = l.ff.weight.data_ptr()
            lp[i].w_rec = l.rec.weight.data_ptr() if l.recurrent else None
            lp[i].lam = lam[i].data_ptr()
            lp[i].theta = theta[i].data_ptr()
            lp[i].packed = None if packed[i] is None else packed[i].data_ptr()
        acts = runner.arena(desc, need_bwd, dev)
        flow = torch.empty((T, B, 2, H, W), dtype=torch.float32, device=dev)
        states = net._states
        sp, keep = _state_ptrs(states, (B, C, H, W))
        pw, pb = net.pred.conv2d.weight, net.pred.conv2d.bias
        _lib.check(L.snnflow_net_forward(ctypes.byref(desc), lp, pw.data_ptr(), None if pb is None else pb.data_ptr(),
                                         cnt.data_ptr(), sp, acts.data_ptr(), flow.data_ptr(), int(need_bwd),
                                         _lib.stream()), "snnflow_net_forward")
        # new states: zero-copy views into the arena ([v | z] of the last bin)
        n = B * C * H * W
        new_states = []
        for i in range(N_LAYERS):
            off = ((i * T + (T - 1)) * 3 * n) if need_bwd else ((i * 2 + ((T - 1) & 1)) * 2 * n)
            new_states.append(acts[off:off + 2 * n].view(2, B, C, H, W))
        runner.new_states = new_states
        if need_bwd:
            ctx.runner, ctx.desc, ctx.lam, ctx.theta, ctx.packed = runner, desc, lam, theta, packed
            ctx.cnt, ctx.acts, ctx.flow, ctx.states_in = cnt, acts, flow, list(states)
        return flow

    @staticmethod
    def backward(ctx, g_flow):
        L = _lib.lib()
        runner, desc, lam, theta = ctx.runner, ctx.desc, ctx.lam, ctx.theta
        layers, net = runner.layers, runner.net
        dev = g_flow.device
        C = desc.C
        g_flow = g_flow.float().contiguous()
        dlam = torch.zeros_like(lam)
        dtheta = torch.zeros_like(theta)
        dws = []
        lp = (LayerPtrs * N_LAYERS)()
        for i, l in enumerate(layers):
            dwf = torch.zeros_like(l.ff.weight)
            dwr = torch.zeros_like(l.rec.weight) if l.recurrent else None
            dws.append((dwf, dwr))
            lp[i].w_ff = l.ff.weight.data_ptr()
            lp[i].w_rec = l.rec.weight.data_ptr() if l.recurrent else None
            lp[i].lam = lam[i].data_ptr()
            lp[i].theta = theta[i].data_ptr()
            lp[i].packed = None if ctx.packed[i] is None else ctx.packed[i].data_ptr()
            lp[i].dw_ff = dwf.data_ptr()
            lp[i].dw_rec = None if dwr is None else dwr.data_ptr()
            lp[i].dlam = dlam[i].data_ptr()
            lp[i].dtheta = dtheta[i].data_ptr()
        pw, pb = net.pred.conv2d.weight, net.pred.conv2d.bias
        d_pw = torch.zeros_like(pw)
        d_pb = torch.zeros(2, dtype=torch.float32, device=dev)
        ws = runner.workspace(desc, dev)
        sp, keep = _state_ptrs(ctx.states_in)
        _lib.check(L.snnflow_net_backward(ctypes.byref(desc), lp, pw.data_ptr(), ctx.cnt.data_ptr(), sp, ctx.acts.data_ptr(),
                                          ctx.flow.data_ptr(), g_flow.data_ptr(), d_pw.data_ptr(), d_pb.data_ptr(),
                                          ws.data_ptr(), ws.numel(), _lib.stream()), "snnflow_net_backward")
        d_leak = dlam * lam * (1.0 - lam)                                           # sigmoid'
        thr = torch.stack([l.thresh.detach().reshape(-1) for l in layers])
        d_thresh = dtheta * (thr >= 0.01).float()                                   # clamp_min'
        grads = []
        for i, l in enumerate(layers):
            grads += [dws[i][0], dws[i][1], d_leak[i].reshape(l.leak.shape), d_thresh[i].reshape(l.thresh.shape)]
        grads += [d_pw, d_pb if pb is not None else None]
        return (None, None) + tuple(grads)

_LM_WIDTHS = (16, 32, 64)   # channel counts the layer-major engine is built for


def _lm_channels(c):
    """Width the layer-major engine runs a C-channel network at: narrower networks (the shipped configs use
    base_num_channels = 8, configs/train_SNN.yml:19) are zero-padded - padded neurons have zero weights, never spike and
    get zero gradients, so the real channels are bit-identical to an unpadded run."""
    for w in _LM_WIDTHS:
        if c <= w:
            return w
    return c


def _pad_dim(t, dim, n, value=0.0):
    if t is None or t.shape[dim] == n:
        return t
    pad = [0, 0] * (t.dim() - 1 - dim) + [0, n - t.shape[dim]]
    return torch.nn.functional.pad(t, pad, value=value).contiguous()


def _make_desc(runner, cnt, channels=None):
    layers = runner.layers
    T, B, nb, H, W = cnt.shape
    first = layers[0]
    flags = (_lib.HARD_RESET if first.hard_reset else 0) | (_lib.DETACH_RESET if first.detach else 0)
    if not first.use_tensor_cores:
        flags |= _lib.NO_TENSOR_CORES
    mask = sum(1 << i for i, l in enumerate(layers) if l.recurrent)
    return NetDesc(B, channels or first.hidden_size, H, W, T, nb, mask, flags, _lib.SURROGATE_ID[first.activation], first._act_width)


def _desc_key(d):
    return (d.B, d.C, d.H, d.W, d.T, d.num_bins, d.recurrent_mask, d.flags, d.surrogate, d.act_width)


def _lm_forward(runner, cnt, need_bwd):
    """One window through snnflow_window_forward.  Returns (flow, saved) where `saved` is what _lm_backward needs."""
    L = _lib.lib()
    _bind(L)
    net, layers = runner.net, runner.layers
    T, B, nb, H, W = cnt.shape
    dev = cnt.device
    cnt = cnt.float().contiguous()
    Cr = layers[0].hidden_size                 # the network's width ...
    C = _lm_channels(Cr)                       # ... and the (zero-padded) width the engine runs it at
    desc = _make_desc(runner, cnt, C)
    lam, theta = _effective_params(runner, layers)
    lam_e, theta_e = _pad_dim(lam, 1, C, 0.5), _pad_dim(theta, 1, C, 1.0)
    w_e = []
    for i, l in enumerate(layers):
        wf = _pad_dim(l.ff.weight.detach(), 0, C)
        if i > 0:
            wf = _pad_dim(wf, 1, C)
        wr = _pad_dim(_pad_dim(l.rec.weight.detach(), 0, C), 1, C) if l.recurrent else None
        w_e.append((wf, wr))
    lp = (LayerPtrs * N_LAYERS)()
    for i, l in enumerate(layers):
        lp[i].w_ff = w_e[i][0].data_ptr()
        lp[i].w_rec = w_e[i][1].data_ptr() if l.recurrent else None
        lp[i].lam = lam_e[i].data_ptr()
        lp[i].theta = theta_e[i].data_ptr()
    arena = runner.lm_arena(desc, need_bwd, dev)
    flow = torch.empty((T, B, 2, H, W), dtype=torch.float32, device=dev)
    _state_ptrs(net._states, (B, Cr, H, W), contiguous=C == Cr)   # shape check on the network's own states
    states = runner.lm_states_in(net._states, Cr, C)
    sp, keep = _state_ptrs(states, (B, C, H, W))
    pw, pb = net.pred.conv2d.weight, net.pred.conv2d.bias
    pw_e = _pad_dim(pw.detach(), 1, C)
    _lib.check(L.snnflow_window_forward(ctypes.byref(desc), lp, pw_e.data_ptr(), None if pb is None else pb.data_ptr(),
                                        cnt.data_ptr(), sp, arena.data_ptr(), flow.data_ptr(), int(need_bwd),
                                        _lib.stream()), "snnflow_window_forward")
    # the arena's sticky "input was not bf16-exact" word (per arena, i.e. per runner and shape): polled here, and
    # handed to the fused optimizer as its update gate (train.TrainWindow) so that graph replays cannot apply an
    # update computed from rounded inputs before the host has looked
    foff = L.snnflow_window_flags_offset(ctypes.byref(desc), int(need_bwd))
    runner.input_flag = arena[foff:foff + 4].view(torch.int32)
    runner._lm_calls = getattr(runner, "_lm_calls", 0) + 1
    capturing = torch.cuda.is_current_stream_capturing()
    if not capturing and (runner.validate_input or runner._lm_calls == 1 or runner._lm_calls % runner.validate_every == 0):
        runner.check_input_flag()
    offs = (ctypes.c_size_t * N_LAYERS)()
    _lib.check(L.snnflow_window_state_offsets(ctypes.byref(desc), int(need_bwd), offs), "snnflow_window_state_offsets")
    nbytes = 2 * B * C * H * W * 4
    full = [arena[offs[i]:offs[i] + nbytes].view(torch.float32).view(2, B, C, H, W) for i in range(N_LAYERS)]
    runner._lm_full_states = full
    runner.new_states = full if C == Cr else [f[:, :, :Cr] for f in full]   # the network's channels of the padded state
    saved = None
    if need_bwd:
        saved = dict(desc=desc, lam=lam, theta=theta, lam_e=lam_e, theta_e=theta_e, w_e=w_e, pw_e=pw_e, Cr=Cr, arena=arena,
                     flow=flow, states_in=list(states))
    return flow, saved


def _lm_backward(runner, saved, g_flow, dst=None):
    """snnflow_window_backward for a window run by _lm_forward(..., need_bwd=True).
    dst None: returns [(dw_ff, dw_rec)] x 7, dlam [7,C], dtheta [7,C], d_pred_w, d_pred_b as slices of ONE fresh zero-filled
    buffer.  dst = dict(dw=[(ff, rec)], dlam, dtheta, d_pw, d_pb) of caller-owned ZEROED tensors (un-padded width only):
    the kernels accumulate straight into them (the direct training step: slices of the optimizer's flat gradient)."""
    L = _lib.lib()
    desc, layers, net = saved["desc"], runner.layers, runner.net
    dev = g_flow.device
    g_flow = g_flow.float().contiguous()
    lam_e, theta_e, w_e, pw_e = saved["lam_e"], saved["theta_e"], saved["w_e"], saved["pw_e"]
    if dst is None:
        # every gradient of the window is a slice of ONE zero-filled buffer (one fill instead of one per tensor)
        shapes = [lam_e.shape, theta_e.shape]
        for i, l in enumerate(layers):
            shapes.append(w_e[i][0].shape)
            if l.recurrent:
                shapes.append(w_e[i][1].shape)
        shapes += [pw_e.shape, (2,)]
        sizes = [(math.prod(sh) + 63) // 64 * 64 for sh in shapes]
        gbuf = torch.zeros(sum(sizes), dtype=torch.float32, device=dev)
        carved, o = [], 0
        for sh, n in zip(shapes, sizes):
            carved.append(gbuf[o:o + math.prod(sh)].view(sh))
            o += n
        carved = iter(carved)
        dlam, dtheta = next(carved), next(carved)
        dws = []
        for i, l in enumerate(layers):
            dwf = next(carved)
            dws.append((dwf, next(carved) if l.recurrent else None))
        d_pw, d_pb = next(carved), next(carved)
    else:
        dws, dlam, dtheta, d_pw, d_pb = dst["dw"], dst["dlam"], dst["dtheta"], dst["d_pw"], dst["d_pb"]
    lp = (LayerPtrs * N_LAYERS)()
    for i, l in enumerate(layers):
        lp[i].w_ff = w_e[i][0].data_ptr()
        lp[i].w_rec = w_e[i][1].data_ptr() if l.recurrent else None
        lp[i].lam = lam_e[i].data_ptr()
        lp[i].theta = theta_e[i].data_ptr()
        lp[i].dw_ff = dws[i][0].data_ptr()
        lp[i].dw_rec = None if dws[i][1] is None else dws[i][1].data_ptr()
        lp[i].dlam = None if dlam is None else dlam[i].data_ptr()
        lp[i].dtheta = None if dtheta is None else dtheta[i].data_ptr()
        if dst is not None and dst.get("d_leak") is not None:   # chain rule to the raw parameters inside the reduction launch
            lp[i].thresh_raw = l.thresh.data_ptr()
            lp[i].d_leak = dst["d_leak"][i].data_ptr()
            lp[i].d_thresh = dst["d_thresh"][i].data_ptr()
    ws = runner.lm_workspace(desc, dev)
    sp, keep = _state_ptrs(saved["states_in"])
    _lib.check(L.snnflow_window_backward(ctypes.byref(desc), lp, pw_e.data_ptr(), sp, saved["arena"].data_ptr(),
                                         saved["flow"].data_ptr(), g_flow.data_ptr(), d_pw.data_ptr(), d_pb.data_ptr(),
                                         ws.data_ptr(), ws.numel(), _lib.stream()), "snnflow_window_backward")
    return dws, dlam, dtheta, d_pw, d_pb


class _LayerMajorWindowFn(torch.autograd.Function):
    """The window on the layer-major engine (snnflow_window_forward / snnflow_window_backward, include/snnflow.h)."""

    @staticmethod
    def forward(ctx, runner, cnt, *params):
        need_bwd = any(ctx.needs_input_grad)
        flow, saved = _lm_forward(runner, cnt, need_bwd)
        if need_bwd:
            ctx.runner, ctx.saved = runner, saved
        return flow

    @staticmethod
    def backward(ctx, g_flow):
        runner, saved = ctx.runner, ctx.saved
        layers, net = runner.layers, runner.net
        lam, theta, Cr, C = saved["lam"], saved["theta"], saved["Cr"], saved["desc"].C
        dws, dlam, dtheta, d_pw, d_pb = _lm_backward(runner, saved, g_flow)
        pb = net.pred.conv2d.bias
        if C != Cr:   # drop the padded neurons (their gradients are exactly zero)
            dlam, dtheta, d_pw = dlam[:, :Cr], dtheta[:, :Cr], d_pw[:, :Cr].contiguous()
            dws = [(f[:Cr, :(l.ff.weight.shape[1])].contiguous(), None if r is None else r[:Cr, :Cr].contiguous())
                   for (f, r), l in zip(dws, layers)]
        d_leak = dlam * lam * (1.0 - lam)                                           # sigmoid'
        thr = torch.stack([l.thresh.detach().reshape(-1) for l in layers])
        d_thresh = dtheta * (thr >= 0.01).float()                                   # clamp_min'
        grads = []
        for i, l in enumerate(layers):
            grads += [dws[i][0], dws[i][1], d_leak[i].reshape(l.leak.shape), d_thresh[i].reshape(l.thresh.shape)]
        grads += [d_pw, d_pb if pb is not None else None]
        return (None, None) + tuple(grads)


class WindowRunner:
    """Runs windows of T bins through a snnflow LIFFireNet / LIFFireFlowNet with one C call per direction."""

    engine = "auto"          # "auto": layer-major engine when it covers the shape, else per-step; "per_step"; "layer_major"
    validate_input = True    # check after EVERY eagerly launched window (one host sync) that all inputs were bf16-exact;
    validate_every = 64      # False: only on the first window and then every `validate_every`-th one.  CUDA-graph
                             # replays cannot sync: there the arena's sticky flag gates the fused optimizer update and
                             # train.TrainWindow polls it (check_input_flag) every `validate_every` replays
    param_override = None    # (lam [7,C], theta [7,C]) evaluated by the caller instead of sigmoid / clamp_min on the device
    input_flag = None        # int32 view of the arena's status word of the last layer-major window

    def __init__(self, net):
        self.net = net
        self._lm = {}
        self.layers = [net.head, net.G1, net.R1a, net.R1b, net.G2, net.R2a, net.R2b]
        self._arenas = {}
        self._flip = 0
        self._ws = None
        self.new_states = None

    def check_input_flag(self):
        """Raise if any window run in this runner's arena saw input values that one bf16 term cannot represent (host sync)."""
        if self.input_flag is None:
            return
        bad = int(self.input_flag.item())
        if bad:
            self.input_flag.zero_()
            raise _lib.SnnflowError(f"window engine: {bad} input values were not exactly representable in bfloat16 (event "
                                    "counts above 256 or fractional values, e.g. a voxel encoding): the results of this window "
                                    "are rounded and any gated optimizer update was skipped - use engine='per_step'")

    def supported(self):
        l0 = self.layers[0]
        same = all((l.hard_reset, l.detach, l.activation, l._act_width) ==
                   (l0.hard_reset, l0.detach, l0.activation, l0._act_width) for l in self.layers)
        return same and not self.net.residual and not self.net.norm_input and not self.layers[0].recurrent

    def arena(self, desc, save, dev):
        # two arenas per shape: the states of window k (views into arena k%2) feed window k+1 (arena (k+1)%2)
        n = _lib.lib().snnflow_net_acts_floats(ctypes.byref(desc), int(save))
        self._flip ^= 1
        key = (self._flip, bool(save))
        buf = self._arenas.get(key)
        if buf is None or buf.numel() < n or buf.device != dev:
            buf = torch.empty(n, dtype=torch.float32, device=dev)
            self._arenas[key] = buf
        return buf

    def workspace(self, desc, dev):
        n = _lib.lib().snnflow_net_bwd_workspace_bytes(ctypes.byref(desc))
        if self._ws is None or self._ws.numel() < n or self._ws.device != dev:
            self._ws = torch.empty(n, dtype=torch.uint8, device=dev)
        return self._ws

    def lm_states_in(self, states, Cr, C):
        """The network's states [2,B,Cr,H,W] as the engine's (zero-padded) [2,B,C,H,W] tensors.  States this runner
        handed out last window are channel slices of padded state blocks in the arena: those are passed back as they are."""
        if C == Cr:
            return states
        full = getattr(self, "_lm_full_states", None)
        out = []
        for i, st in enumerate(states):
            if st is None or st.shape[2] == C:
                out.append(st)
                continue
            f = full[i] if full is not None else None
            if f is not None and st.data_ptr() == f.data_ptr() and st.shape[2] == Cr and st.stride() == f[:, :, :Cr].stride():
                out.append(f)
            else:
                out.append(_pad_dim(st.detach(), 2, C))
        return out

    def layer_major_ok(self, cnt_window, backward):
        L = _lib.lib()
        _bind(L)
        desc = _make_desc(self, cnt_window, _lm_channels(self.layers[0].hidden_size))
        return bool(L.snnflow_window_supported(ctypes.byref(desc), int(backward)))

    def lm_arena(self, desc, save, dev):
        # zero-filled once per shape (the plane borders are never written).  ONE arena per shape: the forward call
        # copies the incoming state before it overwrites the state block, so the addresses never change (CUDA graphs)
        key = ("arena", _desc_key(desc), bool(save), str(dev))
        buf = self._lm.get(key)
        if buf is None:
            n = _lib.lib().snnflow_window_arena_bytes(ctypes.byref(desc), int(save))
            for k in [k for k in self._lm if k[0] == "arena" and k[1] != key[1]]:
                del self._lm[k]
            self._stream = None          # a streamed state lived in one of the arenas just dropped: the NCHW list is the truth
            self.stream_live = False
            buf = torch.zeros(n, dtype=torch.uint8, device=dev)
            self._lm[key] = buf
        return buf

    def lm_workspace(self, desc, dev):
        key = ("ws", _desc_key(desc), str(dev))
        buf = self._lm.get(key)
        if buf is None:
            n = _lib.lib().snnflow_window_workspace_bytes(ctypes.byref(desc))
            for k in [k for k in self._lm if k[0] == "ws"]:
                del self._lm[k]
            buf = torch.zeros(n, dtype=torch.uint8, device=dev)
            self._lm[key] = buf
        return buf

    # ---- streaming inference: one call per time bin, states kept in the engine's layout between calls --------------------
    # (the reference's eval loop, eval_flow.py:220; models/model.py:172-182 hands 7 NCHW states back in per call)
    stream_live = False     # the arena holds a state that is newer than the network's NCHW state list
    _stream = None

    def stream_ok(self, x):
        """x [B,num_bins,H,W]: the layer-major engine covers this shape at the network's own width, under no_grad."""
        if not x.is_cuda or x.dim() != 4 or self.engine == "per_step" or not self.supported():
            return False
        if getattr(self.net, "encoding", "cnt") != "cnt" or _lm_channels(self.layers[0].hidden_size) != self.layers[0].hidden_size:
            return False
        key = ("stream_ok", tuple(x.shape))
        ok = self._lm.get(key)
        if ok is None:
            ok = self._lm[key] = self.layer_major_ok(x[None], False)
        return ok

    def _stream_desc(self, x, extra_flags=0):
        desc = _make_desc(self, x[None])
        desc.flags |= _lib.STATE_INTERNAL | extra_flags
        return desc

    def stream_forward(self, x, states, graph=False):
        """One time bin through snnflow_window_forward (T = 1, SNNFLOW_STATE_INTERNAL).  `states`: the network's NCHW state
        list - imported into the arena only when it is newer than the arena's own state (first call, reset_states(), a
        state assigned by the caller).  Returns flow [B,2,H,W].
        graph=True (model.graph_forward()): the launches of a bin are replayed as a CUDA graph - one per phase of the
        recurrent layers' ping-pong slots - from a static copy of the input; the flow map returned is the graph's own
        buffer (overwritten two calls later).  A call that has to re-pack the weights runs eagerly."""
        L = _lib.lib()
        _bind(L)
        layers = self.layers
        B, nb, H, W = x.shape
        dev = x.device
        x = x.float().contiguous()
        desc0 = self._stream_desc(x)
        arena = self.lm_arena(desc0, False, dev)      # (a NEW arena drops self._stream: the list is imported below)
        st = self._stream
        wkey = tuple((p.data_ptr(), p._version) for l in layers for p in ((l.ff.weight, l.rec.weight, l.leak, l.thresh) if l.recurrent
                                                                        else (l.ff.weight, l.leak, l.thresh)))
        if st is None or st["shape"] != tuple(x.shape) or st["dev"] != dev:
            st = self._stream = dict(shape=tuple(x.shape), dev=dev, phase=0, wkey=None, lam=None, theta=None, graphs=None)
            self.stream_live = False
        if not self.stream_live:      # the NCHW list is the truth: bring it into the arena
            C = layers[0].hidden_size
            _state_ptrs(states, (B, C, H, W))
            arr = (ctypes.c_void_p * N_LAYERS)()
            keep = []
            for i, s_ in enumerate(states):
                if s_ is not None:
                    s_ = s_.detach().float().contiguous()
                    keep.append(s_)
                    arr[i] = s_.data_ptr()
            _lib.check(L.snnflow_window_import_state(ctypes.byref(desc0), arr, arena.data_ptr(), _lib.stream()),
                       "snnflow_window_import_state")
            st["phase"] = 0
        reuse = st["wkey"] == wkey
        if not reuse:
            st["lam"], st["theta"] = _effective_params(self, layers)
            st["wkey"] = wkey
        # the graphs bake device pointers in (weights, flow head): a parameter that moved to other storage drops them
        pred = self.net.pred.conv2d
        pkey = tuple(k[0] for k in wkey) + (pred.weight.data_ptr(), None if pred.bias is None else pred.bias.data_ptr())
        if st.get("pkey") != pkey:
            st["pkey"], st["graphs"] = pkey, None
        if graph and reuse and not torch.cuda.is_current_stream_capturing():
            gs = st.get("graphs")
            if gs is None:
                gs = st["graphs"] = self._stream_capture(x, st, arena)
            ph = st["phase"]
            gs["x"].copy_(x)
            gs["graph"][ph].replay()
            flow = gs["flow"][ph]
            st["phase"] = ph ^ 1
        else:
            flow = self._stream_launch(x, st, arena, reuse)
        foff = L.snnflow_window_flags_offset(ctypes.byref(desc0), 0)
        self.input_flag = arena[foff:foff + 4].view(torch.int32)
        st["calls"] = st.get("calls", 0) + 1
        if self.validate_input and not torch.cuda.is_current_stream_capturing() and (st["calls"] == 1 or st["calls"] % self.validate_every == 0):
            self.check_input_flag()               # streaming: polled, not synchronised every frame
        self.stream_live = True
        return flow

    def _stream_launch(self, x, st, arena, reuse):
        """The launches of one streamed bin (pack input, 7 layers, flow head); toggles the phase."""
        L = _lib.lib()
        layers, net = self.layers, self.net
        B, nb, H, W = x.shape
        lam, theta = st["lam"], st["theta"]
        desc = self._stream_desc(x, (_lib.STREAM_PHASE if st["phase"] else 0) | (_lib.REUSE_PACKED if reuse else 0))
        lp = (LayerPtrs * N_LAYERS)()
        for i, l in enumerate(layers):
            lp[i].w_ff = l.ff.weight.data_ptr()
            lp[i].w_rec = l.rec.weight.data_ptr() if l.recurrent else None
            lp[i].lam = lam[i].data_ptr()
            lp[i].theta = theta[i].data_ptr()
        flow = torch.empty((1, B, 2, H, W), dtype=torch.float32, device=x.device)
        pw, pb = net.pred.conv2d.weight, net.pred.conv2d.bias
        _lib.check(L.snnflow_window_forward(ctypes.byref(desc), lp, pw.data_ptr(), None if pb is None else pb.data_ptr(), x.data_ptr(),
                                            None, arena.data_ptr(), flow.data_ptr(), 0, _lib.stream()), "snnflow_window_forward")
        st["phase"] = (st["phase"] + 1) & 1       # T = 1
        return flow[0]

    def _stream_capture(self, x, st, arena):
        """Two CUDA graphs of one streamed bin: phase 0 and phase 1 (the slots the recurrent layers read / write alternate).
        Capturing executes nothing: the arena's state and the phase are unchanged afterwards."""
        sx = torch.empty_like(x)
        p0 = st["phase"]
        graphs, flows = {}, {}
        torch.cuda.synchronize()
        for ph in (p0, p0 ^ 1):
            st["phase"] = ph
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                flows[ph] = self._stream_launch(sx, st, arena, True)
            graphs[ph] = g
        st["phase"] = p0
        return dict(x=sx, graph=graphs, flow=flows)

    def stream_export(self):
        """The arena's state as the reference's list of 7 tensors [2,B,C,H,W] = stack([v, z])."""
        L = _lib.lib()
        st = self._stream
        B, nb, H, W = st["shape"]
        C = self.layers[0].hidden_size
        x_like = torch.empty((1,) + st["shape"], device="meta")
        desc = _make_desc(self, x_like)
        desc.flags |= _lib.STATE_INTERNAL
        arena = self.lm_arena(desc, False, st["dev"])
        out = [torch.empty((2, B, C, H, W), dtype=torch.float32, device=st["dev"]) for _ in range(N_LAYERS)]
        arr = (ctypes.c_void_p * N_LAYERS)(*[o.data_ptr() for o in out])
        _lib.check(L.snnflow_window_export_state(ctypes.byref(desc), arena.data_ptr(), st["phase"], arr, _lib.stream()),
                   "snnflow_window_export_state")
        return out

    # ---- direct (autograd-free) training interface: train.TrainWindow.step_direct ------------------------------------
    def direct_ok(self, cnt_window, optimizer):
        """The direct training step covers: layer-major engine, un-padded width, every parameter trainable and managed by
        a FusedClipAdam whose flat gradient buffer receives the kernels' output."""
        layers = self.layers
        if not (self.supported() and self.engine != "per_step" and getattr(self.net, "encoding", "cnt") == "cnt"):
            return False
        if _lm_channels(layers[0].hidden_size) != layers[0].hidden_size or not hasattr(optimizer, "grad_view"):
            return False
        ps = [p for l in layers for p in ((l.ff.weight, l.rec.weight, l.leak, l.thresh) if l.recurrent else (l.ff.weight, l.leak, l.thresh))]
        ps += [self.net.pred.conv2d.weight, self.net.pred.conv2d.bias]
        managed = {id(p) for p in optimizer.params}
        if any(p is None or not isinstance(p, torch.nn.Parameter) or id(p) not in managed for p in ps) or len(managed) != len(ps):
            return False
        return self.layer_major_ok(cnt_window, True)

    def direct_forward(self, cnt_window):
        flow, self._direct_saved = _lm_forward(self, cnt_window, True)
        self.net._states = self.new_states
        return flow

    def direct_backward(self, g_flow, optimizer):
        """BPTT of the window run by direct_forward(): every parameter gradient is accumulated by the kernels straight into
        its slice of ``optimizer.grad`` (zeroed here); the sigmoid / clamp_min chain rule of d leak / d thresh is applied by
        the final reduction launch (snnflow_layer_ptrs.d_leak / d_thresh)."""
        saved, layers, net = self._direct_saved, self.layers, self.net
        self._direct_saved = None
        optimizer.grad.zero_()
        dst = dict(dw=[(optimizer.grad_view(l.ff.weight), optimizer.grad_view(l.rec.weight) if l.recurrent else None) for l in layers],
                   dlam=None, dtheta=None, d_pw=optimizer.grad_view(net.pred.conv2d.weight),
                   d_pb=optimizer.grad_view(net.pred.conv2d.bias),
                   d_leak=[optimizer.grad_view(l.leak) for l in layers], d_thresh=[optimizer.grad_view(l.thresh) for l in layers])
        _lm_backward(self, saved, g_flow, dst)

    def __call__(self, cnt_window):
        """cnt_window [T,B,num_bins,H,W] -> flow [T,B,2,H,W]; updates net._states like T forward calls would."""
        if not cnt_window.is_cuda:
            raise _lib.SnnflowError("snnflow WindowRunner runs on CUDA tensors only (no CPU fallback)")
        self.net._states                      # (getter: materialises a streamed state before the arenas change hands)
        params = []
        for l in self.layers:
            params += [l.ff.weight, l.rec.weight if l.recurrent else None, l.leak, l.thresh]
        params += [self.net.pred.conv2d.weight, self.net.pred.conv2d.bias]
        need_bwd = torch.is_grad_enabled() and any(p is not None and p.requires_grad for p in params)
        # the layer-major engine carries inputs as ONE bf16 term: event counts are exact, a voxel grid (fractional
        # weights, encodings.py:48-67) is not - such networks run on the per-step engine (exact fp32 head layer)
        exact_input = getattr(self.net, "encoding", "cnt") == "cnt"
        use_lm = self.engine != "per_step" and (exact_input or self.engine == "layer_major") and self.layer_major_ok(cnt_window, need_bwd)
        if self.engine == "layer_major" and not use_lm:
            raise _lib.SnnflowError("the layer-major window engine does not cover this shape / these options")
        if not use_lm and self.engine == "auto" and not getattr(self, "_warned_fallback", False):
            import warnings
            self._warned_fallback = True
            warnings.warn("snnflow: this window runs on the per-step engine (shape / options / input encoding outside the "
                          "layer-major engine's envelope: C in {16, 32}; W <= 256 or a multiple of 128 without autograd, W <= 128 with it "
                          "(256 for feed-forward networks); detached reset; count encoding)")
        fn = _LayerMajorWindowFn if use_lm else _WindowFn
        if not torch.is_grad_enabled():      # Function.forward sees requires_grad flags, not the grad mode: nothing to save
            params = [p if p is None else p.detach() for p in params]
        flow = fn.apply(self, cnt_window, *params)
        self.net._states = self.new_states
        return flow
