"""Flow prediction head: host-side mirror of models/submodules.py:16-113 (ConvLayer) for the one
configuration LIFFireNet uses (models/model.py:105-107): 1x1 convolution + bias + tanh.
State-dict keys stay ``conv2d.weight`` / ``conv2d.bias``.  The arithmetic is ``snnflow_pred_fwd`` /
``snnflow_pred_bwd`` (include/snnflow.h)."""
import torch
import torch.nn as nn

from . import _lib
from .spiking_submodules import _f32c, _workspace


class _PredHead(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias):
        L = _lib.lib()
        x = _f32c(x)
        B, C, H, W = x.shape
        w2 = _f32c(weight).reshape(2, C)
        b = _f32c(bias) if bias is not None else None
        flow = torch.empty((B, 2, H, W), dtype=torch.float32, device=x.device)
        _lib.check(L.snnflow_pred_fwd(_lib.ptr(x), _lib.ptr(w2), _lib.ptr(b), _lib.ptr(flow), B, C, H, W,
                                      _lib.stream()), "snnflow_pred_fwd")
        ctx.save_for_backward(x, w2, flow)
        ctx.has_bias = bias is not None
        ctx.wshape = weight.shape
        return flow

    @staticmethod
    def backward(ctx, g_flow):
        L = _lib.lib()
        x, w2, flow = ctx.saved_tensors
        B, C, H, W = x.shape
        g_flow = _f32c(g_flow)
        g_x = torch.empty_like(x)
        dw = torch.zeros_like(w2)
        db = torch.zeros(2, dtype=torch.float32, device=x.device)
        ws = _workspace(x.device, L.snnflow_pred_bwd_workspace_bytes(B, C, H, W))
        _lib.check(L.snnflow_pred_bwd(_lib.ptr(x), _lib.ptr(w2), _lib.ptr(flow), _lib.ptr(g_flow), _lib.ptr(g_x),
                                      _lib.ptr(dw), _lib.ptr(db), ws.data_ptr(), ws.numel(), B, C, H, W,
                                      _lib.stream()), "snnflow_pred_bwd")
        return g_x, dw.reshape(ctx.wshape), (db if ctx.has_bias else None)


class ConvLayer(nn.Module):
    """1x1 conv + bias + tanh flow head (models/submodules.py:16-113 restricted to what FireNet uses)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, activation="relu", norm=None,
                 BN_momentum=0.1, w_scale=None, quantization_config=None, exporting=False):
        super().__init__()
        if kernel_size != 1 or stride != 1 or out_channels != 2 or activation != "tanh" or norm is not None:
            raise NotImplementedError("snnflow ConvLayer covers the FireNet flow head only: 1x1, 2 channels, tanh")
        if quantization_config and quantization_config.get("enabled", False):
            raise NotImplementedError("snnflow ConvLayer: quantised head is out of scope")
        self.conv2d = nn.Conv2d(in_channels, out_channels, kernel_size, stride, 0, bias=True)
        if w_scale is not None:
            nn.init.uniform_(self.conv2d.weight, -w_scale, w_scale)
            nn.init.zeros_(self.conv2d.bias)

    def forward(self, x):
        if not x.is_cuda:
            raise _lib.SnnflowError("snnflow ConvLayer runs on CUDA tensors only (no CPU fallback)")
        return _PredHead.apply(x, self.conv2d.weight, self.conv2d.bias)
