"""Device-side event branch of the loader: raw event windows -> the tensors of one batch, in one C call.

Mirrors the event-facing half of ``BaseDataLoader`` / ``H5Loader.__getitem__`` (dataloader/base.py:14-46,53-126,
160-256; dataloader/h5.py:285-331,375-410; collate base.py:261-278): same configuration keys, same per-slot
augmentation flags (drawn from ``np.random`` in the reference's order, so a seeded run flips the same slots), same
hot-pixel filter state, same output dictionary.  The HDF5 file handling (which events form a window) stays with the
caller: it hands over the ``B`` windows of a batch as ``[B,N]`` device tensors, e.g. slices of an event stream that
is already resident in HBM.  No CPU fallback.
"""
import ctypes
from ctypes import c_float, c_int32, c_int64

import numpy as np
import torch

from . import _lib


class LoaderDesc(ctypes.Structure):   # snnflow_loader_desc (include/snnflow.h)
    _fields_ = [("B", c_int32), ("H", c_int32), ("W", c_int32), ("N", c_int64), ("num_bins", c_int32),
                ("round_ts", c_int32), ("pool_h", c_int32), ("pool_w", c_int32), ("target_h", c_int32),
                ("target_w", c_int32), ("hot_enabled", c_int32), ("hot_max_px", c_int32), ("hot_min_obvs", c_int32),
                ("hot_max_rate", c_float)]


_bound = False


def _fn():
    global _bound
    lib = _lib.lib()
    if not _bound:
        P = ctypes.c_void_p
        lib.snnflow_format_window_workspace_bytes.restype = ctypes.c_size_t
        lib.snnflow_format_window_workspace_bytes.argtypes = [ctypes.POINTER(LoaderDesc)]
        lib.snnflow_format_window.restype = ctypes.c_int
        lib.snnflow_format_window.argtypes = [ctypes.POINTER(LoaderDesc), P, P, P, ctypes.c_int] + [P] * 11 + [
            ctypes.c_size_t, P]
        _bound = True
    return lib


class EventWindowFormatter:
    """``EventWindowFormatter(config, num_bins, round_encoding=False, device="cuda")``.

    ``config`` is the reference's loader configuration (configs/parser.py): ``data.mode``, ``loader.resolution``,
    ``loader.std_resolution``, ``loader.batch_size``, ``loader.augment`` / ``augment_prob``, ``hot_filter.*``.
    """

    def __init__(self, config, num_bins, round_encoding=False, device="cuda"):
        self.config = config
        self.num_bins = num_bins
        self.round_encoding = round_encoding
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.SnnflowError("EventWindowFormatter runs on a CUDA device only (no CPU fallback)")
        self.seq_num = 0
        self._hot_reset = set()
        # base.py:24-27: events are encoded at loader.resolution in "events" mode, at std_resolution otherwise ...
        if config["data"]["mode"] == "events":
            self.resolution = list(config["loader"]["resolution"])
        else:
            self.resolution = list(config["loader"]["std_resolution"])
        # ... and down-sampled to loader.resolution when that is smaller (h5.py:294-316,375-382)
        self.target = list(config["loader"]["resolution"])
        H, W = self.resolution
        if self.target[0] < H or self.target[1] < W:
            self.pool = (H // self.target[0], W // self.target[1])
            if self.pool[0] == 0 or self.pool[1] == 0:
                raise ValueError(f"Invalid pooling kernel size: pool_h={self.pool[0]}, pool_w={self.pool[1]}. "
                                 f"Original size: ({H}, {W}), Target size: ({self.target[0]}, {self.target[1]})")
        else:
            self.pool = (1, 1)
        B = config["loader"]["batch_size"]
        self.batch_size = B
        # base.py:29-38: per-slot augmentation flags, drawn mechanism-major from np.random
        self.batch_augmentation = {m: [False] * B for m in config["loader"]["augment"]}
        for i, m in enumerate(config["loader"]["augment"]):
            for b in range(B):
                if np.random.random() < config["loader"]["augment_prob"][i]:
                    self.batch_augmentation[m][b] = True
        self._flips = None
        # base.py:40-45: hot-pixel filter state
        hf = config["hot_filter"]
        self.hot_enabled = bool(hf["enabled"])
        self.hot_idx = self.hot_events = None      # device state, allocated by the first format_batch()
        self._ws = None

    def _hot_state(self):
        if self.hot_idx is None:
            B, (H, W) = self.batch_size, self.resolution
            self.hot_idx = torch.zeros(B, dtype=torch.int32, device=self.device)
            self.hot_events = torch.zeros((B, H, W), dtype=torch.float32, device=self.device) if self.hot_enabled else None
            self._hot_reset.clear()
        for b in sorted(self._hot_reset):
            self.hot_idx[b] = 0
            if self.hot_events is not None:
                self.hot_events[b].zero_()
        self._hot_reset.clear()

    # ---- sequence bookkeeping (base.py:53-69) ----
    def reset_sequence(self, batch):
        self.seq_num += 1
        if self.hot_enabled:
            self._hot_reset.add(batch)             # applied on the device before the next window is formatted
        for i, m in enumerate(self.config["loader"]["augment"]):
            self.batch_augmentation[m][batch] = bool(np.random.random() < self.config["loader"]["augment_prob"][i])
        self._flips = None

    def _flip_tensor(self):
        if self._flips is None:
            rows = [[int(self.batch_augmentation.get(m, [False] * self.batch_size)[b])
                     for m in ("Horizontal", "Vertical", "Polarity")] for b in range(self.batch_size)]
            self._flips = torch.tensor(rows, dtype=torch.int32, device=self.device)
        return self._flips

    def _desc(self, N):
        hf = self.config["hot_filter"]
        return LoaderDesc(B=self.batch_size, H=self.resolution[0], W=self.resolution[1], N=N, num_bins=self.num_bins,
                          round_ts=int(bool(self.round_encoding)), pool_h=self.pool[0], pool_w=self.pool[1],
                          target_h=self.target[0], target_w=self.target[1], hot_enabled=int(self.hot_enabled),
                          hot_max_px=int(hf.get("max_px", 100)), hot_min_obvs=int(hf.get("min_obvs", 5)),
                          hot_max_rate=float(hf.get("max_rate", 0.8)))

    def format_batch(self, xs, ys, ts, ps, t0=None):
        """``xs, ys, ps``: ``[B,N]`` CUDA tensors (sensor coordinates, raw polarity in {0,1}; any real dtype);
        ``ts``: ``[B,N]`` float64 (absolute seconds, ``t0`` = ``[B]`` sequence start times subtracted in float64 like
        ``H5Loader.get_events``) or float32 (already relative).  Returns the collated batch dictionary
        ``event_cnt [B,2,h,w]``, ``event_voxel [B,num_bins,h,w]``, ``event_mask [B,1,h,w]``, ``event_list [B,N,4]``,
        ``event_list_pol_mask [B,N,2]``."""
        for t in (xs, ys, ts, ps):
            if not t.is_cuda:
                raise _lib.SnnflowError("EventWindowFormatter needs CUDA tensors (no CPU fallback)")
        if xs.dim() != 2 or xs.shape[0] != self.batch_size or not (xs.shape == ys.shape == ts.shape == ps.shape):
            raise ValueError(f"expected four [B={self.batch_size}, N] tensors, got {tuple(xs.shape)}, {tuple(ys.shape)}, "
                             f"{tuple(ts.shape)}, {tuple(ps.shape)}")
        B, N = xs.shape
        dev = xs.device
        xs, ys, ps = xs.float().contiguous(), ys.float().contiguous(), ps.float().contiguous()
        ts64 = ts.dtype == torch.float64
        ts = ts.contiguous() if ts64 else ts.float().contiguous()
        if t0 is not None:
            if not ts64:
                raise ValueError("t0 is subtracted in float64: pass float64 timestamps with it")
            t0 = torch.as_tensor(t0, dtype=torch.float64, device=dev).reshape(B).contiguous()
        self._hot_state()
        d = self._desc(N)
        H, W = self.resolution
        h, w = H // self.pool[0], W // self.pool[1]
        lib = _fn()
        need = lib.snnflow_format_window_workspace_bytes(ctypes.byref(d))
        if self._ws is None or self._ws.numel() < need or self._ws.device != dev:
            self._ws = torch.empty(need, dtype=torch.uint8, device=dev)
        out = {
            "event_cnt": torch.empty((B, 2, h, w), dtype=torch.float32, device=dev),
            "event_voxel": torch.empty((B, self.num_bins, h, w), dtype=torch.float32, device=dev),
            "event_mask": torch.empty((B, 1, h, w), dtype=torch.float32, device=dev),
            "event_list": torch.empty((B, N, 4), dtype=torch.float32, device=dev),
            "event_list_pol_mask": torch.empty((B, N, 2), dtype=torch.float32, device=dev),
        }
        p = _lib.ptr
        _lib.check(lib.snnflow_format_window(
            ctypes.byref(d), p(xs), p(ys), p(ts), int(ts64), p(t0), p(ps), p(self._flip_tensor()),
            p(self.hot_events), p(self.hot_idx), p(out["event_cnt"]), p(out["event_voxel"]), p(out["event_mask"]),
            p(out["event_list"]) if N else None, p(out["event_list_pol_mask"]) if N else None, p(self._ws),
            self._ws.numel(), _lib.stream()), "snnflow_format_window")
        return out
