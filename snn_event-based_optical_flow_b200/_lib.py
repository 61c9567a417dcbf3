"""ctypes binding of libsnnflow.so (include/snnflow.h).  No fallback: a missing library is an error."""
import ctypes
import os
from ctypes import c_float, c_int, c_int64, c_size_t, c_uint, c_uint64, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libsnnflow.so")

HARD_RESET = 1
DETACH_RESET = 2
NO_TENSOR_CORES = 4
INPUT_EXACT16 = 8
STATE_INTERNAL = 16
STREAM_PHASE = 32
REUSE_PACKED = 64
SURROGATE_ID = {"arctanspike": 0, "superspike": 1, "trianglespike": 2, "mgspike": 3}

P = c_void_p
_PROTOS = {
    "snnflow_abi_version": (c_int, []),
    "snnflow_last_error": (ctypes.c_char_p, []),
    "snnflow_launch_count": (c_uint64, []),
    "snnflow_profile_enable": (c_int, [c_int]),
    "snnflow_profile_summary": (c_int, [ctypes.c_char_p, c_size_t]),
    "snnflow_convlif_fwd": (c_int, [P] * 12 + [c_int] * 5 + [c_uint, P]),
    "snnflow_convlif_packed_bytes": (c_size_t, [c_int] * 3),
    "snnflow_convlif_pack": (c_int, [P] * 3 + [c_int, c_int, P]),
    "snnflow_convlif_fwd_tc": (c_int, [P, P, c_int] + [P] * 9 + [c_int] * 5 + [c_uint, P]),
    "snnflow_tc_inexact_count": (c_uint, [c_int]),
    "snnflow_convlif_bwd_workspace_bytes": (c_size_t, [c_int] * 6),
    "snnflow_convlif_bwd": (c_int, [P] * 19 + [P, c_size_t] + [c_int] * 5 + [c_uint, c_int, c_float, P]),
    "snnflow_pred_fwd": (c_int, [P] * 4 + [c_int] * 4 + [P]),
    "snnflow_pred_bwd_workspace_bytes": (c_size_t, [c_int] * 4),
    "snnflow_pred_bwd": (c_int, [P] * 7 + [P, c_size_t] + [c_int] * 4 + [P]),
    "snnflow_encode_cnt": (c_int, [P] * 4 + [c_int64, c_int, c_int, c_int, P]),
    "snnflow_encode_image": (c_int, [P] * 5 + [c_int64, c_int, c_int, c_int, P]),
    "snnflow_encode_voxel": (c_int, [P] * 6 + [c_int64, c_int, c_int, c_int, c_int, P]),
    "snnflow_flow_gather_fwd": (c_int, [P] * 3 + [c_int, c_int64, c_int, c_int, P]),
    "snnflow_flow_gather_bwd": (c_int, [P] * 3 + [c_int, c_int64, c_int, c_int, P]),
    "snnflow_iwe_splat_fwd": (c_int, [P] * 5 + [c_int, c_int64, c_int, c_int, c_float, c_float, c_int, c_int, c_float,
                                                c_int, P]),
    "snnflow_leaky_fwd": (c_int, [P] * 7 + [c_int] * 5 + [P]),
    "snnflow_leaky_bwd_workspace_bytes": (c_size_t, [c_int] * 4),
    "snnflow_leaky_bwd": (c_int, [P] * 8 + [P, c_size_t] + [c_int] * 5 + [P]),
    "snnflow_clip_adam_partials": (c_int, [c_int64]),
    "snnflow_clip_adam": (c_int, [P] * 4 + [c_int64] + [P] * 7),
    "snnflow_dp_allreduce_ctas": (c_int, []),
    "snnflow_dp_allreduce_sum": (c_int, [P, P, P, P, c_int, c_int, c_size_t, P]),
    "snnflow_dp_clip_adam_ctas": (c_int, []),
    "snnflow_dp_clip_adam_max_n": (c_int64, []),
    "snnflow_dp_clip_adam": (c_int, [P, P, P, P, c_int, c_int, c_int64, c_int64] + [P] * 11 + [P]),
    "snnflow_dp_clip_adam_emulated": (c_int, [P, P, P, c_int, c_int64, c_int64, P]),
    "snnflow_window_loss_workspace_bytes": (c_size_t, [c_int, c_int, c_int64, c_int, c_int]),
    "snnflow_window_loss": (c_int, [P] * 7 + [c_size_t, c_int, c_int, c_int64, c_int, c_int, c_float, c_float, c_int, P]),
    "snnflow_iwe_splat_bwd": (c_int, [P] * 5 + [c_int, c_int64, c_int, c_int, c_float, c_float, c_int, c_int, c_float,
                                                P]),
}

_lib = None


class SnnflowError(RuntimeError):
    pass


_ENGINE_SYMBOLS = ["snnflow_net_acts_floats", "snnflow_net_bwd_workspace_bytes", "snnflow_net_forward",
                   "snnflow_net_backward", "snnflow_window_supported", "snnflow_window_arena_bytes",
                   "snnflow_window_workspace_bytes", "snnflow_window_state_offsets", "snnflow_window_flags_offset",
                   "snnflow_window_forward", "snnflow_window_backward", "snnflow_window_import_state",
                   "snnflow_window_export_state",
                   "snnflow_format_window_workspace_bytes", "snnflow_format_window"]   # struct-taking entry points, bound in engine.py / loader.py


def exported_symbols():
    return sorted(list(_PROTOS) + _ENGINE_SYMBOLS)


def lib():
    """Load libsnnflow.so once.  Raises if it has not been built (python __graft_entry__.py / build.py)."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise SnnflowError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
                "There is no CPU fallback.")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _PROTOS.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().snnflow_last_error().decode(errors="replace")
        raise SnnflowError(f"{what} failed ({rc}): {msg}")


def ptr(t):
    """Device pointer of a tensor (None -> NULL).  The tensor must be a contiguous fp32/int CUDA tensor."""
    if t is None:
        return None
    if not t.is_cuda:
        raise SnnflowError("snnflow kernels need CUDA tensors (no CPU fallback)")
    if not t.is_contiguous():
        raise SnnflowError("snnflow kernels need contiguous tensors")
    return t.data_ptr()


def stream():
    import torch

    return torch.cuda.current_stream().cuda_stream


def launch_count():
    return int(lib().snnflow_launch_count())


def profile(on=True):
    check(lib().snnflow_profile_enable(int(on)), "snnflow_profile_enable")


def profile_summary():
    """{kernel: dict(launches, ms, bytes, flops)} for the launches recorded since profile(True)."""
    buf = ctypes.create_string_buffer(1 << 16)
    check(lib().snnflow_profile_summary(buf, len(buf)), "snnflow_profile_summary")
    out = {}
    for line in buf.value.decode().splitlines():
        name, n, ms, by, fl = line.split()
        out[name] = dict(launches=int(n), ms=float(ms), bytes=float(by), flops=float(fl))
    return out
