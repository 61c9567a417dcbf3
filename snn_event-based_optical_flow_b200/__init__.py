"""snnflow: B200-native hot path of spiking-FireNet event-based optical flow.

Drop-in cells (``ConvLIF`` / ``ConvLIFRecurrent``), event encodings, image-of-warped-events kernels and
the contrast loss, each mirroring the interface of LSquarzoni/SNN_Event-based_Optical_Flow and running as
hand-written sm_100a CUDA behind the C ABI of ``include/snnflow.h`` (``libsnnflow.so``).
There is no CPU or PyTorch fallback: using these ops without the built library or on CPU tensors raises.
"""
from . import _lib  # noqa: F401
from .spiking_submodules import ConvLIF, ConvLIFRecurrent  # noqa: F401
from .snntorch_submodules import SNNtorch_ConvLIF, SNNtorch_ConvLIFRecurrent  # noqa: F401
from .submodules import ConvLayer  # noqa: F401
from .model import LIFFireNet, LIFFireFlowNet, SNNtorchLIFFireNet  # noqa: F401
from . import encodings, iwe  # noqa: F401
from .flow_loss import EventWarping  # noqa: F401
from .loader import EventWindowFormatter  # noqa: F401
from .optim import FusedClipAdam  # noqa: F401

__all__ = ["ConvLIF", "ConvLIFRecurrent", "SNNtorch_ConvLIF", "SNNtorch_ConvLIFRecurrent", "ConvLayer", "LIFFireNet",
           "LIFFireFlowNet", "SNNtorchLIFFireNet", "EventWarping",
           "EventWindowFormatter", "FusedClipAdam", "encodings", "iwe"]
