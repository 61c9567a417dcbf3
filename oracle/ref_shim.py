"""Import shim that makes the read-only reference importable in the build container.

The reference imports snntorch / brevitas / matplotlib / h5py / hdf5plugin / progress
at module top but only uses them on branches outside the hot path (quantised layers,
snn.Leaky cells, plotting, HDF5 IO).  None of them is installed here, so they are
replaced by empty stub modules.  Used by ``oracle/make_golden.py``, by the tests that compare the
oracle / the CUDA cells with the live reference, and by the reference arm of ``bench.py``.  On the GPU
box ``/root/reference`` does not exist: there the shim resolves to the unmodified copy of the hot-path
files that ``oracle/stage_reference.py`` stages under the git-ignored ``baseline/_ref/``.
"""
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
STAGED_ROOT = os.path.join(os.path.dirname(_HERE), "baseline", "_ref")   # written by oracle/stage_reference.py


def _pick_root():
    """The reference checkout: $SNNFLOW_REFERENCE_ROOT, /root/reference (build container), or the unmodified copy of the
    hot-path files staged under the git-ignored baseline/_ref (the only one that exists on the GPU box)."""
    for cand in (os.environ.get("SNNFLOW_REFERENCE_ROOT"), "/root/reference", STAGED_ROOT):
        if cand and os.path.isfile(os.path.join(cand, "models", "spiking_submodules.py")):
            return cand
    return "/root/reference"


REFERENCE_ROOT = _pick_root()


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "models", "spiking_submodules.py"))


def full_checkout() -> bool:
    """True for a complete checkout (configs, train/eval scripts), False for the staged hot-path subset."""
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "train_flow.py"))


class _Raising:
    def __init__(self, *a, **k):
        raise NotImplementedError("stubbed third-party class (not on the hot path)")


def _stub(name, **attrs):
    if name in sys.modules and not getattr(sys.modules[name], "__snnflow_stub__", False):
        return sys.modules[name]  # a real install wins
    mod = types.ModuleType(name)
    mod.__snnflow_stub__ = True
    mod.__path__ = []  # behave like a package
    for k, v in attrs.items():
        setattr(mod, k, v)
    sys.modules[name] = mod
    return mod


def install():
    """Install the stubs and put the reference root on sys.path. Idempotent."""
    if not reference_available():
        raise RuntimeError(f"reference not found at {REFERENCE_ROOT}")
    _stub("snntorch", Leaky=_Raising, surrogate=types.SimpleNamespace(atan=lambda *a, **k: None))
    _stub("snntorch.functional", quant=None)
    _stub("brevitas")
    _stub("brevitas.nn", QuantConv2d=_Raising, QuantIdentity=_Raising, QuantTanh=_Raising, QuantReLU=_Raising)
    _stub("brevitas.quant", Int8WeightPerTensorFloat=object, Int8ActPerTensorFloat=object, Int8Bias=object,
          Uint8ActPerTensorFloat=object)
    _stub("brevitas.nn.quant_layer", QuantLayerMixin=object)
    _stub("brevitas.core")
    _stub("brevitas.core.quant", QuantType=object)
    _stub("matplotlib", use=lambda *a, **k: None)
    _stub("matplotlib.pyplot")
    _stub("matplotlib.cm")
    _stub("matplotlib.patches")
    _stub("h5py")
    _stub("hdf5plugin")
    _stub("progress")
    _stub("progress.bar", Bar=object)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


def load():
    """Return a namespace with the reference symbols on the hot path."""
    install()
    import importlib

    ss = importlib.import_module("models.spiking_submodules")
    su = importlib.import_module("models.spiking_util")
    mm = importlib.import_module("models.model")
    enc = importlib.import_module("dataloader.encodings")
    iwe = importlib.import_module("utils.iwe")
    fl = importlib.import_module("loss.flow")

    def adapt(cell):
        # class-attribute seam of models/model.py:37-39: swallow the kwargs LIFFireNet passes
        # (exporting, tebn, num_timesteps, mpbn) and the forward kwargs residual=/timestep=.
        import inspect
        takes_residual = "residual" in inspect.signature(cell.forward).parameters

        class Adapted(cell):
            def __init__(self, *a, exporting=False, tebn=False, num_timesteps=4, mpbn=False,
                         quantization_config=None, **k):
                if not quantization_config:
                    quantization_config = None
                super().__init__(*a, quantization_config=quantization_config, **k)

            def forward(self, input_, prev_state, residual=0, timestep=None):
                if takes_residual:
                    return super().forward(input_, prev_state, residual=residual)
                return super().forward(input_, prev_state)

        Adapted.__name__ = "Adapted" + cell.__name__
        return Adapted

    return types.SimpleNamespace(
        ConvLIF=ss.ConvLIF, ConvLIFRecurrent=ss.ConvLIFRecurrent, spiking_util=su, model=mm,
        encodings=enc, iwe=iwe, flow=fl, adapt=adapt,
    )
