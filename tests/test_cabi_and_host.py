"""CPU: the C-ABI library builds/loads and exports every symbol include/snnflow.h declares; the host-side
mirrors match the reference's interface (constructor, parameter names, RNG-identical init)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "snnflow.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(snnflow_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as ge
    ge.build()
    from snnflow_b200 import _lib
    handle = ctypes.CDLL(_lib.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(handle, s), s
    assert set(syms) == set(_lib.exported_symbols()), set(syms) ^ set(_lib.exported_symbols())
    handle.snnflow_abi_version.restype = ctypes.c_int
    assert handle.snnflow_abi_version() == 1


def test_no_cpu_fallback():
    import snnflow_b200 as snnflow
    from snnflow_b200 import _lib
    layer = snnflow.ConvLIF(2, 4, 3)
    with pytest.raises(_lib.SnnflowError):
        layer(torch.zeros(1, 2, 8, 8), None)
    with pytest.raises(_lib.SnnflowError):
        snnflow.encodings.events_to_channels(torch.zeros(3), torch.zeros(3), torch.ones(3), (4, 4))
    with pytest.raises(_lib.SnnflowError):
        snnflow.ConvLayer(4, 2, 1, activation="tanh")(torch.zeros(1, 4, 8, 8))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "snn_event-based_optical_flow_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "oracle" not in src.replace("the CPU oracle", ""), f"{f} references the oracle"


def test_unsupported_options_raise():
    import snnflow_b200 as snnflow
    for kw in (dict(stride=2), dict(norm="weight"), dict(norm="group"), dict(activation="no_such_spike"),
               dict(quantization_config={"enabled": True})):
        with pytest.raises(NotImplementedError):
            snnflow.ConvLIF(4, 4, 3, **kw)
    with pytest.raises(NotImplementedError):
        snnflow.ConvLIFRecurrent(4, 4, 5)
    snnflow.ConvLIF(4, 4, 3, activation="mgspike")   # all four surrogates of models/spiking_util.py:96-109
    for kw in (dict(tebn=True), dict(mpbn=True), dict(detach=False), dict(norm="group")):
        with pytest.raises(NotImplementedError):
            snnflow.SNNtorch_ConvLIF(4, 4, 3, **kw)
    # LIFFireNet passes quantization_config={} and these extra kwargs (models/model.py:71,83)
    snnflow.ConvLIF(4, 4, 3, quantization_config={}, exporting=False, tebn=False, num_timesteps=4, mpbn=False)


from oracle import ref_shim

needs_ref = pytest.mark.skipif(not ref_shim.reference_available(), reason="/root/reference not present")


@needs_ref
@pytest.mark.parametrize("rec", [False, True])
def test_cell_init_identical_to_reference(rec):
    import snnflow_b200 as snnflow
    ref = ref_shim.load()
    kw = dict(leak=(0.0, 1.0), thresh=(0.3, 0.1), learn_thresh=False)
    torch.manual_seed(3)
    a = (ref.ConvLIFRecurrent if rec else ref.ConvLIF)(5, 8, 3, **kw)
    torch.manual_seed(3)
    b = (snnflow.ConvLIFRecurrent if rec else snnflow.ConvLIF)(5, 8, 3, **kw)
    sa, sb = a.state_dict(), b.state_dict()
    assert list(sa) == list(sb)
    for k in sa:
        assert torch.equal(sa[k], sb[k]), k
    assert [n for n, _ in a.named_parameters()] == [n for n, _ in b.named_parameters()]
    assert (a.input_size, a.hidden_size, a.hard_reset, a.detach) == (b.input_size, b.hidden_size, b.hard_reset, b.detach)


@needs_ref
@pytest.mark.parametrize("kind", ["LIFFireNet", "LIFFireFlowNet"])
def test_network_seam_and_state_dict(kind):
    """Our cells install under the reference's own network classes through the class attributes
    (models/model.py:37-39), and our mirror network has the same state_dict as the reference's."""
    import snnflow_b200 as snnflow
    ref = ref_shim.load()
    base = getattr(ref.model, kind)

    class Net(base):
        head_neuron = snnflow.ConvLIF
        ff_neuron = snnflow.ConvLIF
        rec_neuron = snnflow.ConvLIFRecurrent if kind == "LIFFireNet" else snnflow.ConvLIF

    cfg = dict(num_bins=2, encoding="cnt", mask_output=False, spiking_neuron=None, base_num_channels=8,
               kernel_size=3, activations=["arctanspike", "arctanspike"], quantization={"enabled": False})
    torch.manual_seed(11)
    seam = Net(dict(cfg))
    A, AR = ref.adapt(ref.ConvLIF), ref.adapt(ref.ConvLIFRecurrent)

    class RefNet(base):
        head_neuron = A
        ff_neuron = A
        rec_neuron = AR if kind == "LIFFireNet" else A

    torch.manual_seed(11)
    refnet = RefNet(dict(cfg))
    torch.manual_seed(11)
    mirror = getattr(snnflow, kind)(dict(num_bins=2, encoding="cnt", base_num_channels=8, kernel_size=3))
    for other in (seam, mirror):
        sa, sb = refnet.state_dict(), other.state_dict()
        assert list(sa) == list(sb)
        for k in sa:
            assert torch.equal(sa[k], sb[k]), k
    # forward on CPU must refuse loudly (no silent fallback) rather than compute something else
    from snnflow_b200 import _lib
    with pytest.raises(_lib.SnnflowError):
        seam(None, torch.zeros(1, 2, 8, 8))


@pytest.mark.skipif(not ref_shim.reference_available(), reason="needs the reference checkout")
def test_window_formatter_host_state_matches_reference_loader():
    """EventWindowFormatter draws its per-slot augmentation flags from np.random in the reference's order
    (dataloader/base.py:29-38 at construction, :64-69 on reset_sequence): with the same seed the same slots flip.
    Host logic only - nothing is computed here (the object allocates device memory on its first format_batch)."""
    import importlib
    import numpy as np
    import snnflow_b200 as snnflow
    ref_shim.install()
    base = importlib.import_module("dataloader.base")

    class RefLoader(base.BaseDataLoader):        # the abstract methods are not needed for the bookkeeping
        def __getitem__(self, index):
            raise NotImplementedError

        def get_events(self, history):
            raise NotImplementedError

    cfg = {"data": {"mode": "gtflow_dt1"}, "loader": {"resolution": [128, 128], "std_resolution": [256, 256],
                                                      "batch_size": 5, "augment": ["Horizontal", "Vertical", "Polarity"],
                                                      "augment_prob": [0.5, 0.3, 0.7]},
           "hot_filter": {"enabled": True, "max_px": 100, "min_obvs": 5, "max_rate": 0.8}}
    np.random.seed(123)
    ref = RefLoader(cfg, 5)
    for b in (3, 0, 3):
        ref.reset_sequence(b)
    np.random.seed(123)
    ours = snnflow.EventWindowFormatter(cfg, 5)
    for b in (3, 0, 3):
        ours.reset_sequence(b)
    assert ours.batch_augmentation == ref.batch_augmentation
    assert ours.seq_num == ref.seq_num == 3
    assert list(ours.resolution) == list(ref.resolution) == [256, 256] and ours.pool == (2, 2)
    with pytest.raises(ValueError):
        snnflow.EventWindowFormatter(dict(cfg, loader=dict(cfg["loader"], resolution=[300, 128])), 5).pool
    with pytest.raises(_lib_mod().SnnflowError):
        snnflow.EventWindowFormatter(cfg, 5, device="cpu")


def _lib_mod():
    from snnflow_b200 import _lib
    return _lib


def test_header_is_plain_c_and_links_against_the_library(tmp_path):
    """include/snnflow.h is a C header (no C++, no torch types): a C99 translation unit that includes it compiles with
    -pedantic, links against libsnnflow.so and can call the entry points that need no GPU."""
    import shutil
    import subprocess
    import __graft_entry__ as ge
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    ge.build()
    from snnflow_b200 import _lib
    src = tmp_path / "abi.c"
    src.write_text('#include <stdio.h>\n#include "snnflow.h"\n'
                   "int main(void) {\n"
                   "  snnflow_loader_desc d = {0};\n"
                   "  d.B = 2; d.H = 16; d.W = 16; d.N = 100; d.num_bins = 5; d.pool_h = d.pool_w = 1;\n"
                   "  size_t ws = snnflow_format_window_workspace_bytes(&d);\n"
                   '  printf("%d %d %zu %d\\n", snnflow_abi_version(), SNNFLOW_ABI_VERSION, ws, snnflow_clip_adam_partials(5000));\n'
                   "  return 0;\n}\n")
    exe = tmp_path / "abi"
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", f"-I{os.path.join(ROOT, 'include')}", str(src), "-o", str(exe),
                    f"-L{libdir}", "-lsnnflow", f"-Wl,-rpath,{libdir}"], check=True, capture_output=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    assert out[0] == out[1] == "1"
    assert int(out[2]) > 2 * 16 * 16 * (2 * 4 + 4 + 5 * 8)          # counts + last index + fixed-point voxels, per slot
    assert int(out[3]) == 5                                          # ceil(5000 / 1024) partial sums
