"""Stage the reference files of the hot path under the git-ignored ``baseline/_ref/`` (TEST / BASELINE INFRASTRUCTURE).

    python -m oracle.stage_reference

``/root/reference`` exists only in the build container.  The GPU box receives a snapshot of the working tree, and
``baseline/_ref/`` is git-ignored but NOT gpurun-ignored, so an UNMODIFIED copy of the files the hot path imports
 placed there travels with it while staying out of the history.  On the GPU box it lets

  * ``tests/test_gpu_reference_seam.py`` run the reference's own ``LIFFireNet`` / ``LIFFireFlowNet`` with the class
    attributes ``head_neuron / ff_neuron / rec_neuron`` (models/model.py:37-39) pointed at the CUDA cells, and
  * ``bench.py --impl reference`` / ``cpu_baseline`` time the reference itself on the host cores
    (``cpu_baseline.kind = "reference"``).

Nothing in the product package reads ``baseline/_ref``.  Files are copied byte for byte (a manifest with their
SHA-256 is written next to them, ``MANIFEST.json``).
"""
import hashlib
import json
import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SOURCE = os.environ.get("SNNFLOW_REFERENCE_SOURCE", "/root/reference")
STAGED = os.path.join(ROOT, "baseline", "_ref")

# what models/model.py, loss/flow.py, utils/iwe.py and dataloader/encodings.py import (closed under `import`)
FILES = [
    "models/__init__.py", "models/base.py", "models/model.py", "models/model_util.py", "models/spiking_submodules.py",
    "models/spiking_util.py", "models/submodules.py", "models/unet.py", "models/SNNtorch_spiking_submodules.py",
    "loss/__init__.py", "loss/flow.py",
    "utils/__init__.py", "utils/iwe.py",
    "dataloader/__init__.py", "dataloader/encodings.py", "dataloader/base.py", "dataloader/h5.py", "dataloader/utils.py",
]


def _sha(path):
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def source_available():
    return os.path.isfile(os.path.join(SOURCE, "models", "spiking_submodules.py"))


def staged_available():
    return os.path.isfile(os.path.join(STAGED, "MANIFEST.json")) and os.path.isfile(
        os.path.join(STAGED, "models", "spiking_submodules.py"))


def stage(verbose=False):
    """Copy FILES from SOURCE to baseline/_ref (no-op when SOURCE is absent).  Returns the staged root or None."""
    if not source_available():
        return STAGED if staged_available() else None
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(SOURCE, rel), os.path.join(STAGED, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not os.path.isfile(dst) or _sha(dst) != _sha(src):
            shutil.copyfile(src, dst)
        manifest[rel] = _sha(dst)
    with open(os.path.join(STAGED, "MANIFEST.json"), "w") as f:
        json.dump({"source": SOURCE, "sha256": manifest}, f, indent=1, sort_keys=True)
    if verbose:
        print(f"staged {len(FILES)} reference files under {STAGED}")
    return STAGED


def verify():
    """True when every staged file still has the digest recorded at staging time (nothing edited it)."""
    if not staged_available():
        return False
    with open(os.path.join(STAGED, "MANIFEST.json")) as f:
        m = json.load(f)["sha256"]
    return all(os.path.isfile(os.path.join(STAGED, rel)) and _sha(os.path.join(STAGED, rel)) == d for rel, d in m.items())


if __name__ == "__main__":
    print(stage(verbose=True))
