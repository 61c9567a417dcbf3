"""Importable alias: the package directory is named ``snn_event-based_optical_flow_b200`` (not a valid
Python identifier), so ``import snnflow_b200 as snnflow`` is the way to spell it in code."""
import importlib
import sys

_pkg = importlib.import_module("snn_event-based_optical_flow_b200")
sys.modules[__name__] = _pkg
