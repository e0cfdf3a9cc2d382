"""TEST INFRASTRUCTURE ONLY -- loader for the *real* reference modules.

Imports ``/root/reference/model.py`` and ``weight_initialization.py`` unmodified, with the
un-installable third-party ``ultralytics`` package stubbed in ``sys.modules`` (the only import
blockers are reference ``model.py:3-4`` and ``weight_initialization.py:6``).  The torch-only
classes on the hot path (``ConvBlock``, ``DownBlock``, ``UpBlock``, ``ConvLSTM2d``, ``TemporalUNet``,
``initialize_weights``) then run exactly as the reference wrote them.

``/root/reference`` exists only in the build container, never on the GPU box, so this module is
used (a) by ``tests/golden/make_golden.py`` to generate the committed fixtures and (b) by CPU
tests that are skipped when the reference tree is absent.  Nothing in the product imports it.
"""
import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("SNN_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "model.py"))


def _install_ultralytics_stub():
    import torch.nn as nn

    if "ultralytics" in sys.modules and not getattr(sys.modules["ultralytics"], "_snn_stub", False):
        return  # a real ultralytics is importable: use it
    u = types.ModuleType("ultralytics")
    u._snn_stub = True
    u.YOLO = object
    nn_mod = types.ModuleType("ultralytics.nn")
    mods = types.ModuleType("ultralytics.nn.modules")
    head = types.ModuleType("ultralytics.nn.modules.head")

    class Detect(nn.Module):  # placeholder type; never instantiated by the oracle path
        pass

    head.Detect = Detect
    utils = types.ModuleType("ultralytics.utils")
    loss = types.ModuleType("ultralytics.utils.loss")
    loss.v8DetectionLoss = object
    sys.modules.update({
        "ultralytics": u, "ultralytics.nn": nn_mod, "ultralytics.nn.modules": mods,
        "ultralytics.nn.modules.head": head, "ultralytics.utils": utils,
        "ultralytics.utils.loss": loss,
    })


def load_reference():
    """Return (model_module, weight_initialization_module) of the unmodified reference."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    _install_ultralytics_stub()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    # the reference's module is literally called `model`; keep it under a private alias too
    ref_model = importlib.import_module("model")
    ref_init = importlib.import_module("weight_initialization")
    return ref_model, ref_init
