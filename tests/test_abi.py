"""CPU: the C-ABI library builds, loads, and exports every symbol include/snn_b200.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built_lib():
    import __graft_entry__ as g
    g.build()
    from snn_object_detectionddp_b200 import _lib
    return _lib


def test_header_symbols_are_exported(built_lib):
    hdr = open(os.path.join(ROOT, "include", "snn_b200.h")).read()
    declared = set(re.findall(r"\b(snn_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 15
    so = ctypes.CDLL(built_lib.LIB_PATH)
    missing = [s for s in sorted(declared) if not hasattr(so, s)]
    assert not missing, missing
    assert set(built_lib.exported_symbols()) == declared


def test_version_and_error_string(built_lib):
    L = built_lib.lib()
    assert L.snn_version() >= 100
    assert isinstance(L.snn_last_error(), bytes)


def test_product_has_no_cpu_fallback(built_lib):
    import torch
    from snn_object_detectionddp_b200 import kernels
    with pytest.raises(built_lib.SnnKernelError):
        kernels.nchw_to_nhwc(torch.zeros(1, 8, 2, 2))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "snn_object_detectionddp_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                assert "oracle" not in src.replace("no oracle", ""), f"{f} references the oracle"
