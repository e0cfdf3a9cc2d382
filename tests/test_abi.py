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


def test_host_side_planning_queries(built_lib):
    """Workspace / geometry queries of the ABI are pure host logic (no GPU): fused-BatchNorm-statistics planning of the
    conv epilogue, NMS key workspace, standalone-statistics workspace."""
    L = built_lib.lib()
    g = ctypes.c_int(0)
    # 32x32 maps: one image = 8 tiles of 128 pixels -> 4 groups per tile; T=4, B=2 -> NB=8
    assert L.snn_conv_stats_groups(0, 8, 32, 32, 2, ctypes.byref(g)) == 4 * 8 * 8 and g.value == 4 * 8 * 2
    # stride-2 3x3: the tile domain is the OUTPUT map (16x16 -> 2 tiles per image)
    assert L.snn_conv_stats_groups(1, 8, 32, 32, 2, ctypes.byref(g)) == 4 * 2 * 8 and g.value == 4 * 2 * 2
    # 4x4 maps: a 128-pixel tile spans 8 images -> only available when the frames of a timestep fill whole tiles
    assert L.snn_conv_stats_groups(0, 8, 4, 4, 2, ctypes.byref(g)) == 0
    assert L.snn_conv_stats_groups(0, 16, 4, 4, 8, ctypes.byref(g)) == 4 * 2 and g.value == 4
    # transposed conv has no BatchNorm behind it on the path
    assert L.snn_conv_stats_groups(3, 8, 16, 16, 2, ctypes.byref(g)) == 0
    # NMS keys: next power of two of the candidate enumeration (anchor x class when multi_label)
    assert L.snn_nms_workspace_keys(8, 1344, 1) == 16384 and L.snn_nms_workspace_keys(8, 1344, 0) == 2048
    assert L.snn_nms_workspace_keys(1, 1344, 1) == 2048            # a single class ignores multi_label
    # standalone statistics: T x blocks x 2 x C floats, independent of how the T*B batch was folded
    w1, w4 = L.snn_bn_stats_workspace_floats(1, 65536, 128), L.snn_bn_stats_workspace_floats(4, 65536, 128)
    assert w1 > 0 and w4 == 4 * w1 and L.snn_bn_stats_workspace_floats(1, 64, 6) == 0
