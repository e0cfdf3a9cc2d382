"""CPU: the host-side launch planning of the tensor-core conv kernels (pixel box, N tile, CTA pairs, row-strip mode, pipeline
stages, K split), read back through snn_conv_plan without launching anything (no GPU: the SM count defaults to 148, a B200's).
Pins the decisions DESIGN.md section 3 describes and the measurements in profiles/README.md (r2) were taken with."""
import pytest

G31, G32, G11, GT = 0, 1, 2, 3
SMEM_MAX = 227 * 1024


@pytest.fixture(scope="module")
def K():
    from snn_object_detectionddp_b200 import kernels
    return kernels


@pytest.fixture
def knobs():
    from snn_object_detectionddp_b200 import _lib
    L = _lib.lib()
    yield L
    for k in (6, 12, 13, 14):
        L.snn_debug_set(k, 0)
    L.snn_set_deterministic(0)


# (Cin, Cout, map) of the 3x3 stride-1 ConvBlocks of the default U-Net at 256x256 input (configs[1]: NB = T*B = 256)
UNET_S1 = [(144, 128, 32), (128, 128, 32), (256, 128, 32), (256, 256, 16), (400, 256, 16), (512, 256, 16), (512, 512, 8), (656, 512, 8),
           (1024, 512, 8), (1024, 1024, 4), (1024, 4096, 4)]


@pytest.mark.parametrize("cin,cout,hw", UNET_S1)
@pytest.mark.parametrize("kind", ["fprop", "dgrad", "wgrad"])
def test_every_unet_layer_has_a_valid_plan(K, kind, cin, cout, hw):
    p = K.conv_plan(kind, G31, 256, hw, hw, cin, cout, out_f32=(kind != "dgrad"), frames_per_step=64 if kind == "fprop" else 0)
    assert p["stages"] >= 2 and p["smem_bytes"] <= SMEM_MAX and p["stages"] * p["stage_bytes"] < p["smem_bytes"]
    assert p["bn"] * p["bh"] * p["bw"] == (64 if kind == "wgrad" else 128)
    assert 1 <= p["ctas"] <= (74 if p["pair"] else 148) and p["ctas"] <= p["items"]
    assert p["hnw"] == p["strip"]                       # the (h, n, w) row order exists for the strip views only
    if kind != "wgrad":
        assert p["tma_out"] == 1 and p["small_k"] == 0 and p["epi_groups"] == 1


@pytest.mark.parametrize("kind", ["fprop", "dgrad"])
def test_row_strip_mode_is_chosen_where_it_was_measured_to_win(K, kind):
    f32 = kind == "fprop"
    # (the 144-channel case is the head's first conv: forward only matters -- its dgrad would need a 144-column MN-major
    #  weight tile that cannot be split between two CTAs, 96 KB per stage, and stays tap by tap)
    for cin, cout, hw in [(128, 128, 32), (256, 256, 16), (512, 512, 8)] + ([(144, 64, 32)] if f32 else [(64, 144, 32)]):
        p = K.conv_plan(kind, G31, 256, hw, hw, cin, cout, out_f32=f32)
        assert p["strip"] == 1, (cin, cout, hw, p)
        # one activation box of bh + 2 rows for three taps, three weight slices
        assert p["a_bytes"] == (p["bh"] + 2) * p["bn"] * p["bw"] * 128 and p["stage_bytes"] == p["a_bytes"] + 3 * p["b_bytes"]
    # 8x8 maps: two images per box; 4x4 maps (a third of the box would be padding rows) stay tap by tap
    assert K.conv_plan(kind, G31, 256, 8, 8, 512, 512, out_f32=f32)["bn"] == 2
    p4 = K.conv_plan(kind, G31, 256, 4, 4, 1024, 1024, out_f32=f32)
    assert p4["strip"] == 0 and p4["bn"] == 8 and p4["stage_bytes"] == 16384 + p4["b_bytes"] and p4["stages"] >= 5
    # stride 2, 1x1 and transposed convs have no stencil column to share
    for geom, cin, cout in [(G32, 128, 256), (G11, 128, 144), (GT, 256, 128)]:
        assert K.conv_plan(kind, geom, 256, 32, 32, cin, cout, out_f32=f32)["strip"] == 0


def test_strip_stage_budget(K):
    """128-column pair tiles: 24 KB strip + 3 x 8 KB weight slices = 48 KB -> 3 stages; 256-column tiles 24 + 3 x 16 = 72 KB -> 2."""
    a = K.conv_plan("fprop", G31, 256, 32, 32, 128, 128)
    assert (a["pair"], a["n_tile"], a["stage_bytes"], a["stages"]) == (1, 128, 49152, 3)
    b = K.conv_plan("fprop", G31, 256, 16, 16, 256, 256)
    assert (b["pair"], b["n_tile"], b["stages"]) == (1, 256, 2) and b["stage_bytes"] == (8 + 2) * 16 * 128 + 3 * 16384


def test_strip_knobs(K, knobs):
    knobs.snn_debug_set(12, 1)
    assert K.conv_plan("fprop", G31, 256, 32, 32, 128, 128)["strip"] == 0
    knobs.snn_debug_set(12, 2)          # only where three stages fit
    assert K.conv_plan("fprop", G31, 256, 32, 32, 128, 128)["strip"] == 1
    assert K.conv_plan("fprop", G31, 256, 16, 16, 256, 256)["strip"] == 0
    knobs.snn_debug_set(12, 3)          # only boxes inside one image
    assert K.conv_plan("fprop", G31, 256, 8, 8, 512, 512)["strip"] == 0
    knobs.snn_debug_set(12, 0)
    knobs.snn_debug_set(13, 1)
    assert K.conv_plan("wgrad", G31, 256, 32, 32, 128, 128)["strip"] == 0


def test_wgrad_plan(K):
    """Strip wgrad: cin tiles of 128 (three 128-column accumulators in TMEM), one round of work items; cin = 144 (not a multiple of
    128) and 4x4 maps stay tap by tap; Cout <= 128 runs single-CTA."""
    p = K.conv_plan("wgrad", G31, 256, 32, 32, 128, 128)
    assert (p["strip"], p["pair"], p["n_tile"]) == (1, 0, 128) and p["items"] <= 148 and p["items"] == p["ctas"]
    p = K.conv_plan("wgrad", G31, 256, 16, 16, 256, 256)
    assert (p["strip"], p["pair"], p["n_tile"], p["n_blocks"]) == (1, 1, 128, 2) and p["items"] <= 74 and p["stages"] >= 4
    assert p["items"] == 3 * 2 * p["ksplit"]            # 3 stencil columns x 2 cin tiles x K split
    assert K.conv_plan("wgrad", G31, 256, 32, 32, 144, 128)["strip"] == 0
    assert K.conv_plan("wgrad", G31, 256, 4, 4, 1024, 1024)["strip"] == 0
    assert K.conv_plan("wgrad", G31, 256, 8, 8, 512, 512)["strip"] == 1


def test_small_m_dgrad_splits_k_and_deterministic_mode_does_not(K, knobs):
    """ConvLSTM recurrent dgrad (64 frames of 4x4, 1024 <- 4096 channels, fp32 out): 16 tiles for 74 CTA pairs -> K split over
    the 9 taps; deterministic mode takes every K split out (one contribution per output element and launch)."""
    p = K.conv_plan("dgrad", G31, 64, 4, 4, 1024, 4096, out_f32=True)
    assert p["ksplit"] == 9 and p["items"] == 144
    assert K.conv_plan("dgrad", G31, 256, 16, 16, 256, 256, out_f32=False)["ksplit"] == 1
    assert K.conv_plan("wgrad", G31, 256, 32, 32, 128, 128)["ksplit"] > 1
    knobs.snn_set_deterministic(1)
    assert K.conv_plan("dgrad", G31, 64, 4, 4, 1024, 4096, out_f32=True)["ksplit"] == 1
    assert K.conv_plan("wgrad", G31, 256, 32, 32, 128, 128)["ksplit"] == 1
    assert K.conv_plan("wgrad", G31, 256, 32, 32, 144, 128)["ksplit"] == 1


def test_small_k_convs_run_single_cta_with_two_epilogue_groups(K):
    for geom, cin, cout in [(G11, 128, 144), (GT, 256, 128), (G11, 144, 8)]:
        p = K.conv_plan("fprop", geom, 256, 32, 32, cin, cout)
        assert (p["small_k"], p["pair"], p["epi_groups"]) == (1, 0, 2), (geom, cin, cout, p)


def test_fused_statistics_box_is_chosen_per_timestep(K):
    """The pixel box of a conv with fused BatchNorm statistics depends on B (frames per timestep), not on NB = T*B: the fused
    T-step launch and T per-frame launches then sum the same 32-pixel groups."""
    for hw, c in [(8, 512), (4, 1024), (16, 256)]:
        a = K.conv_plan("fprop", G31, 4 * 8, hw, hw, c, c, frames_per_step=8)
        b = K.conv_plan("fprop", G31, 8, hw, hw, c, c, frames_per_step=8)
        assert (a["bn"], a["bh"], a["bw"], a["strip"]) == (b["bn"], b["bh"], b["bw"], b["strip"]), (hw, a, b)


# configs[2]: T = 8, B = 16, 512x512 frames -> NB = 128, maps 64 / 32 / 16 / 8
UNET_S1_CFG3 = [(144, 128, 64), (128, 128, 64), (256, 128, 64), (256, 256, 32), (400, 256, 32), (512, 512, 16), (656, 512, 16),
                (1024, 1024, 8), (1024, 4096, 8)]


@pytest.mark.parametrize("cin,cout,hw", UNET_S1_CFG3)
@pytest.mark.parametrize("kind", ["fprop", "dgrad", "wgrad"])
def test_configs2_layers_have_valid_plans_and_use_the_strip_mode(K, kind, cin, cout, hw):
    p = K.conv_plan(kind, G31, 128, hw, hw, cin, cout, out_f32=(kind != "dgrad"), frames_per_step=16 if kind == "fprop" else 0)
    assert p["stages"] >= 2 and p["smem_bytes"] <= SMEM_MAX
    if kind == "wgrad":
        assert p["strip"] == (1 if cin % 128 == 0 else 0), p
    elif not (kind == "dgrad" and cin % 128 != 0 and cin > 256):
        # every map of this config is >= 8x8: row-strip mode everywhere (dgrad towards a ragged wide input -- 400 / 656 channels,
        # single-CTA MN-major weight tiles -- is the exception: its stage would not fit twice)
        assert p["strip"] == 1 or (kind == "dgrad" and cin == 144), p
