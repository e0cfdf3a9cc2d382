"""-m gpu: row-strip mode of the 3x3 stride-1 convs (fprop and dgrad): one activation box of bh + 2 rows feeds the three taps
of a stencil column through row-shifted shared-memory views.  Same products, different fp32 summation order (column-major
over the stencil instead of row-major): strip vs. tap-by-tap within 1e-5 relative, both within 1e-5 (fp32 out) / 4e-3 (bf16
out) of torch fp32 convs on the same bf16 operands."""
import pytest
import torch

from tests.gpu_util import describe_mismatch, ref_conv, rel_err, setup_exact

pytestmark = pytest.mark.gpu
G31 = 0


def _k():
    from snn_object_detectionddp_b200 import kernels
    return kernels


def _mk(nb, h, w, c, seed, spikes=False):
    g = torch.Generator(device="cuda").manual_seed(seed)
    if spikes:
        return (torch.rand(nb, h, w, c, device="cuda", generator=g) < 0.3).to(torch.bfloat16)
    return torch.randn(nb, h, w, c, device="cuda", generator=g).to(torch.bfloat16)


def _mkw(rows, k, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(rows, 9, k, device="cuda", generator=g) * (1.0 / (9 * k) ** 0.5)).to(torch.bfloat16)


@pytest.fixture
def knobs():
    from snn_object_detectionddp_b200 import _lib
    L = _lib.lib()
    yield L
    for k in (4, 5, 6, 12, 13):
        L.snn_debug_set(k, 0)


CASES = [
    # nb, h, w, c0, c1, cout
    (6, 32, 32, 128, 0, 128),      # enc / decoder 128-channel layers
    (5, 32, 32, 144, 0, 64),       # head cv2[0]: ragged K (144 = 2*64 + 16), N = 64
    (3, 16, 16, 256, 144, 256),    # concat of two sources, 256-column tile (two 72 KB stages; tap-by-tap with knob 12 = 2)
    (4, 16, 16, 64, 0, 144),       # N = 144
    (2, 64, 64, 128, 0, 128),      # bw = 64, bh = 2
    (3, 12, 20, 128, 0, 128),      # non power-of-two map: partly out-of-bounds boxes
    (1, 16, 8, 64, 0, 64),         # exactly one tile per image
    (7, 16, 16, 128, 0, 8),        # N = 8
    (8, 8, 8, 512, 0, 512),        # 8x8 maps: a box spans 2 images, rows ordered (h, n, w)
    (5, 8, 8, 256, 144, 256),      # ... odd image count (the second image of the last box is out of range), two sources
    (16, 4, 4, 1024, 0, 512),      # 4x4 maps: 8 images per box
    (3, 4, 4, 128, 0, 128),        # ... mostly out-of-range box
    (40, 2, 2, 64, 0, 64),         # 2x2 maps: 32 images per box
]


@pytest.mark.parametrize("nb,h,w,c0,c1,cout", CASES)
@pytest.mark.parametrize("single,cap", [(0, 0), (1, 0), (0, 2)])
def test_fprop_strip_mode(knobs, nb, h, w, c0, c1, cout, single, cap):
    setup_exact()
    K = _k()
    x0, x1 = _mk(nb, h, w, c0, 1, spikes=True), (_mk(nb, h, w, c1, 2) if c1 else None)
    wgt = _mkw(cout, c0 + c1, 3)
    bias = torch.randn(cout, device="cuda")
    ref = ref_conv(G31, x0 if x1 is None else torch.cat([x0, x1], 3), wgt, bias)
    knobs.snn_debug_set(6, single)
    knobs.snn_debug_set(5, cap)
    outs = {}
    for mode in (1, 0, 2):
        knobs.snn_debug_set(12, mode)
        outs[mode] = K.conv_fprop(G31, x0, wgt, cout, x1=x1, bias=bias)
        ob = K.conv_fprop(G31, x0, wgt, cout, x1=x1, bias=bias, out_dtype=torch.bfloat16)
        assert rel_err(outs[mode], ref) < 1e-5, (mode, describe_mismatch(outs[mode], ref))
        assert rel_err(ob, ref) < 4e-3, mode
    assert rel_err(outs[0], outs[1]) < 1e-5 and rel_err(outs[2], outs[1]) < 1e-5


@pytest.mark.parametrize("T,B,h,w,cin,cout", [(4, 2, 32, 32, 144, 128), (2, 3, 16, 16, 128, 64), (3, 2, 12, 20, 64, 144),
                                              (2, 4, 8, 8, 256, 512), (3, 8, 4, 4, 512, 256)])
def test_fprop_strip_mode_fused_statistics(knobs, T, B, h, w, cin, cout):
    """The BN partial sums come out of the epilogue, which the strip mode does not touch: sums == a pass over the same y."""
    setup_exact()
    K = _k()
    x = _mk(T * B, h, w, cin, 5, spikes=True)
    wgt = _mkw(cout, cin, 6)
    for mode in (1, 0):
        knobs.snn_debug_set(12, mode)
        y, sums = K.conv_fprop_stats(G31, x, wgt, cout, T)
        assert sums is not None
        yt = y.double().reshape(T, -1, cout)
        ref = torch.stack([yt.sum(1), (yt * yt).sum(1)], 1)
        scale = torch.stack([yt.abs().sum(1), (yt * yt).sum(1)], 1)
        assert bool(((sums - ref).abs() <= 1e-6 * scale + 1e-9).all()), mode
        assert rel_err(y, ref_conv(G31, x, wgt)) < 1e-5, mode


@pytest.mark.parametrize("nb,h,w,cin,cout", [(6, 32, 32, 128, 128), (3, 16, 16, 256, 256), (4, 16, 16, 144, 64), (3, 12, 20, 128, 128),
                                             (2, 64, 64, 64, 128), (5, 16, 16, 128, 144), (8, 8, 8, 512, 512), (5, 8, 8, 256, 128),
                                             (16, 4, 4, 1024, 1024), (3, 4, 4, 128, 256)])
@pytest.mark.parametrize("single,cap", [(0, 0), (1, 0), (0, 2)])
def test_dgrad_strip_mode(knobs, nb, h, w, cin, cout, single, cap):
    setup_exact()
    K = _k()
    dy = _mk(nb, h, w, cout, 7)
    wgt = _mkw(cout, cin, 8)
    x = torch.zeros(nb, cin, h, w, device="cuda", requires_grad=True)
    wt = wgt.float().reshape(cout, 3, 3, cin).permute(0, 3, 1, 2).contiguous()
    torch.nn.functional.conv2d(x, wt, padding=1).backward(dy.float().permute(0, 3, 1, 2))
    ref = x.grad.permute(0, 2, 3, 1).contiguous()
    knobs.snn_debug_set(6, single)
    knobs.snn_debug_set(5, cap)
    outs = {}
    for mode in (1, 0, 2):
        knobs.snn_debug_set(12, mode)
        outs[mode] = K.conv_dgrad(G31, dy, wgt, (h, w), cin, out_dtype=torch.float32)
        assert rel_err(outs[mode], ref) < 1e-5, (mode, describe_mismatch(outs[mode], ref))
        ob = K.conv_dgrad(G31, dy, wgt, (h, w), cin)
        assert rel_err(ob, ref) < 4e-3, mode
    assert rel_err(outs[0], outs[1]) < 1e-5 and rel_err(outs[2], outs[1]) < 1e-5


WGRAD_CASES = [
    # nb, h, w, cin, cout
    (8, 32, 32, 128, 128),     # single-CTA (Cout <= 128), two X boxes per stage
    (6, 16, 16, 256, 256),     # CTA pair, two cin tiles
    (8, 8, 8, 512, 512),       # 8x8 maps: box = one whole image, bw = 8
    (4, 32, 32, 256, 128),     # decoder conv on the concatenated input's first source
    (3, 16, 16, 64, 144),      # 64-channel input: NT = 64; Cout = 144 (pair with a ragged second CTA)
    (2, 64, 64, 128, 256),     # bw = 64, bh = 1
    (5, 12, 20, 128, 128),     # non power-of-two map
    (4, 16, 16, 144, 128),     # cin not a multiple of 128: stays tap-by-tap (same result either way)
    (16, 4, 4, 1024, 1024),    # 4x4 maps: the 64-pixel box spans 4 images, rows ordered (h, n, w) in both operands
    (7, 4, 4, 256, 512),       # ... image count not a multiple of the box
]


@pytest.mark.parametrize("nb,h,w,cin,cout", WGRAD_CASES)
@pytest.mark.parametrize("single,cap,ks", [(0, 0, 0), (1, 0, 0), (0, 3, 0), (0, 0, 5)])
def test_wgrad_strip_mode(knobs, nb, h, w, cin, cout, single, cap, ks):
    """dW of one stencil column (three taps) per work item from ONE staged dY tile and the row-shifted views of one X box.
    Same products as the tap-by-tap kernel; fp32 sums differ by the split-K order only (2e-5 vs torch fp32, 1e-5 between the
    two modes); accumulating twice doubles the result."""
    setup_exact()
    K = _k()
    wgt = _mkw(cout, cin, 13).float().requires_grad_(True)
    x = _mk(nb, h, w, cin, 14, spikes=(cin % 128 == 0))
    y = ref_conv(G31, x, wgt)
    dy = torch.randn(y.shape, device="cuda").to(torch.bfloat16)
    (gw_ref,) = torch.autograd.grad(y, wgt, dy.float())
    knobs.snn_debug_set(6, single)
    knobs.snn_debug_set(5, cap)
    knobs.snn_debug_set(4, ks)
    outs = {}
    try:
        for off in (1, 0):
            knobs.snn_debug_set(13, off)
            dw = torch.zeros(cout, 9, cin, device="cuda")
            K.conv_wgrad(G31, x, dy, dw)
            assert rel_err(dw, gw_ref) < 2e-5, (off, describe_mismatch(dw, gw_ref))
            outs[off] = dw.clone()
            K.conv_wgrad(G31, x, dy, dw)
            assert rel_err(dw, 2 * gw_ref) < 2e-5, off
    finally:
        knobs.snn_debug_set(13, 0)
        knobs.snn_debug_set(4, 0)
    assert rel_err(outs[0], outs[1]) < 1e-5


def test_wgrad_strip_mode_into_channel_offset(knobs):
    """Second source of a concatenated input: gradient lands at its channel offset of the weight, nothing else is touched."""
    setup_exact()
    K = _k()
    x = _mk(3, 16, 16, 128, 15)
    dy = _mk(3, 16, 16, 256, 16)
    dw = torch.zeros(256, 9, 384, device="cuda")
    K.conv_wgrad(G31, x, dy, dw, w_coff=256)
    w0 = torch.zeros(256, 9, 128, device="cuda", requires_grad=True)
    (ref,) = torch.autograd.grad(ref_conv(G31, x, w0), w0, dy.float())
    assert rel_err(dw[:, :, 256:], ref) < 2e-5
    assert float(dw[:, :, :256].abs().max()) == 0
