"""-m gpu: tcgen05 implicit-GEMM convs (through the C ABI) vs torch fp32 convs on the same bf16-rounded operands.

Tolerance: operands are identical bf16 values, products are exact in fp32, so only fp32 summation order
differs -> rel 1e-5 for fp32 outputs; bf16 outputs add one rounding (rel 2^-9 per element -> 4e-3).
"""
import pytest
import torch

from tests.gpu_util import describe_mismatch, ref_conv, rel_err, setup_exact, w_from_torch

pytestmark = pytest.mark.gpu

G31, G32, G11, GT = 0, 1, 2, 3


def _k():
    from snn_object_detectionddp_b200 import kernels
    return kernels


def _mk(nb, h, w, c, seed, spikes=False):
    g = torch.Generator(device="cuda").manual_seed(seed)
    if spikes:
        return (torch.rand(nb, h, w, c, device="cuda", generator=g) < 0.3).to(torch.bfloat16)
    return torch.randn(nb, h, w, c, device="cuda", generator=g).to(torch.bfloat16)


def _mkw(rows, taps, k, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(rows, taps, k, device="cuda", generator=g) * (1.0 / (taps * k) ** 0.5)).to(torch.bfloat16)


FPROP_CASES = [
    # geom, NB, H, W, Cin, Cout
    (G31, 2, 32, 32, 64, 128),
    (G31, 2, 32, 32, 144, 128),     # enc1: K chunk tail (144 = 2*64 + 16)
    (G31, 3, 16, 16, 256, 256),
    (G31, 4, 8, 8, 512, 512),       # 2 images per pixel box
    (G31, 16, 4, 4, 1024, 1024),    # bottleneck: 8 images per pixel box
    (G31, 2, 8, 8, 144, 64),        # head cv2[0]
    (G31, 5, 4, 4, 128, 256),       # NB not a multiple of the box
    (G31, 1, 12, 20, 64, 64),       # non power-of-two map
    (G32, 2, 32, 32, 128, 256),     # down1.conv1
    (G32, 4, 8, 8, 512, 1024),      # down3.conv1
    (G11, 2, 32, 32, 128, 144),     # out_p3 (N = 144)
    (G11, 2, 8, 8, 512, 144),
    (G11, 2, 16, 16, 144, 8),       # head cls 1x1 (N = 8 -> UMMA N 16)
    (GT, 8, 4, 4, 1024, 512),       # up1.up
    (GT, 2, 16, 16, 256, 128),      # up3.up
]


@pytest.mark.parametrize("geom,nb,h,w,cin,cout", FPROP_CASES)
def test_fprop(geom, nb, h, w, cin, cout):
    setup_exact()
    K = _k()
    taps = {G31: 9, G32: 9, G11: 1, GT: 4}[geom]
    x = _mk(nb, h, w, cin, 1)
    wgt = _mkw(cout, taps, cin, 2)
    bias = torch.randn(cout, device="cuda")
    out = K.conv_fprop(geom, x, wgt, cout, bias=bias)
    ref = ref_conv(geom, x, wgt, bias)
    torch.cuda.synchronize()
    assert rel_err(out, ref) < 1e-5, describe_mismatch(out, ref)


@pytest.mark.parametrize("geom,nb,h,w,cin,cout", [(G31, 6, 32, 32, 128, 256), (G31, 5, 16, 16, 400, 512), (G32, 4, 16, 16, 256, 512),
                                                  (G11, 3, 32, 32, 128, 144), (GT, 8, 8, 8, 512, 256)])
def test_fprop_cta_pair_equals_single_cta(geom, nb, h, w, cin, cout):
    """The cta_group::2 (two CTAs per 256-pixel tile) and the single-CTA kernels accumulate every output element over
    the same K order -> bit-identical results; odd m-tile counts exercise the all-out-of-bounds tail of a pair.  Same
    for the two epilogues (swizzled staging + TMA tensor store vs per-thread row stores)."""
    setup_exact()
    K = _k()
    from snn_object_detectionddp_b200 import _lib
    taps = {G31: 9, G32: 9, G11: 1, GT: 4}[geom]
    x, wgt = _mk(nb, h, w, cin, 11), _mkw(cout, taps, cin, 12)
    bias = torch.randn(cout, device="cuda")
    L = _lib.lib()
    outs = {}
    try:
        # tap-by-tap K order on both sides (the row-strip mode, chosen by how many pipeline stages fit, may apply to the pair
        # tile and not to the single-CTA tile; its own pair / single-CTA checks are in test_gpu_strip.py)
        L.snn_debug_set(12, 1)
        for name, single, legacy in (("pair+tma", 0, 0), ("single+tma", 1, 0), ("pair+rows", 0, 1), ("single+rows", 1, 1)):
            L.snn_debug_set(6, single)
            L.snn_debug_set(0, legacy)
            for dt in (torch.float32, torch.bfloat16):
                outs[(name, dt)] = K.conv_fprop(geom, x, wgt, cout, bias=bias, out_dtype=dt)
    finally:
        L.snn_debug_set(6, 0)
        L.snn_debug_set(0, 0)
        L.snn_debug_set(12, 0)
    torch.cuda.synchronize()
    for dt in (torch.float32, torch.bfloat16):
        for name in ("single+tma", "pair+rows", "single+rows"):
            assert torch.equal(outs[("pair+tma", dt)], outs[(name, dt)]), (name, dt)
    assert rel_err(outs[("pair+tma", torch.float32)], ref_conv(geom, x, wgt, bias)) < 1e-5


@pytest.mark.parametrize("geom,T,B,h,w,cin,cout,c1", [(G31, 4, 2, 32, 32, 144, 128, 0), (G31, 3, 2, 16, 16, 256, 256, 144),
                                                      (G32, 2, 4, 16, 16, 128, 256, 0), (G31, 2, 8, 4, 4, 256, 512, 0),
                                                      (G31, 2, 8, 12, 20, 64, 144, 0), (G11, 2, 2, 16, 16, 144, 64, 0),
                                                      (G31, 4, 2, 4, 4, 128, 128, 0)])
def test_fprop_fused_bn_statistics(geom, T, B, h, w, cin, cout, c1):
    """Per-timestep BatchNorm sums produced by the conv epilogue == a separate pass over the conv output (fp64 sums of
    the same fp32 values: 1e-6 of sum |y|), the conv output itself is unchanged, and two launches agree bit for bit.  The last
    case (B = 2 on 4x4 maps: a 128-pixel tile spans 4 timesteps) must report 'not available'."""
    setup_exact()
    K = _k()
    taps = {G31: 9, G32: 9, G11: 1}[geom]
    x0 = _mk(T * B, h, w, cin, 21, spikes=True)
    x1 = _mk(T * B, h, w, c1, 22) if c1 else None
    wgt = _mkw(cout, taps, cin + c1, 23)
    y, sums = K.conv_fprop_stats(geom, x0, wgt, cout, T, x1=x1)
    y_ref = K.conv_fprop(geom, x0, wgt, cout, x1=x1)
    assert torch.equal(y, y_ref)
    if (h, w, B) == (4, 4, 2):
        assert sums is None
        return
    assert sums is not None
    yt = y_ref.double().reshape(T, -1, cout)
    ref = torch.stack([yt.sum(1), (yt * yt).sum(1)], 1)
    # 32-row partials are fp32 sums: error is relative to sum |y| (not to a possibly cancelling total)
    scale = torch.stack([yt.abs().sum(1), (yt * yt).sum(1)], 1)
    assert bool(((sums - ref).abs() <= 1e-6 * scale + 1e-9).all()), float(((sums - ref).abs() / (scale + 1e-9)).max())
    y2, sums2 = K.conv_fprop_stats(geom, x0, wgt, cout, T, x1=x1)
    assert torch.equal(sums, sums2) and torch.equal(y, y2)


def test_fprop_concat_two_sources():
    """enc2 = conv(cat([down1(x1), p4])) (reference model.py:126) without materialising the cat."""
    setup_exact()
    K = _k()
    a, b = _mk(2, 16, 16, 256, 3, spikes=True), _mk(2, 16, 16, 144, 4)
    wgt = _mkw(256, 9, 400, 5)
    out = K.conv_fprop(G31, a, wgt, 256, x1=b)
    ref = ref_conv(G31, torch.cat([a, b], 3), wgt)
    assert rel_err(out, ref) < 1e-5, describe_mismatch(out, ref)


def test_fprop_bf16_out_and_accumulate_and_weight_slices():
    """ConvLSTM split: gates = Wx*x + Wh*h (reference model.py:66), second call accumulates."""
    setup_exact()
    K = _k()
    x, hprev = _mk(8, 4, 4, 128, 6, spikes=True), _mk(8, 4, 4, 128, 7)
    wgt = _mkw(512, 9, 256, 8)
    bias = torch.randn(512, device="cuda")
    gates = K.conv_fprop(G31, x, wgt, 512, bias=bias, w_coff=0)
    K.conv_fprop(G31, hprev, wgt, 512, out=gates, w_coff=128, accumulate=True)
    ref = ref_conv(G31, torch.cat([x, hprev], 3), wgt, bias)
    assert rel_err(gates, ref) < 1e-5, describe_mismatch(gates, ref)
    ob = K.conv_fprop(G31, x, wgt, 512, out_dtype=torch.bfloat16, w_coff=0)
    refb = ref_conv(G31, x, wgt[:, :, :128].contiguous())
    assert rel_err(ob, refb) < 4e-3, describe_mismatch(ob.float(), refb)
    # row slice of the weights (N offset)
    o2 = K.conv_fprop(G31, x, wgt, 128, w_row_off=256)
    assert rel_err(o2, refb[..., 256:384]) < 1e-5


def test_fprop_channel_slice_views():
    setup_exact()
    K = _k()
    big = _mk(2, 8, 8, 256, 9)
    x = big[..., 64:192]                      # ld = 256, C = 128
    wgt = _mkw(128, 9, 128, 10)
    outbuf = torch.zeros(2, 8, 8, 320, device="cuda")
    K.conv_fprop(G31, x, wgt, 128, out=outbuf[..., 64:192])
    ref = ref_conv(G31, x.contiguous(), wgt)
    assert rel_err(outbuf[..., 64:192], ref) < 1e-5
    assert float(outbuf[..., :64].abs().max()) == 0 and float(outbuf[..., 192:].abs().max()) == 0


DGRAD_CASES = [
    (G31, 2, 32, 32, 128, 128), (G31, 16, 4, 4, 1024, 512), (G31, 2, 8, 8, 64, 144),
    (G32, 2, 32, 32, 128, 256), (G32, 4, 8, 8, 512, 1024),
    (G11, 2, 16, 16, 256, 144), (G11, 2, 16, 16, 144, 8),
    (GT, 8, 4, 4, 1024, 512), (GT, 2, 16, 16, 256, 128),
]


@pytest.mark.parametrize("geom,nb,h,w,cin,cout", DGRAD_CASES)
def test_dgrad(geom, nb, h, w, cin, cout):
    setup_exact()
    K = _k()
    taps = {G31: 9, G32: 9, G11: 1, GT: 4}[geom]
    wgt = _mkw(cout, taps, cin, 11)
    x = _mk(nb, h, w, cin, 12).float().requires_grad_(True)
    y = ref_conv(geom, x, wgt)
    dy = torch.randn(y.shape, device="cuda").to(torch.bfloat16)
    (gx_ref,) = torch.autograd.grad(y, x, dy.float())
    gx = K.conv_dgrad(geom, dy, wgt, (h, w), cin, out_dtype=torch.float32)
    assert rel_err(gx, gx_ref) < 1e-5, describe_mismatch(gx, gx_ref)
    gxb = K.conv_dgrad(geom, dy, wgt, (h, w), cin)
    assert rel_err(gxb, gx_ref) < 4e-3
    if cin >= 128:   # channel sub-range (concat inputs get separate dgrads)
        part = K.conv_dgrad(geom, dy, wgt, (h, w), 64, ci_off=64, out_dtype=torch.float32)
        assert rel_err(part, gx_ref[..., 64:128]) < 1e-5


WGRAD_CASES = [
    (G31, 2, 32, 32, 144, 128), (G31, 4, 16, 16, 256, 256), (G31, 16, 4, 4, 256, 512), (G31, 6, 8, 8, 64, 64),
    (G31, 2, 8, 8, 144, 72),
    (G32, 2, 32, 32, 128, 256), (G32, 4, 8, 8, 512, 256),
    (G11, 2, 32, 32, 128, 144), (G11, 2, 16, 16, 144, 8),
    (GT, 8, 4, 4, 512, 256), (GT, 2, 16, 16, 256, 128),
]


@pytest.mark.parametrize("geom,nb,h,w,cin,cout", WGRAD_CASES)
def test_wgrad(geom, nb, h, w, cin, cout):
    setup_exact()
    K = _k()
    taps = {G31: 9, G32: 9, G11: 1, GT: 4}[geom]
    wgt = _mkw(cout, taps, cin, 13).float().requires_grad_(True)
    x = _mk(nb, h, w, cin, 14, spikes=(cin % 128 == 0))
    y = ref_conv(geom, x, wgt)
    dy = torch.randn(y.shape, device="cuda").to(torch.bfloat16)
    (gw_ref,) = torch.autograd.grad(y, wgt, dy.float())
    dw = torch.zeros(cout, taps, cin, device="cuda")
    K.conv_wgrad(geom, x, dy, dw)
    assert rel_err(dw, gw_ref) < 2e-5, describe_mismatch(dw, gw_ref)
    K.conv_wgrad(geom, x, dy, dw)               # accumulates
    assert rel_err(dw, 2 * gw_ref) < 2e-5


def test_wgrad_into_channel_offset():
    setup_exact()
    K = _k()
    x = _mk(2, 16, 16, 144, 15)
    dy = _mk(2, 16, 16, 256, 16)
    dw = torch.zeros(256, 9, 400, device="cuda")
    K.conv_wgrad(G31, x, dy, dw, w_coff=256)
    w0 = torch.zeros(256, 9, 144, device="cuda", requires_grad=True)
    (ref,) = torch.autograd.grad(ref_conv(G31, x, w0), w0, dy.float())
    assert rel_err(dw[:, :, 256:], ref) < 2e-5
    assert float(dw[:, :, :256].abs().max()) == 0


def test_weight_prep():
    K = _k()
    w = torch.randn(96, 9, 144, device="cuda")
    wf, wt = K.weight_prep(w)
    assert torch.equal(wf, w.to(torch.bfloat16))
    assert torch.equal(wt, w.to(torch.bfloat16).permute(2, 1, 0).contiguous())


@pytest.mark.parametrize("geom,nb,h,w,cin,cout", [(G31, 4, 16, 16, 128, 128), (G32, 4, 16, 16, 128, 256), (G11, 6, 16, 16, 256, 144), (GT, 4, 8, 8, 256, 128)])
def test_dgrad_accumulates_into_bf16_buffer(geom, nb, h, w, cin, cout):
    """Second gradient of a fan-out tensor added in the dgrad epilogue (bf16 TMA reduce-add): G += dgrad(dy) must equal
    bf16(G + bf16(dgrad)) -- what autograd's add of two bf16 gradients gives -- on plain, stride-2 (phase-view store),
    1x1 and transposed geometries, and on a tail slice of a larger buffer (live frames only)."""
    setup_exact()
    K = _k()
    taps = {G31: 9, G32: 9, G11: 1, GT: 4}[geom]
    wgt = _mkw(cout, taps, cin, 91)
    ho, wo = K.out_hw(geom, h, w)
    dy = _mk(nb, ho, wo, cout, 92)
    g0 = _mk(nb + 2, h, w, cin, 93)
    ref = (g0[2:].float() + K.conv_dgrad(geom, dy, wgt, (h, w), cin).float()).to(torch.bfloat16)
    buf = g0.clone()
    K.conv_dgrad(geom, dy, wgt, (h, w), cin, out=buf[2:], accumulate=True)
    torch.cuda.synchronize()
    assert torch.equal(buf[:2], g0[:2])
    assert torch.equal(buf[2:], ref), describe_mismatch(buf[2:].float(), ref.float())
