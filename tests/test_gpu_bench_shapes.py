"""-m gpu: the persistent multi-tile regime of the tcgen05 conv kernels -- the regime every bench-sized layer runs in --
against torch fp32 convolutions of the same bf16 operands.

(1) every conv layer shape of BASELINE.json configs[1] (T=4, B=64, 256x256 -> NB = 256 folded frames @ 32/16/8/4) and the
    large ones of configs[2] (T=8, B=16, 512x512 -> NB = 128 @ 64/32/16/8): fprop (+ fused BN statistics), dgrad, wgrad.
    These walk 3-14 tiles per persistent CTA pair: shared-memory ring wrap across tiles, TMEM double-buffer hand-off
    (parity first flips at the 3rd tile of a worker), rotating epilogue staging tiles.
(2) the small cases of tests/test_gpu_conv.py re-run with the persistent grid capped to 1 and 3 workers
    (snn_debug_set(5, cap)) so that ONE worker walks every tile of the problem, with the static and the dynamic
    (snn_set_tile_scheduling(1)) tile scheduler.
(3) fused BN statistics on maps the pixel box does not divide (10x10, 12x20 at B=2, 24x40, 15x20): rows of out-of-range
    pixels must not reach the sums.

Tolerances: identical bf16 operands, fp32 accumulation -> 1e-5 (fp32 out) / 4e-3 (bf16 out = one 2^-9 rounding);
split-K wgrad / dgrad add fp32 partial sums in a different order -> 2e-5.
"""
import pytest
import torch

from tests.gpu_util import describe_mismatch, ref_conv, rel_err, setup_exact
from tests.test_gpu_conv import DGRAD_CASES, FPROP_CASES, WGRAD_CASES, _k, _mk, _mkw

pytestmark = pytest.mark.gpu

# fp32 outputs: both sides accumulate up to K = 9 * 4096 products in fp32, in different orders (torch's reference
# carries the same error); measured 1e-6 (K = 1296) ... 3.3e-5 (K = 36864)
TOL32 = 5e-5
G31, G32, G11, GT = 0, 1, 2, 3
TAPS = {G31: 9, G32: 9, G11: 1, GT: 4}

# geom, NB, H, W, C0, C1 (concat source), Cout, bn_frames_per_step (0 = no BatchNorm behind it), note
FPROP_BENCH = [
    (G31, 256, 32, 32, 144, 0, 128, 64, "enc1"),
    (G32, 256, 32, 32, 128, 0, 256, 64, "down1.conv1"),
    (G31, 256, 16, 16, 256, 0, 256, 64, "down1.conv2"),
    (G31, 256, 16, 16, 256, 144, 256, 64, "enc2 (concat 256+144 = 400)"),
    (G32, 256, 16, 16, 256, 0, 512, 64, "down2.conv1"),
    (G31, 256, 8, 8, 512, 0, 512, 64, "down2.conv2"),
    (G31, 256, 8, 8, 512, 144, 512, 64, "enc3 (concat 656)"),
    (G32, 256, 8, 8, 512, 0, 1024, 64, "down3.conv1"),
    (G31, 256, 4, 4, 1024, 0, 1024, 64, "down3.conv2 / bottleneck"),
    (G31, 256, 8, 8, 512, 512, 512, 64, "up1.conv1 (concat 1024)"),
    (G31, 256, 16, 16, 256, 256, 256, 64, "up2.conv1"),
    (G31, 256, 32, 32, 128, 128, 128, 64, "up3.conv1"),
    (G31, 256, 32, 32, 128, 0, 128, 64, "up3.conv2"),
    (GT, 256, 4, 4, 1024, 0, 512, 0, "up1.up"),
    (GT, 256, 8, 8, 512, 0, 256, 0, "up2.up"),
    (GT, 256, 16, 16, 256, 0, 128, 0, "up3.up"),
    (G11, 256, 32, 32, 128, 0, 144, 0, "out_p3"),
    (G11, 256, 8, 8, 512, 0, 144, 0, "out_p5"),
    (G31, 256, 32, 32, 144, 0, 64, 64, "head cv2[0][0]"),
    (G11, 256, 32, 32, 144, 0, 144, 64, "head cv3 pointwise"),
    (G31, 128, 64, 64, 144, 0, 128, 16, "cfg3 enc1"),
    (G32, 128, 64, 64, 128, 0, 256, 16, "cfg3 down1.conv1"),
    (G31, 128, 32, 32, 256, 144, 256, 16, "cfg3 enc2"),
    (G31, 128, 8, 8, 1024, 0, 1024, 16, "cfg3 bottleneck"),
    (GT, 128, 32, 32, 256, 0, 128, 0, "cfg3 up3.up"),
    (G11, 128, 64, 64, 128, 0, 144, 0, "cfg3 out_p3"),
]


@pytest.mark.parametrize("geom,nb,h,w,c0,c1,cout,bpt,note", FPROP_BENCH, ids=[c[-1] for c in FPROP_BENCH])
def test_fprop_bench_shape(geom, nb, h, w, c0, c1, cout, bpt, note):
    setup_exact()
    K = _k()
    x0 = _mk(nb, h, w, c0, 31, spikes=True)
    x1 = _mk(nb, h, w, c1, 32) if c1 else None
    wgt = _mkw(cout, TAPS[geom], c0 + c1, 33)
    with torch.no_grad():
        ref = ref_conv(geom, x0 if x1 is None else torch.cat([x0, x1], 3), wgt)
    if bpt:
        T = nb // bpt
        y, sums = K.conv_fprop_stats(geom, x0, wgt, cout, T, x1=x1)
        assert sums is not None, "fused statistics must be available at the bench shapes"
        yt = y.double().reshape(T, -1, cout)             # sums of the kernel's OWN output (fp64 reference of the same fp32 values)
        want = torch.stack([yt.sum(1), (yt * yt).sum(1)], 1)
        scale = torch.stack([yt.abs().sum(1), (yt * yt).sum(1)], 1)
        assert bool(((sums - want).abs() <= 2e-6 * scale + 1e-9).all()), float(((sums - want).abs() / (scale + 1e-9)).max())
    else:
        bias = torch.randn(cout, device="cuda")
        y = K.conv_fprop(geom, x0, wgt, cout, x1=x1, bias=bias)
        ref = ref + bias
        yb = K.conv_fprop(geom, x0, wgt, cout, x1=x1, bias=bias, out_dtype=torch.bfloat16)
        assert rel_err(yb, ref) < 4e-3, describe_mismatch(yb.float(), ref)
    assert rel_err(y, ref) < TOL32, describe_mismatch(y, ref)


def test_fprop_convlstm_split_at_bench_shape():
    """gates = W_x * x (folded T*B = 256 frames, weight columns [0,1024) of the 2048-wide kernel) + W_h * h (64 frames,
    accumulating epilogue, columns [1024, 2048)) -- reference model.py:66 at the configs[1] bottleneck."""
    setup_exact()
    K = _k()
    x = _mk(256, 4, 4, 1024, 41, spikes=True)
    hprev = _mk(64, 4, 4, 1024, 42)
    wgt = _mkw(4096, 9, 2048, 43)
    bias = torch.randn(4096, device="cuda")
    gates = K.conv_fprop(G31, x, wgt, 4096, bias=bias, w_coff=0)
    with torch.no_grad():
        ref = ref_conv(G31, x, wgt[:, :, :1024].contiguous(), bias)
    assert rel_err(gates, ref) < TOL32, describe_mismatch(gates, ref)
    g1 = gates[64:128]
    K.conv_fprop(G31, hprev, wgt, 4096, out=g1, w_coff=1024, accumulate=True)
    with torch.no_grad():
        ref1 = ref[64:128] + ref_conv(G31, hprev, wgt[:, :, 1024:].contiguous())
    assert rel_err(g1, ref1) < TOL32, describe_mismatch(g1, ref1)


# geom, NB, H, W (conv INPUT dims), Ci (this launch), ci_off, Cin_total, Cout, fp32_out, note
DGRAD_BENCH = [
    (G31, 256, 32, 32, 128, 0, 128, 128, 0, "up3.conv2"),
    (G31, 256, 16, 16, 256, 0, 256, 256, 0, "down1.conv2"),
    (G31, 256, 16, 16, 256, 0, 400, 256, 0, "enc2 part 1 of the concat"),
    (G31, 256, 16, 16, 144, 256, 400, 256, 0, "enc2 part 2 (ci_off 256)"),
    (G31, 256, 8, 8, 512, 0, 512, 512, 0, "down2.conv2"),
    (G31, 256, 4, 4, 1024, 0, 1024, 1024, 0, "bottleneck"),
    (G31, 256, 4, 4, 1024, 0, 2048, 4096, 0, "convlstm x-part"),
    (G31, 64, 4, 4, 1024, 1024, 2048, 4096, 1, "convlstm recurrent (fp32 out, K split over taps)"),
    (G32, 256, 32, 32, 128, 0, 128, 256, 0, "down1.conv1"),
    (G32, 256, 8, 8, 512, 0, 512, 1024, 0, "down3.conv1"),
    (GT, 256, 4, 4, 1024, 0, 1024, 512, 0, "up1.up"),
    (GT, 256, 16, 16, 256, 0, 256, 128, 0, "up3.up"),
    (G11, 256, 32, 32, 128, 0, 128, 144, 0, "out_p3"),
    (G31, 128, 64, 64, 128, 0, 128, 128, 0, "cfg3 up3.conv2"),
    (G31, 128, 16, 16, 512, 0, 512, 512, 0, "cfg3 down2.conv2"),
]


@pytest.mark.parametrize("geom,nb,h,w,ci,ci_off,cin_tot,cout,f32,note", DGRAD_BENCH, ids=[c[-1] for c in DGRAD_BENCH])
def test_dgrad_bench_shape(geom, nb, h, w, ci, ci_off, cin_tot, cout, f32, note):
    setup_exact()
    K = _k()
    wgt = _mkw(cout, TAPS[geom], cin_tot, 51)
    x = torch.zeros(nb, h, w, ci, device="cuda", requires_grad=True)
    y = ref_conv(geom, x, wgt[:, :, ci_off:ci_off + ci].contiguous())
    dy = torch.randn(y.shape, device="cuda", generator=torch.Generator(device="cuda").manual_seed(52)).to(torch.bfloat16)
    (gx_ref,) = torch.autograd.grad(y, x, dy.float())
    del y
    gx = K.conv_dgrad(geom, dy, wgt, (h, w), ci, ci_off=ci_off, out_dtype=torch.float32 if f32 else torch.bfloat16)
    tol = TOL32 if f32 else 4e-3
    assert rel_err(gx, gx_ref) < tol, describe_mismatch(gx.float(), gx_ref)
    if not f32:
        gxf = K.conv_dgrad(geom, dy, wgt, (h, w), ci, ci_off=ci_off, out_dtype=torch.float32)
        assert rel_err(gxf, gx_ref) < TOL32, describe_mismatch(gxf, gx_ref)


# geom, NB, H, W, Ci, w_coff, wK, Cout, note
WGRAD_BENCH = [
    (G31, 256, 32, 32, 144, 0, 144, 128, "enc1"),
    (G31, 256, 32, 32, 128, 0, 128, 128, "up3.conv2"),
    (G31, 256, 16, 16, 256, 0, 256, 256, "down1.conv2"),
    (G31, 256, 16, 16, 144, 256, 400, 256, "enc2 second source (w_coff 256)"),
    (G31, 256, 8, 8, 512, 0, 512, 512, "down2.conv2"),
    (G31, 256, 4, 4, 1024, 0, 1024, 1024, "bottleneck"),
    (G31, 256, 4, 4, 1024, 0, 2048, 4096, "convlstm x-part"),
    (G31, 192, 4, 4, 1024, 1024, 2048, 4096, "convlstm recurrent part (3 of 4 steps)"),
    (G32, 256, 32, 32, 128, 0, 128, 256, "down1.conv1"),
    (GT, 256, 4, 4, 1024, 0, 1024, 512, "up1.up"),
    (G11, 256, 32, 32, 128, 0, 128, 144, "out_p3"),
    (G31, 128, 64, 64, 128, 0, 128, 128, "cfg3 up3.conv2"),
]


@pytest.mark.parametrize("geom,nb,h,w,ci,w_coff,wk,cout,note", WGRAD_BENCH, ids=[c[-1] for c in WGRAD_BENCH])
def test_wgrad_bench_shape(geom, nb, h, w, ci, w_coff, wk, cout, note):
    setup_exact()
    K = _k()
    x = _mk(nb, h, w, ci, 61, spikes=(ci % 128 == 0))
    w0 = torch.zeros(cout, TAPS[geom], ci, device="cuda", requires_grad=True)
    y = ref_conv(geom, x, w0)
    dy = torch.randn(y.shape, device="cuda", generator=torch.Generator(device="cuda").manual_seed(62)).to(torch.bfloat16)
    (gw_ref,) = torch.autograd.grad(y, w0, dy.float())
    del y
    dw = torch.zeros(cout, TAPS[geom], wk, device="cuda")
    K.conv_wgrad(geom, x, dy, dw, w_coff=w_coff)
    got = dw[:, :, w_coff:w_coff + ci]
    assert rel_err(got, gw_ref) < TOL32, describe_mismatch(got, gw_ref)
    if wk > ci:
        rest = torch.cat([dw[:, :, :w_coff], dw[:, :, w_coff + ci:]], 2)
        assert float(rest.abs().max()) == 0, "wgrad wrote outside its channel range"


# ------------------------------------------------------------------------------------------------
# (2) one / three persistent workers walk the whole problem
# ------------------------------------------------------------------------------------------------
@pytest.fixture(params=[(1, 0), (3, 0), (3, 1), (0, 1)], ids=["cap1", "cap3", "cap3-dynamic", "dynamic"])
def grid_cap(request):
    """(cap on persistent workers, tile scheduling): the dynamic scheduler (atomic counter + mbarrier ring, DSMEM hand-off
    to the peer CTA; what data-parallel training runs) must give the same results as the static walk."""
    from snn_object_detectionddp_b200 import _lib
    L = _lib.lib()
    cap, dyn = request.param
    L.snn_debug_set(5, cap)
    L.snn_set_tile_scheduling(dyn)
    yield request.param
    L.snn_debug_set(5, 0)
    L.snn_set_tile_scheduling(0)


@pytest.mark.parametrize("geom,nb,h,w,cin,cout", FPROP_CASES)
def test_fprop_single_worker_walks_all_tiles(grid_cap, geom, nb, h, w, cin, cout):
    setup_exact()
    K = _k()
    x, wgt = _mk(nb, h, w, cin, 1), _mkw(cout, TAPS[geom], cin, 2)
    bias = torch.randn(cout, device="cuda")
    ref = ref_conv(geom, x, wgt, bias)
    out = K.conv_fprop(geom, x, wgt, cout, bias=bias)
    assert rel_err(out, ref) < 1e-5, describe_mismatch(out, ref)
    outb = K.conv_fprop(geom, x, wgt, cout, bias=bias, out_dtype=torch.bfloat16)
    assert rel_err(outb, ref) < 4e-3, describe_mismatch(outb.float(), ref)


@pytest.mark.parametrize("single", [0, 1], ids=["pair", "single-cta"])
@pytest.mark.parametrize("geom,T,B,h,w,cin,cout", [(G31, 4, 8, 16, 16, 128, 256), (G32, 2, 8, 16, 16, 128, 128), (G11, 4, 4, 16, 16, 144, 144),
                                                   (G31, 8, 8, 8, 8, 256, 512)])
def test_fprop_stats_single_worker_walks_all_tiles(grid_cap, single, geom, T, B, h, w, cin, cout):
    """Fused statistics + TMEM double buffer + rotating staging tiles with >= 8 tiles per worker, CTA pair and single CTA."""
    setup_exact()
    K = _k()
    from snn_object_detectionddp_b200 import _lib
    x, wgt = _mk(T * B, h, w, cin, 3, spikes=True), _mkw(cout, TAPS[geom], cin, 4)
    ref = ref_conv(geom, x, wgt)
    try:
        _lib.lib().snn_debug_set(6, single)
        y, sums = K.conv_fprop_stats(geom, x, wgt, cout, T)
    finally:
        _lib.lib().snn_debug_set(6, 0)
    assert sums is not None
    assert rel_err(y, ref) < 1e-5, describe_mismatch(y, ref)
    yt = y.double().reshape(T, -1, cout)
    want = torch.stack([yt.sum(1), (yt * yt).sum(1)], 1)
    scale = torch.stack([yt.abs().sum(1), (yt * yt).sum(1)], 1)
    assert bool(((sums - want).abs() <= 2e-6 * scale + 1e-9).all())


@pytest.mark.parametrize("geom,nb,h,w,cin,cout", DGRAD_CASES)
def test_dgrad_single_worker_walks_all_tiles(grid_cap, geom, nb, h, w, cin, cout):
    setup_exact()
    K = _k()
    wgt = _mkw(cout, TAPS[geom], cin, 11)
    x = _mk(nb, h, w, cin, 12).float().requires_grad_(True)
    y = ref_conv(geom, x, wgt)
    dy = torch.randn(y.shape, device="cuda").to(torch.bfloat16)
    (gx_ref,) = torch.autograd.grad(y, x, dy.float())
    gx = K.conv_dgrad(geom, dy, wgt, (h, w), cin, out_dtype=torch.float32)
    assert rel_err(gx, gx_ref) < 2e-5, describe_mismatch(gx, gx_ref)
    gxb = K.conv_dgrad(geom, dy, wgt, (h, w), cin)
    assert rel_err(gxb, gx_ref) < 4e-3


@pytest.mark.parametrize("geom,nb,h,w,cin,cout", WGRAD_CASES)
def test_wgrad_single_worker_walks_all_items(grid_cap, geom, nb, h, w, cin, cout):
    setup_exact()
    K = _k()
    wgt = _mkw(cout, TAPS[geom], cin, 13).float().requires_grad_(True)
    x = _mk(nb, h, w, cin, 14, spikes=(cin % 128 == 0))
    y = ref_conv(geom, x, wgt)
    dy = torch.randn(y.shape, device="cuda").to(torch.bfloat16)
    (gw_ref,) = torch.autograd.grad(y, wgt, dy.float())
    dw = torch.zeros(cout, TAPS[geom], cin, device="cuda")
    K.conv_wgrad(geom, x, dy, dw)
    assert rel_err(dw, gw_ref) < 2e-5, describe_mismatch(dw, gw_ref)


# ------------------------------------------------------------------------------------------------
# (3) fused statistics when the pixel box does not divide the map
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("geom,T,B,h,w,cin,cout", [(G31, 2, 8, 10, 10, 128, 128), (G31, 2, 2, 12, 20, 64, 64), (G31, 2, 1, 24, 40, 64, 128),
                                                   (G31, 2, 4, 15, 20, 128, 128), (G32, 2, 8, 20, 20, 128, 128), (G11, 2, 2, 12, 20, 144, 64),
                                                   (G31, 1, 8, 10, 10, 256, 512)])
def test_fused_statistics_on_maps_the_pixel_box_does_not_divide(geom, T, B, h, w, cin, cout):
    """A 3x3 tap of an out-of-range output row/column still reads in-range input, so its accumulator is NOT zero: such
    rows must be masked out of the fused sums (they were not in round 1: wrong batch statistics on e.g. 640x640 input)."""
    setup_exact()
    K = _k()
    x = _mk(T * B, h, w, cin, 71)
    wgt = _mkw(cout, TAPS[geom], cin, 72)
    y, sums = K.conv_fprop_stats(geom, x, wgt, cout, T)
    ref = ref_conv(geom, x, wgt)
    assert rel_err(y, ref) < 1e-5
    if sums is None:            # planner declined (falls back to snn_bn_stats): allowed, but then there is nothing to check
        pytest.skip("fused statistics not offered for this geometry")
    yt = y.double().reshape(T, -1, cout)
    want = torch.stack([yt.sum(1), (yt * yt).sum(1)], 1)
    scale = torch.stack([yt.abs().sum(1), (yt * yt).sum(1)], 1)
    assert bool(((sums - want).abs() <= 2e-6 * scale + 1e-9).all()), float(((sums - want).abs() / (scale + 1e-9)).max())
    # and the standalone statistics kernel agrees
    s2 = K.bn_stats(y, T)
    assert bool(((s2 - want).abs() <= 2e-6 * scale + 1e-9).all())


@pytest.mark.parametrize("which", ["fprop", "dgrad", "wgrad"])
def test_dynamic_tile_scheduler_at_bench_shapes(which):
    """Dynamic tile scheduling == static walk, bit for bit (fprop / non-split dgrad: every output element is produced by
    exactly one tile) or to summation order (split-K wgrad), on a configs[1] layer with 14 tiles per worker."""
    setup_exact()
    K = _k()
    from snn_object_detectionddp_b200 import _lib
    L = _lib.lib()
    x = _mk(256, 16, 16, 256, 81, spikes=True)
    wgt = _mkw(256, 9, 256, 82)
    dy = _mk(256, 16, 16, 256, 83)
    outs = []
    try:
        for dyn in (0, 1, 1):
            L.snn_set_tile_scheduling(dyn)
            if which == "fprop":
                y, sums = K.conv_fprop_stats(G31, x, wgt, 256, 4)
                outs.append((y, sums))
            elif which == "dgrad":
                outs.append((K.conv_dgrad(G31, dy, wgt, (16, 16), 256),))
            else:
                dw = torch.zeros(256, 9, 256, device="cuda")
                K.conv_wgrad(G31, x, dy, dw)
                outs.append((dw,))
    finally:
        L.snn_set_tile_scheduling(0)
    torch.cuda.synchronize()
    for o in outs[1:]:
        for a, b in zip(outs[0], o):
            if which == "wgrad":
                assert rel_err(a, b) < 2e-6
            else:
                assert torch.equal(a, b)
