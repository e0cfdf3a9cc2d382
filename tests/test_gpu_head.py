"""-m gpu: depthwise 3x3 kernels (Detect head class branch) and the frame packer vs torch fp32."""
import pytest
import torch
import torch.nn.functional as F

from tests.gpu_util import rel_err, setup_exact

pytestmark = pytest.mark.gpu


def _k():
    from snn_object_detectionddp_b200 import kernels
    return kernels


@pytest.mark.parametrize("nb,h,w,c", [(2, 8, 8, 144), (3, 16, 12, 64), (1, 4, 4, 1024), (5, 2, 2, 8)])
def test_dw3x3_fprop_dgrad_wgrad(nb, h, w, c):
    setup_exact()
    K = _k()
    x = torch.randn(nb, h, w, c, device="cuda").to(torch.bfloat16)
    w9 = torch.randn(9, c, device="cuda") * 0.3
    xr = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    wr = w9.reshape(3, 3, c).permute(2, 0, 1).unsqueeze(1).contiguous().requires_grad_(True)   # [C,1,3,3]
    yr = F.conv2d(xr, wr, padding=1, groups=c)
    y = K.dw3x3_fprop(x, w9)
    assert rel_err(y, yr.permute(0, 2, 3, 1)) < 1e-6
    dy = torch.randn(nb, h, w, c, device="cuda").to(torch.bfloat16)
    gx_ref, gw_ref = torch.autograd.grad(yr, (xr, wr), dy.float().permute(0, 3, 1, 2))
    gx = K.dw3x3_dgrad(dy, w9)
    assert rel_err(gx, gx_ref.permute(0, 2, 3, 1)) < 4e-3
    dw = torch.zeros(9, c, device="cuda")
    K.dw3x3_wgrad(x, dy, dw)
    assert rel_err(dw, gw_ref.squeeze(1).permute(1, 2, 0).reshape(9, c)) < 1e-5
    K.dw3x3_wgrad(x, dy, dw)
    assert rel_err(dw, 2 * gw_ref.squeeze(1).permute(1, 2, 0).reshape(9, c)) < 1e-5


@pytest.mark.parametrize("T,B,h,w,c", [(4, 8, 32, 32, 144), (2, 3, 16, 20, 144), (1, 5, 4, 4, 64), (3, 2, 15, 7, 1024), (4, 64, 8, 8, 144)])
def test_dw3x3_fprop_with_fused_bn_statistics(T, B, h, w, c):
    """Column-walking depthwise kernel at head shapes (incl. the configs[1] batch): y == torch's grouped conv, the
    per-timestep partial sums reduce to the fp64 sums of y (1e-6 of sum |y|), two launches agree bit for bit."""
    setup_exact()
    K = _k()
    x = torch.randn(T * B, h, w, c, device="cuda").to(torch.bfloat16)
    w9 = torch.randn(9, c, device="cuda") * 0.3
    wr = w9.reshape(3, 3, c).permute(2, 0, 1).unsqueeze(1).contiguous()
    yr = F.conv2d(x.float().permute(0, 3, 1, 2), wr, padding=1, groups=c).permute(0, 2, 3, 1)
    y, part, gpt = K.dw3x3_fprop(x, w9, T=T)
    assert rel_err(y, yr) < 1e-6 and tuple(part.shape) == (T, gpt, 2, c)
    assert torch.equal(y, K.dw3x3_fprop(x, w9))
    yt = y.double().reshape(T, -1, c)
    want = torch.stack([yt.sum(1), (yt * yt).sum(1)], 1)
    scale = torch.stack([yt.abs().sum(1), (yt * yt).sum(1)], 1)
    got = part.double().sum(1)
    assert bool(((got - want).abs() <= 2e-6 * scale + 1e-9).all()), float(((got - want).abs() / (scale + 1e-9)).max())
    y2, part2, _ = K.dw3x3_fprop(x, w9, T=T)
    assert torch.equal(part, part2) and torch.equal(y, y2)


def test_space_to_depth8_layout():
    K = _k()
    B, T, H, W = 2, 3, 64, 128
    fr = torch.rand(B, T, 3, H, W, device="cuda")
    out = K.space_to_depth8(fr, B, T)
    assert out.shape == (T * B, H // 8, W // 8, 192)
    ref = fr.permute(1, 0, 2, 3, 4).reshape(T * B, 3, H // 8, 8, W // 8, 8).permute(0, 2, 4, 1, 3, 5).reshape(T * B, H // 8, W // 8, 192)
    assert torch.equal(out, ref.to(torch.bfloat16))
