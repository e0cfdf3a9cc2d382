"""-m gpu: fused BN + LIF scan kernels (through the C ABI) vs the build-defined LIF oracle (oracle/snn_oracle.py).

LIF is BUILD-DEFINED (the reference has no spiking neuron) -> "parity unpinned" w.r.t.
the reference; the oracle is the specification.  Spikes / masks must match bit-exactly when both sides get
the same scale/shift; gradients within 1e-5 relative (fp32) or 4e-3 where the kernel emits bf16.
"""
import pytest
import torch

from oracle import snn_oracle as O
from tests.gpu_util import rel_err, setup_exact

pytestmark = pytest.mark.gpu
LIF, SILU = 0, 1


def _k():
    from snn_object_detectionddp_b200 import kernels
    return kernels


def _data(T, B, H, W, C, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    y = torch.randn(T * B, H, W, C, device="cuda", generator=g) * 1.5
    gamma = torch.rand(C, device="cuda", generator=g) + 0.5
    beta = torch.randn(C, device="cuda", generator=g) * 0.3 + 0.3
    return y, gamma, beta


@pytest.mark.parametrize("T,B,H,W,C", [(4, 2, 16, 16, 128), (8, 1, 8, 8, 256), (5, 3, 4, 4, 1024), (16, 1, 8, 8, 64),
                                       (4, 2, 8, 8, 144), (1, 2, 8, 8, 128)])
def test_bn_stats_and_finalize(T, B, H, W, C):
    K = _k()
    y, gamma, beta = _data(T, B, H, W, C)
    P = B * H * W
    sums = K.bn_stats(y, T)
    yt = y.reshape(T, P, C).double()
    assert torch.allclose(sums[:, 0], yt.sum(1), rtol=1e-5, atol=1e-4)  # fp32 partials of <=64 values, fp64 combine
    assert torch.allclose(sums[:, 1], (yt * yt).sum(1), rtol=1e-5, atol=1e-4)
    rm, rv = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda")
    scale, shift, mean, invstd = K.bn_finalize(sums, gamma, beta, rm, rv, T, C, P, 1e-5, 0.1, True)
    # torch BatchNorm2d called once per timestep (reference train.py:64-66 -> model.py:14)
    bn = torch.nn.BatchNorm2d(C).cuda().train()
    with torch.no_grad():
        bn.weight.copy_(gamma); bn.bias.copy_(beta)
    for t in range(T):
        xt = y.reshape(T, B, H, W, C)[t].permute(0, 3, 1, 2)
        ref = bn(xt).permute(0, 2, 3, 1)
        mine = y.reshape(T, B, H, W, C)[t] * scale[t] + shift[t]
        assert (mine - ref).abs().max().item() < 2e-5
    assert torch.allclose(rm, bn.running_mean, rtol=1e-5, atol=1e-6)
    assert torch.allclose(rv, bn.running_var, rtol=1e-5, atol=1e-6)
    # eval mode: running statistics, one scale/shift for all t
    es, eh, _, _ = K.bn_finalize(None, gamma, beta, rm, rv, T, C, P, 1e-5, 0.1, False)
    bn.eval()
    ref = bn(y.permute(0, 3, 1, 2)).permute(0, 2, 3, 1)
    assert ((y * es[0] + eh[0]) - ref).abs().max().item() < 2e-5


@pytest.mark.parametrize("T,B,H,W,C", [(4, 2, 16, 16, 128), (8, 2, 8, 8, 256), (5, 2, 4, 4, 1024), (16, 1, 8, 8, 64),
                                       (4, 1, 8, 8, 144), (1, 2, 8, 8, 128)])
def test_lif_forward_bit_exact(T, B, H, W, C):
    K = _k()
    y, gamma, beta = _data(T, B, H, W, C, seed=1)
    scale = (gamma * 0.9).repeat(T, 1) * (1 + 0.05 * torch.arange(T, device="cuda")[:, None])
    shift = beta.repeat(T, 1)
    v0 = torch.rand(B, H, W, C, device="cuda") * 0.5
    out, mask, vf = K.bn_act_fwd(LIF, y, scale, shift, T, v_init=v0, want_v_final=True)
    yt = y.reshape(T, B, H, W, C)
    x = yt * scale[:, None, None, None, :] + shift[:, None, None, None, :]
    s_ref, v_ref, u_ref = O.lif_sequence(x, v0)
    assert torch.equal(out.reshape(T, B, H, W, C).float(), s_ref), "spikes differ from the oracle"
    assert torch.equal(vf.reshape(B, H, W, C), v_ref), "final membrane differs"
    assert torch.equal(mask.reshape(-1), O.pack_spike_mask(s_ref.reshape(T, -1)).reshape(-1)), "bit-packed mask differs"
    rate = float(s_ref.mean())
    assert 0.02 < rate < 0.9, f"degenerate test data (rate {rate})"
    # zero initial membrane (v_init = NULL)
    out0, _, _ = K.bn_act_fwd(LIF, y, scale, shift, T)
    s0, _, _ = O.lif_sequence(x)
    assert torch.equal(out0.reshape(T, B, H, W, C).float(), s0)


def test_silu_forward():
    K = _k()
    T, B, H, W, C = 4, 2, 8, 8, 128
    y, gamma, beta = _data(T, B, H, W, C, seed=2)
    scale, shift = gamma.repeat(T, 1), beta.repeat(T, 1)
    out, _, _ = K.bn_act_fwd(SILU, y, scale, shift, T)
    x = y.reshape(T, -1, C) * scale[:, None] + shift[:, None]
    ref = torch.nn.functional.silu(x).reshape(out.shape)
    assert rel_err(out, ref) < 4e-3


@pytest.mark.parametrize("T,B,H,W,C", [(4, 2, 8, 8, 128), (8, 2, 4, 4, 256), (5, 2, 4, 4, 1024), (16, 1, 4, 4, 64),
                                       (4, 1, 8, 8, 144)])
def test_lif_backward_vs_oracle_autograd(T, B, H, W, C):
    """Frozen-statistics path: dy = dL/dy through x = y*scale+shift -> LIF(T steps), incl. membrane carry in/out."""
    setup_exact()
    K = _k()
    y, gamma, beta = _data(T, B, H, W, C, seed=3)
    scale, shift = gamma.repeat(T, 1), beta.repeat(T, 1)
    v0 = (torch.rand(B, H, W, C, device="cuda") * 0.5)
    gs = torch.randn(T * B, H, W, C, device="cuda").to(torch.bfloat16)
    gvf = torch.randn(B, H, W, C, device="cuda")
    yr = y.clone().requires_grad_(True)
    v0r = v0.clone().requires_grad_(True)
    x = yr.reshape(T, B, H, W, C) * scale[:, None, None, None, :] + shift[:, None, None, None, :]
    s, vfin, _ = O.lif_sequence(x, v0r)
    loss = (s * gs.float().reshape(s.shape)).sum() + (vfin * gvf).sum()
    gy_ref, gv0_ref = torch.autograd.grad(loss, (yr, v0r))
    _, dy, gv0, _ = K.bn_act_bwd(LIF, False, y, scale, shift, None, None, gs, T, v_init=v0, gv_final=gvf.reshape(-1),
                                 want_gv_init=True)
    assert rel_err(dy, gy_ref) < 4e-3
    assert rel_err(gv0.reshape(gv0_ref.shape), gv0_ref) < 1e-5


def test_lif_backward_train_mode_bn():
    """Batch-statistics path: gx, reductions and the BN input gradient vs torch autograd through batch_norm+LIF.
    Neurons whose oracle membrane is within 1e-5 of threshold are excluded (flip-rate protocol)."""
    setup_exact()
    K = _k()
    T, B, H, W, C = 4, 2, 8, 8, 128
    P = B * H * W
    y, gamma, beta = _data(T, B, H, W, C, seed=4)
    gs = torch.randn(T * B, H, W, C, device="cuda").to(torch.bfloat16)
    sums = K.bn_stats(y, T)
    scale, shift, mean, invstd = K.bn_finalize(sums, gamma, beta, None, None, T, C, P, 1e-5, 0.1, True)
    out, mask, _ = K.bn_act_fwd(LIF, y, scale, shift, T)
    gx, _, _, red = K.bn_act_bwd(LIF, True, y, scale, shift, mean, invstd, gs, T)
    dgamma, dbeta = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    dy = K.bn_bwd_dx(red, gamma, gx, y, scale, mean, invstd, dgamma, dbeta, T)
    # oracle: autograd through per-timestep batch-stat BN + LIF
    yr = y.clone().requires_grad_(True)
    g_r, b_r = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    yt = yr.reshape(T, P, C)
    m = yt.mean(1, keepdim=True)
    var = yt.var(1, unbiased=False, keepdim=True)
    x = (yt - m) / torch.sqrt(var + 1e-5) * g_r + b_r
    s, _, u = O.lif_sequence(x)
    near = ((u.detach() - 1.0).abs() < 1e-5)
    flips = (s.detach() != out.reshape(s.shape).float())
    assert int((flips & ~near).sum()) == 0, "spike flip away from threshold"
    loss = (s * gs.float().reshape(s.shape)).sum()
    gy_ref, gg_ref, gb_ref = torch.autograd.grad(loss, (yr, g_r, b_r))
    if int(flips.sum()) == 0:
        assert rel_err(dy, gy_ref) < 4e-3
        assert rel_err(dgamma, gg_ref) < 1e-4 and rel_err(dbeta, gb_ref) < 1e-4
    # internal consistency of the fused reductions
    xh = (y.reshape(T, P, C) - mean[:, None]) * invstd[:, None]
    gxt = gx.reshape(T, P, C)
    assert torch.allclose(red[:, 0], gxt.sum(1), rtol=1e-4, atol=1e-3)
    assert torch.allclose(red[:, 1], (gxt * xh).sum(1), rtol=1e-4, atol=1e-3)


def test_silu_backward_train_mode_matches_reference_convblock_tail():
    """SiLU variant reproduces the reference's bn->silu backward (model.py:14-18) for one timestep group."""
    setup_exact()
    K = _k()
    T, B, H, W, C = 2, 2, 8, 8, 128
    P = B * H * W
    y, gamma, beta = _data(T, B, H, W, C, seed=5)
    gs = torch.randn(T * B, H, W, C, device="cuda").to(torch.bfloat16)
    sums = K.bn_stats(y, T)
    scale, shift, mean, invstd = K.bn_finalize(sums, gamma, beta, None, None, T, C, P, 1e-5, 0.1, True)
    gx, _, _, red = K.bn_act_bwd(SILU, True, y, scale, shift, mean, invstd, gs, T)
    dgamma, dbeta = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    dy = K.bn_bwd_dx(red, gamma, gx, y, scale, mean, invstd, dgamma, dbeta, T)
    yr = y.clone().requires_grad_(True)
    g_r, b_r = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    outs = []
    for t in range(T):
        xt = yr.reshape(T, B, H, W, C)[t].permute(0, 3, 1, 2)
        z = torch.nn.functional.batch_norm(xt, None, None, g_r, b_r, True, 0.1, 1e-5)
        outs.append(torch.nn.functional.silu(z).permute(0, 2, 3, 1))
    loss = (torch.stack(outs) * gs.float().reshape(T, B, H, W, C)).sum()
    gy_ref, gg_ref, gb_ref = torch.autograd.grad(loss, (yr, g_r, b_r))
    assert rel_err(dy, gy_ref) < 4e-3
    assert rel_err(dgamma, gg_ref) < 1e-4 and rel_err(dbeta, gb_ref) < 1e-4


def test_lstm_gates_fwd_bwd():
    setup_exact()
    K = _k()
    P, Ch = 2 * 4 * 4, 256
    gates = torch.randn(P, 4 * Ch, device="cuda")
    c_prev = torch.randn(P, Ch, device="cuda")
    h, c, hb = K.lstm_gates_fwd(gates, c_prev, Ch)
    gr, cr = gates.clone().requires_grad_(True), c_prev.clone().requires_grad_(True)
    i, f, g, o = torch.split(gr, Ch, dim=1)                    # reference model.py:67-69
    c_ref = torch.sigmoid(f) * cr + torch.sigmoid(i) * torch.tanh(g)
    h_ref = torch.sigmoid(o) * torch.tanh(c_ref)
    assert rel_err(c, c_ref) < 1e-5 and rel_err(h, h_ref) < 1e-5 and rel_err(hb, h_ref) < 4e-3
    dh, dc = torch.randn(P, Ch, device="cuda"), torch.randn(P, Ch, device="cuda")
    gg_ref, gc_ref = torch.autograd.grad((h_ref * dh).sum() + (c_ref * dc).sum(), (gr, cr))
    dg, dcp = K.lstm_gates_bwd(gates, c_prev, c, dh, dc, Ch)
    assert rel_err(dg, gg_ref) < 4e-3 and rel_err(dcp, gc_ref) < 1e-5
    # zero initial state (c_prev = NULL)
    h0, c0, _ = K.lstm_gates_fwd(gates, None, Ch)
    assert rel_err(c0, torch.sigmoid(i) * torch.tanh(g)) < 1e-5


def test_layout_roundtrip():
    K = _k()
    x = torch.randn(3, 144, 8, 12, device="cuda")
    n = K.nchw_to_nhwc(x, dtype=torch.float32)
    assert torch.equal(n, x.permute(0, 2, 3, 1).contiguous())
    assert torch.equal(K.nhwc_to_nchw(n), x)
    nb = K.nchw_to_nhwc(x)
    assert torch.equal(nb, x.permute(0, 2, 3, 1).to(torch.bfloat16))


def test_clip_adamw_matches_torch():
    """clip_grad_norm_(10) + AdamW (reference train.py:77-78, 156-160) on a flat buffer."""
    K = _k()
    n = 4096 * 5
    p0 = torch.randn(n, device="cuda")
    g = torch.randn(n, device="cuda") * 0.5        # norm ~ 71 > 10 -> clipping active
    pt = torch.nn.Parameter(p0.clone())
    opt = torch.optim.AdamW([pt], lr=3e-4, betas=(0.93, 0.999), weight_decay=5e-4)
    p, m, v = p0.clone(), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    shadow = torch.empty(n, device="cuda", dtype=torch.bfloat16)
    acc = torch.zeros(1, device="cuda", dtype=torch.float64)
    gn = torch.zeros(1, device="cuda")
    for step in range(1, 4):
        pt.grad = g.clone() * step
        ref_norm = torch.nn.utils.clip_grad_norm_([pt], 10.0)
        opt.step()
        hp = torch.tensor([3e-4, 0.93, 0.999, 1e-8, 5e-4, 1 - 0.93 ** step, 1 - 0.999 ** step, 10.0], device="cuda")
        gg = g * step
        K.grad_sumsq(gg, acc)
        K.adamw_step(p, gg, m, v, shadow, hp, acc, gn)
        assert abs(float(gn) - float(ref_norm)) < 1e-3 * float(ref_norm)
        assert rel_err(p, pt.data) < 1e-6
    assert torch.equal(shadow, p.to(torch.bfloat16))


@pytest.mark.parametrize("act", [LIF, SILU])
@pytest.mark.parametrize("T,B,H,W,C", [(4, 2, 8, 8, 128), (8, 2, 4, 4, 256), (5, 2, 4, 4, 1024), (16, 1, 4, 4, 64),
                                       (4, 1, 8, 8, 144), (16, 2, 2, 2, 1024), (1, 2, 8, 8, 128), (4, 3, 5, 7, 72)])
def test_recompute_backward_matches_autograd(act, T, B, H, W, C):
    """2-pass recompute backward (snn_bn_act_bwd2) vs torch autograd through per-timestep batch-stat BN + LIF|SiLU."""
    setup_exact()
    K = _k()
    P = B * H * W
    y, gamma, beta = _data(T, B, H, W, C, seed=6)
    y = y + 0.7                                  # non-zero channel means: exercises the (y - mean) accumulation
    gs = torch.randn(T * B, H, W, C, device="cuda").to(torch.bfloat16)
    v0 = torch.rand(B, H, W, C, device="cuda") * 0.5 if act == LIF else None
    gvf = torch.randn(B, H, W, C, device="cuda") if act == LIF else None
    sums = K.bn_stats(y, T)
    scale, shift, mean, invstd = K.bn_finalize(sums, gamma, beta, None, None, T, C, P, 1e-5, 0.1, True)
    out, _, _ = K.bn_act_fwd(act, y, scale, shift, T, v_init=v0)
    dgamma, dbeta = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    dy, gv0, red = K.bn_act_bwd_train(act, y, scale, shift, mean, invstd, beta, gs, T, dgamma, dbeta, v_init=v0,
                                      gv_final=None if gvf is None else gvf.reshape(-1), want_gv_init=act == LIF)
    yr = y.clone().requires_grad_(True)
    g_r, b_r = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    yt = yr.reshape(T, P, C)
    m = yt.mean(1, keepdim=True)
    var = yt.var(1, unbiased=False, keepdim=True)
    x = (yt - m) / torch.sqrt(var + 1e-5) * g_r + b_r
    if act == LIF:
        v0r = v0.clone().reshape(P, C).requires_grad_(True)
        s, vfin, u = O.lif_sequence(x, v0r)
        flips = int((s.detach() != out.reshape(s.shape).float()).sum())
        loss = (s * gs.float().reshape(s.shape)).sum() + (vfin * gvf.reshape(P, C)).sum()
        gy_ref, gg_ref, gb_ref, gv0_ref = torch.autograd.grad(loss, (yr, g_r, b_r, v0r))
    else:
        s = torch.nn.functional.silu(x)
        flips = 0
        loss = (s * gs.float().reshape(s.shape)).sum()
        gy_ref, gg_ref, gb_ref = torch.autograd.grad(loss, (yr, g_r, b_r))
    if flips == 0:
        assert rel_err(dy, gy_ref) < 4e-3
        assert rel_err(dgamma, gg_ref) < 2e-4 and rel_err(dbeta, gb_ref) < 2e-4
        if act == LIF:
            assert rel_err(gv0.reshape(P, C), gv0_ref) < 1e-4


@pytest.mark.parametrize("B,H,W,C", [(3, 9, 5, 144), (2, 16, 16, 64), (1, 7, 3, 1024), (4, 8, 8, 72)])
def test_silu_t1_fast_backward_equals_generic_kernel(B, H, W, C):
    """T == 1 SiLU layers (Detect head, live frame only) use a 4-pixels-per-thread kernel; same arithmetic as the generic
    two-pass kernel: reductions agree to the fp32-atomic order, dy to a handful of bf16 rounding flips."""
    setup_exact()
    K = _k()
    from snn_object_detectionddp_b200 import _lib
    T, P = 1, B * H * W
    y, gamma, beta = _data(T, B, H, W, C, seed=17)
    y = y + 0.3
    gs = torch.randn(T * B, H, W, C, device="cuda").to(torch.bfloat16)
    sums = K.bn_stats(y, T)
    scale, shift, mean, invstd = K.bn_finalize(sums, gamma, beta, None, None, T, C, P, 1e-3, 0.03, True)
    outs = []
    try:
        for generic in (0, 1):
            _lib.lib().snn_debug_set(8, generic)
            dg, db = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
            dy, _, red = K.bn_act_bwd_train(SILU, y, scale, shift, mean, invstd, beta, gs, T, dg, db)
            outs.append((dy.float(), red.clone(), dg, db))
    finally:
        _lib.lib().snn_debug_set(8, 0)
    (dy0, red0, dg0, db0), (dy1, red1, dg1, db1) = outs
    assert rel_err(red0, red1) < 1e-5 and rel_err(dg0, dg1) < 1e-5 and rel_err(db0, db1) < 1e-5
    assert rel_err(dy0, dy1) < 1e-4 and float((dy0 != dy1).float().mean()) < 1e-3


@pytest.mark.parametrize("act", [LIF, SILU])
@pytest.mark.parametrize("B,H,W,C", [(2, 16, 16, 64), (3, 9, 5, 144), (1, 8, 8, 512), (2, 4, 4, 1024)])
def test_t16_chunked_backward_equals_one_chunk_kernel(act, B, H, W, C):
    """T == 16 walks the steps as two chunks of 8 (forward-only scan of steps 0..7 for the membrane entering t = 8, then
    chunk 1 and chunk 0 backward with the surrogate state carried across): per-step arithmetic is the same instruction
    sequence as the one-chunk kernel (snn_debug_set(9, 1)) -> gv_init bit-equal; the BN reductions differ by the fp32-atomic
    order only (1e-5), hence dy (which reads them) by a handful of bf16 rounding flips."""
    setup_exact()
    K = _k()
    from snn_object_detectionddp_b200 import _lib
    T, P = 16, B * H * W
    y, gamma, beta = _data(T, B, H, W, C, seed=23)
    y = y + 0.4
    gs = torch.randn(T * B, H, W, C, device="cuda").to(torch.bfloat16)
    v0 = torch.rand(B, H, W, C, device="cuda") * 0.5 if act == LIF else None
    gvf = torch.randn(B, H, W, C, device="cuda") if act == LIF else None
    sums = K.bn_stats(y, T)
    scale, shift, mean, invstd = K.bn_finalize(sums, gamma, beta, None, None, T, C, P, 1e-5, 0.1, True)
    outs = []
    try:
        for one_chunk in (0, 1):
            _lib.lib().snn_debug_set(9, one_chunk)
            dg, db = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
            dy, gv0, red = K.bn_act_bwd_train(act, y, scale, shift, mean, invstd, beta, gs, T, dg, db, v_init=v0,
                                              gv_final=None if gvf is None else gvf.reshape(-1), want_gv_init=act == LIF)
            outs.append((dy.float(), red.clone(), dg, db, None if gv0 is None else gv0.clone()))
    finally:
        _lib.lib().snn_debug_set(9, 0)
    (dy0, red0, dg0, db0, gv00), (dy1, red1, dg1, db1, gv01) = outs
    assert rel_err(red0, red1) < 1e-5 and rel_err(dg0, dg1) < 1e-5 and rel_err(db0, db1) < 1e-5
    assert rel_err(dy0, dy1) < 1e-4 and float((dy0 != dy1).float().mean()) < 1e-3
    if act == LIF:
        # gv_init = beta * gx[0] depends on gs, gv_final and the recomputed membranes only -- not on the reductions
        assert torch.equal(gv00, gv01)
