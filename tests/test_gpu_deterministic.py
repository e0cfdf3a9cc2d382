"""-m gpu: deterministic mode (snn_set_deterministic): no split-K in wgrad / small-M dgrad, every block-level reduction
(BatchNorm backward sums, bias column sums, depthwise wgrad, gradient norm, loss sums) through per-block partials added in
block order.  Two runs from the same state must give BIT-IDENTICAL parameters, Adam moments and gradient norm -- eagerly and
through the captured graph -- and stay within the usual reordering band (2e-3) of the default (atomics) mode."""
import pytest
import torch

from tests.gpu_util import rel_err, setup_exact
from tests.test_gpu_fullwidth import DEV, _product, _sync_state

pytestmark = pytest.mark.gpu


@pytest.fixture
def det():
    from snn_object_detectionddp_b200 import _lib
    L = _lib.lib()
    before = L.snn_get_deterministic()
    yield L
    L.snn_set_deterministic(before)


def _state(tr):
    return [getattr(tr.store, n).clone() for n in ("flat_p", "flat_m", "flat_v", "shadow")] + [b.clone() for b in tr.model.buffers()]


@pytest.mark.parametrize("neuron", ["lif", "silu"])
def test_training_steps_bit_reproducible(det, neuron):
    setup_exact()
    L = det
    from snn_object_detectionddp_b200.data import synthetic_batch
    from snn_object_detectionddp_b200.trainer import Trainer
    B, T, HW = 8, 4, 256
    L.snn_set_deterministic(1)
    assert L.snn_get_deterministic() == 1
    a = Trainer(_product(neuron, seed=6), total_steps=50, device=DEV)
    b = Trainer(_product(neuron, seed=6), total_steps=50, device=DEV)
    c = Trainer(_product(neuron, seed=6), total_steps=50, device=DEV)       # default (atomics) mode, for the distance
    for step in range(5):
        frames, labels = synthetic_batch(B, T, HW, HW, seed=400 + step)
        frames = frames.to(DEV)
        batch = {"padded": tuple(t.to(DEV) for t in a.prepare_batch(labels, B, max_boxes=8)["padded"])}
        _sync_state(a, b)
        _sync_state(a, c)
        L.snn_set_deterministic(1)
        _, it_a = a.train_step(frames, batch)
        it_a, gn_a = it_a.clone(), float(a.grad_norm)
        _, it_b = b.train_step_graphed(frames, batch)          # eager for the first calls, then the captured graph
        it_b, gn_b = it_b.clone(), float(b.grad_norm)
        torch.cuda.synchronize()
        assert torch.equal(it_a, it_b) and gn_a == gn_b, (step, it_a, it_b, gn_a, gn_b)
        for x, y, name in zip(_state(a), _state(b), ("flat_p", "flat_m", "flat_v", "shadow") + ("buffer",) * 1000):
            assert torch.equal(x, y), (step, name, float((x.float() - y.float()).abs().max()))
        L.snn_set_deterministic(0)
        _, it_c = c.train_step(frames, batch)
        torch.cuda.synchronize()
        assert torch.equal(it_a, it_c.clone())                      # the forward never depended on the mode
        e_m = rel_err(c.store.flat_m, a.store.flat_m)
        print(f"{neuron} step {step}: graphed={b._graph is not None} gn {gn_a:.6f}  default-mode exp_avg rel {e_m:.2e}")
        assert e_m < 2e-3 and abs(float(c.grad_norm) - gn_a) < 1e-4 * gn_a
    assert b._graph is not None and not b._graph_failed


def test_wgrad_and_reductions_repeat_bit_exactly(det):
    """Kernel level: ten launches of wgrad (K split off), the LIF / SiLU BatchNorm-backward reductions (T = 1, 4, 16), the bias
    column sum and the depthwise wgrad give the same bits every time in deterministic mode."""
    setup_exact()
    L = det
    from snn_object_detectionddp_b200 import kernels as K
    L.snn_set_deterministic(1)
    g = torch.Generator(device="cuda").manual_seed(3)
    x = (torch.rand(64, 32, 32, 128, device="cuda", generator=g) < 0.3).to(torch.bfloat16)
    dy = torch.randn(64, 32, 32, 128, device="cuda", generator=g).to(torch.bfloat16)
    ref = None
    for _ in range(10):
        dw = torch.zeros(128, 9, 128, device="cuda")
        K.conv_wgrad(0, x, dy, dw)
        ref = dw if ref is None else ref
        assert torch.equal(dw, ref)
    for act, T in ((0, 4), (0, 16), (1, 1), (1, 4)):
        Bn, H, W, C = 4, 16, 16, 128
        P = Bn * H * W
        y = torch.randn(T * Bn, H, W, C, device="cuda", generator=g) + 0.3
        gs = torch.randn(T * Bn, H, W, C, device="cuda", generator=g).to(torch.bfloat16)
        gamma, beta = torch.rand(C, device="cuda", generator=g) + 0.5, torch.randn(C, device="cuda", generator=g) * 0.1
        sums = K.bn_stats(y, T)
        scale, shift, mean, invstd = K.bn_finalize(sums, gamma, beta, None, None, T, C, P, 1e-5, 0.1, True)
        first = None
        for _ in range(10):
            dg, db = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
            dyo, _, red = K.bn_act_bwd_train(act, y, scale, shift, mean, invstd, beta, gs, T, dg, db)
            cur = (dyo.clone(), red.clone(), dg, db)
            first = cur if first is None else first
            assert all(torch.equal(u, v) for u, v in zip(cur, first)), (act, T)
    first = None
    for _ in range(10):
        acc = torch.zeros(128, device="cuda")
        K.colsum_accumulate(dy, acc)
        first = acc if first is None else first
        assert torch.equal(acc, first)
