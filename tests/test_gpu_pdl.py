"""-m gpu: programmatic dependent launch (snn_set_dependent_launch): kernels become resident while their predecessor drains
and wait (griddepcontrol.wait) before touching global memory -> results must be identical to serialized launches.

A missing wait in any kernel would show as a read of half-written data (O(1) differences, run-to-run varying), so the
checks are bit-equality of everything that has a fixed summation order (forward: loss items, spikes; eval outputs) and the
usual reordering-noise bound on what is accumulated with atomics (Adam first moment, gradient norm).
"""
import pytest
import torch

from tests.gpu_util import rel_err, setup_exact
from tests.test_gpu_fullwidth import DEV, _product, _sync_state

pytestmark = pytest.mark.gpu


@pytest.fixture
def dependent_launch():
    from snn_object_detectionddp_b200 import _lib
    L = _lib.lib()
    before = L.snn_get_dependent_launch()
    yield L
    L.snn_set_dependent_launch(before)


def test_switch_round_trips(dependent_launch):
    L = dependent_launch
    L.snn_set_dependent_launch(1)
    assert L.snn_get_dependent_launch() == 1
    L.snn_set_dependent_launch(0)
    assert L.snn_get_dependent_launch() == 0


@pytest.mark.parametrize("neuron", ["lif", "silu"])
def test_eval_forward_identical_with_dependent_launch(dependent_launch, neuron):
    """Whole eval forward (extractor -> U-Net over T frames -> Detect decode): ~250 back-to-back dependent kernels."""
    setup_exact()
    L = dependent_launch
    from snn_object_detectionddp_b200.data import synthetic_batch
    net = _product(neuron, seed=5).eval()
    frames, _ = synthetic_batch(8, 4, 256, 256, seed=31)
    frames = frames.to(DEV)
    outs = {}
    for on in (0, 1, 0, 1):
        L.snn_set_dependent_launch(on)
        with torch.no_grad():
            det, _ = net.forward_sequence(frames)
        y = torch.cat([t.float().reshape(-1) for t in list(det.box) + list(det.cls)])
        torch.cuda.synchronize()
        outs.setdefault(on, []).append(y.clone())
    assert torch.equal(outs[0][0], outs[0][1])
    for y in outs[1]:
        assert torch.equal(outs[0][0], y), float((outs[0][0] - y).abs().max())


@pytest.mark.parametrize("graphed", [False, True])
def test_training_steps_identical_with_dependent_launch(dependent_launch, graphed):
    """Trainer a: serialized launches; trainer b: dependent launches (eager, and captured into a CUDA graph -- programmatic
    edges between kernel nodes).  Same state in, same batch: forward loss items bit-equal on every step; Adam first moment /
    gradient norm within the fp32-atomics reordering noise (2e-3, as graphed-vs-eager in test_gpu_fullwidth)."""
    setup_exact()
    L = dependent_launch
    from snn_object_detectionddp_b200.data import synthetic_batch
    from snn_object_detectionddp_b200.trainer import Trainer
    B, T, HW = 16, 4, 256
    L.snn_set_dependent_launch(0)
    a = Trainer(_product("lif", seed=4), total_steps=50, device=DEV)
    b = Trainer(_product("lif", seed=4), total_steps=50, device=DEV)
    worst = 0.0
    for step in range(6):
        frames, labels = synthetic_batch(B, T, HW, HW, seed=300 + step)
        frames = frames.to(DEV)
        batch = {"padded": tuple(t.to(DEV) for t in a.prepare_batch(labels, B, max_boxes=8)["padded"])}
        _sync_state(a, b)
        L.snn_set_dependent_launch(0)
        _, it_a = a.train_step(frames, batch)
        it_a = it_a.clone()
        torch.cuda.synchronize()
        L.snn_set_dependent_launch(1)
        _, it_b = (b.train_step_graphed if graphed else b.train_step)(frames, batch)
        it_b = it_b.clone()
        torch.cuda.synchronize()
        assert torch.equal(it_a, it_b), (step, it_a, it_b)
        e_m = rel_err(b.store.flat_m, a.store.flat_m)
        e_n = abs(float(a.grad_norm) - float(b.grad_norm)) / float(a.grad_norm)
        print(f"step {step}: graphed={b._graph is not None} loss {it_a.tolist()} exp_avg rel {e_m:.2e} norm rel {e_n:.2e}")
        worst = max(worst, e_m, e_n)
        assert float(b.store.flat_g.abs().max()) == 0
    if graphed:
        assert b._graph is not None and not b._graph_failed, "the step was never captured with dependent launches on"
    assert worst < 2e-3, worst
