"""-m gpu (needs >= 2 GPUs, skipped otherwise): data-parallel training over NCCL.

The reference has no distributed code; the contract tested here is stock-DDP semantics:
  * ranks fed the SAME batch end up with the parameters of a single-process run (mean of identical gradients),
  * ranks fed DIFFERENT batches stay bit-identical to each other after every step (same averaged gradient everywhere),
  * the bucketed all-reduce launched from inside backward gives the same result as one all-reduce after backward,
  * the whole step (collectives included) replays from a CUDA graph.
"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
HYP = {"box": 7.5, "cls": 1.0, "dfl": 2.5, "reg_max": 16}
WIDTHS = (64, 128, 256, 512)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _build(dev):
    from snn_object_detectionddp_b200.model import TemporalUNet, YOLOTemporalUNet
    from snn_object_detectionddp_b200.weight_initialization import initialize_model
    torch.manual_seed(0)
    net = YOLOTemporalUNet(num_classes=8, hyp=HYP, neuron="lif")
    net.temporal_unet = TemporalUNet([144, 144, 144], neuron="lif", widths=WIDTHS)
    initialize_model(net)
    return net.to(dev)


def _backward_only(tr, frames, batch):
    """Trainer.train_step without the optimizer: returns the (all-reduced) flat gradient and the loss items."""
    tr.model.train()
    tr.store.zero_grad()
    if tr.bucketer is not None:
        tr.bucketer.begin_step()
    det, _ = tr.model.forward_sequence(frames)
    loss, items = tr.loss_fn(det, {"padded": tuple(t.to(frames.device) for t in batch["padded"])})
    loss.sum().backward()
    if tr.bucketer is not None:
        tr.bucketer.finish()
    torch.cuda.synchronize()
    return tr.store.flat_g.clone(), items.clone()


def _worker(rank, world, port, mode, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from snn_object_detectionddp_b200.data import synthetic_batch
    from snn_object_detectionddp_b200.trainer import Trainer
    net = _build(dev)
    tr = Trainer(net, total_steps=20, device=dev, bucket_mb=1)
    assert tr.bucketer is not None and len(tr.bucketer.buckets) > 3
    B, T, HW = 2, 2, 128
    losses = []
    if mode == "grad":
        # ONE backward with the bucketed all-reduce, no optimizer step: the averaged gradient of identical per-rank batches
        frames, labels = synthetic_batch(B, T, HW, HW, seed=100)
        g, items = _backward_only(tr, frames.to(dev), tr.prepare_batch(labels, B, max_boxes=8))
        launched = list(tr.bucketer.launch_order)
        dist.barrier()
        if rank == 0:
            out.put((True, g.cpu().numpy(), [items.cpu().numpy()], launched, False, len(tr.bucketer.buckets)))
        os._exit(0)
    for step in range(4):
        seed = 100 + step + (0 if mode == "same" else 17 * rank)
        frames, labels = synthetic_batch(B, T, HW, HW, seed=seed)
        batch = tr.prepare_batch(labels, B, max_boxes=8)
        fn = tr.train_step_graphed if mode == "graph" else tr.train_step
        _, items = fn(frames.to(dev), {"padded": tuple(t.to(dev) for t in batch["padded"])})
        losses.append(items.clone())
    torch.cuda.synchronize()
    flat = tr.store.flat_p.detach().clone()
    gathered = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    same = all(torch.equal(gathered[0], g) for g in gathered[1:])
    launched = list(tr.bucketer.launch_order)
    graphed = bool(tr._graph is not None)
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    if rank == 0:
        # plain numpy payloads: torch tensors travel through mp queues as shared-memory handles that die with this process
        out.put((same, flat.cpu().numpy(), [l.cpu().numpy() for l in losses], launched, graphed, len(tr.bucketer.buckets)))
    # leave without NCCL teardown: destroy_process_group() can block for minutes while a captured graph still references the
    # communicator's kernels (seen on the 2-GPU box); SimpleQueue.put is synchronous, so the payload is already in the pipe
    os._exit(0)


def _run(mode):
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, mode, q)) for r in range(2)]
    for p in procs:
        p.start()
    import time
    t0 = time.time()
    while q.empty():         # poll instead of a blocking get(): a crashed worker must fail the test, not hang it
        if any(p.exitcode not in (None, 0) for p in procs) or time.time() - t0 > 240:
            for p in procs:
                if p.is_alive():
                    p.kill()
            raise AssertionError(f"DDP worker failed or timed out: exit codes {[p.exitcode for p in procs]}")
        time.sleep(0.2)
    res = q.get()
    for p in procs:
        p.join(60)
        if p.is_alive():
            p.kill()
        assert p.exitcode == 0
    same, flat, losses, launched, graphed, nb = res
    return same, torch.from_numpy(flat), [torch.from_numpy(l) for l in losses], launched, graphed, nb


needs2 = pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")


@needs2
def test_ddp_same_batch_equals_single_process():
    """Ranks fed the SAME batch: the bucketed, backward-overlapped all-reduce (AVG) must reproduce the single-process
    gradient.  The backward is not bit-reproducible (fp32 atomics in split-K wgrad / BN reductions) and this tiny
    geometry amplifies that noise, so the bound is the run-to-run noise of the single-process gradient itself; the
    forward is deterministic, so the loss must match exactly."""
    _, g_ddp, losses, launched, _, nb = _run("grad")
    assert len(launched) == nb and launched[0] == 0      # bucket 0 holds the LAST tensors: launched first, during backward
    from snn_object_detectionddp_b200.data import synthetic_batch
    from snn_object_detectionddp_b200.trainer import Trainer
    from tests.gpu_util import rel_err
    ref = []
    for _ in range(2):
        net = _build("cuda:0")
        tr = Trainer(net, total_steps=20, device="cuda:0")
        frames, labels = synthetic_batch(2, 2, 128, 128, seed=100)
        ref.append(_backward_only(tr, frames.cuda(), tr.prepare_batch(labels, 2, max_boxes=8)))
    (g1, it1), (g2, it2) = ref
    assert torch.equal(it1, it2) and torch.equal(it1.cpu(), losses[0])
    noise = float(rel_err(g1, g2))
    err = float(rel_err(g_ddp.cuda(), g1))
    assert float(g1.abs().max()) > 0 and err < 3 * noise + 1e-5, (err, noise)


@needs2
def test_ddp_same_batch_is_bit_exact_in_deterministic_mode():
    """Deterministic mode (snn_set_deterministic: fixed-order reductions, no split-K): the single-process gradient is
    bit-reproducible, the all-reduce averages two identical fp32 gradients ((g + g) / 2 == g exactly), so the DDP gradient
    must EQUAL the single-process one bit for bit -- any lost, doubled or reordered contribution in the bucketed exchange
    shows.  (Workers inherit SNN_DETERMINISTIC from the environment.)"""
    from snn_object_detectionddp_b200 import _lib
    from snn_object_detectionddp_b200.data import synthetic_batch
    from snn_object_detectionddp_b200.trainer import Trainer
    L = _lib.lib()
    before, env_before = L.snn_get_deterministic(), os.environ.get("SNN_DETERMINISTIC")
    os.environ["SNN_DETERMINISTIC"] = "1"
    L.snn_set_deterministic(1)
    try:
        _, g_ddp, losses, launched, _, nb = _run("grad")
        ref = []
        for _ in range(2):
            net = _build("cuda:0")
            tr = Trainer(net, total_steps=20, device="cuda:0")
            frames, labels = synthetic_batch(2, 2, 128, 128, seed=100)
            ref.append(_backward_only(tr, frames.cuda(), tr.prepare_batch(labels, 2, max_boxes=8)))
    finally:
        L.snn_set_deterministic(before)
        if env_before is None:
            os.environ.pop("SNN_DETERMINISTIC", None)
        else:
            os.environ["SNN_DETERMINISTIC"] = env_before
    (g1, it1), (g2, it2) = ref
    assert torch.equal(it1, it2) and torch.equal(it1.cpu(), losses[0])
    assert float(g1.abs().max()) > 0 and torch.equal(g1, g2)
    assert torch.equal(g_ddp.cuda(), g1), float((g_ddp.cuda() - g1).abs().max())


@needs2
def test_ddp_different_batches_keep_ranks_in_lockstep():
    same, _, losses, _, _, _ = _run("diff")
    assert same and all(torch.isfinite(l).all() for l in losses)


@needs2
def test_ddp_step_with_collectives_replays_from_cuda_graph():
    same, _, losses, _, graphed, _ = _run("graph")
    assert same and graphed and all(torch.isfinite(l).all() for l in losses)
