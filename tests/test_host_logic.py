"""CPU: host-side logic of the product that needs no GPU -- OneCycle/AdamW hyper-parameter table vs torch's own
scheduler, label padding, the dense (synchronisation-free) task-aligned assigner vs the restated ultralytics
assigner, the synthetic-batch generator, and bucketed gradient all-reduce over gloo with world_size 2."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import detect_oracle as D
from oracle import model_oracle as MO


def test_one_cycle_table_matches_torch_scheduler():
    """Reference train.py:156-169: AdamW(default lr) + OneCycleLR(max_lr=1e-4, pct_start .3, cos), beta1 cycled."""
    from snn_object_detectionddp_b200.trainer import one_cycle_table
    for total in (10, 37, 200):
        p = torch.nn.Parameter(torch.zeros(1))
        opt = torch.optim.AdamW([p], weight_decay=5e-4)
        sch = torch.optim.lr_scheduler.OneCycleLR(opt, max_lr=1e-4, total_steps=total, pct_start=0.3, anneal_strategy="cos")
        tab = one_cycle_table(total, 1e-4, 5e-4)
        for k in range(total):
            g = opt.param_groups[0]
            assert abs(g["lr"] - float(tab[k, 0])) <= 1e-12 + 1e-9 * g["lr"], (total, k)
            assert abs(g["betas"][0] - float(tab[k, 1])) <= 1e-12, (total, k)
            assert float(tab[k, 2]) == g["betas"][1] and float(tab[k, 3]) == g["eps"] and float(tab[k, 4]) == g["weight_decay"]
            assert abs(float(tab[k, 5]) - (1 - g["betas"][0] ** (k + 1))) < 1e-12
            p.grad = torch.ones(1)
            opt.step()
            if k + 1 < total:
                sch.step()


def test_pad_targets_layout_and_empty():
    from snn_object_detectionddp_b200.loss import pad_targets
    lab = torch.tensor([[1, 3, .5, .5, .2, .2], [0, 1, .3, .3, .1, .1], [1, 2, .6, .6, .1, .3]])
    cls, box, valid = pad_targets(lab, 3)
    assert cls.shape == (3, 2) and valid.tolist() == [[True, False], [True, True], [False, False]]
    assert cls[1].tolist() == [3, 2] and torch.allclose(box[1, 1], torch.tensor([.6, .6, .1, .3]))
    cls, box, valid = pad_targets(torch.zeros(0, 6), 2)
    assert cls.shape == (2, 1) and not valid.any()


def test_collate_and_synthetic_batch_match_oracle_generator():
    from snn_object_detectionddp_b200.data import custom_collate_fn, synthetic_batch
    f1, l1 = synthetic_batch(3, 2, 64, 64, seed=7)
    f2, l2 = MO.synthetic_batch(3, 2, 64, 64, seed=7)
    assert torch.equal(f1, f2) and torch.equal(l1, l2)
    assert l1.shape[1] == 6 and float(f1.min()) >= 0 and float(f1.max()) < 1
    imgs, labs = custom_collate_fn([(torch.zeros(2, 3, 8, 8), torch.tensor([[1., .5, .5, .1, .1]])),
                                    (torch.zeros(2, 3, 8, 8), torch.zeros(0, 5))])
    assert imgs.shape == (2, 2, 3, 8, 8) and torch.allclose(labs, torch.tensor([[0, 1, .5, .5, .1, .1]]))
    _, labs = custom_collate_fn([(torch.zeros(1, 3, 8, 8), torch.zeros(0, 5))])
    assert labs.shape == (0, 6)


@pytest.mark.parametrize("seed,B,nmax", [(0, 4, 6), (1, 2, 1), (2, 3, 12)])
def test_dense_assigner_matches_restated_ultralytics(seed, B, nmax):
    from snn_object_detectionddp_b200.loss import task_aligned_assign
    g = torch.Generator().manual_seed(seed)
    nc, strides, hw = 8, (8.0, 16.0, 32.0), 128
    maps = [torch.zeros(B, 1, hw // int(s), hw // int(s)) for s in strides]
    anchors, st = D.make_anchors(maps, strides)
    A = anchors.shape[0]
    pd_scores = torch.rand(B, A, nc, generator=g) * 0.5
    ctr = (anchors * st)[None].expand(B, -1, -1)
    half = torch.rand(B, A, 2, generator=g) * 30 + 4
    jit = (torch.rand(B, A, 2, generator=g) - 0.5) * 6
    pd_bboxes = torch.cat((ctr + jit - half, ctr + jit + half), -1)
    n = torch.randint(0, nmax + 1, (B,), generator=g)
    n[0] = nmax
    gt_labels = torch.randint(0, nc, (B, nmax, 1), generator=g).float()
    c = torch.rand(B, nmax, 2, generator=g) * 90 + 19
    wh = torch.rand(B, nmax, 2, generator=g) * 40 + 8
    gt_bboxes = torch.cat((c - wh / 2, c + wh / 2), -1)
    valid = torch.arange(nmax)[None] < n[:, None]
    gt_bboxes = gt_bboxes * valid[..., None]
    gt_labels = gt_labels * valid[..., None]
    mask_gt = gt_bboxes.sum(2, keepdim=True).gt_(0.0)
    ref = D.TaskAlignedAssigner(topk=10, num_classes=nc)(pd_scores, pd_bboxes, anchors * st, gt_labels, gt_bboxes, mask_gt)
    _, r_boxes, r_scores, r_fg, _ = ref
    t_boxes, t_scores, fg = task_aligned_assign(pd_scores, pd_bboxes, anchors * st, gt_labels.squeeze(-1).long(), gt_bboxes,
                                                mask_gt.squeeze(-1).bool(), nc)
    assert torch.equal(fg.bool(), r_fg)
    assert int(r_fg.sum()) > 0
    assert torch.equal(t_scores, r_scores)
    assert torch.equal(t_boxes[r_fg], r_boxes[r_fg])


# ------------------------------------------------------------------------------------------------
# bucketed all-reduce over gloo, world_size 2
# ------------------------------------------------------------------------------------------------
def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _ddp_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from snn_object_detectionddp_b200.ddp import GradBucketer, broadcast_module_state
    spans = [(0, 24), (24, 8), (32, 104), (136, 16), (152, 8)]
    flat = torch.arange(160, dtype=torch.float32) * (rank + 1)
    bk = GradBucketer(flat, spans, bucket_bytes=64, group=None)
    assert bk.world == 2 and len(bk.buckets) >= 3
    covered = sorted((lo, hi) for lo, hi, _ in bk.buckets)
    assert covered[0][0] == 0 and covered[-1][1] == 160 and all(a[1] == b[0] for a, b in zip(covered, covered[1:]))
    for step in range(2):
        flat.copy_(torch.arange(160, dtype=torch.float32) * (rank + 1) * (step + 1))
        bk.begin_step()
        for i in (4, 3, 2):          # backward order: last tensors first; tensors 0 and 1 never report (no grad)
            bk.entry_ready(i)
        bk.entry_ready(3)            # duplicate notifications are ignored
        bk.finish()
        expect = torch.arange(160, dtype=torch.float32) * 1.5 * (step + 1)       # mean over ranks of (rank+1)
        assert torch.allclose(flat, expect), (rank, step)
        assert bk.launch_order[0] == bk.bucket_of[4]
    lin = torch.nn.Linear(4, 4)
    torch.manual_seed(rank)
    with torch.no_grad():
        lin.weight.normal_()
    broadcast_module_state(lin)
    ws = [torch.zeros_like(lin.weight) for _ in range(world)]
    dist.all_gather(ws, lin.weight.data)
    assert torch.equal(ws[0], ws[1])
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        q.put("ok")


def test_grad_bucketer_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_ddp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get() == "ok"
