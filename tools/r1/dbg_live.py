import sys, torch
sys.path.insert(0, ".")
from tests.test_gpu_train import _models, MO, DEV
from tests.gpu_util import rel_err, setup_exact
from snn_object_detectionddp_b200.params import store_for
from snn_object_detectionddp_b200.loss import v8DetectionLoss
setup_exact()
res = []
for skip in (True, False, False):
    _, net = _models("lif", seed=7)
    net.skip_dead_backward = skip
    net.train()
    B, T, HW = 2, 3, 128
    frames, labels = MO.synthetic_batch(B, T, HW, HW, seed=13)
    frames, labels = frames.to(DEV), labels.to(DEV)
    st = store_for(net, DEV)
    st.zero_grad()
    det, _ = net.forward_sequence(frames)
    loss, items = v8DetectionLoss(net)(det, {"batch_idx": labels[:, 0], "cls": labels[:, 1], "bboxes": labels[:, 2:]})
    loss.sum().backward()
    torch.cuda.synchronize()
    res.append((st, st.flat_g.clone()))
(st, ga), (_, gb), (_, gc) = res
print("skip vs noskip", rel_err(ga, gb), " noskip vs noskip", rel_err(gb, gc))
rows = []
for e in st.entries:
    a, b, c = (g[e.offset:e.offset + e.numel] for g in (ga, gb, gc))
    rows.append((float(rel_err(a, b)), float(rel_err(b, c)), e.name, float(b.norm())))
for r in sorted(rows, reverse=True)[:25]:
    print("%.2e %.2e %-55s %.3e" % r)
