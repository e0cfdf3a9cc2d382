"""Generate tests/golden/nms_golden.pt (run in the build container, CPU):

    python tests/golden/make_golden_nms.py

The reference calls ultralytics' non_max_suppression (visualize.py:73-78, eval_2.py:108); ultralytics is not installable
offline, but the part of it that does the arithmetic -- torchvision.ops.nms, the real third-party kernel -- IS here
(torchvision 0.26).  The fixture records, per case, the prediction tensor [B, 4+nc, A] and the rows produced by the
published candidate-selection logic (restated in oracle/detect_oracle.py) around the REAL torchvision.ops.nms, plus a
direct torchvision.ops.batched_nms cross-check of the kept set.  It pins both the oracle restatement (CPU test) and the
snn_nms kernel (GPU test) without either of them present at comparison time.
"""
import os
import sys

import torch
import torchvision

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import detect_oracle as D  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "nms_golden.pt")


def make_pred(B, nc, A, seed, ties=False):
    g = torch.Generator().manual_seed(seed)
    pred = torch.zeros(B, 4 + nc, A)
    for b in range(B):
        centers = torch.rand(6, 2, generator=g) * 128
        sizes = 8 + torch.rand(6, 2, generator=g) * 40
        which = torch.randint(0, 6, (A,), generator=g)
        jit = torch.randn(A, 4, generator=g) * 2.0
        pred[b, 0:2] = (centers[which] + jit[:, :2]).t()
        pred[b, 2:4] = (sizes[which] + jit[:, 2:].abs()).t()
        sc = torch.rand(A, nc, generator=g) ** 3
        if ties:
            sc = (sc * 8).round() / 8
        pred[b, 4:] = sc.t()
    return pred


def main():
    cases = []
    for name, (B, nc, A, seed, ties), kw in [
        ("visualize", (2, 8, 336, 1, False), dict(conf_thres=0.3, iou_thres=0.45, multi_label=True)),       # visualize.py:73-78
        ("eval", (2, 8, 336, 2, False), dict(conf_thres=0.001, iou_thres=0.6, multi_label=False)),          # eval_2.py:108
        ("ties_agnostic", (1, 4, 336, 3, True), dict(conf_thres=0.25, iou_thres=0.5, multi_label=True, agnostic=True)),
    ]:
        pred = make_pred(B, nc, A, seed, ties)
        rows = D.non_max_suppression(pred, max_det=300, **kw)
        # independent cross-check with torchvision.ops.batched_nms for the class-aware single-label case
        if not kw.get("multi_label") and not kw.get("agnostic"):
            for b in range(B):
                x = pred[b].t()
                conf, j = x[:, 4:].max(1)
                keep = conf > kw["conf_thres"]
                boxes = D.xywh2xyxy(x[keep, :4])
                k = torchvision.ops.batched_nms(boxes, conf[keep], j[keep], kw["iou_thres"])[:300]
                assert torch.equal(boxes[k], rows[b][:, :4]) and torch.equal(conf[keep][k], rows[b][:, 4])
        cases.append(dict(name=name, pred=pred, kwargs=kw, rows=[r.clone() for r in rows]))
        print(name, [tuple(r.shape) for r in rows])
    torch.save(dict(torchvision=torchvision.__version__, torch=torch.__version__, cases=cases), OUT)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
