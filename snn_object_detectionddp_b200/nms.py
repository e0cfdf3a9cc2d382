"""Detection post-processing with the call surface of ``ultralytics.utils.nms.non_max_suppression`` (the reference calls
it at visualize.py:73-78 with conf 0.3 / iou 0.45 / multi_label=True and at eval_2.py:108 with conf 0.001 / iou 0.6).

PARITY UNPINNED: ultralytics is un-vendored; the published algorithm is restated as a CPU checker in the
test infrastructure and the kernel (csrc/nms.cu, ``snn_nms``) is bit-compared against it (tests/test_gpu_infer.py).
Everything runs on the device in one launch for the whole batch; the only host read is the per-image row count, and
``non_max_suppression_padded`` avoids even that (streaming inference keeps the step free of host syncs).
"""
import torch

from . import _lib
from ._lib import call, ptr, require_cuda, stream_ptr


def non_max_suppression_padded(prediction, conf_thres=0.25, iou_thres=0.45, agnostic=False, multi_label=False, max_det=300,
                               max_nms=30000, max_wh=7680):
    """prediction fp32 [B, 4+nc, A] (eval-mode Detect output).  Returns device tensors
    (rows [B, max_det, 6] = x1 y1 x2 y2 conf cls, kept [B, max_det] int32 enumeration index, counts [B] int32);
    rows beyond counts[b] are undefined."""
    require_cuda(prediction)
    if prediction.dtype != torch.float32:
        prediction = prediction.float()
    prediction = prediction.contiguous()
    b, no, a = prediction.shape
    nc = no - 4
    if not (0.0 <= conf_thres <= 1.0 and 0.0 <= iou_thres <= 1.0):
        raise ValueError("conf_thres and iou_thres must be in [0, 1]")
    if not 1 <= max_det <= 512:
        raise ValueError("max_det must be in [1, 512]")
    multi = bool(multi_label) and nc > 1
    cap = int(_lib.lib().snn_nms_workspace_keys(nc, a, int(multi)))
    dev = prediction.device
    keys = torch.empty((b, cap), device=dev, dtype=torch.int64)
    rows = torch.empty((b, max_det, 6), device=dev, dtype=torch.float32)
    kept = torch.empty((b, max_det), device=dev, dtype=torch.int32)
    counts = torch.empty((b,), device=dev, dtype=torch.int32)
    call("snn_nms", ptr(prediction), b, nc, a, float(conf_thres), float(iou_thres), int(multi), int(bool(agnostic)), int(max_det),
         int(max_nms), float(max_wh), ptr(keys), cap, ptr(rows), ptr(kept), ptr(counts), stream_ptr())
    return rows, kept, counts


def non_max_suppression(prediction, conf_thres=0.25, iou_thres=0.45, classes=None, agnostic=False, multi_label=False,
                        labels=(), max_det=300, nc=0, max_time_img=0.05, max_nms=30000, max_wh=7680, in_place=True,
                        rotated=False, end2end=False, return_idxs=False):
    """Drop-in for ultralytics' function (the arguments the reference uses; `classes`, `labels`, `rotated`, `end2end`
    are not on the path and must keep their defaults).  Accepts the eval-mode model output -- the decoded tensor or the
    ``(decoded, maps)`` tuple -- and returns a list of [n, 6] tensors (x1, y1, x2, y2, conf, cls), one per image."""
    if isinstance(prediction, (list, tuple)):
        prediction = prediction[0]
    if classes is not None or len(labels) or rotated or end2end:
        raise NotImplementedError("classes / labels / rotated / end2end are outside the B200 path")
    rows, kept, counts = non_max_suppression_padded(prediction, conf_thres, iou_thres, agnostic, multi_label, max_det, max_nms,
                                                    max_wh)
    n = counts.tolist()                       # the one host read
    out = [rows[i, :n[i]] for i in range(len(n))]
    if return_idxs:
        return out, [kept[i, :n[i]].long() for i in range(len(n))]
    return out
