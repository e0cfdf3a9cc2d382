"""TEST INFRASTRUCTURE ONLY -- plain-PyTorch restatement of the ultralytics pieces the reference calls.

PARITY UNPINNED.  The reference imports these from the third-party package ``ultralytics`` (un-vendored, version
unpinned -- no requirements file anywhere in the tree; ``yolo11m.pt`` + ``ultralytics.utils.nms`` bracket it to
8.3.x).  The package is not installable offline and the reference holds no test or golden vector at
this boundary, so what follows restates the PUBLISHED ultralytics 8.3.x algorithms, anchored on the reference's call
sites:

* ``Detect(nc, ch)``                      built at reference model.py:186-192, called model.py:209
* ``v8DetectionLoss(model)(preds, batch)`` built at train.py:155, called train.py:74,126 (hyp from config.yaml:33-37)
* ``non_max_suppression(pred, conf, iou, multi_label=True)`` called visualize.py:73-78 and eval_2.py:108-112

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this file.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F


# ----------------------------------------------------------------------------------------------
# Detect head (ultralytics/nn/modules/head.py, non-legacy YOLO11 variant) and helpers
# ----------------------------------------------------------------------------------------------
class Conv(nn.Module):
    """ultralytics Conv: Conv2d(bias=False) + BatchNorm2d(eps=1e-3, momentum=0.03) + SiLU."""

    def __init__(self, c1, c2, k=1, g=1):
        super().__init__()
        self.conv = nn.Conv2d(c1, c2, k, 1, k // 2, groups=g, bias=False)
        self.bn = nn.BatchNorm2d(c2, eps=1e-3, momentum=0.03)

    def forward(self, x):
        return F.silu(self.bn(self.conv(x)))


class _DFL(nn.Module):
    def __init__(self, c1=16):
        super().__init__()
        self.conv = nn.Conv2d(c1, 1, 1, bias=False).requires_grad_(False)
        self.conv.weight.data[:] = torch.arange(c1, dtype=torch.float).view(1, c1, 1, 1)


class OracleDetect(nn.Module):
    def __init__(self, nc=80, ch=(), reg_max=16):
        super().__init__()
        self.nc, self.nl, self.reg_max = nc, len(ch), reg_max
        self.no = nc + reg_max * 4
        self.stride = torch.zeros(self.nl)
        c2, c3 = max((16, ch[0] // 4, reg_max * 4)), max(ch[0], min(nc, 100))
        self.cv2 = nn.ModuleList(nn.Sequential(Conv(x, c2, 3), Conv(c2, c2, 3), nn.Conv2d(c2, 4 * reg_max, 1)) for x in ch)
        self.cv3 = nn.ModuleList(
            nn.Sequential(nn.Sequential(Conv(x, x, 3, g=x), Conv(x, c3, 1)), nn.Sequential(Conv(c3, c3, 3, g=c3), Conv(c3, c3, 1)),
                          nn.Conv2d(c3, nc, 1)) for x in ch)
        self.dfl = _DFL(reg_max)      # frozen arange projection; kept for state_dict key compatibility

    def forward(self, x):
        maps = [torch.cat((self.cv2[i](x[i]), self.cv3[i](x[i])), 1) for i in range(self.nl)]
        if self.training:
            return maps
        return decode(maps, self.stride, self.nc, self.reg_max), maps


def make_anchors(feats, strides, offset=0.5):
    pts, st = [], []
    for f, s in zip(feats, strides):
        h, w = f.shape[2:]
        sx = torch.arange(w, device=f.device, dtype=f.dtype) + offset
        sy = torch.arange(h, device=f.device, dtype=f.dtype) + offset
        sy, sx = torch.meshgrid(sy, sx, indexing="ij")
        pts.append(torch.stack((sx, sy), -1).view(-1, 2))
        st.append(torch.full((h * w, 1), float(s), device=f.device, dtype=f.dtype))
    return torch.cat(pts), torch.cat(st)


def dist2bbox(distance, anchor_points, xywh=True, dim=-1):
    lt, rb = distance.chunk(2, dim)
    x1y1, x2y2 = anchor_points - lt, anchor_points + rb
    if xywh:
        return torch.cat(((x1y1 + x2y2) / 2, x2y2 - x1y1), dim)
    return torch.cat((x1y1, x2y2), dim)


def bbox2dist(anchor_points, bbox, reg_max):
    x1y1, x2y2 = bbox.chunk(2, -1)
    return torch.cat((anchor_points - x1y1, x2y2 - anchor_points), -1).clamp_(0, reg_max - 0.01)


def decode(maps, strides, nc, reg_max):
    """Detect._inference: [B, 4+nc, A] with xywh boxes in pixels and sigmoid scores."""
    b = maps[0].shape[0]
    x_cat = torch.cat([m.view(b, nc + 4 * reg_max, -1) for m in maps], 2)
    box, cls = x_cat.split((reg_max * 4, nc), 1)
    anchors, st = make_anchors(maps, strides, 0.5)
    proj = torch.arange(reg_max, dtype=box.dtype, device=box.device)
    a = box.shape[-1]
    dist = (box.view(b, 4, reg_max, a).transpose(2, 1).softmax(1) * proj.view(1, -1, 1, 1)).sum(1)     # [B,4,A]
    dbox = dist2bbox(dist, anchors.transpose(0, 1).unsqueeze(0), xywh=True, dim=1) * st.transpose(0, 1)
    return torch.cat((dbox, cls.sigmoid()), 1)


def xywh2xyxy(x):
    y = torch.empty_like(x)
    xy, wh = x[..., :2], x[..., 2:] / 2
    y[..., :2], y[..., 2:] = xy - wh, xy + wh
    return y


def bbox_iou_ciou(box1, box2, eps=1e-7):
    """ultralytics utils/metrics.py bbox_iou(xywh=False, CIoU=True)."""
    b1_x1, b1_y1, b1_x2, b1_y2 = box1.chunk(4, -1)
    b2_x1, b2_y1, b2_x2, b2_y2 = box2.chunk(4, -1)
    w1, h1 = b1_x2 - b1_x1, b1_y2 - b1_y1 + eps
    w2, h2 = b2_x2 - b2_x1, b2_y2 - b2_y1 + eps
    inter = (b1_x2.minimum(b2_x2) - b1_x1.maximum(b2_x1)).clamp_(0) * (b1_y2.minimum(b2_y2) - b1_y1.maximum(b2_y1)).clamp_(0)
    union = w1 * h1 + w2 * h2 - inter + eps
    iou = inter / union
    cw = b1_x2.maximum(b2_x2) - b1_x1.minimum(b2_x1)
    ch = b1_y2.maximum(b2_y2) - b1_y1.minimum(b2_y1)
    c2 = cw.pow(2) + ch.pow(2) + eps
    rho2 = ((b2_x1 + b2_x2 - b1_x1 - b1_x2).pow(2) + (b2_y1 + b2_y2 - b1_y1 - b1_y2).pow(2)) / 4
    v = (4 / math.pi ** 2) * ((w2 / h2).atan() - (w1 / h1).atan()).pow(2)
    with torch.no_grad():
        alpha = v / (v - iou + (1 + eps))
    return iou - (rho2 / c2 + v * alpha)


# ----------------------------------------------------------------------------------------------
# TaskAlignedAssigner (ultralytics/utils/tal.py), topk=10, alpha=0.5, beta=6.0
# ----------------------------------------------------------------------------------------------
class TaskAlignedAssigner:
    def __init__(self, topk=10, num_classes=80, alpha=0.5, beta=6.0, eps=1e-9):
        self.topk, self.num_classes, self.alpha, self.beta, self.eps = topk, num_classes, alpha, beta, eps

    @torch.no_grad()
    def __call__(self, pd_scores, pd_bboxes, anc_points, gt_labels, gt_bboxes, mask_gt):
        self.bs, self.n_max_boxes = pd_scores.shape[0], gt_bboxes.shape[1]
        if self.n_max_boxes == 0:
            return (torch.full_like(pd_scores[..., 0], self.num_classes), torch.zeros_like(pd_bboxes),
                    torch.zeros_like(pd_scores), torch.zeros_like(pd_scores[..., 0]).bool(), torch.zeros_like(pd_scores[..., 0]))
        mask_pos, align_metric, overlaps = self.get_pos_mask(pd_scores, pd_bboxes, gt_labels, gt_bboxes, anc_points, mask_gt)
        target_gt_idx, fg_mask, mask_pos = self.select_highest_overlaps(mask_pos, overlaps, self.n_max_boxes)
        target_labels, target_bboxes, target_scores = self.get_targets(gt_labels, gt_bboxes, target_gt_idx, fg_mask)
        align_metric *= mask_pos
        pos_align_metrics = align_metric.amax(dim=-1, keepdim=True)
        pos_overlaps = (overlaps * mask_pos).amax(dim=-1, keepdim=True)
        norm_align_metric = (align_metric * pos_overlaps / (pos_align_metrics + self.eps)).amax(-2).unsqueeze(-1)
        target_scores = target_scores * norm_align_metric
        return target_labels, target_bboxes, target_scores, fg_mask.bool(), target_gt_idx

    def get_pos_mask(self, pd_scores, pd_bboxes, gt_labels, gt_bboxes, anc_points, mask_gt):
        mask_in_gts = self.select_candidates_in_gts(anc_points, gt_bboxes)
        align_metric, overlaps = self.get_box_metrics(pd_scores, pd_bboxes, gt_labels, gt_bboxes, mask_in_gts * mask_gt)
        mask_topk = self.select_topk_candidates(align_metric, topk_mask=mask_gt.expand(-1, -1, self.topk).bool())
        return mask_topk * mask_in_gts * mask_gt, align_metric, overlaps

    def get_box_metrics(self, pd_scores, pd_bboxes, gt_labels, gt_bboxes, mask_gt):
        na = pd_bboxes.shape[-2]
        mask_gt = mask_gt.bool()
        overlaps = torch.zeros([self.bs, self.n_max_boxes, na], dtype=pd_bboxes.dtype, device=pd_bboxes.device)
        bbox_scores = torch.zeros([self.bs, self.n_max_boxes, na], dtype=pd_scores.dtype, device=pd_scores.device)
        ind = torch.zeros([2, self.bs, self.n_max_boxes], dtype=torch.long, device=pd_scores.device)
        ind[0] = torch.arange(end=self.bs, device=pd_scores.device).view(-1, 1).expand(-1, self.n_max_boxes)
        ind[1] = gt_labels.squeeze(-1)
        bbox_scores[mask_gt] = pd_scores[ind[0], :, ind[1]][mask_gt]
        pd_boxes = pd_bboxes.unsqueeze(1).expand(-1, self.n_max_boxes, -1, -1)[mask_gt]
        gt_boxes = gt_bboxes.unsqueeze(2).expand(-1, -1, na, -1)[mask_gt]
        overlaps[mask_gt] = bbox_iou_ciou(gt_boxes, pd_boxes).squeeze(-1).clamp_(0)
        align_metric = bbox_scores.pow(self.alpha) * overlaps.pow(self.beta)
        return align_metric, overlaps

    def select_topk_candidates(self, metrics, topk_mask=None):
        topk_metrics, topk_idxs = torch.topk(metrics, self.topk, dim=-1, largest=True)
        if topk_mask is None:
            topk_mask = (topk_metrics.max(-1, keepdim=True)[0] > self.eps).expand_as(topk_idxs)
        topk_idxs.masked_fill_(~topk_mask, 0)
        count_tensor = torch.zeros(metrics.shape, dtype=torch.int8, device=topk_idxs.device)
        ones = torch.ones_like(topk_idxs[:, :, :1], dtype=torch.int8, device=topk_idxs.device)
        for k in range(self.topk):
            count_tensor.scatter_add_(-1, topk_idxs[:, :, k:k + 1], ones)
        count_tensor.masked_fill_(count_tensor > 1, 0)
        return count_tensor.to(metrics.dtype)

    def get_targets(self, gt_labels, gt_bboxes, target_gt_idx, fg_mask):
        batch_ind = torch.arange(end=self.bs, dtype=torch.int64, device=gt_labels.device)[..., None]
        target_gt_idx = target_gt_idx + batch_ind * self.n_max_boxes
        target_labels = gt_labels.long().flatten()[target_gt_idx]
        target_bboxes = gt_bboxes.view(-1, gt_bboxes.shape[-1])[target_gt_idx]
        target_labels.clamp_(0)
        target_scores = torch.zeros((target_labels.shape[0], target_labels.shape[1], self.num_classes), dtype=torch.int64,
                                    device=target_labels.device)
        target_scores.scatter_(2, target_labels.unsqueeze(-1), 1)
        fg_scores_mask = fg_mask[:, :, None].repeat(1, 1, self.num_classes)
        target_scores = torch.where(fg_scores_mask > 0, target_scores, 0)
        return target_labels, target_bboxes, target_scores

    @staticmethod
    def select_candidates_in_gts(xy_centers, gt_bboxes, eps=1e-9):
        n_anchors = xy_centers.shape[0]
        bs, n_boxes, _ = gt_bboxes.shape
        lt, rb = gt_bboxes.view(-1, 1, 4).chunk(2, 2)
        bbox_deltas = torch.cat((xy_centers[None] - lt, rb - xy_centers[None]), dim=2).view(bs, n_boxes, n_anchors, -1)
        return bbox_deltas.amin(3).gt_(eps)

    @staticmethod
    def select_highest_overlaps(mask_pos, overlaps, n_max_boxes):
        fg_mask = mask_pos.sum(-2)
        if fg_mask.max() > 1:
            mask_multi_gts = (fg_mask.unsqueeze(1) > 1).expand(-1, n_max_boxes, -1)
            max_overlaps_idx = overlaps.argmax(1)
            is_max_overlaps = torch.zeros(mask_pos.shape, dtype=mask_pos.dtype, device=mask_pos.device)
            is_max_overlaps.scatter_(1, max_overlaps_idx.unsqueeze(1), 1)
            mask_pos = torch.where(mask_multi_gts, is_max_overlaps, mask_pos).float()
            fg_mask = mask_pos.sum(-2)
        target_gt_idx = mask_pos.argmax(-2)
        return target_gt_idx, fg_mask, mask_pos


# ----------------------------------------------------------------------------------------------
# v8DetectionLoss (ultralytics/utils/loss.py)
# ----------------------------------------------------------------------------------------------
class OracleV8DetectionLoss:
    """loss_fn(preds, batch) -> (loss[3] * batch_size, loss[3].detach());  loss = (box, cls, dfl) * hyp gains."""

    def __init__(self, model, tal_topk=10):
        m = model.model[-1]
        self.hyp, self.stride, self.nc, self.reg_max = model.args, m.stride, m.nc, m.reg_max
        self.no = m.nc + m.reg_max * 4
        self.assigner = TaskAlignedAssigner(topk=tal_topk, num_classes=self.nc, alpha=0.5, beta=6.0)

    def preprocess(self, targets, batch_size, scale_tensor):
        nl, ne = targets.shape
        if nl == 0:
            return torch.zeros(batch_size, 0, ne - 1, device=targets.device)
        i = targets[:, 0]
        _, counts = i.unique(return_counts=True)
        out = torch.zeros(batch_size, int(counts.max()), ne - 1, device=targets.device)
        for j in range(batch_size):
            matches = i == j
            n = int(matches.sum())
            if n:
                out[j, :n] = targets[matches, 1:]
        out[..., 1:5] = xywh2xyxy(out[..., 1:5].mul_(scale_tensor))
        return out

    def bbox_decode(self, anchor_points, pred_dist):
        b, a, c = pred_dist.shape
        proj = torch.arange(self.reg_max, dtype=pred_dist.dtype, device=pred_dist.device)
        pred_dist = pred_dist.view(b, a, 4, c // 4).softmax(3).matmul(proj)
        return dist2bbox(pred_dist, anchor_points, xywh=False)

    def __call__(self, preds, batch):
        feats = preds[1] if isinstance(preds, tuple) else preds
        dev = feats[0].device
        loss = torch.zeros(3, device=dev)
        pred_distri, pred_scores = torch.cat([xi.view(feats[0].shape[0], self.no, -1) for xi in feats], 2).split(
            (self.reg_max * 4, self.nc), 1)
        pred_scores = pred_scores.permute(0, 2, 1).contiguous()
        pred_distri = pred_distri.permute(0, 2, 1).contiguous()
        dtype, batch_size = pred_scores.dtype, pred_scores.shape[0]
        imgsz = torch.tensor(feats[0].shape[2:], device=dev, dtype=dtype) * self.stride[0]
        anchor_points, stride_tensor = make_anchors(feats, self.stride, 0.5)
        targets = torch.cat((batch["batch_idx"].view(-1, 1), batch["cls"].view(-1, 1), batch["bboxes"]), 1)
        targets = self.preprocess(targets.to(dev), batch_size, scale_tensor=imgsz[[1, 0, 1, 0]])
        gt_labels, gt_bboxes = targets.split((1, 4), 2)
        mask_gt = gt_bboxes.sum(2, keepdim=True).gt_(0.0)
        pred_bboxes = self.bbox_decode(anchor_points, pred_distri)
        _, target_bboxes, target_scores, fg_mask, _ = self.assigner(
            pred_scores.detach().sigmoid(), (pred_bboxes.detach() * stride_tensor).type(gt_bboxes.dtype),
            anchor_points * stride_tensor, gt_labels, gt_bboxes, mask_gt)
        target_scores_sum = max(target_scores.sum(), 1)
        loss[1] = F.binary_cross_entropy_with_logits(pred_scores, target_scores.to(dtype), reduction="none").sum() / target_scores_sum
        if fg_mask.sum():
            target_bboxes = target_bboxes / stride_tensor
            weight = target_scores.sum(-1)[fg_mask].unsqueeze(-1)
            iou = bbox_iou_ciou(pred_bboxes[fg_mask], target_bboxes[fg_mask])
            loss[0] = ((1.0 - iou) * weight).sum() / target_scores_sum
            target_ltrb = bbox2dist(anchor_points, target_bboxes, self.reg_max - 1)
            loss[2] = (self._dfl(pred_distri[fg_mask].view(-1, self.reg_max), target_ltrb[fg_mask]) * weight).sum() / target_scores_sum
        loss[0] *= self.hyp.box
        loss[1] *= self.hyp.cls
        loss[2] *= self.hyp.dfl
        return loss * batch_size, loss.detach()

    def _dfl(self, pred_dist, target):
        target = target.clamp_(0, self.reg_max - 1 - 0.01)
        tl = target.long()
        tr = tl + 1
        wl = tr - target
        wr = 1 - wl
        return (F.cross_entropy(pred_dist, tl.view(-1), reduction="none").view(tl.shape) * wl
                + F.cross_entropy(pred_dist, tr.view(-1), reduction="none").view(tl.shape) * wr).mean(-1, keepdim=True)


# ----------------------------------------------------------------------------------------------
# non_max_suppression (ultralytics/utils/nms.py), the options the reference uses
# ----------------------------------------------------------------------------------------------
def non_max_suppression(prediction, conf_thres=0.25, iou_thres=0.45, multi_label=False, agnostic=False, max_det=300,
                        max_nms=30000, max_wh=7680):
    """prediction [B, 4+nc, A] (xywh pixels + scores) -> list of [n, 6] (x1, y1, x2, y2, conf, cls) per image."""
    import torchvision
    bs, nc = prediction.shape[0], prediction.shape[1] - 4
    multi_label &= nc > 1
    xc = prediction[:, 4:4 + nc].amax(1) > conf_thres
    prediction = prediction.transpose(-1, -2)
    prediction = torch.cat((xywh2xyxy(prediction[..., :4]), prediction[..., 4:]), dim=-1)
    output = [torch.zeros((0, 6), device=prediction.device)] * bs
    for xi, x in enumerate(prediction):
        x = x[xc[xi]]
        if not x.shape[0]:
            continue
        box, cls = x.split((4, nc), 1)
        if multi_label:
            i, j = torch.where(cls > conf_thres)
            x = torch.cat((box[i], x[i, 4 + j, None], j[:, None].float()), 1)
        else:
            conf, j = cls.max(1, keepdim=True)
            x = torch.cat((box, conf, j.float()), 1)[conf.view(-1) > conf_thres]
        n = x.shape[0]
        if not n:
            continue
        if n > max_nms:
            x = x[x[:, 4].argsort(descending=True)[:max_nms]]
        c = x[:, 5:6] * (0 if agnostic else max_wh)
        scores = x[:, 4]
        boxes = x[:, :4] + c
        i = torchvision.ops.nms(boxes, scores, iou_thres)
        output[xi] = x[i[:max_det]]
    return output
