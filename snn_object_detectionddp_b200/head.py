"""YOLO `Detect` head, restated (PARITY UNPINNED).

The reference instantiates ultralytics' ``Detect(nc, ch)`` (reference model.py:186-192); ultralytics is an
un-vendored, unpinned third-party dependency that cannot be installed offline, so its published
structure (ultralytics/nn/modules/head.py, v8.3.x "non-legacy" head used by YOLO11) is restated here with the
same attribute names, hence the same ``detection_head.*`` state_dict keys:

    cv2[i] = Conv(ch,c2,3) -> Conv(c2,c2,3) -> Conv2d(c2, 4*reg_max, 1)                      (box / DFL logits)
    cv3[i] = [DWConv(ch,ch,3) -> Conv(ch,c3,1)] -> [DWConv(c3,c3,3) -> Conv(c3,c3,1)] -> Conv2d(c3,nc,1)  (class logits)
    Conv = Conv2d(bias=False) + BatchNorm2d(eps=1e-3, momentum=0.03) + SiLU ;  c2 = max(16, ch//4, 4*reg_max), c3 = max(ch, min(nc,100))

All convs run on the libsnnb200 kernels; the head keeps SiLU (it is not part of the spiking backbone).
"""
import torch
import torch.nn as nn

from . import kernels as K
from .model import ConvBlock
from .ops import ConvBiasFn, NeuronCfg
from ._lib import GEOM_1x1


class DFL(nn.Module):
    """Integral module of Distribution Focal Loss: expectation over reg_max bins (frozen arange 1x1 conv)."""

    def __init__(self, c1=16):
        super().__init__()
        self.conv = nn.Conv2d(c1, 1, 1, bias=False).requires_grad_(False)
        self.conv.weight.data[:] = torch.arange(c1, dtype=torch.float).view(1, c1, 1, 1)
        self.c1 = c1

    def forward(self, x):  # x [B, 4*c1, A]
        b, _, a = x.shape
        proj = self.conv.weight.view(-1).to(x.dtype)
        return (x.view(b, 4, self.c1, a).softmax(2) * proj.view(1, 1, -1, 1)).sum(2)


def make_anchors(shapes, strides, offset=0.5, device=None, dtype=torch.float32):
    """Anchor centres (grid units) and per-anchor stride for maps of spatial `shapes` [(h,w),...]."""
    pts, st = [], []
    for (h, w), s in zip(shapes, strides):
        sx = torch.arange(w, device=device, dtype=dtype) + offset
        sy = torch.arange(h, device=device, dtype=dtype) + offset
        yy, xx = torch.meshgrid(sy, sx, indexing="ij")
        pts.append(torch.stack((xx, yy), -1).view(-1, 2))
        st.append(torch.full((h * w, 1), float(s), device=device, dtype=dtype))
    return torch.cat(pts), torch.cat(st)


def dist2bbox(distance, anchor_points, xywh=True, dim=-1):
    lt, rb = distance.chunk(2, dim)
    x1y1 = anchor_points - lt
    x2y2 = anchor_points + rb
    if xywh:
        return torch.cat(((x1y1 + x2y2) / 2, x2y2 - x1y1), dim)
    return torch.cat((x1y1, x2y2), dim)


class HeadOut:
    """Head outputs of one timestep: per scale `box` [B,h,w,4*reg_max] and `cls` [B,h,w,nc], fp32 NHWC.

    The per-scale tensors are views of two SCALE-MAJOR buffers `flat_box` [B*A, 4*reg_max] / `flat_cls` [B*A, nc] (scale i
    occupies rows [B*a_off[i], B*a_off[i+1]), image-major inside) that the closing 1x1 convs write directly, so the loss
    kernels read the predictions without any torch.cat (they index rows through `a_off`)."""

    def __init__(self, box, cls, training, flat_box=None, flat_cls=None, a_off=None):
        self.box, self.cls, self.training = box, cls, training
        self.flat_box, self.flat_cls, self.a_off = flat_box, flat_cls, a_off

    def shapes(self):
        return [tuple(b.shape[1:3]) for b in self.box]

    def flat(self):
        """(pred_distri [B,A,4*reg_max], pred_scores [B,A,nc]) -- anchors ordered scale-major, row-major."""
        b = self.box[0].shape[0]
        return (torch.cat([x.reshape(b, -1, x.shape[-1]) for x in self.box], 1),
                torch.cat([x.reshape(b, -1, x.shape[-1]) for x in self.cls], 1))

    def maps_nchw(self):
        return [torch.cat([b, c], -1).permute(0, 3, 1, 2) for b, c in zip(self.box, self.cls)]

    def as_reference(self, head):
        """What ultralytics' Detect.forward returns: list of [B,no,h,w] (train) or (decoded [B,4+nc,A], maps) (eval)."""
        maps = self.maps_nchw()
        if self.training:
            return maps
        return head.decode(self), maps


class Detect(nn.Module):
    def __init__(self, nc=80, ch=()):
        super().__init__()
        self.nc, self.nl, self.reg_max = nc, len(ch), 16
        self.no = nc + self.reg_max * 4
        self.stride = torch.zeros(self.nl)
        c2, c3 = max((16, ch[0] // 4, self.reg_max * 4)), max(ch[0], min(self.nc, 100))
        silu = NeuronCfg("silu")

        def conv(a, b, k):
            return ConvBlock(a, b, kernel_size=k, padding=k // 2, neuron=silu, bn_eps=1e-3, bn_momentum=0.03)

        def dw(a):
            return ConvBlock(a, a, kernel_size=3, groups=a, neuron=silu, bn_eps=1e-3, bn_momentum=0.03)

        self.cv2 = nn.ModuleList(nn.Sequential(conv(x, c2, 3), conv(c2, c2, 3), nn.Conv2d(c2, 4 * self.reg_max, 1)) for x in ch)
        self.cv3 = nn.ModuleList(
            nn.Sequential(nn.Sequential(dw(x), conv(x, c3, 1)), nn.Sequential(dw(c3), conv(c3, c3, 1)), nn.Conv2d(c3, self.nc, 1))
            for x in ch)
        self.dfl = DFL(self.reg_max)

    def forward_seq(self, rc, feats, B, last_only=True):
        """feats: 3 bf16 NHWC maps [T*B,h,w,ch].  Every timestep runs (BatchNorm running statistics advance once per
        frame exactly like the reference's per-frame calls); the returned HeadOut holds the last step (train.py:66)."""
        boxes, clss = [], []
        # scale-major prediction buffers (see HeadOut): the frames that are returned, all scales
        n_out = B if last_only else feats[0].shape[0]
        hw = [f.shape[1] * f.shape[2] for f in feats]
        a_off = [0]
        for v in hw:
            a_off.append(a_off[-1] + v)
        dev = feats[0].device
        flat_box = torch.empty((n_out * a_off[-1], 4 * self.reg_max), device=dev, dtype=torch.float32)
        flat_cls = torch.empty((n_out * a_off[-1], self.nc), device=dev, dtype=torch.float32)
        for i in range(self.nl):
            x = feats[i]
            if x.dtype != torch.bfloat16:
                x = x.to(torch.bfloat16)
            dead = dict(dead_frames_ok=last_only)
            b, _ = self.cv2[i][0].forward_seq(rc, x, **dead)
            b, _ = self.cv2[i][1].forward_seq(rc, b, **dead)
            c, _ = self.cv3[i][0][0].forward_seq(rc, x, **dead)
            c, _ = self.cv3[i][0][1].forward_seq(rc, c, **dead)
            c, _ = self.cv3[i][1][0].forward_seq(rc, c, **dead)
            c, _ = self.cv3[i][1][1].forward_seq(rc, c, **dead)
            if last_only:
                # the closing 1x1 convs have no BatchNorm (no per-frame side effect): only the frames that are returned
                b, c = b[-B:], c[-B:]
            h, w = x.shape[1], x.shape[2]
            rows = slice(n_out * a_off[i], n_out * a_off[i + 1])
            cfg_b = dict(store=rc.store, geom=GEOM_1x1, out_dtype=torch.float32, out=flat_box[rows].view(n_out, h, w, 4 * self.reg_max))
            cfg_c = dict(store=rc.store, geom=GEOM_1x1, out_dtype=torch.float32, out=flat_cls[rows].view(n_out, h, w, self.nc))
            b = ConvBiasFn.apply(b, self.cv2[i][2].weight, self.cv2[i][2].bias, cfg_b)
            c = ConvBiasFn.apply(c, self.cv3[i][2].weight, self.cv3[i][2].bias, cfg_c)
            boxes.append(b)
            clss.append(c)
        return HeadOut(boxes, clss, self.training, flat_box, flat_cls, a_off)

    def decode(self, out):
        """Eval-mode Detect._inference: DFL expectation -> dist2bbox(xywh) * stride, sigmoid class scores -> [B,4+nc,A]."""
        distri, scores = out.flat()                                   # [B,A,64], [B,A,nc]
        anchors, strides = make_anchors(out.shapes(), self.stride.tolist(), 0.5, device=distri.device)
        b, a, _ = distri.shape
        proj = torch.arange(self.reg_max, device=distri.device, dtype=distri.dtype)
        dist = (distri.view(b, a, 4, self.reg_max).softmax(3) * proj).sum(3)    # [B,A,4]
        dbox = dist2bbox(dist, anchors.unsqueeze(0), xywh=True, dim=2) * strides
        return torch.cat((dbox, scores.sigmoid()), 2).permute(0, 2, 1)

    def forward(self, x):
        """Drop-in: list of 3 NCHW fp32 maps -> ultralytics-style output."""
        from .model import RunCtx, _to_nhwc_bf16
        from .params import store_for
        st = store_for(self, x[0].device)
        st.refresh_operands()
        out = self.forward_seq(RunCtx(st, 1), [_to_nhwc_bf16(f) for f in x], x[0].shape[0])
        return out.as_reference(self)
