"""TEST INFRASTRUCTURE ONLY -- whole-detector oracle: stand-in extractor -> OracleTemporalUNet -> OracleDetect, and the
reference's training step around it (train.py:48-80, 155-169), all plain PyTorch (runs on CPU or GPU).

Reference-pinned parts: TemporalUNet / init / step semantics / AdamW + OneCycleLR (oracle/snn_oracle.py, golden
fixtures).  PARITY UNPINNED parts: the LIF neuron (build-defined), Detect head + loss (ultralytics restated,
oracle/detect_oracle.py) and the feature extractor -- the reference's frozen pretrained YOLO11m (model.py:74-98)
cannot exist offline, so BOTH the product and this oracle use the same documented stand-in pyramid
(deterministic frozen random weights, seed 1234; see snn_object_detectionddp_b200/model.py YOLOFeatureExtractor).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this file.
"""
import math
from types import SimpleNamespace

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import detect_oracle as D
from . import snn_oracle as O


def _bf16(x):
    return x.to(torch.bfloat16).to(torch.float32)


class OracleFeatureExtractor(nn.Module):
    """Stand-in pyramid: 8x8 space-to-depth -> 1x1 (192->128)+SiLU -> P3; 3x3 s2 +SiLU -> P4; 3x3 s2 +SiLU -> P5;
    each level projected 1x1 -> 144.  Weights generated exactly like the product's buffers (same generator order)."""
    WIDTH, OUT = 128, 144

    def __init__(self, seed=1234):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        w, o = self.WIDTH, self.OUT

        def mk(rows, taps, k):
            return (torch.randn(rows, taps, k, generator=g) * math.sqrt(2.0 / (taps * k))).to(torch.bfloat16).float()

        def to_conv(t):   # [rows][taps][k] -> [rows, k, kh, kw]
            rows, taps, k = t.shape
            ks = 3 if taps == 9 else 1
            return t.reshape(rows, ks, ks, k).permute(0, 3, 1, 2).contiguous()

        for name, t in (("w_stem", mk(w, 1, 192)), ("w_d4", mk(w, 9, w)), ("w_d5", mk(w, 9, w)),
                        ("w_p3", mk(o, 1, w)), ("w_p4", mk(o, 1, w)), ("w_p5", mk(o, 1, w))):
            self.register_buffer(name, to_conv(t), persistent=False)

    @torch.no_grad()
    def forward(self, x):
        """x fp32 [N,3,H,W] in [0,1) -> three fp32 NCHW maps [N,144,H/8..H/32] (values bf16-representable)."""
        n, _, h, w = x.shape
        s = _bf16(x).reshape(n, 3, h // 8, 8, w // 8, 8).permute(0, 1, 3, 5, 2, 4).reshape(n, 192, h // 8, w // 8)
        f3 = _bf16(F.silu(F.conv2d(s, self.w_stem)))
        f4 = _bf16(F.silu(F.conv2d(f3, self.w_d4, stride=2, padding=1)))
        f5 = _bf16(F.silu(F.conv2d(f4, self.w_d5, stride=2, padding=1)))
        return _bf16(F.conv2d(f3, self.w_p3)), _bf16(F.conv2d(f4, self.w_p4)), _bf16(F.conv2d(f5, self.w_p5))


class OracleYOLOTemporalUNet(nn.Module):
    """reference model.py:148-211 with the stand-in extractor and the restated Detect head."""

    def __init__(self, num_classes=80, use_conv_lstm=True, hyp=None, neuron="lif", emulate_bf16=False,
                 widths=(128, 256, 512, 1024)):
        super().__init__()
        hyp = hyp or {"box": 7.5, "cls": 0.5, "dfl": 1.5, "reg_max": 16}
        self.args = SimpleNamespace(**hyp)
        self.nc = num_classes
        self.feature_extractor = OracleFeatureExtractor()
        self.temporal_unet = O.OracleTemporalUNet([144, 144, 144], neuron=neuron, emulate_bf16=emulate_bf16, widths=widths)
        self.detection_head = D.OracleDetect(nc=num_classes, ch=[144, 144, 144], reg_max=self.args.reg_max)
        self.detection_head.stride = torch.tensor([8.0, 16.0, 32.0])
        self.model = nn.ModuleList([self.detection_head])

    def forward(self, x, hidden_state=None):
        feats = self.feature_extractor(x)
        feats, new_hidden = self.temporal_unet(feats, hidden_state)
        return self.detection_head(list(feats)), new_hidden


def initialize_model_oracle(model):
    model.temporal_unet.apply(O.initialize_weights_oracle)      # weight_initialization.py:62-83
    return model


def synthetic_batch(B, T, H, W, nc=8, seed=42, device="cpu"):
    """SURVEY.md 8d synthetic inputs: frames U[0,1) [B,T,3,H,W]; labels [M,6] = (batch_idx, cls, cx, cy, w, h)."""
    g = torch.Generator().manual_seed(seed)
    frames = torch.rand(B, T, 3, H, W, generator=g)
    rows = []
    for b in range(B):
        n = int(torch.randint(0, 8, (1,), generator=g))
        for _ in range(n):
            c = int(torch.randint(0, nc, (1,), generator=g))
            cx, cy = (torch.rand(2, generator=g) * 0.8 + 0.1).tolist()
            w, h = (torch.rand(2, generator=g) * 0.25 + 0.05).tolist()
            rows.append([b, c, cx, cy, w, h])
    labels = torch.tensor(rows, dtype=torch.float32).reshape(-1, 6)
    return frames.to(device), labels.to(device)


def reference_train_step(model, loss_fn, optimizer, scheduler, frames, labels):
    """reference train.py:58-80, verbatim semantics."""
    optimizer.zero_grad(set_to_none=True)
    hidden = None
    for t in range(frames.shape[1]):
        preds, hidden = model(frames[:, t], hidden)
    batch = {"batch_idx": labels[:, 0], "cls": labels[:, 1], "bboxes": labels[:, 2:]}
    loss, items = loss_fn(preds, batch)
    loss.sum().backward()
    gn = torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=10.0)
    optimizer.step()
    if scheduler is not None:
        scheduler.step()
    return loss.detach(), items, gn


def make_reference_trainer(model, total_steps, max_lr=1e-4, weight_decay=5e-4):
    """reference train.py:155-169: v8DetectionLoss, AdamW(weight_decay) with default lr, OneCycleLR(max_lr, cos)."""
    loss_fn = D.OracleV8DetectionLoss(model)
    opt = torch.optim.AdamW(model.parameters(), weight_decay=weight_decay)
    sched = torch.optim.lr_scheduler.OneCycleLR(opt, max_lr=max_lr, total_steps=total_steps, pct_start=0.3,
                                                anneal_strategy="cos")
    return loss_fn, opt, sched
