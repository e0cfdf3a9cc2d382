"""Batch layout of the reference input pipeline and the synthetic-input generator used by bench / tests.

`custom_collate_fn` mirrors reference train.py:10-44: images stacked to [B,S,3,H,W]; labels -> [M,6] rows
(batch_idx, cls, cx, cy, w, h), empty -> [0,6].  `synthetic_batch` is SURVEY.md 8d's recipe (frames U[0,1) like
dataset.py:152's /255 scaling; 0-7 boxes per sample, normalised cxcywh).
"""
import torch


def custom_collate_fn(batch):
    images, labels = zip(*batch)
    images = torch.stack(images, 0)
    rows = []
    for i, lab in enumerate(labels):
        if lab is not None and lab.numel():
            lab = lab.reshape(-1, 5).float()
            rows.append(torch.cat((torch.full((lab.shape[0], 1), float(i)), lab), 1))
    return images, (torch.cat(rows, 0) if rows else torch.zeros((0, 6)))


def synthetic_batch(B, T, H, W, nc=8, seed=42, device="cpu"):
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
