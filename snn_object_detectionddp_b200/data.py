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


class DevicePrefetcher:
    """Double-buffered host -> device pipeline for (frames, padded labels) batches (the reference moves
    each batch synchronously with `.to(device)` at train.py:59-60 after dividing by 255 on the host, dataset.py:152).

    Frames stay uint8 on the host (4x fewer PCIe bytes; the `/255` runs in the frame-packer kernel, bit-identical),
    live in pinned staging buffers and are copied on a side stream while the previous step computes.  Usage:

        pf = DevicePrefetcher(device)
        pf.stage(frames_u8, padded)            # first batch
        for next_batch in loader:
            frames_d, padded_d = pf.take()     # waits (on the compute stream) for the staged copy
            pf.stage(*next_batch)              # overlaps with the step below
            trainer.train_step_graphed(frames_d, {"padded": padded_d})
            pf.release()                       # the step's kernels are queued: the buffer may be refilled after them
    """

    def __init__(self, device, depth=2):
        self.device = torch.device(device)
        self.depth = depth
        self.copy_stream = torch.cuda.Stream(self.device)
        self.ready = [torch.cuda.Event() for _ in range(depth)]
        self.consumed = [torch.cuda.Event() for _ in range(depth)]
        self.pin = [None] * depth
        self.dev = [None] * depth
        self._w = self._r = 0
        main = torch.cuda.current_stream(self.device)
        for e in self.consumed:
            e.record(main)

    @staticmethod
    def _like(tensors, **kw):
        return tuple(torch.empty(t.shape, dtype=t.dtype, **kw) for t in tensors)

    def stage(self, frames, padded):
        """Queue the copy of one host batch (frames uint8|fp32 [B,T,3,H,W], padded = pad_targets(...) tuple)."""
        s = self._w % self.depth
        self._w += 1
        host = (frames,) + tuple(padded)
        if self.dev[s] is None or any(d.shape != h.shape or d.dtype != h.dtype for d, h in zip(self.dev[s], host)):
            self.dev[s] = self._like(host, device=self.device)
            self.pin[s] = None
        src = host
        if not all(h.is_pinned() for h in host):
            # pageable batch: stage it in pinned memory first (a loader that decodes into staging_buffers() skips this copy)
            if self.pin[s] is None:
                self.pin[s] = tuple(t.pin_memory() for t in self._like(host))
            self.ready[s].synchronize()                            # the previous H2D out of this staging slot has finished
            for p, h in zip(self.pin[s], host):
                if p.data_ptr() != h.data_ptr():
                    p.copy_(h)
            src = self.pin[s]
        self._bytes = sum(t.numel() * t.element_size() for t in host)
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.consumed[s])          # the step that read this device buffer has run
            for d, h in zip(self.dev[s], src):
                d.copy_(h, non_blocking=True)
            self.ready[s].record(self.copy_stream)

    def staging_buffers(self, like=None):
        """Pinned host tensors of the NEXT stage() slot (frames, *padded): a loader can decode straight into them and
        pass them to stage().  `like` = (frames, *padded) example tensors to (re)allocate the slot."""
        s = self._w % self.depth
        if like is not None and (self.pin[s] is None or any(p.shape != h.shape or p.dtype != h.dtype for p, h in zip(self.pin[s], like))):
            self.pin[s] = tuple(t.pin_memory() for t in self._like(tuple(like)))
        if self.pin[s] is not None:
            self.ready[s].synchronize()
        return self.pin[s]

    def take(self):
        s = self._r % self.depth
        torch.cuda.current_stream(self.device).wait_event(self.ready[s])
        return self.dev[s][0], self.dev[s][1:]

    def release(self):
        s = self._r % self.depth
        self._r += 1
        self.consumed[s].record(torch.cuda.current_stream(self.device))

    def bytes_per_batch(self):
        return getattr(self, "_bytes", 0)
