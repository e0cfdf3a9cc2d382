"""Data-parallel gradient synchronisation: bucketed all-reduce of contiguous ranges of the flat gradient buffer,
launched while backward is still running.

The reference has NO distributed code despite its name; this is new functionality whose contract is
stock-DDP semantics: one process per GPU, per-rank batch, gradients averaged over ranks before clip + AdamW
(train.py:77-78), BatchNorm statistics per rank (no SyncBN in the reference).

Why it is cheap here: parameters live in ONE flat fp32 gradient buffer (params.py) laid out in forward order, so a
bucket is just a slice -- no flatten/unflatten copies.  The backward kernels accumulate weight gradients straight
into that buffer and call ``entry_ready`` when a tensor's gradient is final; when every tensor of a bucket is ready
the slice is all-reduced asynchronously (NCCL runs it on its own stream, ordered after the kernels already queued on
the compute stream, so it overlaps the rest of backward).  ``finish`` launches whatever is left and makes the compute
stream wait for all of them.

Works with any torch.distributed backend; CPU tests drive it over gloo with world_size 2.
"""
import torch
import torch.distributed as dist


class GradBucketer:
    def __init__(self, flat_g, spans, bucket_bytes=32 << 20, group=None, average=True):
        """spans: list of (offset, numel) of every tensor in flat-buffer order (elements of flat_g)."""
        self.flat_g, self.group, self.average = flat_g, group, average
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.spans = list(spans)
        es = flat_g.element_size()
        # buckets are built from the END of the buffer (gradients of the last layers are final first)
        self.buckets = []          # (lo, hi, [entry indices])
        self.bucket_of = [None] * len(self.spans)
        cur, cur_bytes, hi = [], 0, None
        for i in range(len(self.spans) - 1, -1, -1):
            off, n = self.spans[i]
            if hi is None:
                hi = off + n
            cur.append(i)
            cur_bytes += n * es
            if cur_bytes >= bucket_bytes or i == 0:
                self.buckets.append((off, hi, cur))
                cur, cur_bytes, hi = [], 0, None
        for b, (_, _, idxs) in enumerate(self.buckets):
            for i in idxs:
                self.bucket_of[i] = b
        self._use_avg = False
        if self.world > 1 and average and flat_g.is_cuda and dist.get_backend(group) == "nccl":
            self._use_avg = True
        self.begin_step()

    # ------------------------------------------------------------------------------------------
    def begin_step(self):
        self._pending = [len(idxs) for _, _, idxs in self.buckets]
        self._seen = [False] * len(self.spans)
        self._launched = [False] * len(self.buckets)
        self._works = []
        self.launch_order = []

    def entry_ready(self, i):
        """Gradient of tensor i is final (called from the backward kernels' host code)."""
        if self.world == 1 or self._seen[i]:
            return
        self._seen[i] = True
        b = self.bucket_of[i]
        self._pending[b] -= 1
        if self._pending[b] == 0:
            self._launch(b)

    def _launch(self, b):
        lo, hi, _ = self.buckets[b]
        self._launched[b] = True
        self.launch_order.append(b)
        view = self.flat_g[lo:hi]
        op = dist.ReduceOp.AVG if self._use_avg else dist.ReduceOp.SUM
        self._works.append((dist.all_reduce(view, op=op, group=self.group, async_op=True), view))

    def finish(self):
        """Launch the buckets that never completed (tensors without a gradient this step), then wait for all."""
        if self.world == 1:
            return
        for b in range(len(self.buckets)):
            if not self._launched[b]:
                self._launch(b)
        for work, view in self._works:
            work.wait()
            if self.average and not self._use_avg:
                view.div_(self.world)
        self._works = []


def broadcast_module_state(module, flat_p=None, group=None, src=0):
    """Identical initial weights and buffers on every rank (rank `src` wins), as DDP's constructor does."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    if flat_p is not None:
        dist.broadcast(flat_p, src=src, group=group)
    else:
        for p in module.parameters():
            dist.broadcast(p.data, src=src, group=group)
    for b in module.buffers():
        if b.is_floating_point() or b.dtype in (torch.int64, torch.int32):
            dist.broadcast(b, src=src, group=group)
