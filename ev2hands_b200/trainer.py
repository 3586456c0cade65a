"""Data-parallel training step (BASELINE.json configs[3]): one process per GPU, per-replica BatchNorm statistics
like the reference's ``nn.DataParallel`` (``train.py:68``), gradients averaged over ranks with bucketed NCCL
all-reduces that are launched from backward hooks, so the exchange overlaps the rest of the backward pass.

Reference loop: ``src/Ev2Hands/train.py:70-92`` (forward, criterion, ``loss = sum(losses.values())``,
``zero_grad``, ``backward``, ``step``; Adam lr 1e-3, ``:23,56``).

Layout: every parameter's ``.grad`` is a VIEW into one flat fp32 buffer (allocated once), in reverse registration
order - the order backward produces them - cut into a few contiguous buckets.  A post-accumulate-grad hook counts a
bucket's parameters; when the last one has its gradient, the bucket's slice of the flat buffer is all-reduced in
place (``ReduceOp.AVG``, asynchronous: NCCL's own stream) - no ``torch.cat``, no copy back.  ``finish()`` flushes
buckets whose parameters received no gradient this step and waits for the handles before the optimiser reads them.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


class BucketedGradReducer:
    def __init__(self, params, n_buckets: int = 4, group=None):
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        order = list(reversed(self.params))                      # backward reaches the last layers first
        total = sum(p.numel() for p in order)
        dev = order[0].device
        self.flat = torch.zeros(total, dtype=order[0].dtype, device=dev)
        target = (total + n_buckets - 1) // n_buckets
        self.buckets = []            # [lo, hi, n_params]
        self._bucket_of = {}
        off, lo, count = 0, 0, 0
        for p in order:
            n = p.numel()
            p.grad = self.flat[off:off + n].view_as(p)
            self._bucket_of[p] = len(self.buckets)
            off += n
            count += 1
            if off - lo >= target or p is order[-1]:
                self.buckets.append([lo, off, count])
                lo, count = off, 0
        self._ready = [0] * len(self.buckets)
        self._launched = [False] * len(self.buckets)
        self._handles = []
        self.bytes = total * self.flat.element_size()
        self._hooks = [p.register_post_accumulate_grad_hook(self._on_grad) for p in self.params]

    # -- helpers --------------------------------------------------------------------------------
    def _active(self):
        return dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1

    def _launch(self, b):
        self._launched[b] = True
        if self._active():
            lo, hi, _ = self.buckets[b]
            seg = self.flat[lo:hi]
            if dist.get_backend(self.group) == "nccl":
                self._handles.append(dist.all_reduce(seg, op=dist.ReduceOp.AVG, group=self.group, async_op=True))
            else:                                                # gloo (CPU tests) has no AVG
                h = dist.all_reduce(seg, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
                self._handles.append((h, seg))

    def _on_grad(self, p):
        if p.grad.data_ptr() < self.flat.data_ptr() or p.grad.data_ptr() >= self.flat.data_ptr() + self.bytes:
            # something replaced .grad (zero_grad(set_to_none=True) + a fresh accumulation): move it back into the flat buffer
            b = self._bucket_of[p]
            view = self._view_of(p)
            view.copy_(p.grad)
            p.grad = view
        b = self._bucket_of[p]
        self._ready[b] += 1
        if self._ready[b] == self.buckets[b][2] and not self._launched[b]:
            self._launch(b)

    def _view_of(self, p):
        off = 0
        for q in reversed(self.params):
            if q is p:
                return self.flat[off:off + q.numel()].view_as(q)
            off += q.numel()
        raise KeyError("parameter not managed by this reducer")

    # -- per step -------------------------------------------------------------------------------
    def zero_grad(self):
        """zero the flat buffer (gradients stay views of it) and re-arm the buckets"""
        self.flat.zero_()
        for p in self.params:
            if p.grad is None:
                p.grad = self._view_of(p)
        self._ready = [0] * len(self.buckets)
        self._launched = [False] * len(self.buckets)
        self._handles = []

    def finish(self):
        """after backward: reduce the buckets no hook completed, wait for every all-reduce"""
        for b in range(len(self.buckets)):
            if not self._launched[b]:
                self._launch(b)
        world = dist.get_world_size(self.group) if self._active() else 1
        for h in self._handles:
            if isinstance(h, tuple):
                h[0].wait()
                h[1].div_(world)
            else:
                h.wait()
        self._handles = []
        return self.bytes

    def remove(self):
        for h in self._hooks:
            h.remove()


def train_step(net, hands, batch, optimizer, reducer: BucketedGradReducer, losses_fn):
    """one iteration of the reference's loop (train.py:70-92) on this rank's shard of the batch -> (loss, losses)"""
    reducer.zero_grad()
    outs = net(batch["events"], hands)
    losses = losses_fn(outs, batch, hands)
    loss = sum(losses.values())
    loss.backward()
    reducer.finish()
    optimizer.step()
    return loss, losses
