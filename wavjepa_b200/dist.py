"""Data-parallel gradient exchange: one process per GPU, NCCL over NVLink 5 / NVSwitch (torch.distributed plumbing).

The reference trains with Lightning's DDP strategy (train.py:174-179): a bucketed sum-all-reduce of the 111 M
trainable-parameter gradients (444 MB fp32) overlapped with backward, averaged over ranks.  Here the hand-written
backward fills ONE flat fp32 gradient buffer from its END towards its START (the flat layout is the reverse of the
order in which gradients become final), so a bucket is simply the next contiguous slice: `ready(offset)` is called
by the backward whenever gradients [offset:] are final, and every time >= bucket_bytes of new gradients are
available an asynchronous all-reduce of that slice is enqueued on NCCL's stream while the main stream keeps
computing.  The 1/world_size averaging is folded into the fused clip+AdamW kernel (grad_scale), so buckets are
plain sums.  There is no other collective on the path.
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.distributed as dist


class BucketedAllReduce:
    def __init__(self, bucket_bytes: int = 48 << 20, group: Optional["dist.ProcessGroup"] = None):
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        self.group = group
        self.world_size = dist.get_world_size(group)
        self.bucket_elems = max(1, bucket_bytes // 4)
        self._flat = None
        self._hi = 0
        self._handles: List = []
        self.launched: List[tuple] = []  # (lo, hi) of every bucket of the last step, for tests / reporting

    def begin(self, flat_grads: torch.Tensor) -> None:
        self._flat = flat_grads
        self._hi = flat_grads.numel()
        self._handles = []
        self.launched = []

    def _launch(self, lo: int) -> None:
        if lo >= self._hi:
            return
        chunk = self._flat[lo:self._hi]
        self._handles.append(dist.all_reduce(chunk, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        self.launched.append((lo, self._hi))
        self._hi = lo

    def ready(self, offset: int) -> None:
        """Gradients flat[offset:] are final.  Launches a bucket when enough new gradients accumulated (or at 0)."""
        if self._flat is None:
            return
        if offset == 0 or self._hi - offset >= self.bucket_elems:
            self._launch(offset)

    def finish(self) -> None:
        """Flushes the remainder and makes the current stream wait for every bucket."""
        self._launch(0)
        for h in self._handles:
            h.wait()
        self._handles = []
        self._flat = None
