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

import os

import torch
import torch.distributed as dist

_DRAIN = os.environ.get("WJ_DDP_NO_DRAIN") != "1"   # (opt-out for measurements)


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

    # ---- setup-time helpers (not on the step's critical path)
    def broadcast_(self, tensors) -> None:
        """In-place broadcast of rank 0's tensors to every rank of the group."""
        for t in tensors:
            dist.broadcast(t, src=dist.get_global_rank(self.group, 0) if self.group is not None else 0, group=self.group)

    def broadcast_int(self, value: int) -> int:
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(self.group) == "nccl" else torch.device("cpu")
        t = torch.tensor([int(value)], dtype=torch.int64, device=dev)
        self.broadcast_([t])
        return int(t.item())

    def any_rank(self, flag: bool) -> bool:
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(self.group) == "nccl" else torch.device("cpu")
        t = torch.tensor([1 if flag else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return bool(t.item())

    def begin(self, flat_grads: torch.Tensor) -> None:
        self._flat = flat_grads
        self._hi = flat_grads.numel()
        self._handles = []
        self.launched = []

    def _launch(self, lo: int) -> None:
        if lo >= self._hi:
            return
        chunk = self._flat[lo:self._hi]
        if _DRAIN and chunk.is_cuda:
            # The bucket is handed to NCCL only after the compute stream has drained up to here.  With the event-based
            # ordering alone (what all_reduce(async_op=True) sets up between the current stream and NCCL's) two of two
            # 8-GPU runs ended with one rank's weight gradients of the first bucket differing from the other ranks'
            # (bench.py dp_check names the tensors); with the drain the replicas stayed bit-identical and the step was
            # no slower (107.0 vs 107.9-108.2 ms at N = 8: the host is a whole backward ahead of the GPU anyway).
            torch.cuda.current_stream().synchronize()
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
