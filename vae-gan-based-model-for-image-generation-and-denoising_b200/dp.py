"""Batch-sharded data parallelism for the fused step: one process per GPU, bucketed gradient all-reduce.

The reference's VAE-GAN path is single-device (vaegan_code.py:28); its semantic extension is
"N replicas of the reference step on disjoint batch shards, gradients averaged": every loss term is a per-sample mean
(BCELoss mean :46, MSELoss mean :47, KL / batch_size :114), so with equal shards the global gradient is the mean of
the per-rank gradients.  BatchNorm statistics stay per rank (local BN), exactly N reference replicas.

Host-side pieces (device agnostic, exercised on CPU with gloo in tests/test_dp_cpu.py):
  * shard_range      - which rows of a global batch a rank owns
  * plan_buckets     - contiguous ranges of a flat gradient buffer, in REVERSE parameter order (the order backward
                       produces gradients), each ~bucket_bytes
  * BucketedAllReduce- issues one all-reduce per bucket as soon as every parameter in it has its gradient, on a
                       communication stream when running on CUDA so NCCL overlaps the remaining backward kernels;
                       the 1/world factor is folded into the fused Adam (grad_scale), not applied here.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(global_batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Rows [lo, hi) of the global batch owned by `rank` (equal shards; the global batch must divide)."""
    if global_batch % world != 0:
        raise ValueError(f"global batch {global_batch} is not divisible by world size {world}")
    per = global_batch // world
    return rank * per, (rank + 1) * per


@dataclass(frozen=True)
class Bucket:
    lo: int                    # element range [lo, hi) of the flat gradient buffer
    hi: int
    params: Tuple[int, ...]    # indices (into the optimizer's parameter list) whose gradients live in the range


def plan_buckets(offsets: Sequence[int], sizes: Sequence[int], bucket_bytes: int = 8 << 20,
                 elem_bytes: int = 4) -> List[Bucket]:
    """Group parameters (given by their offsets / padded sizes in the flat buffer, forward order) into buckets,
    walking from the LAST parameter to the first.  A parameter is never split; a bucket closes once it holds at least
    `bucket_bytes`."""
    buckets: List[Bucket] = []
    cur: List[int] = []
    cur_elems = 0
    for i in range(len(offsets) - 1, -1, -1):
        cur.append(i)
        cur_elems += sizes[i]
        if cur_elems * elem_bytes >= bucket_bytes or i == 0:
            lo = offsets[cur[-1]]
            hi = offsets[cur[0]] + sizes[cur[0]]
            buckets.append(Bucket(lo, hi, tuple(cur)))
            cur, cur_elems = [], 0
    return buckets


class BucketedAllReduce:
    """Sum-all-reduce of one flat gradient buffer, bucket by bucket, overlapped with the producer.

    `mark_ready(i)` is called (from the backward of layer i's kernels) once parameter i's gradient is complete; when
    the last parameter of a bucket arrives the bucket's all-reduce is launched.  `finish()` makes the consumer
    (Adam) wait for every bucket and resets the state for the next backward pass.
    """

    def __init__(self, flat: torch.Tensor, offsets: Sequence[int], sizes: Sequence[int], group=None,
                 bucket_bytes: int = 8 << 20, comm_stream: Optional["torch.cuda.Stream"] = None,
                 producer_streams: Sequence["torch.cuda.Stream"] = ()):
        """`producer_streams`: side streams that also write gradients into `flat` (the fused step issues its weight
        gradients there); a bucket's all-reduce waits for them as well as for the current stream."""
        self.flat = flat
        self.producer_streams = list(producer_streams)
        self.group = group
        self.buckets = plan_buckets(offsets, sizes, bucket_bytes, flat.element_size())
        self.owner = {}
        for b_idx, b in enumerate(self.buckets):
            for p in b.params:
                self.owner[p] = b_idx
        self.pending = [len(b.params) for b in self.buckets]
        self.launched = [False] * len(self.buckets)
        self.cuda = flat.is_cuda
        self.comm_stream = comm_stream if self.cuda else None
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1

    def reset(self):
        self.pending = [len(b.params) for b in self.buckets]
        self.launched = [False] * len(self.buckets)

    def _launch(self, b_idx: int):
        b = self.buckets[b_idx]
        self.launched[b_idx] = True
        if self.world == 1:
            return
        view = self.flat[b.lo:b.hi]
        if self.comm_stream is not None:
            self.comm_stream.wait_stream(torch.cuda.current_stream())     # gradients of this bucket are complete
            for st in self.producer_streams:
                self.comm_stream.wait_stream(st)
            with torch.cuda.stream(self.comm_stream):
                dist.all_reduce(view, group=self.group)
        else:
            dist.all_reduce(view, group=self.group)

    def mark_ready(self, param_index: int):
        b_idx = self.owner[param_index]
        if self.launched[b_idx]:
            return
        self.pending[b_idx] -= 1
        if self.pending[b_idx] == 0 and not self.launched[b_idx]:
            self._launch(b_idx)

    def finish(self):
        """Launch whatever was not triggered (parameters that received no gradient), then join."""
        for b_idx in range(len(self.buckets)):
            if not self.launched[b_idx]:
                self._launch(b_idx)
        if self.comm_stream is not None and self.world > 1:
            torch.cuda.current_stream().wait_stream(self.comm_stream)
        self.reset()
