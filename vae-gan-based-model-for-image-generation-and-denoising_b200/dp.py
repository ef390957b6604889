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
                 producer_streams: Sequence["torch.cuda.Stream"] = (),
                 reducer: Optional[Callable[[Bucket, bool], None]] = None):
        """`producer_streams`: side streams that also write gradients into `flat` (the fused step issues its weight
        gradients there); a bucket's all-reduce waits for them as well as for the current stream.
        `reducer(bucket, first)`: what to launch for a ready bucket instead of an NCCL all-reduce - the fused step's
        peer-memory kernel (PeerAdam.reduce_and_step: reduce-scatter + sharded Adam + all-gather); `first` marks the
        first bucket of a backward pass."""
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
        self.last_event = None
        self.reducer = reducer

    def reset(self):
        self.pending = [len(b.params) for b in self.buckets]
        self.launched = [False] * len(self.buckets)

    def _launch(self, b_idx: int):
        b = self.buckets[b_idx]
        self.launched[b_idx] = True
        if self.world == 1 and self.reducer is None:
            return
        view = self.flat[b.lo:b.hi]
        if self.comm_stream is not None:
            self.comm_stream.wait_stream(torch.cuda.current_stream())     # gradients of this bucket are complete
            for st in self.producer_streams:
                self.comm_stream.wait_stream(st)
            with torch.cuda.stream(self.comm_stream):
                if self.reducer is not None:
                    self.reducer(b, sum(self.launched) == 1)
                else:
                    dist.all_reduce(view, group=self.group)
                # consumers wait for THIS buffer's last bucket, not for whatever else shares the communication stream
                self.last_event = torch.cuda.Event()
                self.last_event.record(self.comm_stream)
        elif self.reducer is not None:
            self.reducer(b, sum(self.launched) == 1)
        else:
            dist.all_reduce(view, group=self.group)

    def mark_ready(self, param_index: int):
        b_idx = self.owner[param_index]
        if self.launched[b_idx]:
            return
        self.pending[b_idx] -= 1
        if self.pending[b_idx] == 0 and not self.launched[b_idx]:
            self._launch(b_idx)

    def flush(self):
        """Launch whatever the backward pass did not trigger (parameters that received no gradient); does not wait."""
        for b_idx in range(len(self.buckets)):
            if not self.launched[b_idx]:
                self._launch(b_idx)

    def wait(self):
        """Make the current stream wait for every launched bucket and reset the state for the next backward pass."""
        if self.comm_stream is not None and self.last_event is not None:
            torch.cuda.current_stream().wait_event(self.last_event)
        self.last_event = None
        self.reset()

    def finish(self):
        self.flush()
        self.wait()


class PeerAdam:
    """The data-parallel optimizer step of ONE network over NVLink / NVSwitch peer memory (csrc/dp_comm.cu): per
    gradient bucket one kernel does reduce-scatter(gradients) -> Adam on this rank's 1/world slice -> all-gather
    (parameters), with multimem instructions when the fabric exposes a multicast mapping and peer loads / stores
    otherwise.  `grads` / `params` must come from torch.distributed._symmetric_memory.empty(); the optimizer state
    (exp_avg, exp_avg_sq) is thereby sharded: each rank keeps only its slices current (gather_moments() before a
    checkpoint).  Replaces `dist.all_reduce` + the replicated Adam kernel (vaegan_code.py:105,134-135 on N replicas)."""

    def __init__(self, grads, params, exp_avg, exp_avg_sq, step_count, hyper, group, shared, write_grads: bool):
        import ctypes
        import torch.distributed._symmetric_memory as symm_mem
        from . import _lib
        self._lib, self._ctypes = _lib, ctypes
        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        if self.world > _lib.DP_MAX_RANKS:
            raise RuntimeError(f"peer-memory optimizer step supports up to {_lib.DP_MAX_RANKS} ranks")
        self.grads, self.params, self.m, self.v, self.step_count = grads, params, exp_avg, exp_avg_sq, step_count
        self.lr, self.betas, self.eps = hyper
        self.write_grads = int(write_grads)
        self.shared = shared
        hg, hp = symm_mem.rendezvous(grads, self.group), symm_mem.rendezvous(params, self.group)
        for h, t in ((hg, grads), (hp, params)):
            if int(h.buffer_ptrs[self.rank]) != t.data_ptr():
                raise RuntimeError("symmetric-memory handle does not map the tensor at offset 0")
        self._handles = (hg, hp)                    # keep the mappings alive
        c = _lib.VgDpComm()
        mc_g, mc_p = int(hg.multicast_ptr or 0), int(hp.multicast_ptr or 0)
        use_mc = bool(mc_g and mc_p) and shared.allow_multicast
        c.mc_grads, c.mc_params = (mc_g, mc_p) if use_mc else (None, None)
        for r in range(self.world):
            c.peer_grads[r], c.peer_params[r] = int(hg.buffer_ptrs[r]), int(hp.buffer_ptrs[r])
            c.peer_sig[r] = int(shared.sig_handle.buffer_ptrs[r])
        c.epoch, c.rank, c.world = shared.epoch.data_ptr(), self.rank, self.world
        self.comm, self.multicast = c, use_mc
        self.max_blocks = int(_lib.load().vg_dp_max_blocks())

    def reduce_and_step(self, bucket: Bucket, first: bool):
        ct, lib = self._ctypes, self._lib
        stream = ct.c_void_p(torch.cuda.current_stream().cuda_stream)
        P = lambda t: ct.c_void_p(t.data_ptr())
        if first:                                   # one tick per optimizer step, before its first bucket
            lib.call("vg_adam_tick", P(self.step_count), stream)
        n = bucket.hi - bucket.lo
        per_rank_vec = (n // 4 + self.world - 1) // self.world
        blocks = max(1, min(self.max_blocks, (per_rank_vec + 2047) // 2048))
        lib.call("vg_dp_adam_bucket", ct.byref(self.comm), bucket.lo, n, P(self.m), P(self.v), float(self.lr),
                 float(self.betas[0]), float(self.betas[1]), float(self.eps), P(self.step_count), 1.0 / self.world,
                 blocks, self.write_grads, P(self.shared.err), stream)

    def gather_moments(self, buckets: Sequence[Bucket]):
        """Make exp_avg / exp_avg_sq complete on every rank (each slice from its owner) - before state_dict()."""
        for b in buckets:
            nvec = (b.hi - b.lo) // 4
            per = (nvec + self.world - 1) // self.world
            for r in range(self.world):
                lo, hi = b.lo + 4 * min(nvec, per * r), b.lo + 4 * min(nvec, per * (r + 1))
                if hi > lo:
                    src = dist.get_global_rank(self.group, r)
                    dist.broadcast(self.m[lo:hi], src, group=self.group)
                    dist.broadcast(self.v[lo:hi], src, group=self.group)


class PeerShared:
    """Per-step state shared by the PeerAdam objects of the three networks: the cross-GPU signal pad (symmetric
    memory), the local epoch counter and the barrier-timeout flag."""

    def __init__(self, device, group, allow_multicast: bool = True):
        import torch.distributed._symmetric_memory as symm_mem
        from . import _lib
        group = group if group is not None else dist.group.WORLD
        world = dist.get_world_size(group)
        blocks = int(_lib.load().vg_dp_max_blocks())
        self.sig = symm_mem.empty(blocks * world, dtype=torch.int32, device=device)
        self.sig.zero_()
        self.sig_handle = symm_mem.rendezvous(self.sig, group)
        self.epoch = torch.zeros(2, dtype=torch.int32, device=device)
        self.err = torch.zeros(1, dtype=torch.int32, device=device)
        self.allow_multicast = allow_multicast
        torch.cuda.synchronize(device)
        dist.barrier(group)                 # every pad is zero before anyone signals into it

    def check(self):
        if int(self.err) != 0:
            raise RuntimeError("peer-memory optimizer step: a cross-GPU barrier timed out (a rank is missing or "
                               "launched a different sequence of buckets)")


class LocalAdam:
    """One GPU: the optimizer step of a network issued bucket by bucket from inside its backward pass (the reducer hook
    of BucketedAllReduce with nothing to reduce) - Adam on the ranges whose gradients are complete runs on a side
    stream under the remaining dgrad / wgrad launches; only the last bucket's update is left when backward ends."""

    def __init__(self, opt):
        from . import _lib
        self._lib, self.opt = _lib, opt

    def reduce_and_step(self, bucket: Bucket, first: bool):
        import ctypes
        o, lib = self.opt, self._lib
        stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        P = lambda t, off=0: ctypes.c_void_p(t.data_ptr() + 4 * off)
        if first:
            lib.call("vg_adam_tick", P(o.step_count), stream)
        lo, n = bucket.lo, bucket.hi - bucket.lo
        lib.call("vg_adam_apply", P(o.params, lo), P(o.grads, lo), P(o.exp_avg, lo), P(o.exp_avg_sq, lo), n,
                 float(o.lr), float(o.betas[0]), float(o.betas[1]), float(o.eps), ctypes.c_void_p(o.step_count.data_ptr()),
                 1.0, stream)
