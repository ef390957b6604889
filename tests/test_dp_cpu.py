"""CPU, world_size 2 and 4, gloo: the data-parallel host logic (sharding, bucket planning, bucketed all-reduce with the
1/world factor applied by the optimizer) reproduces "N reference replicas on disjoint shards, gradients averaged"
 - checked against the oracle run on the whole batch by hand-averaging."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _dp():
    from importlib import import_module
    import vaegan_b200  # noqa: F401
    return import_module("vaegan_b200.dp")


def test_shard_range_and_bucket_plan():
    dp = _dp()
    assert dp.shard_range(2048, 3, 8) == (768, 1024)
    with pytest.raises(ValueError):
        dp.shard_range(10, 0, 4)
    sizes = [100, 4, 300, 8, 50]
    offsets = [0, 100, 104, 404, 412]
    buckets = dp.plan_buckets(offsets, sizes, bucket_bytes=4 * 300)
    # reverse order, contiguous, complete, no parameter split
    covered = sorted(p for b in buckets for p in b.params)
    assert covered == [0, 1, 2, 3, 4]
    assert buckets[0].params[0] == 4 and buckets[0].hi == 462
    for b in buckets:
        assert b.lo == offsets[min(b.params)] and b.hi == offsets[max(b.params)] + sizes[max(b.params)]
    assert sum(b.hi - b.lo for b in buckets) == 462


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    from importlib import import_module
    import vaegan_b200  # noqa: F401
    dp = import_module("vaegan_b200.dp")
    from oracle import vaegan_oracle as vo
    hw, nz, gb = 64, 128, 4
    enc, gen, dis = vo.build_nets(vo.NetConfig(hw=hw, nz=nz))
    real, eps, n_real, n_fake = vo.make_inputs(gb, hw, nz)
    lo, hi = dp.shard_range(gb, rank, world)
    # one discriminator-side backward on this rank's shard (local BN statistics), as in vaegan_code.py:96-104
    bce = torch.nn.BCELoss()
    p_real = dis(real[lo:hi] + 0.05 * n_real[lo:hi])
    loss = bce(p_real, torch.full((hi - lo,), 0.9))
    loss.backward()
    params = list(dis.parameters())
    sizes = [(p.numel() + 3) // 4 * 4 for p in params]
    offsets = [sum(sizes[:i]) for i in range(len(sizes))]
    flat = torch.zeros(sum(sizes))
    for p, o in zip(params, offsets):
        flat[o:o + p.numel()] = p.grad.flatten()
    local = flat.clone()
    ar = dp.BucketedAllReduce(flat, offsets, sizes, bucket_bytes=1 << 20)
    for i in reversed(range(len(params))):      # backward order
        ar.mark_ready(i)
    ar.finish()
    gathered = [torch.zeros_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    want = sum(gathered)
    # (with more than two ranks the collective's summation order differs from sum(gathered): fp32 round-off)
    tol = dict(rtol=1e-6, atol=1e-7) if world == 2 else dict(rtol=1e-5, atol=1e-6 * float(want.abs().max()))
    ok = torch.allclose(flat, want, **tol) and len(ar.buckets) > 1
    # averaged gradient == mean of the replicas' gradients (what Adam sees with grad_scale = 1/world)
    ok = ok and torch.allclose(flat / world, sum(gathered) / world, **tol)
    # a second round after reset must work too
    flat.copy_(local)
    for i in reversed(range(len(params))):
        ar.mark_ready(i)
    ar.finish()
    ok = ok and torch.allclose(flat, want, **tol)
    # a custom per-bucket reducer (what the peer-memory transport plugs in): sharded "optimizer step" on CPU - every
    # rank reduces all buckets, applies p -= lr * g on ITS slice only and broadcasts the slice; flush() / wait() split
    flat.copy_(local)
    p_flat = torch.arange(flat.numel(), dtype=torch.float32) * 1e-3
    firsts, order = [], []

    def reducer(bucket, first):
        firsts.append(first)
        order.append(bucket.lo)
        view = flat[bucket.lo:bucket.hi]
        dist.all_reduce(view)
        n = bucket.hi - bucket.lo
        per = (n + world - 1) // world
        for r in range(world):
            a, b = bucket.lo + min(n, per * r), bucket.lo + min(n, per * (r + 1))
            if b > a:
                if r == rank:
                    p_flat[a:b] -= 0.1 * flat[a:b] / world
                dist.broadcast(p_flat[a:b], r)

    ar2 = dp.BucketedAllReduce(flat, offsets, sizes, bucket_bytes=1 << 20, reducer=reducer)
    for i in reversed(range(len(params))[1:]):   # the first parameter never reports: flush() must launch its bucket
        ar2.mark_ready(i)
    n_before = len(order)
    ar2.flush()
    ar2.wait()
    ok = ok and len(order) == len(ar2.buckets) and n_before == len(ar2.buckets) - 1
    ok = ok and firsts == [True] + [False] * (len(ar2.buckets) - 1) and order == sorted(order, reverse=True)
    want_p = torch.arange(flat.numel(), dtype=torch.float32) * 1e-3 - 0.1 * want / world
    ok = ok and torch.allclose(p_flat, want_p, **tol) and ar2.pending == [len(b.params) for b in ar2.buckets]
    out.put((rank, bool(ok), len(ar.buckets)))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_bucketed_allreduce_gloo(world):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    results = [out.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in results), results
