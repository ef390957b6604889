"""Launched by tests/test_dp_gpu.py under torchrun (one rank per GPU, NCCL): the data-parallel fused step (bucketed
all-reduce launched from inside the backward passes, captured in the CUDA graph) against N independent single-GPU
steps on the same shards whose gradients are summed by hand - SURVEY.md section 8(e): "N replicas of the reference
step on disjoint batch shards, gradients averaged"."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from importlib import import_module
    import vaegan_b200 as vb
    from oracle import vaegan_oracle as vo
    VAEGANStep = import_module("vaegan_b200.step").VAEGANStep
    # fp32 mode: run-to-run reproducible to ~1e-6 (the bf16 mode's 1-ulp avalanche, DESIGN.md section 6, would hide a
    # missing or doubled bucket behind its own 5e-2 noise); the bucketing / stream logic under test is the same
    hw, nz, per, prec = 64, 128, 16, "fp32"
    groups = [dist.new_group([r]) for r in range(world)]          # every rank creates every group (collective call)
    solo_group = groups[rank]

    def nets(seed):
        torch.manual_seed(seed)
        e = vb.Encoder([3, hw, hw], nz, precision=prec)
        g = vb.Generator(nz=nz, hw=hw, precision=prec)
        d = vb.Discriminator(hw=hw, precision=prec)
        g.apply(vb.weights_init)
        d.apply(vb.weights_init)
        return [m.to(dev) for m in (e, g, d)]

    real, eps, n_real, n_fake = (t.to(dev) for t in vo.make_inputs(per, hw, nz, seed=100 + rank))   # this rank's shard
    ok = True
    report = []
    # transports: the peer-memory kernel (multimem instructions when the fabric has a multicast mapping), the same
    # kernel restricted to peer loads / stores, and NCCL all-reduce + replicated Adam - all must give the same step
    final_params = {}
    for transport in ("peer", "peer-nomc", "nccl"):
      for graph in (False, True):
        dp_step = VAEGANStep(*nets(7), use_cuda_graph=graph, capture_grads=True, dp_transport=transport)
        assert dp_step.world == world and len(dp_step.buckets["G"].buckets) >= 2 and len(dp_step.buckets["D"].buckets) >= 2
        if transport != "nccl":
            assert dp_step.peer is not None and len(dp_step.peer_adam) == 3
            if rank == 0 and not graph:
                print(f"transport {transport}: multicast = {dp_step.peer_adam['G'].multicast}", flush=True)
        solo = VAEGANStep(*nets(7), use_cuda_graph=graph, capture_grads=True, process_group=solo_group)
        assert solo.world == 1
        l_dp = dp_step.step(real, 50, eps, n_real, n_fake)
        l_solo = solo.step(real, 50, eps, n_real, n_fake)
        torch.cuda.synchronize()
        if dp_step.peer is not None:
            dp_step.peer.check()
        # d_loss_0 / recon / kl are computed before any all-reduced update: identical to the solo replica's
        for k in ("d_loss_0", "recon", "kl"):
            a, b = float(l_dp[k]), float(l_solo[k])
            if not abs(a - b) <= 1e-3 * abs(b) + 1e-6:
                ok = False
                print(f"rank {rank} {transport} graph {graph}: loss {k} dp {a} solo {b}", flush=True)
        g_dp, g_solo = dp_step.gradients(), solo.gradients()
        # first discriminator update: DP gradient (summed over ranks) == sum of the solo replicas' gradients
        mine = torch.cat([v.flatten() for v in g_solo["D"][0].values()])
        total = mine.clone()
        dist.all_reduce(total)
        got = torch.cat([v.flatten() for v in g_dp["D"][0].values()])
        cos = float(torch.dot(got.double(), total.double()) / (got.double().norm() * total.double().norm()))
        rel = float((got - total).norm() / total.norm())
        report.append((transport, graph, cos, rel))
        if not (cos > 0.999999 and rel < 1e-3):
            ok = False
            print(f"rank {rank} {transport} graph {graph}: D gradient cos {cos} rel {rel}", flush=True)
        # replicas stay in lock-step: parameters identical on every rank after the step
        for opt in (dp_step.opt_E, dp_step.opt_G, dp_step.opt_D):
            ref = opt.params.clone()
            dist.broadcast(ref, 0)
            if not torch.equal(ref, opt.params):
                ok = False
                print(f"rank {rank} {transport} graph {graph}: parameters differ from rank 0 "
                      f"(max {float((ref - opt.params).abs().max())})", flush=True)
        # ... and every transport takes the same step: Adam's first update is ~lr * sign(g), so all but the elements
        # whose summed gradient is ~0 (its sign is summation-order noise) must agree far inside one lr
        flat = torch.cat([o.params.clone() for o in (dp_step.opt_E, dp_step.opt_G, dp_step.opt_D)])
        prev = final_params.setdefault(graph, (transport, flat))
        if prev[0] != transport:
            bad = int(((flat - prev[1]).abs() > 1e-4).sum())
            if bad > 0.002 * flat.numel():
                ok = False
                print(f"rank {rank} graph {graph}: {bad}/{flat.numel()} parameters differ between transports "
                      f"{prev[0]} and {transport}", flush=True)
        # sharded optimizer state: after gather_moments every rank holds every slice
        if transport != "nccl" and not graph:
            sd = dp_step.state_dict()
            mom = torch.cat([t.flatten() for t in sd["opt_G"]["exp_avg"]])
            ref = mom.clone()
            dist.broadcast(ref, 0)
            if not torch.equal(ref, mom) or float(mom.abs().sum()) == 0.0:
                ok = False
                print(f"rank {rank} {transport}: gathered Adam moments differ across ranks", flush=True)
        del dp_step, solo
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("DP_CHECK", "OK" if int(flag) == 1 else "FAIL", report, flush=True)
    sys.stdout.flush()
    os._exit(0 if int(flag) == 1 else 1)


if __name__ == "__main__":
    main()
