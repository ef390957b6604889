"""Kernel timeline of the data-parallel graph-replayed step on rank 0 (torch.profiler): NCCL kernels, what they
overlap with, per-stream busy time.  Launch under torchrun."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from torch.profiler import profile, ProfilerActivity
from importlib import import_module
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
os.environ.setdefault("TORCH_NCCL_ASYNC_ERROR_HANDLING", "0")
dist.init_process_group("nccl", device_id=dev)
import vaegan_b200 as vb
VAEGANStep = import_module("vaegan_b200.step").VAEGANStep
torch.manual_seed(42)
enc = vb.Encoder([3, 64, 64], 128); gen = vb.Generator(nz=128, hw=64); dis = vb.Discriminator(hw=64)
gen.apply(vb.weights_init); dis.apply(vb.weights_init)
for m in (enc, gen, dis): m.to(dev)
step = VAEGANStep(enc, gen, dis, use_cuda_graph=True, seed=1234 + rank)
real = (torch.rand(256, 3, 64, 64) * 2 - 1).to(dev)
for _ in range(5): step.step(real, 50)
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): step.step(real, 50)
e1.record(); torch.cuda.synchronize()
if rank == 0: print(f"graph replay: {e0.elapsed_time(e1)/10:.3f} ms/step at world {world}; buckets: " +
                    ", ".join(f"{k}:{[(b.hi - b.lo) * 4 >> 10 for b in v.buckets]} KB" for k, v in step.buckets.items()), flush=True)
dist.barrier()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3): step.step(real, 50)
    torch.cuda.synchronize()
if rank == 0:
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    per = len(evs) // 3
    last = evs[-per:]
    t0 = last[0].time_range.start
    span = last[-1].time_range.end - t0
    print(f"last step: span {span/1e3:.3f} ms, {per} events")
    def short(n):
        return n.replace("void ", "").replace("vg::(anonymous namespace)::", "").replace("vg::", "").split("(")[0][:40]
    nccl = [e for e in last if "nccl" in e.name.lower()]
    print("NCCL kernels (start us, dur us):", [(round(e.time_range.start - t0, 1), round(e.time_range.elapsed_us(), 1)) for e in nccl])
    print("NCCL total us:", round(sum(e.time_range.elapsed_us() for e in nccl), 1))
    # what runs while each NCCL kernel runs
    for e in nccl:
        ov = [short(o.name) + f":{o.time_range.elapsed_us():.0f}" for o in last if o is not e and o.time_range.start < e.time_range.end and o.time_range.end > e.time_range.start]
        print(f"  nccl @{e.time_range.start - t0:.0f} +{e.time_range.elapsed_us():.0f}us overlaps: {ov[:12]}")
    path = os.environ.get("TIMELINE")
    if path:
        with open(path, "w") as f:
            for e in last:
                f.write(f"{e.time_range.start - t0:9.1f} {e.time_range.elapsed_us():8.1f} {short(e.name)}\n")
step._graph = None; del step
import gc; gc.collect(); torch.cuda.synchronize()
dist.destroy_process_group()
