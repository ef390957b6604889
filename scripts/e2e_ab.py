"""A/B of the end-to-end loop: direct host copy on the compute stream vs VAEGANStep.prefetch() on a copy stream."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from importlib import import_module
import vaegan_b200 as vb
VAEGANStep = import_module("vaegan_b200.step").VAEGANStep
torch.manual_seed(42)
enc = vb.Encoder([3, 64, 64], 128); gen = vb.Generator(nz=128, hw=64); dis = vb.Discriminator(hw=64)
gen.apply(vb.weights_init); dis.apply(vb.weights_init)
for m in (enc, gen, dis): m.cuda()
step = VAEGANStep(enc, gen, dis, use_cuda_graph=True)
B = 256
host = [(torch.rand(B, 3, 64, 64) * 2 - 1).pin_memory() for _ in range(4)]
devb = [h.cuda() for h in host]
for i in range(5): step.step(devb[i % 4], 50)
torch.cuda.synchronize()
def loop(mode, n=40):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    if mode == "prefetch": step.prefetch(host[0])
    for i in range(n):
        if mode == "resident": l = step.step(devb[i % 4], 50)
        else: l = step.step(host[i % 4], 50)
        if mode == "prefetch": step.prefetch(host[(i + 1) % 4])
        if mode == "prefetch_late":
            pass
        _ = float(l["total"])
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3
for mode in ("resident", "direct", "prefetch", "resident", "direct", "prefetch"):
    print(mode, round(loop(mode), 3), "ms/step")
# H2D alone, idle vs under load
stage = torch.empty_like(devb[0]); cs = torch.cuda.Stream()
def h2d_time(under_load):
    torch.cuda.synchronize()
    if under_load:
        for i in range(3): step.step(devb[i % 4], 50)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(cs):
        a.record(cs); stage.copy_(host[0], non_blocking=True); b.record(cs)
    torch.cuda.synchronize()
    return a.elapsed_time(b)
print("H2D 12.6 MB idle", [round(h2d_time(False), 3) for _ in range(3)], "ms; under load", [round(h2d_time(True), 3) for _ in range(3)])
