"""Kernel timeline of the CUDA-graph-replayed step via torch.profiler (CUPTI): per-kernel totals + GPU busy fraction."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from importlib import import_module
import vaegan_b200 as vb
VAEGANStep = import_module("vaegan_b200.step").VAEGANStep
HW, NZ, B = int(os.environ.get("HW", "64")), int(os.environ.get("NZ", "128")), int(os.environ.get("BATCH", "256"))
torch.manual_seed(42)
WIDTH = int(os.environ.get("WIDTH", "1"))
enc = vb.Encoder([3, HW, HW], NZ, width=WIDTH); gen = vb.Generator(nz=NZ, ngf=64 * WIDTH, hw=HW)
dis = vb.Discriminator(ndf=64 * WIDTH, hw=HW)
gen.apply(vb.weights_init); dis.apply(vb.weights_init)
for m in (enc, gen, dis): m.cuda()
step = VAEGANStep(enc, gen, dis, use_cuda_graph=True)
real = (torch.rand(B, 3, HW, HW) * 2 - 1).cuda()
for _ in range(5): step.step(real, 50)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): step.step(real, 50)
e1.record(); torch.cuda.synchronize()
print(f"graph replay: {e0.elapsed_time(e1)/10:.3f} ms/step, launches/step {step.launches_per_step}")
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3): step.step(real, 50)
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
tot = {}
for e in evs:
    k = e.name[:90]
    t = tot.setdefault(k, [0.0, 0]); t[0] += e.time_range.elapsed_us(); t[1] += 1
span = evs[-1].time_range.end - evs[0].time_range.start
busy = sum(v[0] for v in tot.values())
print(f"3 steps: span {span/1e3:.2f} ms, kernel-busy {busy/1e3:.2f} ms ({100*busy/span:.1f}%), {len(evs)} device events")
for k, (us, n) in sorted(tot.items(), key=lambda kv: -kv[1][0])[:28]:
    print(f"{us/3:9.1f} us/step  n/step {n/3:6.1f}  avg {us/n:8.1f}  {k}")
# gaps
gaps = [(evs[i+1].time_range.start - evs[i].time_range.end) for i in range(len(evs)-1)]
big = sorted(gaps, reverse=True)[:8]
print("largest gaps (us):", [round(g,1) for g in big], " sum of gaps >2us:", round(sum(g for g in gaps if g > 2)/3,1), "us/step")
# per-launch listing of the igemm kernels in the last profiled step (launch order)
names = [e for e in evs if 'igemm' in e.name or 'wgrad_reduce' in e.name]
per = len(names) // 3
print("igemm launches of one step (order, us):")
print(" ".join(f"{'W' if 'wgrad_k' in e.name else ('R' if 'reduce' in e.name else 'F')}{e.time_range.elapsed_us():.0f}" for e in names[-per:]))

for pat, tag in (("channel_reduce_kernel<__nv_bfloat16, 1>", "BN-bwd reduce"), ("bn_act_bwd_apply", "BN-bwd apply"),
                 ("channel_reduce_kernel<__nv_bfloat16, 0>", "BN stats"), ("scale_shift_act_vec", "BN apply")):
    sel = [e for e in evs if pat in e.name]
    per = len(sel) // 3
    print(tag, "per-launch us:", " ".join(f"{e.time_range.elapsed_us():.0f}" for e in sel[-per:]))

# full timeline of the last profiled step: offset, duration, stream, short kernel name
if os.environ.get("TIMELINE"):
    per = len(evs) // 3
    last = evs[-per:]
    t0 = last[0].time_range.start
    def short(n):
        n = n.replace("void ", "").replace("vg::(anonymous namespace)::", "").replace("vg::", "")
        return n.split("(")[0][:48]
    with open(os.environ["TIMELINE"], "w") as f:
        for e in last:
            stream = getattr(e, "stream", None)
            f.write(f"{e.time_range.start - t0:9.1f} {e.time_range.elapsed_us():8.1f} {short(e.name)}\n")
    print("timeline written:", os.environ["TIMELINE"], per, "events")
