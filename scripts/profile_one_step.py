"""One eagerly launched cfg-2 step (B = 256, bf16) between cudaProfilerStart / Stop, for
    ncu --profile-from-start off ... python scripts/profile_one_step.py
(the launch list and the `--set full` captures under profiles/ come from this command).  SERIAL=1 puts the weight
gradients on the main stream too (the order ncu serialises them in anyway)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from importlib import import_module
import vaegan_b200 as vb

VAEGANStep = import_module("vaegan_b200.step").VAEGANStep
HW, NZ, B = int(os.environ.get("HW", "64")), int(os.environ.get("NZ", "128")), int(os.environ.get("BATCH", "256"))
WIDTH = int(os.environ.get("WIDTH", "1"))
torch.manual_seed(42)
enc = vb.Encoder([3, HW, HW], NZ, width=WIDTH)
gen = vb.Generator(nz=NZ, ngf=64 * WIDTH, hw=HW)
dis = vb.Discriminator(ndf=64 * WIDTH, hw=HW)
gen.apply(vb.weights_init)
dis.apply(vb.weights_init)
for m in (enc, gen, dis):
    m.cuda()
step = VAEGANStep(enc, gen, dis, use_cuda_graph=False, overlap_wgrad=os.environ.get("SERIAL", "0") != "1")
real = (torch.rand(B, 3, HW, HW) * 2 - 1).cuda()
for _ in range(3):
    step.step(real, 50)
torch.cuda.synchronize()
torch.cuda.profiler.start()
losses = step.step(real, 50)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("launches/step", step.launches_per_step, "total loss", float(losses["total"]))
