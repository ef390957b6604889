"""Bisect the stream-capture failure: fused step with / without wgrad overlap."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from importlib import import_module
import vaegan_b200 as vb
VAEGANStep = import_module("vaegan_b200.step").VAEGANStep
HW, NZ, B = 64, 128, 32
for overlap in (False, True):
    torch.manual_seed(1)
    enc = vb.Encoder([3, HW, HW], NZ); gen = vb.Generator(nz=NZ, hw=HW); dis = vb.Discriminator(hw=HW)
    for m in (enc, gen, dis): m.cuda()
    step = VAEGANStep(enc, gen, dis, use_cuda_graph=True, overlap_wgrad=overlap)
    real = (torch.rand(B, 3, HW, HW) * 2 - 1).cuda()
    try:
        for _ in range(3): step.step(real, 50)
        torch.cuda.synchronize()
        print("overlap", overlap, "OK")
    except Exception as e:
        print("overlap", overlap, "FAILED:", str(e).splitlines()[0])
        break
