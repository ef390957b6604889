"""A small bf16 step (batch 8, eager + graph) for compute-sanitizer:
    compute-sanitizer --tool memcheck python scripts/sanitize_step.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from importlib import import_module
import vaegan_b200 as vb

VAEGANStep = import_module("vaegan_b200.step").VAEGANStep
torch.manual_seed(0)
enc, gen, dis = vb.Encoder([3, 64, 64], 128), vb.Generator(nz=128, hw=64), vb.Discriminator(hw=64)
gen.apply(vb.weights_init)
dis.apply(vb.weights_init)
for m in (enc, gen, dis):
    m.cuda()
step = VAEGANStep(enc, gen, dis, use_cuda_graph=os.environ.get("GRAPH", "0") == "1")
real = (torch.rand(8, 3, 64, 64) * 2 - 1).cuda()
for _ in range(2):
    losses = step.step(real, 50)
torch.cuda.synchronize()
print("ok", {k: round(float(v), 4) for k, v in losses.items()})
gen.eval()
with torch.no_grad():
    img = gen(torch.randn(4, 128, 1, 1, device="cuda"))
torch.cuda.synchronize()
print("gen ok", tuple(img.shape))
