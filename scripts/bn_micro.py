"""Micro-run of the BN statistics kernel on a D-sized tensor (for ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from importlib import import_module
import vaegan_b200
fn = import_module("vaegan_b200.functional")
rows, C = 256 * 8 * 8, 256
x = torch.randn(1, 1, rows, C, device="cuda").bfloat16()
g, b = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
for _ in range(5):
    st = fn.bn_train_fwd(x, g, b, None, None, None, 0.1, 1e-5)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    st = fn.bn_train_fwd(x, g, b, None, None, None, 0.1, 1e-5)
e1.record(); torch.cuda.synchronize()
print("bn_train_fwd rows", rows, "C", C, ":", e0.elapsed_time(e1) / 20 * 1e3, "us")
