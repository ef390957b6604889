"""One cfg-2 dgrad with a fused epilogue, a few launches (target for `ncu -k regex:igemm_fprop`).
usage: fuse_one.py <layer index in fuse_micro.LAYERS> <mode: 0 plain, 2 bn_bwd, 3 act_bwd, 1 fwd+stats>"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from importlib import import_module
import vaegan_b200  # noqa: F401

fn = import_module("vaegan_b200.functional")
LAYERS = [
    ("G 256->128 16->32", "up", 256, 16, 256, 128, 4, 2, 1, 1),
    ("G 128->64 32->64", "up", 256, 32, 128, 64, 4, 2, 1, 1),
    ("D 128->256 16->8 pair", "down", 512, 16, 256, 128, 4, 2, 1, 2),
]
idx, mode = int(sys.argv[1]), int(sys.argv[2])
name, kind, B, hw, sc, bc, k, s, p, groups = LAYERS[idx]
dev, dt = torch.device("cuda"), torch.bfloat16
spec = fn.ConvSpec(kind, sc, bc, k, s, p)
g = spec.geom(B, hw, hw)
w = torch.randn(sc, bc, k, k, device=dev) * 0.05
wd, wu = fn.pack_weights(w, g)
small = torch.randn(B, g.small_h, g.small_w, sc, device=dev).to(dt)
big = torch.randn(B, g.big_h, g.big_w, bc, device=dev).to(dt)
C_out, C_in = (sc, bc) if kind == "down" else (bc, sc)
raw = big if kind == "down" else small
stats = torch.rand(groups, 4, C_in, device=dev)
sums = torch.zeros(groups * 2 * max(C_in, C_out), device=dev)
ep = {0: None, 1: fn.make_epilogue(fn.EPI_BN_STATS, groups, C_out, sums=sums),
      2: fn.make_epilogue(fn.EPI_BN_BWD, groups, C_in, 2, 0.2, sums, raw, stats),
      3: fn.make_epilogue(fn.EPI_ACT_BWD, 1, C_in, 2, 0.2, None, raw, None)}[mode]
if mode == 1:
    f = (lambda: fn.conv_down(big, wd, g, ep=ep)) if kind == "down" else (lambda: fn.conv_up(small, wu, g, ep=ep))
else:
    f = (lambda: fn.conv_up(small, wu, g, ep=ep)) if kind == "down" else (lambda: fn.conv_down(big, wd, g, ep=ep))
for _ in range(8):
    f()
torch.cuda.synchronize()
print("done", name, mode)
