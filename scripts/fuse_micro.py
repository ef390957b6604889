"""Isolated cost of the fused convolution epilogues on the cfg-2 layer shapes (bf16): forward with / without
VG_EPI_BN_STATS, dgrad with / without VG_EPI_BN_BWD.  Kernels are timed back to back with CUDA events."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from importlib import import_module
import vaegan_b200  # noqa: F401

fn = import_module("vaegan_b200.functional")
dev = torch.device("cuda")
dt = torch.bfloat16

# name, kind, B, in_hw, small_c, big_c, k, s, p, groups   (ConvSpec(kind, small_c, big_c, ...))
LAYERS = [
    ("E 32->64  31->14", "down", 256, 31, 64, 32, 4, 2, 0, 1),
    ("E 64->128 14->6", "down", 256, 14, 128, 64, 4, 2, 0, 1),
    ("E 128->256 6->2", "down", 256, 6, 256, 128, 4, 2, 0, 1),
    ("G nz->1024 1->4", "up", 256, 1, 128, 1024, 4, 1, 0, 1),
    ("G 1024->512 4->8", "up", 256, 4, 1024, 512, 4, 2, 1, 1),
    ("G 512->256 8->16", "up", 256, 8, 512, 256, 4, 2, 1, 1),
    ("G 256->128 16->32", "up", 256, 16, 256, 128, 4, 2, 1, 1),
    ("G 128->64 32->64", "up", 256, 32, 128, 64, 4, 2, 1, 1),
    ("D 64->128 32->16 pair", "down", 512, 32, 128, 64, 4, 2, 1, 2),
    ("D 128->256 16->8 pair", "down", 512, 16, 256, 128, 4, 2, 1, 2),
    ("D 256->512 8->4 pair", "down", 512, 8, 512, 256, 4, 2, 1, 2),
]


def timeit(f, iters=20):
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        f()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3


for name, kind, B, hw, sc, bc, k, s, p, groups in LAYERS:
    spec = fn.ConvSpec(kind, sc, bc, k, s, p)
    g = spec.geom(B, hw, hw)
    w = torch.randn(sc, bc, k, k, device=dev) * 0.05
    wd, wu = fn.pack_weights(w, g)
    small = torch.randn(B, g.small_h, g.small_w, sc, device=dev).to(dt)
    big = torch.randn(B, g.big_h, g.big_w, bc, device=dev).to(dt)
    # ---- forward + statistics of the output
    C_out = sc if kind == "down" else bc
    sums = torch.zeros(groups * 2 * C_out, device=dev)
    ep1 = fn.make_epilogue(fn.EPI_BN_STATS, groups, C_out, sums=sums)
    fwd = (lambda ep=None: fn.conv_down(big, wd, g, ep=ep)) if kind == "down" else (lambda ep=None: fn.conv_up(small, wu, g, ep=ep))
    ok1 = fn.epilogue_supported(g, kind == "up", ep1)
    t0 = timeit(fwd)
    t1 = timeit(lambda: fwd(ep1)) if ok1 else float("nan")
    # ---- dgrad + BatchNorm-backward reduction of the producer (whose raw output has the dgrad's output shape)
    C_in = bc if kind == "down" else sc
    raw = big if kind == "down" else small
    stats = torch.rand(groups, 4, C_in, device=dev)
    sums2 = torch.zeros(groups * 2 * C_in, device=dev)
    ep2 = fn.make_epilogue(fn.EPI_BN_BWD, groups, C_in, 2, 0.2, sums2, raw, stats)
    ep3 = fn.make_epilogue(fn.EPI_ACT_BWD, 1, C_in, 2, 0.2, None, raw, None)
    bwd = (lambda ep=None: fn.conv_up(small, wu, g, ep=ep)) if kind == "down" else (lambda ep=None: fn.conv_down(big, wd, g, ep=ep))
    ok2 = fn.epilogue_supported(g, kind == "down", ep2)
    t2 = timeit(bwd)
    t3 = timeit(lambda: bwd(ep2)) if ok2 else float("nan")
    t4 = timeit(lambda: bwd(ep3)) if ok2 else float("nan")
    print(f"{name:24s} fwd {t0:7.1f} -> +stats {t1:7.1f} us | dgrad {t2:7.1f} -> +bn_bwd {t3:7.1f}  +act_bwd {t4:7.1f} us",
          flush=True)
