"""Micro-benchmark of the from-sums BatchNorm passes at the cfg-2 shapes (bf16): us and GB/s per launch, CUDA events
over back-to-back launches (small tensors stay L2-resident, as after the convolution that produced them)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from importlib import import_module
import vaegan_b200  # noqa: F401

fn = import_module("vaegan_b200.functional")
SHAPES = [("G 4x4x1024", 256 * 16, 1024, 1), ("G 8x8x512", 256 * 64, 512, 1), ("G 16x16x256", 256 * 256, 256, 1),
          ("G 32x32x128", 256 * 1024, 128, 1), ("G 64x64x64", 256 * 4096, 64, 1), ("D 16x16x128 x2", 512 * 256, 128, 2),
          ("D 8x8x256 x2", 512 * 64, 256, 2), ("D 4x4x512 x2", 512 * 16, 512, 2), ("E 31x31x32", 256 * 961, 32, 1),
          ("E 14x14x64", 256 * 196, 64, 1)]


def timeit(f, iters=30):
    for _ in range(5):
        f()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters):
            f()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) / iters)
    return best * 1e3


for name, rows, C, groups in SHAPES:
    x = torch.randn(rows, C, device="cuda").bfloat16().view(1, 1, rows, C)
    dz = torch.randn(rows, C, device="cuda").bfloat16().view(1, 1, rows, C)
    g, b = torch.rand(C, device="cuda") + 0.5, torch.randn(C, device="cuda")
    xf = x.float().view(groups, -1, C)
    sums = torch.cat([torch.stack([xf[i].sum(0), (xf[i] ** 2).sum(0)]).flatten() for i in range(groups)]).contiguous()
    rm, rv, nbt = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda"), torch.zeros((), dtype=torch.int64, device="cuda")
    y, stats = fn.bn_apply_from_sums(x, sums, groups, g, b, rm, rv, nbt, 0.1, 1e-5, 1, 0.0)
    bsums = torch.randn(groups * 2 * C, device="cuda")
    dg, db = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    t_f = timeit(lambda: fn.bn_apply_from_sums(x, sums, groups, g, b, rm, rv, nbt, 0.1, 1e-5, 1, 0.0))
    t_b = timeit(lambda: fn.bn_bwd_apply_from_sums(dz, x, stats, bsums, groups, dg, db))
    mb = rows * C * 2 / 1e6
    print(f"{name:18s} {mb:7.1f} MB   apply {t_f:6.1f} us {2 * mb / t_f:5.2f} TB/s   bwd apply {t_b:6.1f} us "
          f"{3 * mb / t_b:5.2f} TB/s")
