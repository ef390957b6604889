"""Micro-benchmark of the HBM-bound kernels on every BatchNorm tensor shape of the cfg-2 step (B = 256, bf16):
BN statistics, BN apply, BN backward (reduce + apply), plus the weight re-pack.  Prints us and GB/s of algorithmic
traffic, cold (rotating buffers > L2) and hot (one buffer)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from importlib import import_module
import vaegan_b200 as vb

fn = import_module("vaegan_b200.functional")
dev = torch.device("cuda")
B = 256
SHAPES = [("E 64@32^2", B * 32 * 32, 64), ("E 128@16^2", B * 256, 128), ("E 256@8^2", B * 64, 256),
          ("E 512@4^2", B * 16, 512), ("G 1024@4^2", B * 16, 1024), ("G 512@8^2", B * 64, 512),
          ("G 256@16^2", B * 256, 256), ("G 128@32^2", B * 1024, 128), ("G 64@64^2", B * 4096, 64),
          ("D 128@16^2", B * 256, 128), ("D 256@8^2", B * 64, 256), ("D 512@4^2", B * 16, 512)]


def timeit(f, n, iters=20):
    for i in range(3):
        f(i % n)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(iters):
        f(i % n)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3


tot = {}
for name, rows, C in SHAPES:
    nbytes = rows * C * 2
    nbuf = max(2, int(400e6 // nbytes))
    xs = [torch.randn(rows, C, device=dev).bfloat16() for _ in range(min(nbuf, 12))]
    dys = [torch.randn(rows, C, device=dev).bfloat16() for _ in range(len(xs))]
    outs = [torch.empty_like(x) for x in xs[:2]]
    g, b = torch.rand(C, device=dev) + 0.5, torch.randn(C, device=dev)
    dg, db = torch.zeros(C, device=dev), torch.zeros(C, device=dev)
    stats = fn.bn_train_fwd(xs[0], g, b, None, None, None, 0.1, 1e-5)
    for mode, n in (("cold", len(xs)), ("hot", 1)):
        t_st = timeit(lambda i: fn.bn_train_fwd(xs[i], g, b, None, None, None, 0.1, 1e-5), n)
        t_ap = timeit(lambda i: fn.scale_shift_act(xs[i], stats[2], stats[3], 1, 0.0, out=outs[i % 2]), n)
        t_bw = timeit(lambda i: fn.bn_act_bwd(dys[i], xs[i], stats, 1, 0.0, dg, db, out=outs[i % 2]), n)
        print(f"{name:12s} {mode:4s} {nbytes / 1e6:7.1f} MB | stats {t_st:7.1f} us {nbytes / t_st / 1e3:7.0f} GB/s | "
              f"apply {t_ap:7.1f} us {2 * nbytes / t_ap / 1e3:7.0f} GB/s | bwd(reduce+apply) {t_bw:7.1f} us "
              f"{5 * nbytes / t_bw / 1e3:7.0f} GB/s", flush=True)
        if mode == "cold":
            tot[name] = (t_st, t_ap, t_bw)
    del xs, dys, outs

# step totals: E x1, G x1, D: 2 iterations x 2 groups + generator step x1 = 5
w = {"E": 1, "G": 1, "D": 5}
s = [sum(v[k] * w[n[0]] for n, v in tot.items()) for k in range(3)]
print(f"per-step estimate (cold): stats {s[0]:.0f} us, apply {s[1]:.0f} us, bwd {s[2]:.0f} us")

# weight re-pack of the generator (13.2 M parameters) through the one-launch path
gen = vb.Generator(nz=128, hw=64, precision="bf16").to(dev)
layers = gen._layers()
nparam = sum(l.conv.weight.numel() for l in layers)
t = timeit(lambda i: fn.pack_layers(layers, torch.bfloat16), 1)
print(f"pack {len(layers)} generator layers, {nparam / 1e6:.2f} M params: {t:.1f} us, {nparam * 8 / t / 1e3:.0f} GB/s")
