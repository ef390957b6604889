"""Debug: layer-by-layer forward AND backward comparison of the generator against the oracle (test infrastructure).
usage: python scripts/debug_bf16_g.py [bf16|fp32]"""
import copy, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.util import make_pair, rel_err, cosine
from oracle import vaegan_oracle as vo
from importlib import import_module
import vaegan_b200
F_ = import_module("vaegan_b200.functional")

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
dt = torch.bfloat16 if prec == "bf16" else torch.float32
hw, nz, B = 64, 128, 8
o_nets, nets = make_pair(hw, nz, prec)
ref = copy.deepcopy(o_nets[1])
if prec == "bf16":
    vo.attach_bf16_emulation(ref)
mine = nets[1]
z = torch.randn(B, nz, 1, 1, generator=torch.Generator().manual_seed(11))
up = torch.randn(B, 3, hw, hw, generator=torch.Generator().manual_seed(12))

acts_ref, gout_ref = {}, {}
def fwd_hook(i):
    def h(mod, a, out):
        acts_ref[i] = out.detach().clone()
        out.register_hook(lambda g, i=i: gout_ref.__setitem__(i, g.detach().clone()))
    return h
for i, m in enumerate(ref.main):
    m.register_forward_hook(fwd_hook(i))
zr = z.clone().requires_grad_(True)
out_ref = ref(zr)
(out_ref * up).sum().backward()

layers = mine._layers()
h = F_.ToNHWCFn.apply(z.cuda().requires_grad_(True), dt)
outs, gouts = [], {}
idx = 0
mods = []
for li, layer in enumerate(layers):
    last = li == len(layers) - 1
    h = layer(h, True, fuse_act=not last)
    h.register_hook(lambda g, li=li: gouts.__setitem__(li, g.detach().clone()))
    outs.append(h)
    idx += 3 if not last else 1
    mods.append(idx - 1)
out = F_.ToNCHWActFn.apply(h, 3)
(out * up.cuda()).sum().backward()
nchw = lambda t: t.detach().float().permute(0, 3, 1, 2).cpu()
print(f"precision {prec}")
for li, m in enumerate(mods):
    a, g = nchw(outs[li]), nchw(gouts[li])
    ra, rg = acts_ref[m], gout_ref[m]
    print(f"layer {li} (ref main[{m}] output): fwd rel {rel_err(a, ra):.3e} cos {cosine(a, ra):.6f} | "
          f"grad-of-output rel {rel_err(g, rg):.3e} cos {cosine(g, rg):.6f} norm ratio {float(g.norm()/rg.norm()):.5f}")
pm = dict(mine.named_parameters())
for k, p in ref.named_parameters():
    k2 = k.replace("parametrizations.weight.original", "weight")
    g, r = pm[k2].grad, p.grad
    print(f"param {k2}: rel {rel_err(g, r):.3e} cos {cosine(g, r):.6f} norm ratio {float(g.norm().cpu()/r.norm()):.5f}")
