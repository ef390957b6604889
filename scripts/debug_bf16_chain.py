"""Debug: bf16 generator backward chain vs emulated oracle: for every layer compare dy (grad of activated output),
d_raw (grad of conv output) and the weight gradient."""
import copy, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.util import make_pair, rel_err, cosine
from oracle import vaegan_oracle as vo
from importlib import import_module
import vaegan_b200
F_ = import_module("vaegan_b200.functional")
hw, nz, B = 64, 128, 8
o_nets, nets = make_pair(hw, nz, "bf16")
ref = copy.deepcopy(o_nets[1]); vo.attach_bf16_emulation(ref)
mine = nets[1]
z = torch.randn(B, nz, 1, 1, generator=torch.Generator().manual_seed(11))
up = torch.randn(B, 3, hw, hw, generator=torch.Generator().manual_seed(12))
gout_ref = {}
def fwd_hook(i):
    def h(mod, a, out):
        out.register_hook(lambda g, i=i: gout_ref.__setitem__(i, g.detach().clone()))
    return h
for i, m in enumerate(ref.main):
    m.register_forward_hook(fwd_hook(i))       # registered after the emulation hooks -> sees the rounded output
(ref(z.clone()) * up).sum().backward()
calls = []
orig = F_.bn_act_bwd
def spy(dy, x, stats, act, slope, dgamma, dbeta):
    dx = orig(dy, x, stats, act, slope, dgamma, dbeta)
    calls.append((dy.detach().clone(), dx.detach().clone()))
    return dx
F_.bn_act_bwd = spy
(mine(z.cuda()) * up.cuda()).sum().backward()
nchw = lambda t: t.detach().float().permute(0, 3, 1, 2).cpu()
calls = calls[::-1]     # forward order: layer 0..4
for li, (dy, dx) in enumerate(calls):
    conv_idx, relu_idx = 3 * li, 3 * li + 2
    r_dy, r_draw = gout_ref[relu_idx], gout_ref[conv_idx]
    print(f"layer {li}: dy cos {cosine(nchw(dy), r_dy):.6f} rel {rel_err(nchw(dy), r_dy):.2e} | d_raw cos {cosine(nchw(dx), r_draw):.6f} "
          f"rel {rel_err(nchw(dx), r_draw):.2e} | ref: is d_raw bf16-valued? {bool(torch.equal(r_draw, r_draw.bfloat16().float()))} "
          f"is dy bf16-valued? {bool(torch.equal(r_dy, r_dy.bfloat16().float()))} |d_raw|/|dy| {float(r_draw.norm()/r_dy.norm()):.3f}")
