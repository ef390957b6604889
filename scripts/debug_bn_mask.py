"""Debug: are ReLU masks of (my BN affine) vs (torch BN) different on the generator's activations?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.util import make_pair
from importlib import import_module
import vaegan_b200
F_ = import_module("vaegan_b200.functional")
hw, nz, B = 64, 128, 8
o_nets, nets = make_pair(hw, nz, "fp32")
mine = nets[1]
z = torch.randn(B, nz, 1, 1, generator=torch.Generator().manual_seed(11))
h = F_.nchw_to_nhwc(z.cuda(), torch.float32)
for li, layer in enumerate(mine._layers()[:-1]):
    g = layer.spec.geom(h.shape[0], h.shape[1], h.shape[2])
    raw = F_.conv_up(h, layer.conv.weight.detach().contiguous(), g)
    bn = layer.bn
    stats = F_.bn_train_fwd(raw, bn.weight.detach(), bn.bias.detach(), None, None, None, 0.1, 1e-5)
    y = F_.scale_shift_act(raw, stats[2], stats[3], 1, 0.0)
    C = raw.shape[-1]
    x = raw.reshape(-1, C).cpu()
    zt = torch.nn.functional.batch_norm(x, None, None, bn.weight.detach().cpu(), bn.bias.detach().cpu(), True, 0.1, 1e-5)
    z64 = torch.nn.functional.batch_norm(x.double(), None, None, bn.weight.detach().cpu().double(), bn.bias.detach().cpu().double(), True, 0.1, 1e-5)
    zm = (x * stats[2].cpu() + stats[3].cpu())
    m_mine, m_t, m64 = (y.reshape(-1, C).cpu() > 0), (zt > 0), (z64 > 0)
    print(f"layer {li}: C={C} rows={x.shape[0]}  mask(mine)!=mask(torch32): {(m_mine != m_t).sum().item()}  "
          f"mask(mine)!=mask(f64): {(m_mine != m64).sum().item()}  mask(torch32)!=mask(f64): {(m_t != m64).sum().item()}  "
          f"min|z| f64 {z64.abs().min().item():.3e}  #|z|<1e-6: {(z64.abs() < 1e-6).sum().item()} exact-equal-x pairs? uniq frac {x.unique().numel()/x.numel():.4f}")
    h = y
