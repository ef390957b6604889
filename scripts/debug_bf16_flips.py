"""Debug: count ReLU-mask flips between my bf16 generator forward and the emulated oracle's, per layer."""
import copy, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.util import make_pair
from oracle import vaegan_oracle as vo
from importlib import import_module
import vaegan_b200
F_ = import_module("vaegan_b200.functional")
hw, nz, B = 64, 128, 8
o_nets, nets = make_pair(hw, nz, "bf16")
ref = copy.deepcopy(o_nets[1]); vo.attach_bf16_emulation(ref)
mine = nets[1]
z = torch.randn(B, nz, 1, 1, generator=torch.Generator().manual_seed(11))
acts = {}
for i, m in enumerate(ref.main):
    m.register_forward_hook(lambda mod, a, out, i=i: acts.__setitem__(i, out.detach().clone()))
ref(z.clone())
h = F_.nchw_to_nhwc(z.cuda(), torch.bfloat16)
nchw = lambda t: t.detach().float().permute(0, 3, 1, 2).cpu()
for li, layer in enumerate(mine._layers()[:-1]):
    g = layer.spec.geom(h.shape[0], h.shape[1], h.shape[2])
    wd, wu = layer.cache.get(layer.conv.weight, g)
    raw = F_.conv_up(h, wu, g)
    bn = layer.bn
    stats = F_.bn_train_fwd(raw, bn.weight.detach(), bn.bias.detach(), None, None, None, 0.1, 1e-5)
    y = F_.scale_shift_act(raw, stats[2], stats[3], 1, 0.0)
    raw_ref = acts[3 * li]            # conv output after the emulation's rounding hook? (hook order dependent)
    bn_ref = acts[3 * li + 1]         # NOTE: ReLU(inplace) ran on this tensor after the clone was taken -> pre-ReLU values
    C = raw.shape[-1]
    zm = nchw(raw) * stats[2].cpu().view(1, C, 1, 1) + stats[3].cpu().view(1, C, 1, 1)
    flips = ((zm > 0) != (bn_ref > 0))
    dr = (nchw(raw) - raw_ref)
    nz_diff = (dr != 0).float().mean().item()
    print(f"layer {li}: raw_ref bf16-valued {bool(torch.equal(raw_ref, raw_ref.bfloat16().float()))}; frac(raw differs) {nz_diff:.4f}; "
          f"max|raw diff|/max|raw| {float(dr.abs().max()/raw_ref.abs().max()):.2e}; flips {int(flips.sum())} of {flips.numel()} "
          f"({flips.float().mean().item():.2e}); |z_ref| at flips: max {float(bn_ref[flips].abs().max()) if flips.any() else 0:.2e}; "
          f"mean err {float((stats[0].cpu() - raw_ref.mean((0,2,3))).abs().max()):.2e}")
    h = y
