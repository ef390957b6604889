"""Debug: layer-by-layer comparison of the bf16 generator against the bf16-emulated oracle (test infrastructure)."""
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
# forward, layer by layer (reference: hook outputs of every module)
acts_ref = {}
for i, m in enumerate(ref.main):
    m.register_forward_hook(lambda mod, a, out, i=i: acts_ref.__setitem__(i, out.detach().clone()))
zr = z.clone().requires_grad_(True)
out_ref = ref(zr)
layers = mine._layers()
h = F_.nchw_to_nhwc(z.cuda(), torch.bfloat16)
idx = 0
for li, layer in enumerate(layers):
    last = li == len(layers) - 1
    h = layer(h, True, fuse_act=not last)
    idx += 3 if not last else 1
    ref_act = acts_ref[idx - 1]
    mine_nchw = h.detach().float().permute(0, 3, 1, 2).cpu()
    print(f"layer {li}: ref module {idx-1} {tuple(ref_act.shape)} rel_err {rel_err(mine_nchw, ref_act.to(torch.bfloat16).float()):.3e} "
          f"cos {cosine(mine_nchw, ref_act):.6f}")
