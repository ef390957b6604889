"""Debug: capture the real inputs of bn_act_bwd inside the generator backward (fp32) and compare each call's outputs
with a float64 torch recomputation."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.util import make_pair, rel_err
from importlib import import_module
import vaegan_b200
F_ = import_module("vaegan_b200.functional")
prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
hw, nz, B = 64, 128, 8
o_nets, nets = make_pair(hw, nz, prec)
mine = nets[1]
calls = []
orig = F_.bn_act_bwd
def spy(dy, x, stats, act, slope, dgamma, dbeta):
    dx = orig(dy, x, stats, act, slope, dgamma, dbeta)
    torch.cuda.synchronize()
    calls.append((dy.detach().clone(), x.detach().clone(), stats.detach().clone(), act, slope, dgamma.clone(), dbeta.clone(), dx.detach().clone()))
    return dx
F_.bn_act_bwd = spy
z = torch.randn(B, nz, 1, 1, generator=torch.Generator().manual_seed(11))
up = torch.randn(B, 3, hw, hw, generator=torch.Generator().manual_seed(12))
out = mine(z.cuda().requires_grad_(True))
(out * up.cuda()).sum().backward()
for (dy, x, stats, act, slope, dg, db, dx) in calls:
    C = x.shape[-1]
    X, DY = x.reshape(-1, C).double().cpu(), dy.reshape(-1, C).double().cpu()
    mean, rstd, scale, shift = [s.double().cpu() for s in stats]
    n = X.shape[0]
    zz = X * scale + shift
    dz = DY * (zz > 0)
    xhat = (X - mean) * rstd
    s0, s1 = dz.sum(0), (dz * xhat).sum(0)
    dx_ref = scale * (dz - s0 / n - xhat * s1 / n)
    # also check the stats themselves against exact
    m_ex, v_ex = X.mean(0), X.var(0, unbiased=False)
    print(f"C={C} rows={n}: dbeta rel {rel_err(db, s0):.3e}  dgamma rel {rel_err(dg, s1):.3e}  dx rel {rel_err(dx.reshape(-1, C), dx_ref):.3e} | "
          f"mean abs err {float((mean - m_ex).abs().max()):.3e} rstd rel {float(((rstd - 1/torch.sqrt(v_ex + 1e-5))/rstd).abs().max()):.3e} | "
          f"sum|dz|/|sum dz| median {float((dz.abs().sum(0)/(s0.abs()+1e-30)).median()):.1f}  max|s0| {float(s0.abs().max()):.3e} worst-channel abs err {float((db.double().cpu()-s0).abs().max()):.3e}")
