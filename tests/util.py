"""Shared helpers for the parity tests (test infrastructure)."""
import copy

import torch


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max |a-b| relative to max |b| (b = oracle)."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def cosine(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.detach().double().cpu().flatten(), b.detach().double().cpu().flatten()
    return float((a @ b) / (a.norm() * b.norm() + 1e-300))


def make_pair(hw=64, nz=128, precision="fp32", seed=42, device="cuda", width=1):
    """(oracle nets on CPU, vaegan_b200 nets on `device`) with identical weights and buffers."""
    import vaegan_b200 as vb
    from oracle import vaegan_oracle as vo
    o_nets = vo.build_nets(vo.NetConfig(hw=hw, nz=nz, width=width, seed=seed))
    e = vb.Encoder([3, hw, hw], nz, width=width, precision=precision)
    g = vb.Generator(nz=nz, ngf=64 * width, hw=hw, precision=precision)
    d = vb.Discriminator(ndf=64 * width, hw=hw, precision=precision)
    for mine, ref in zip((e, g, d), o_nets):
        mine.load_state_dict(copy.deepcopy(ref.state_dict()))
        mine.to(device)
    return o_nets, (e, g, d)
