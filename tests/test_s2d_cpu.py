"""CPU: the index tables of functional.S2DWeightMap - the space-to-depth rewrite of the three image-side
convolutions (main_vae.py:37 first ConvBlock, gan_code.py:49 last ConvTranspose2d, gan_code.py:59 first Conv2d) -
checked with torch CPU convolutions: same outputs, and the folded-back weight gradient equals autograd's."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F


def _fn():
    from importlib import import_module
    import vaegan_b200  # noqa: F401
    return import_module("vaegan_b200.functional")


def s2d(img, o):
    """[B, C, H, W] -> [B, H/2+o, W/2+o, 64]: slot (sy*2+sx)*16 + c of block (Y, X) = pixel (2Y-o+sy, 2X-o+sx)."""
    B, C, H, W = img.shape
    out = torch.zeros(B, H // 2 + o, W // 2 + o, 64, dtype=img.dtype)
    for Y in range(H // 2 + o):
        for X in range(W // 2 + o):
            for sy in range(2):
                for sx in range(2):
                    py, px = 2 * Y - o + sy, 2 * X - o + sx
                    if 0 <= py < H and 0 <= px < W:
                        out[:, Y, X, (sy * 2 + sx) * 16:(sy * 2 + sx) * 16 + C] = img[:, :, py, px]
    return out


def _gather(master, idx, shape):
    flat = torch.cat([master.reshape(-1), master.new_zeros(1)])          # index -1 -> the appended zero
    return flat[torch.from_numpy(idx).long()].reshape(shape)


@pytest.mark.parametrize("pad", [0, 1])
@pytest.mark.parametrize("n", [32, 16])
def test_down_k4s2_equals_k2s1_on_s2d(pad, n):
    fn = _fn()
    g = torch.Generator().manual_seed(7 + pad)
    img = torch.randn(2, 3, 12, 8, generator=g)
    w = torch.randn(n, 3, 4, 4, generator=g)
    m = fn.S2DWeightMap(fn.ConvSpec("down", n, 3, 4, 2, pad))
    assert m.origin == pad and m.eq_spec == fn.ConvSpec("down", n, 64, 2, 1, 0) and m.fan == 1
    weq = _gather(w, m._fwd_np, (n, 64, 2, 2))
    y_ref = F.conv2d(img, w, None, 2, pad)
    y_eq = F.conv2d(s2d(img, pad).permute(0, 3, 1, 2), weq, None, 1, 0)
    assert torch.allclose(y_eq, y_ref, atol=1e-5)
    # gradient fold-back == autograd through the gather
    wr = w.clone().requires_grad_(True)
    dweq = torch.randn(weq.shape, generator=g)
    (_gather(wr, m._fwd_np, weq.shape) * dweq).sum().backward()
    folded = _gather(dweq, m._bwd_np, (w.numel(), m.fan)).sum(1).reshape(w.shape)
    assert torch.allclose(folded, wr.grad, atol=1e-6)


@pytest.mark.parametrize("m_ch", [64, 16])
def test_up_k3s1p1_equals_k4s2p1_into_s2d(m_ch):
    fn = _fn()
    g = torch.Generator().manual_seed(11)
    x = torch.randn(2, m_ch, 8, 6, generator=g)
    w = torch.randn(m_ch, 3, 3, 3, generator=g)                           # ConvTranspose2d.weight [Cin, Cout, 3, 3]
    m = fn.S2DWeightMap(fn.ConvSpec("up", m_ch, 3, 3, 1, 1))
    assert m.origin == 0 and m.eq_spec == fn.ConvSpec("down", 64, m_ch, 4, 2, 1) and m.fan == 4
    weq = _gather(w, m._fwd_np, (64, m_ch, 4, 4))
    y_ref = F.conv_transpose2d(x, w, None, 1, 1)                          # [2, 3, 8, 6]
    y_eq = F.conv2d(x, weq, None, 2, 1).permute(0, 2, 3, 1)               # [2, 4, 3, 64] = s2d(y_ref, 0)
    assert torch.allclose(y_eq, s2d(y_ref, 0), atol=1e-5)
    wr = w.clone().requires_grad_(True)
    dweq = torch.randn(weq.shape, generator=g)
    (_gather(wr, m._fwd_np, weq.shape) * dweq).sum().backward()
    folded = _gather(dweq, m._bwd_np, (w.numel(), m.fan)).sum(1).reshape(w.shape)
    assert torch.allclose(folded, wr.grad, atol=1e-5)


def test_eligibility():
    fn = _fn()
    ok = fn.S2DWeightMap.eligible
    assert ok(fn.ConvSpec("down", 64, 3, 4, 2, 1)) and ok(fn.ConvSpec("down", 32, 3, 4, 2, 0))
    assert ok(fn.ConvSpec("up", 64, 3, 3, 1, 1))
    assert not ok(fn.ConvSpec("down", 64, 64, 4, 2, 1)) and not ok(fn.ConvSpec("down", 64, 3, 3, 1, 1))
    assert not ok(fn.ConvSpec("up", 64, 3, 4, 2, 1))


def test_linear_over_nchw_flatten_is_gemm_over_nhwc_flatten():
    """The identity behind functional.LinearGemmMap / vg_linear_permute: nn.Linear on x.view(B, -1) of an NCHW map
    (main_vae.py:53) == a GEMM on the NHWC flatten with W'[n][tap*C + c] = W[n][c*kk + tap] and zero rows up to n_pad;
    the gradient folds back through the same index map."""
    fn = _fn()
    g = torch.Generator().manual_seed(3)
    B, C, h, nz = 3, 16, 5, 100
    m = fn.LinearGemmMap(fn.ConvSpec("down", nz, C, h, 1, 0))
    assert m.n_pad == 128 and m.eq_spec == fn.ConvSpec("down", 128, C * h * h, 1, 1, 0)
    assert fn.LinearGemmMap.needed(fn.ConvSpec("down", 100, 256, 14, 1, 0))          # reference encoder, 256x256
    assert fn.LinearGemmMap.needed(fn.ConvSpec("down", 128, 256, 14, 1, 0))          # 196 taps
    assert not fn.LinearGemmMap.needed(fn.ConvSpec("down", 128, 256, 2, 1, 0))       # cfg 2 heads stay convolutions
    x = torch.randn(B, C, h, h, generator=g)
    W = torch.randn(nz, C * h * h, generator=g)
    kk = h * h
    Wp = torch.zeros(m.n_pad, kk * C)
    Wp[:nz] = W.view(nz, C, kk).permute(0, 2, 1).reshape(nz, kk * C)                  # what mode 0 of the kernel writes
    y_ref = F.linear(x.view(B, -1), W)
    y_eq = x.permute(0, 2, 3, 1).reshape(B, -1) @ Wp.t()
    assert torch.allclose(y_eq[:, :nz], y_ref, atol=1e-4) and float(y_eq[:, nz:].abs().max()) == 0.0
    dWp = torch.randn(m.n_pad, kk * C, generator=g)
    folded = dWp[:nz].view(nz, kk, C).permute(0, 2, 1).reshape(nz, C * kk)            # what mode 1 adds to dW
    Wr = W.clone().requires_grad_(True)
    Wq = torch.zeros(m.n_pad, kk * C)
    Wq = torch.cat([Wr.view(nz, C, kk).permute(0, 2, 1).reshape(nz, kk * C), torch.zeros(m.n_pad - nz, kk * C)])
    (Wq * dWp).sum().backward()
    assert torch.allclose(folded, Wr.grad, atol=1e-6)


def test_pad_rows_map_shapes():
    fn = _fn()
    m = fn.PadRowsMap(fn.ConvSpec("up", 100, 1024, 4, 1, 0))
    assert m.n_pad == 128 and m.eq_spec == fn.ConvSpec("up", 128, 1024, 4, 1, 0) and m.row == 1024 * 16
    assert fn.PadRowsMap.needed(fn.ConvSpec("up", 100, 1024, 4, 1, 0))
    assert not fn.PadRowsMap.needed(fn.ConvSpec("up", 128, 1024, 4, 1, 0))
    assert not fn.PadRowsMap.needed(fn.ConvSpec("down", 64, 3, 4, 2, 1))
