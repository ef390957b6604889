"""GPU: every C-ABI kernel family through ctypes (vaegan_b200.functional) against torch CPU fp32 references of
the same op.  fp32 path: 1e-5..1e-4 relative; bf16 path: operands are bf16-representable so only the output rounding
(2^-9 relative) separates the tensor-core result from the reference."""
import ctypes

import pytest
import torch
import torch.nn.functional as F

from tests.util import rel_err

pytestmark = pytest.mark.gpu


def _fn():
    from importlib import import_module
    import vaegan_b200  # noqa: F401
    return import_module("vaegan_b200.functional")


def _nhwc(t):      # NCHW cpu -> NHWC contiguous
    return t.permute(0, 2, 3, 1).contiguous()


def _nchw(t):
    return t.permute(0, 3, 1, 2).contiguous()


CONV_CASES = [
    # kind, B, H, W, Cin, Cout, k, s, p
    ("down", 4, 16, 16, 64, 128, 4, 2, 1),      # discriminator stage
    ("down", 3, 31, 31, 32, 64, 4, 2, 0),       # encoder stage, odd extents, SWIZZLE_64B operands
    ("down", 5, 64, 64, 3, 64, 4, 2, 1),        # image layer (CUDA-core path in both modes)
    ("down", 6, 2, 2, 256, 128, 2, 1, 0),       # fc head as a full-extent conv
    ("down", 8, 4, 4, 512, 1, 4, 1, 0),         # final discriminator conv (gemv)
    ("up", 4, 8, 8, 128, 64, 4, 2, 1),          # generator stage (4-phase)
    ("up", 16, 1, 1, 128, 256, 4, 1, 0),        # first generator layer (dense GEMM)
    ("up", 2, 16, 16, 64, 3, 3, 1, 1),          # last generator layer k3
    ("up", 130, 4, 4, 256, 128, 4, 2, 1),       # batch larger than one M tile
    ("up", 8, 32, 32, 128, 64, 4, 2, 1),        # last BN'd generator stage: 32-wide tiles, 8 row tiles per image
    ("down", 4, 64, 64, 3, 32, 4, 2, 0),        # first encoder layer (p=0), 32 output channels
    ("down", 3, 16, 16, 16, 32, 4, 2, 1),       # native-256 discriminator widths (16 / 32 channels)
    ("up", 4, 16, 16, 32, 16, 4, 2, 1),         # native-256 generator tail
]


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
@pytest.mark.parametrize("case", CONV_CASES, ids=lambda c: f"{c[0]}-B{c[1]}-{c[2]}x{c[3]}-{c[4]}to{c[5]}-k{c[6]}s{c[7]}p{c[8]}")
def test_conv_fwd_dgrad_wgrad(case, prec):
    fn = _fn()
    kind, B, H, W, cin, cout, k, s, p = case
    gen = torch.Generator().manual_seed(hash(case) & 0xFFFF)
    rb = (lambda t: t.bfloat16().float()) if prec == "bf16" else (lambda t: t)
    x = rb(torch.randn(B, cin, H, W, generator=gen))
    if kind == "down":
        w = rb(torch.randn(cout, cin, k, k, generator=gen) * 0.1)
        spec = fn.ConvSpec("down", cout, cin, k, s, p)
        y_ref = F.conv2d(x, w, None, s, p)
    else:
        w = rb(torch.randn(cin, cout, k, k, generator=gen) * 0.1)
        spec = fn.ConvSpec("up", cin, cout, k, s, p)
        y_ref = F.conv_transpose2d(x, w, None, s, p)
    dy = rb(torch.randn(y_ref.shape, generator=gen))
    xr, wr = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
    (F.conv2d(xr, wr, None, s, p) if kind == "down" else F.conv_transpose2d(xr, wr, None, s, p)).backward(dy)

    dt = fn.PRECISION_DTYPE[prec]
    big_c = cin if kind == "down" else cout
    cpad = fn.padded_channels(big_c, dt)           # bf16: 3-channel image tensors are stored with 16 channels
    g = spec.geom(B, H, W, cpad)

    def dev(t_nchw, pad_to=None):                  # NCHW cpu -> NHWC device tensor, optionally channel-padded
        t = _nhwc(t_nchw)
        if pad_to is not None and pad_to != t.shape[-1]:
            t = torch.cat([t, torch.zeros(*t.shape[:-1], pad_to - t.shape[-1])], dim=-1)
        return t.contiguous().cuda().to(dt)

    xd = dev(x, cpad if kind == "down" else None)
    dyd = dev(dy, cpad if kind == "up" else None)
    wdv = w.cuda()
    if prec == "bf16":
        wd, wu = fn.pack_weights(wdv, g)
        w_fwd, w_bwd = (wd, wu) if kind == "down" else (wu, wd)
    else:
        w_fwd = w_bwd = wdv
    if kind == "down":
        y = fn.conv_down(xd, w_fwd, g)
        dx = fn.conv_up(dyd, w_bwd, g)
        dw = fn.conv_wgrad(dyd, xd, g)
    else:
        y = fn.conv_up(xd, w_fwd, g)
        dx = fn.conv_down(dyd, w_bwd, g)
        dw = fn.conv_wgrad(xd, dyd, g)
    torch.cuda.synchronize()
    big_out = dx if kind == "down" else y          # the tensor living on the (possibly padded) big side
    if cpad != big_c:
        assert float(big_out[..., big_c:].float().abs().max()) == 0.0     # padded channels stay exactly zero
    crop = lambda t, is_big: t[..., :big_c] if (is_big and cpad != big_c) else t
    tol_out = 1e-5 if prec == "fp32" else 6e-3      # bf16: output rounding only
    assert rel_err(_nchw(crop(y, kind == "up").float().cpu()), y_ref) < tol_out
    assert rel_err(_nchw(crop(dx, kind == "down").float().cpu()), xr.grad) < tol_out
    assert dw.shape == wr.grad.shape
    assert rel_err(dw.cpu(), wr.grad) < 2e-5           # fp32 accumulate + fp32 output in both modes


@pytest.mark.parametrize("sc,bc,bcv,k", [(64, 128, 128, 4), (48, 80, 80, 3), (64, 16, 3, 4), (64, 16, 3, 3),
                                         (1, 512, 512, 4), (130, 34, 34, 5), (128, 512, 512, 4)])
def test_pack_weights_bit_exact(sc, bc, bcv, k):
    """wd[tap][small][big] / wu[tap][big][small] are the bf16 roundings of the master, zero in the padded channels -
    both through the single-layer entry point and through the one-launch multi-layer path."""
    fn = _fn()
    from vaegan_b200 import _lib
    gen = torch.Generator().manual_seed(sc * 7 + bc)
    w = torch.randn(sc, bcv, k, k, generator=gen)
    ref = torch.zeros(sc, bc, k * k)
    ref[:, :bcv] = w.view(sc, bcv, k * k)
    ref = ref.bfloat16()
    wd_ref, wu_ref = ref.permute(2, 0, 1).contiguous(), ref.permute(2, 1, 0).contiguous()
    g = _lib.VgConvGeom(1, 8, 8, bc, 4, 4, sc, k, 1, 0, bcv if bcv != bc else 0)
    wdev = w.cuda()
    wd, wu = fn.pack_weights(wdev, g)
    assert torch.equal(wd.cpu(), wd_ref) and torch.equal(wu.cpu(), wu_ref)
    wd2, wu2 = torch.empty_like(wd).fill_(7.0), torch.empty_like(wu).fill_(7.0)
    w_other = torch.randn(32, 32, 2, 2, generator=gen).cuda()
    od, ou = torch.empty(4, 32, 32, dtype=torch.bfloat16, device="cuda"), torch.empty(4, 32, 32, dtype=torch.bfloat16, device="cuda")
    items = (_lib.VgPackItem * 2)(
        _lib.VgPackItem(w_other.data_ptr(), od.data_ptr(), ou.data_ptr(), 32, 32, 0, 4),
        _lib.VgPackItem(wdev.data_ptr(), wd2.data_ptr(), wu2.data_ptr(), sc, bc, bcv if bcv != bc else 0, k * k))
    _lib.call("vg_pack_weights_multi", items, 2, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    assert torch.equal(wd2.cpu(), wd_ref) and torch.equal(wu2.cpu(), wu_ref)
    o_ref = w_other.cpu().view(32, 32, 4).bfloat16()
    assert torch.equal(od.cpu(), o_ref.permute(2, 0, 1).contiguous())
    assert torch.equal(ou.cpu(), o_ref.permute(2, 1, 0).contiguous())


@pytest.mark.parametrize("origin", [0, 1])
@pytest.mark.parametrize("mode", [0, 1, 2])
def test_s2d_layout_kernels(origin, mode):
    """vg_nchw_to_s2d / vg_s2d_to_nchw against the CPU definition of the space-to-depth image layout
    (tests/test_s2d_cpu.py::s2d), including the fused noise (mode 1) and Tanh-backward (mode 2) forms."""
    from tests.test_s2d_cpu import s2d
    fn = _fn()
    gen = torch.Generator().manual_seed(origin * 3 + mode)
    B, C, H, W = 3, 3, 12, 20
    src, aux = torch.randn(B, C, H, W, generator=gen), torch.rand(B, C, H, W, generator=gen) * 2 - 1
    ref = {0: src, 1: (src + 0.05 * aux).clamp(-1, 1), 2: src * (1 - aux * aux)}[mode]
    got = fn.nchw_to_nhwc(src.cuda(), torch.bfloat16, aux=aux.cuda() if mode else None, mode=mode, sigma=0.05,
                          clamp=(mode == 1), s2d_origin=origin)
    torch.cuda.synchronize()
    assert got.shape == (B, H // 2 + origin, W // 2 + origin, 64)
    assert torch.equal(got.cpu(), s2d(ref, origin).bfloat16())
    back = fn.nhwc_to_nchw(got, fn.ACT_TANH, channels=C, s2d_origin=origin)
    torch.cuda.synchronize()
    assert torch.allclose(back.cpu(), torch.tanh(ref.bfloat16().float()), atol=1e-6)


def test_gather_f32():
    fn = _fn()
    gen = torch.Generator().manual_seed(5)
    src = torch.randn(1000, generator=gen)
    idx = torch.randint(-1, 1000, (300, 4), generator=gen, dtype=torch.int32)
    dst = torch.randn(300, generator=gen)
    ref = dst + torch.where(idx >= 0, src[idx.clamp(min=0).long()], torch.zeros(())).sum(1)
    d, s_dev, i_dev = dst.cuda(), src.cuda(), idx.cuda()        # keep the device copies alive across the call
    fn.call("vg_gather_f32", fn._p(d), fn._p(s_dev), fn._p(i_dev), 300, 4, 1, fn._stream())
    torch.cuda.synchronize()
    assert torch.allclose(d.cpu(), ref, atol=1e-5)


FUSE_CASES = [
    # kind, B, H, W, Cin, Cout, k, s, p, groups, bias
    ("down", 8, 16, 16, 64, 128, 4, 2, 1, 2, False),     # discriminator stage, real/fake pair
    ("down", 6, 31, 31, 32, 64, 4, 2, 0, 1, True),       # encoder stage (bias + BatchNorm), odd extents
    ("up", 8, 8, 8, 128, 64, 4, 2, 1, 1, False),         # generator stage, 4 output phases
    ("up", 256, 1, 1, 128, 256, 4, 1, 0, 2, False),      # first generator layer: dense GEMM, channel = column % C
    ("down", 260, 8, 8, 64, 256, 4, 2, 1, 1, False),     # several tiles per CTA, two N tiles... and a ragged last tile
]


@pytest.mark.parametrize("case", FUSE_CASES, ids=lambda c: f"{c[0]}-B{c[1]}-{c[2]}x{c[3]}-{c[4]}to{c[5]}-g{c[9]}")
def test_fused_bn_stats_epilogue(case):
    """VG_EPI_BN_STATS + vg_bn_apply_from_sums against the stand-alone statistics / apply kernels on the same raw
    convolution output (which must be bit-identical with and without the fused epilogue)."""
    fn = _fn()
    kind, B, H, W, cin, cout, k, s, p, groups, with_bias = case
    gen = torch.Generator().manual_seed(B * 131 + cin)
    dt = torch.bfloat16
    x = torch.randn(B, H, W, cin, generator=gen).cuda().to(dt)
    if kind == "down":
        spec, w = fn.ConvSpec("down", cout, cin, k, s, p), torch.randn(cout, cin, k, k, generator=gen).cuda() * 0.1
    else:
        spec, w = fn.ConvSpec("up", cin, cout, k, s, p), torch.randn(cin, cout, k, k, generator=gen).cuda() * 0.1
    g = spec.geom(B, H, W)
    wd, wu = fn.pack_weights(w, g)
    bias = (torch.randn(cout, generator=gen).cuda() if with_bias else None)
    C = cout
    sums = torch.zeros(groups * 2 * C, device="cuda")
    ep = fn.make_epilogue(fn.EPI_BN_STATS, groups, C, sums=sums)
    assert fn.epilogue_supported(g, kind == "up", ep)
    if kind == "down":
        raw0, raw1 = fn.conv_down(x, wd, g, bias), fn.conv_down(x, wd, g, bias, ep=ep)
    else:
        raw0, raw1 = fn.conv_up(x, wu, g), fn.conv_up(x, wu, g, ep=ep)
    assert torch.equal(raw0, raw1)
    gamma, beta = torch.rand(C, generator=gen).cuda() + 0.5, torch.randn(C, generator=gen).cuda()
    rm0, rv0 = torch.randn(C, generator=gen).cuda(), torch.rand(C, generator=gen).cuda() + 0.5
    rm1, rv1 = rm0.clone(), rv0.clone()
    nbt0, nbt1 = (torch.zeros((), dtype=torch.int64, device="cuda") for _ in range(2))
    y1, st1 = fn.bn_apply_from_sums(raw1, sums, groups, gamma, beta, rm1, rv1, nbt1, 0.1, 1e-5, 2, 0.2)
    rg = raw0.view(groups, -1, C)
    for i in range(groups):
        st0 = fn.bn_train_fwd(rg[i], gamma, beta, rm0, rv0, nbt0, 0.1, 1e-5)
        y0 = fn.scale_shift_act(rg[i], st0[2], st0[3], 2, 0.2)
        torch.cuda.synchronize()
        for j, name in enumerate(("mean", "rstd", "scale", "shift")):
            assert rel_err(st1[i, j].cpu(), st0[j].cpu()) < 2e-5, (i, name)
        assert rel_err(y1.view(groups, -1, C)[i].float().cpu(), y0.float().cpu()) < 8e-3     # one bf16 ulp
    assert rel_err(rm1.cpu(), rm0.cpu()) < 1e-5 and rel_err(rv1.cpu(), rv0.cpu()) < 2e-5
    assert int(nbt1) == int(nbt0) == groups


@pytest.mark.parametrize("act,slope", [(1, 0.0), (2, 0.2)])
@pytest.mark.parametrize("case", [FUSE_CASES[0], FUSE_CASES[2], ("down", 5, 33, 33, 64, 64, 2, 1, 0, 1, False)],
                         ids=lambda c: f"{c[0]}-B{c[1]}-{c[2]}x{c[3]}-{c[4]}to{c[5]}")
def test_fused_activation_forward_epilogue(case, act, slope):
    """VG_EPI_ACT_FWD == convolution followed by the stand-alone activation pass, bit for bit (same fp32 accumulator,
    same single rounding to bf16)."""
    fn = _fn()
    kind, B, H, W, cin, cout, k, s, p, _, _ = case
    gen = torch.Generator().manual_seed(B + cout)
    x = torch.randn(B, H, W, cin, generator=gen).cuda().bfloat16()
    if kind == "down":
        spec, w = fn.ConvSpec("down", cout, cin, k, s, p), torch.randn(cout, cin, k, k, generator=gen).cuda() * 0.1
    else:
        spec, w = fn.ConvSpec("up", cin, cout, k, s, p), torch.randn(cin, cout, k, k, generator=gen).cuda() * 0.1
    g = spec.geom(B, H, W)
    wd, wu = fn.pack_weights(w, g)
    ep = fn.make_epilogue(fn.EPI_ACT_FWD, 1, 0, act, slope)
    assert fn.epilogue_supported(g, kind == "up", ep)
    conv = (lambda e=None: fn.conv_down(x, wd, g, ep=e)) if kind == "down" else (lambda e=None: fn.conv_up(x, wu, g, ep=e))
    raw, fused = conv(), conv(ep)
    ref = fn.scale_shift_act(raw, None, None, act, slope)
    torch.cuda.synchronize()
    assert torch.equal(fused, ref)


@pytest.mark.parametrize("act,slope", [(0, 0.0), (1, 0.0), (2, 0.2)])
@pytest.mark.parametrize("case", [FUSE_CASES[0], FUSE_CASES[2], FUSE_CASES[3], FUSE_CASES[4]],
                         ids=lambda c: f"{c[0]}-B{c[1]}-{c[2]}x{c[3]}-{c[4]}to{c[5]}")
def test_fused_eval_batchnorm_epilogue(case, act, slope):
    """VG_EPI_AFFINE_ACT_FWD (eval-mode BatchNorm + activation in the epilogue, generation / validation passes)
    against the fp32 reference act(conv * scale + shift) - tighter than, and within one bf16 ulp of, the two-pass form
    it replaces (convolution rounded to bf16, then vg_scale_shift_act)."""
    fn = _fn()
    kind, B, H, W, cin, cout, k, s, p, _, _ = case
    gen = torch.Generator().manual_seed(B * 7 + cout)
    x = torch.randn(B, H, W, cin, generator=gen).cuda().bfloat16()
    if kind == "down":
        spec, w = fn.ConvSpec("down", cout, cin, k, s, p), torch.randn(cout, cin, k, k, generator=gen).cuda() * 0.1
    else:
        spec, w = fn.ConvSpec("up", cin, cout, k, s, p), torch.randn(cin, cout, k, k, generator=gen).cuda() * 0.1
    g = spec.geom(B, H, W)
    wd, wu = fn.pack_weights(w, g)
    C = cout
    gamma, beta = torch.rand(C, generator=gen).cuda() + 0.5, torch.randn(C, generator=gen).cuda()
    rm, rv = torch.randn(C, generator=gen).cuda() * 0.1, torch.rand(C, generator=gen).cuda() + 0.5
    stats = fn.bn_eval_coeffs(gamma, beta, rm, rv, 1e-5)
    ep = fn.make_epilogue(fn.EPI_AFFINE_ACT_FWD, 1, C, act, slope, stats=stats)
    assert fn.epilogue_supported(g, kind == "up", ep)
    conv = (lambda e=None: fn.conv_down(x, wd, g, ep=e)) if kind == "down" else (lambda e=None: fn.conv_up(x, wu, g, ep=e))
    raw, fused = conv(), conv(ep)
    two_pass = fn.scale_shift_act(raw, stats[2], stats[3], act, slope)
    # fp32 reference on the CPU from the same bf16 operands
    xc, wc = x.float().cpu().permute(0, 3, 1, 2), w.bfloat16().float().cpu()
    ref = F.conv2d(xc, wc, None, s, p) if kind == "down" else F.conv_transpose2d(xc, wc, None, s, p)
    if kind == "up" and H == 1 and W == 1:
        ref = ref                                      # dense first generator layer: [B, cout, k, k]
    ref = ref * stats[2].cpu().view(1, -1, 1, 1) + stats[3].cpu().view(1, -1, 1, 1)
    ref = ref if act == 0 else (ref.clamp_min(0) if act == 1 else torch.where(ref > 0, ref, ref * slope))
    ref = ref.permute(0, 2, 3, 1)
    torch.cuda.synchronize()
    assert rel_err(fused.float().cpu(), ref) < 5e-3                       # half a bf16 ulp of the largest value
    assert rel_err(fused.float().cpu(), two_pass.float().cpu()) < 1.2e-2    # the old path carries one more rounding


@pytest.mark.parametrize("act,slope", [(1, 0.0), (2, 0.2)])
@pytest.mark.parametrize("case", FUSE_CASES[:3] + [FUSE_CASES[4], ("down", 64, 4, 4, 512, 1, 4, 1, 0, 2, False)],
                         ids=lambda c: f"{c[0]}-B{c[1]}-{c[2]}x{c[3]}-{c[4]}to{c[5]}-g{c[9]}")
def test_fused_bn_bwd_epilogue(case, act, slope):
    """The dgrad of layer L+1 with VG_EPI_BN_BWD (and VG_EPI_ACT_BWD) against dgrad -> vg_bn_act_bwd / vg_act_bwd:
    layer L here is a BatchNorm'd tensor of the dgrad's output shape."""
    fn = _fn()
    kind, B, H, W, cin, cout, k, s, p, groups, _ = case
    gen = torch.Generator().manual_seed(B * 17 + cout)
    dt = torch.bfloat16
    if kind == "down":      # layer L+1 is a Conv2d: its dgrad is `up` and lands on the [B, H, W, cin] input
        spec, w = fn.ConvSpec("down", cout, cin, k, s, p), torch.randn(cout, cin, k, k, generator=gen).cuda() * 0.1
    else:
        spec, w = fn.ConvSpec("up", cin, cout, k, s, p), torch.randn(cin, cout, k, k, generator=gen).cuda() * 0.1
    g = spec.geom(B, H, W)
    oh, ow = spec.out_hw(H, W)
    wd, wu = fn.pack_weights(w, g)
    d_next = torch.randn(B, oh, ow, cout, generator=gen).cuda().to(dt)           # gradient of layer L+1's raw output
    raw_l = (torch.randn(B, H, W, cin, generator=gen) * 1.5 + 0.3).cuda().to(dt)   # layer L's raw conv output
    C = cin
    gamma, beta = torch.rand(C, generator=gen).cuda() + 0.5, torch.randn(C, generator=gen).cuda() * 0.3
    stats = torch.stack([fn.bn_train_fwd(r, gamma, beta, None, None, None, 0.1, 1e-5)
                         for r in raw_l.view(groups, -1, C)])
    dgrad = (lambda ep=None: fn.conv_up(d_next, wu, g, ep=ep)) if kind == "down" else \
            (lambda ep=None: fn.conv_down(d_next, wd, g, ep=ep))
    # reference: plain dgrad, then the stand-alone BatchNorm backward per group
    dy = dgrad()
    dg0, db0 = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    dx0 = torch.empty_like(raw_l)
    for i in range(groups):
        fn.bn_act_bwd(dy.view(groups, -1, C)[i], raw_l.view(groups, -1, C)[i], stats[i], act, slope, dg0, db0,
                      out=dx0.view(groups, -1, C)[i])
    # fused
    sums = torch.zeros(groups * 2 * C, device="cuda")
    ep = fn.make_epilogue(fn.EPI_BN_BWD, groups, C, act, slope, sums, raw_l, stats)
    assert fn.epilogue_supported(g, kind == "down", ep)
    dz = dgrad(ep)
    dg1, db1 = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    dx1 = fn.bn_bwd_apply_from_sums(dz, raw_l, stats, sums, groups, dg1, db1)
    torch.cuda.synchronize()
    assert rel_err(dx1.float().cpu(), dx0.float().cpu()) < 1.5e-2          # dz is rounded to bf16 once more
    assert rel_err(dg1.cpu(), dg0.cpu()) < 3e-3 and rel_err(db1.cpu(), db0.cpu()) < 3e-3
    if cout == 1:
        return          # the single-output head (GEMV kernels) fuses the BatchNorm form only
    # activation-only form
    ep3 = fn.make_epilogue(fn.EPI_ACT_BWD, 1, C, act, slope, None, raw_l, None)
    assert fn.epilogue_supported(g, kind == "down", ep3)
    dz3 = dgrad(ep3)
    ref3 = fn.act_bwd(dy, raw_l, act, slope)
    torch.cuda.synchronize()
    assert rel_err(dz3.float().cpu(), ref3.float().cpu()) < 8e-3


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
@pytest.mark.parametrize("act,slope", [(1, 0.0), (2, 0.2), (2, 0.01)])
@pytest.mark.parametrize("rows,C", [(4 * 31 * 31, 32), (7 * 4 * 4, 512), (2 * 64 * 64, 64), (3, 2048)])
def test_batchnorm_act_fwd_bwd(rows, C, act, slope, prec):
    fn = _fn()
    gen = torch.Generator().manual_seed(rows + C)
    dt = fn.PRECISION_DTYPE[prec]
    rb = (lambda t: t.to(dt).float())
    x = rb(torch.randn(rows, C, generator=gen) * 2 + 0.5)
    dy = rb(torch.randn(rows, C, generator=gen))
    gamma, beta = torch.rand(C, generator=gen) + 0.5, torch.randn(C, generator=gen)
    rm, rv = torch.randn(C, generator=gen), torch.rand(C, generator=gen) + 0.5
    xr, gr, br = x.clone().requires_grad_(True), gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    rm_ref, rv_ref = rm.clone(), rv.clone()
    z = F.batch_norm(xr, rm_ref, rv_ref, gr, br, True, 0.1, 1e-5)
    y_ref = F.relu(z) if act == 1 else F.leaky_relu(z, slope)
    y_ref.backward(dy)

    xd = x.cuda().to(dt).view(1, 1, rows, C)
    rmd, rvd, nbt = rm.cuda(), rv.cuda(), torch.zeros((), dtype=torch.int64, device="cuda")
    stats = fn.bn_train_fwd(xd, gamma.cuda(), beta.cuda(), rmd, rvd, nbt, 0.1, 1e-5)
    y = fn.scale_shift_act(xd, stats[2], stats[3], act, slope)
    dg, db = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    dx = fn.bn_act_bwd(dy.cuda().to(dt).view(1, 1, rows, C), xd, stats, act, slope, dg, db)
    torch.cuda.synchronize()
    out_tol = 1e-5 if prec == "fp32" else 6e-3
    # the fused (single cooperative launch) forms must agree with the two-kernel forms
    rm2, rv2, nbt2 = rm.cuda(), rv.cuda(), torch.zeros((), dtype=torch.int64, device="cuda")
    y2, stats2 = fn.bn_act_train_fwd(xd, gamma.cuda(), beta.cuda(), rm2, rv2, nbt2, 0.1, 1e-5, act, slope)
    dg2, db2 = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    dx2 = fn.bn_act_train_bwd(dy.cuda().to(dt).view(1, 1, rows, C), xd, stats2, act, slope, dg2, db2)
    torch.cuda.synchronize()
    assert int(nbt2) == 1 and rel_err(rm2, rm_ref) < 1e-5 and rel_err(rv2, rv_ref) < 1e-5
    assert rel_err(stats2, stats) < 1e-5
    assert rel_err(y2.float().view(rows, C), y_ref) < out_tol
    assert rel_err(dx2.float().view(rows, C), xr.grad) < (2e-4 if prec == "fp32" else 1e-2) * (10 if rows < 16 else 1)
    assert rel_err(dg2, gr.grad) < 1e-4 and rel_err(db2, br.grad) < 1e-4
    assert int(nbt) == 1
    assert rel_err(rmd, rm_ref) < 1e-5 and rel_err(rvd, rv_ref) < 1e-5
    assert rel_err(y.float().view(rows, C), y_ref) < out_tol
    assert rel_err(dx.float().view(rows, C), xr.grad) < (2e-4 if prec == "fp32" else 1e-2) * (10 if rows < 16 else 1)
    assert rel_err(dg, gr.grad) < 1e-4 and rel_err(db, br.grad) < 1e-4


def test_bn_eval_and_single_value_error():
    fn = _fn()
    C = 64
    g, b, rm, rv = (torch.rand(C) + 0.5, torch.randn(C), torch.randn(C), torch.rand(C) + 0.5)
    x = torch.randn(10, C)
    ref = F.batch_norm(x, rm, rv, g, b, False, 0.1, 1e-5)
    st = fn.bn_eval_coeffs(g.cuda(), b.cuda(), rm.cuda(), rv.cuda(), 1e-5)
    y = fn.scale_shift_act(x.cuda().view(1, 1, 10, C), st[2], st[3], 0, 0.0)
    assert rel_err(y.view(10, C), ref) < 1e-5
    with pytest.raises(RuntimeError, match="more than 1 value per channel"):
        fn.bn_train_fwd(torch.zeros(1, 1, 1, C, device="cuda"), g.cuda(), b.cuda(), None, None, None, 0.1, 1e-5)


def test_layout_edges_and_noise_modes():
    fn = _fn()
    gen = torch.Generator().manual_seed(3)
    x = torch.rand(3, 3, 20, 12, generator=gen) * 2 - 1
    n = torch.randn(3, 3, 20, 12, generator=gen)
    for dt in (torch.float32, torch.bfloat16):
        cp = fn.padded_channels(3, dt)
        a = fn.nchw_to_nhwc(x.cuda(), dt)
        assert a.shape == (3, 20, 12, cp)
        assert torch.equal(a[..., :3].float().cpu(), _nhwc(x).to(dt).float())
        assert float(a[..., 3:].float().abs().sum()) == 0.0 if cp > 3 else True
        b = fn.nchw_to_nhwc(x.cuda(), dt, aux=n.cuda(), mode=1, sigma=0.3, clamp=True)
        assert rel_err(b[..., :3].float(), _nhwc(torch.clamp(x + 0.3 * n, -1, 1)).to(dt).float()) < 1e-6
        back = fn.nhwc_to_nchw(a, 3, channels=3)      # tanh
        assert back.shape == x.shape
        assert rel_err(back, torch.tanh(x.to(dt).float())) < 1e-6
        c = fn.nchw_to_nhwc(n.cuda(), dt, aux=back, mode=2)
        assert rel_err(c[..., :3].float(), _nhwc(n * (1 - torch.tanh(x.to(dt).float()) ** 2)).to(dt).float()) < 1e-5


def test_reparam_kl_bce_mse_against_torch():
    fn = _fn()
    lib = __import__("vaegan_b200").load_library()
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    B, nz = 12, 128
    gen = torch.Generator().manual_seed(5)
    mu = torch.randn(B, nz, generator=gen)
    lv = torch.randn(B, nz, generator=gen) * 6          # some entries beyond the +-10 clamp
    eps = torch.randn(B, nz, generator=gen)
    dz = torch.randn(B, nz, generator=gen)
    mur, lvr = mu.clone().requires_grad_(True), lv.clone().requires_grad_(True)
    lvc = torch.clamp(lvr, min=-10, max=10)
    z_ref = mur + torch.exp(0.5 * lvc) * eps
    kl_ref = -0.5 * torch.sum(1 + lvc - mur.pow(2) - lvc.exp()) / B
    ((z_ref * dz).sum() + 0.07 * kl_ref).backward()
    mud, lvd, epsd = mu.cuda(), lv.cuda(), eps.cuda()
    z, kl = torch.empty(B, nz, device="cuda"), torch.zeros((), device="cuda")
    assert lib.vg_reparam_fwd(P(mud), P(lvd), P(epsd), B, nz, P(z), 0, P(kl), None) == 0
    dmu, dlv = torch.empty_like(mud), torch.empty_like(lvd)
    assert lib.vg_reparam_bwd(P(dz.cuda()), 0, P(mud), P(lvd), P(epsd), B, nz, None, 0.07, P(dmu), P(dlv), None) == 0
    assert rel_err(z, z_ref) < 1e-6 and abs(float(kl) - float(kl_ref)) < 1e-5 * abs(float(kl_ref))
    assert rel_err(dmu, mur.grad) < 1e-5 and rel_err(dlv, lvr.grad) < 1e-5

    p = torch.rand(B, generator=gen).clamp(1e-4, 1 - 1e-4)
    p[0], p[1] = 0.0, 1.0                                  # log clamp at -100 (nn.BCELoss)
    for target in (0.9, 0.1):
        pr = p.clone().requires_grad_(True)
        l_ref = F.binary_cross_entropy(pr, torch.full((B,), target))
        (0.1 * l_ref).backward()
        loss, dp = torch.zeros((), device="cuda"), torch.empty(B, device="cuda")
        assert lib.vg_bce(P(p.cuda()), B, target, 0.1, P(loss), 0, P(dp), None) == 0
        assert abs(float(loss) - float(l_ref)) < 1e-5 * abs(float(l_ref))
        assert rel_err(dp[2:], pr.grad[2:]) < 1e-5

    a, b = torch.randn(3, 3, 17, 19, generator=gen), torch.randn(3, 3, 17, 19, generator=gen)
    gi = torch.randn(3, 3, 17, 19, generator=gen)
    ar = a.clone().requires_grad_(True)
    l_ref = F.mse_loss(ar, b)
    l_ref.backward()
    ws = torch.empty(lib.vg_mse_workspace_bytes() // 4, device="cuda")
    loss, go = torch.zeros((), device="cuda"), torch.empty(a.shape, device="cuda")
    assert lib.vg_mse(P(a.cuda()), P(b.cuda()), a.numel(), 1.0, P(gi.cuda()), P(go), P(loss), P(ws), ws.numel() * 4, None) == 0
    assert abs(float(loss) - float(l_ref)) < 1e-6 * abs(float(l_ref))
    assert rel_err(go, ar.grad + gi) < 1e-6

    # one launch per discriminator update: BCE(real) + BCE(fake) of the stacked vector, both gradient seeds
    pp = torch.cat([p, p.flip(0)])
    pr = pp.clone().requires_grad_(True)
    l_ref = F.binary_cross_entropy(pr[:B], torch.full((B,), 0.9)) + F.binary_cross_entropy(pr[B:], torch.full((B,), 0.1))
    l_ref.backward()
    loss, dp = torch.zeros((), device="cuda"), torch.empty(2 * B, device="cuda")
    assert lib.vg_bce_pair(P(pp.cuda()), B, 0.9, 0.1, 1.0, P(loss), P(dp), None) == 0
    assert abs(float(loss) - float(l_ref)) < 1e-5 * abs(float(l_ref))
    ok = torch.ones(2 * B, dtype=torch.bool)
    ok[[0, 1, 2 * B - 1, 2 * B - 2]] = False             # p = 0 / 1: the clamped logs have no finite torch gradient
    assert rel_err(dp.cpu()[ok], pr.grad[ok]) < 1e-5

    # MSE + gradient + the step's total in one launch (vaegan_code.py:113 + :117), fp32 pixels and bf16 features;
    # called twice on the same workspace: the kernel re-arms it
    ws = torch.zeros(lib.vg_mse_workspace_bytes() // 4, device="cuda")
    kl, adv, wkl = (torch.tensor(v, device="cuda") for v in (15.5, 2.25, 0.07))
    for dt, code in ((torch.float32, 0), (torch.bfloat16, 1)):
        a2, b2, g2 = (torch.randn(4, 3, 16, 24, generator=gen).to(dt) for _ in range(3))
        ar = a2.float().requires_grad_(True)
        l_ref = F.mse_loss(ar, b2.float())
        l_ref.backward()
        for _ in range(2):
            loss, tot = torch.zeros((), device="cuda"), torch.zeros((), device="cuda")
            go = torch.empty(a2.shape, device="cuda", dtype=dt)
            assert lib.vg_mse_total(P(a2.cuda()), P(b2.cuda()), code, a2.numel(), 1.0, P(g2.cuda()), P(go), P(loss), P(kl),
                                    P(adv), P(wkl), 0.1, P(tot), P(ws), ws.numel() * 4, None) == 0
            assert abs(float(loss) - float(l_ref)) < 1e-6 * abs(float(l_ref))
            assert abs(float(tot) - (float(l_ref) + 0.07 * 15.5 + 0.1 * 2.25)) < 1e-6 * float(tot)
            assert rel_err(go, ar.grad + g2.float()) < (1e-6 if dt == torch.float32 else 8e-3)


def test_adam_matches_torch_optim():
    lib = __import__("vaegan_b200").load_library()
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    n = 10007
    gen = torch.Generator().manual_seed(9)
    p0 = torch.randn(n, generator=gen)
    pr = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([pr], lr=2e-4)
    pd, m, v = p0.cuda(), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    step = torch.zeros((), dtype=torch.int64, device="cuda")
    for it in range(5):
        g = torch.randn(n, generator=gen) * (10.0 ** (it - 2))
        pr.grad = g.clone()
        opt.step()
        assert lib.vg_adam_step(P(pd), P((2 * g).cuda()), P(m), P(v), n, 2e-4, 0.9, 0.999, 1e-8, P(step), 0.5, None) == 0
    assert int(step) == 5
    assert float((pd.cpu() - pr.detach()).abs().max()) < 1e-6      # |p| ~ 4: a couple of fp32 ulps
    assert rel_err(m, opt.state[pr]["exp_avg"]) < 1e-6 and rel_err(v, opt.state[pr]["exp_avg_sq"]) < 1e-6


def test_randn_statistics_and_streams():
    lib = __import__("vaegan_b200").load_library()
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    n = 1 << 20
    a, b = torch.empty(n, device="cuda"), torch.empty(n, device="cuda")
    off = torch.zeros((), dtype=torch.int64, device="cuda")
    assert lib.vg_randn(P(a), n, 123, P(off), 1, None) == 0
    assert lib.vg_randn(P(b), n, 123, P(off), 1, None) == 0
    assert int(off) == 2 and not torch.equal(a, b)
    for t in (a, b):
        assert abs(float(t.mean())) < 5e-3 and abs(float(t.std()) - 1) < 5e-3
        assert abs(float((t ** 4).mean()) - 3.0) < 0.05
    c = torch.empty(n, device="cuda")
    off2 = torch.zeros((), dtype=torch.int64, device="cuda")
    assert lib.vg_randn(P(c), n, 123, P(off2), 1, None) == 0
    assert torch.equal(a, c)                      # counter-based: same (seed, offset, stream) -> same draw


def test_error_convention():
    fn = _fn()
    from importlib import import_module
    err = import_module("vaegan_b200._lib").VaeganB200Error
    spec = fn.ConvSpec("down", 8, 8, 4, 1, 0)
    with pytest.raises(RuntimeError, match="Kernel size can't be greater than actual input size"):
        spec.geom(1, 2, 2)
    g = fn.ConvSpec("down", 8, 8, 4, 2, 1).geom(2, 8, 8)
    g.small_h = 7                                    # inconsistent geometry -> VG_ERR_SHAPE, message via vg_last_error
    with pytest.raises(err, match="inconsistent"):
        fn.conv_down(torch.zeros(2, 8, 8, 8, device="cuda"), torch.zeros(8, 8, 4, 4, device="cuda"), g)
