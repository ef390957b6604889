"""Tensor-level wrappers over the C-ABI and the autograd glue that lets the reference's nn.Module call sites
(encoder(x), decoder(z), discriminator(x), loss.backward()) run on the sm_100a kernels.

PyTorch is used for device memory, streams and the autograd tape only; every arithmetic pass below is a kernel of
libvaegan_b200.so.  Internal activations are NHWC tensors `[B, H, W, C]` in the precision's dtype.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import (ACT_LEAKY, ACT_NONE, ACT_RELU, ACT_SIGMOID, ACT_TANH, EPI_ACT_BWD, EPI_ACT_FWD, EPI_AFFINE_ACT_FWD, EPI_BN_BWD,
                   EPI_BN_STATS, VG_BF16, VG_F32, VgConvGeom, VgEpilogue, call)

_DT = {torch.float32: VG_F32, torch.bfloat16: VG_BF16}
PRECISION_DTYPE = {"fp32": torch.float32, "bf16": torch.bfloat16}


def _p(t: Optional[torch.Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise _lib.VaeganB200Error(f"{what}: tensor is on {t.device}; vaegan_b200 runs on sm_100 CUDA devices only "
                                   "(no CPU fallback)")


def _contig(t: torch.Tensor) -> torch.Tensor:
    return t if t.is_contiguous() else t.contiguous()


def padded_channels(c: int, dtype) -> int:
    """Channels per pixel of an internal NHWC tensor.  The bf16 path stores narrow (3-channel image) tensors with 16
    channels, zeros above the real ones, so that they are legal TMA / tcgen05 operands (32-byte rows)."""
    return 16 if (dtype == torch.bfloat16 and c < 16) else c


# --------------------------------------------------------------------------------------------- geometry
@dataclass(frozen=True)
class ConvSpec:
    """One reference conv layer.  kind 'down' = nn.Conv2d (and nn.Linear as a full-extent conv), 'up' =
    nn.ConvTranspose2d.  `small_c` / `big_c` follow include/vaegan_b200.h."""
    kind: str
    small_c: int
    big_c: int
    kernel: int
    stride: int
    pad: int

    def out_hw(self, h: int, w: int) -> Tuple[int, int]:
        k, s, p = self.kernel, self.stride, self.pad
        if self.kind == "down":
            if h + 2 * p < k or w + 2 * p < k:
                raise RuntimeError(f"Calculated padded input size per channel: ({h + 2 * p} x {w + 2 * p}). "
                                   f"Kernel size: ({k} x {k}). Kernel size can't be greater than actual input size")
            return (h + 2 * p - k) // s + 1, (w + 2 * p - k) // s + 1
        return (h - 1) * s - 2 * p + k, (w - 1) * s - 2 * p + k

    def geom(self, batch: int, in_h: int, in_w: int, big_c_tensor: Optional[int] = None) -> VgConvGeom:
        """`big_c_tensor`: channels per pixel of the big-side tensor when it is channel-padded (>= self.big_c)."""
        oh, ow = self.out_hw(in_h, in_w)
        bc = big_c_tensor or self.big_c
        valid = self.big_c if bc != self.big_c else 0
        if self.kind == "down":
            return VgConvGeom(batch, in_h, in_w, bc, oh, ow, self.small_c, self.kernel, self.stride, self.pad, valid)
        return VgConvGeom(batch, oh, ow, bc, in_h, in_w, self.small_c, self.kernel, self.stride, self.pad, valid)


# --------------------------------------------------------------------------------------------- raw ops
def pack_weights(w: torch.Tensor, g: VgConvGeom) -> Tuple[torch.Tensor, torch.Tensor]:
    """fp32 master [small_c, big_c_valid, k, k] -> (wd[tap, small_c, big_c], wu[tap, big_c, small_c]) bf16
    (zero-filled for padded big-side channels)."""
    kk = g.kernel * g.kernel
    wd = torch.empty((kk, g.small_c, g.big_c), dtype=torch.bfloat16, device=w.device)
    wu = torch.empty((kk, g.big_c, g.small_c), dtype=torch.bfloat16, device=w.device)
    call("vg_pack_weights_bf16", ctypes.byref(g), _p(w), _p(wd), _p(wu), _stream())
    return wd, wu


def conv_flops(g: VgConvGeom) -> float:
    """Algorithmic FLOPs of one contraction over this geometry: 2 x MACs over the VALID channels."""
    bc = g.big_c_valid if g.big_c_valid > 0 else g.big_c
    # space-to-depth layers run an EQUIVALENT 64-channel convolution; `algo_scale` (set by ConvLayerFn) brings the
    # count back to the reference layer's own MACs
    return 2.0 * g.batch * g.small_h * g.small_w * g.small_c * bc * g.kernel * g.kernel * getattr(g, "algo_scale", 1.0)


def conv_bytes(g: VgConvGeom, extra_big: int = 0, extra_small: int = 0) -> float:
    """Algorithmic HBM bytes of one bf16 contraction: both activation tensors once (+ the saved tensor a fused
    epilogue reads on the output side) and the packed weights once."""
    big = g.batch * g.big_h * g.big_w * g.big_c
    small = g.batch * g.small_h * g.small_w * g.small_c
    return 2.0 * ((1 + extra_big) * big + (1 + extra_small) * small + g.small_c * g.big_c * g.kernel * g.kernel)


def _conv_tag(g: VgConvGeom, kind: str) -> str:
    """Which kernel family the C-ABI dispatches this geometry to (for bench.py's per-kernel accounting)."""
    if g.small_c == 1 and g.small_h == 1 and g.small_w == 1:
        return "gemv"
    return kind


def make_epilogue(mode: int, groups: int = 1, channels: int = 0, act: int = ACT_NONE, slope: float = 0.0,
                  sums: Optional[torch.Tensor] = None, x: Optional[torch.Tensor] = None,
                  stats: Optional[torch.Tensor] = None) -> VgEpilogue:
    """VgEpilogue for the fused forms of conv_down / conv_up (include/vaegan_b200.h).  The caller keeps the tensors
    alive until the launch has been enqueued."""
    ep = VgEpilogue(mode, groups, channels, act, float(slope), None if sums is None else sums.data_ptr(),
                    None if x is None else x.data_ptr(), None if stats is None else stats.data_ptr())
    return ep


def epilogue_supported(g: VgConvGeom, up: bool, ep: VgEpilogue) -> bool:
    return bool(_lib.load().vg_conv_epilogue_supported(ctypes.byref(g), VG_BF16, int(up), ctypes.byref(ep)))


def conv_down(big: torch.Tensor, w: torch.Tensor, g: VgConvGeom, bias: Optional[torch.Tensor] = None,
              out_f32: bool = False, ep: Optional[VgEpilogue] = None) -> torch.Tensor:
    out_dtype = torch.float32 if out_f32 else big.dtype
    small = torch.empty((g.batch, g.small_h, g.small_w, g.small_c), dtype=out_dtype, device=big.device)
    if ep is not None:
        call("vg_conv_down_ex", ctypes.byref(g), _DT[big.dtype], _p(big), _p(w), _p(bias), _p(small), ctypes.byref(ep),
             _stream(), flops=conv_flops(g), tag=_conv_tag(g, "fprop"), nbytes=conv_bytes(g, 0, int(ep.mode in (2, 3))))
        return small
    ws, nbytes = None, 0
    if big.dtype == torch.bfloat16 and g.batch * g.small_h * g.small_w <= 1024:      # few output tiles: allow split-K
        nbytes = _lib.load().vg_conv_down_workspace_bytes(ctypes.byref(g))
        ws = _ws(nbytes, big.device)
    call("vg_conv_down", ctypes.byref(g), _DT[big.dtype], _p(big), _p(w), _p(bias), _p(small), int(out_f32), _p(ws),
         nbytes, _stream(), flops=conv_flops(g), tag=_conv_tag(g, "fprop"), nbytes=conv_bytes(g))
    return small


def conv_up(small: torch.Tensor, w: torch.Tensor, g: VgConvGeom, ep: Optional[VgEpilogue] = None) -> torch.Tensor:
    big = torch.empty((g.batch, g.big_h, g.big_w, g.big_c), dtype=small.dtype, device=small.device)
    if ep is not None:
        call("vg_conv_up_ex", ctypes.byref(g), _DT[small.dtype], _p(small), _p(w), _p(big), ctypes.byref(ep), _stream(),
             flops=conv_flops(g), tag=_conv_tag(g, "fprop"), nbytes=conv_bytes(g, int(ep.mode in (2, 3)), 0))
    else:
        call("vg_conv_up", ctypes.byref(g), _DT[small.dtype], _p(small), _p(w), _p(big), _stream(),
             flops=conv_flops(g), tag=_conv_tag(g, "fprop"), nbytes=conv_bytes(g))
    return big


def conv_wgrad(small: torch.Tensor, big: torch.Tensor, g: VgConvGeom, dw: Optional[torch.Tensor] = None,
               overwrite: bool = False, dst_zero: bool = False) -> torch.Tensor:
    """dw[small_c, big_c, k, k] += wgrad, or, with `overwrite` / a fresh buffer (`dw` None), dw = wgrad
    (VG_WGRAD_OVERWRITE: the buffer need not be initialised).  `dst_zero`: the caller knows dw is all zeros
    (VG_WGRAD_DST_ZERO)."""
    if dw is None:
        dw = torch.empty((g.small_c, g.big_c_valid or g.big_c, g.kernel, g.kernel), dtype=torch.float32,
                         device=small.device)
        overwrite = True
    nbytes = _lib.load().vg_conv_wgrad_workspace_bytes(ctypes.byref(g), _DT[small.dtype])
    ws = _ws(nbytes, small.device) if nbytes else None
    call("vg_conv_wgrad_ex", ctypes.byref(g), _DT[small.dtype], _p(small), _p(big), _p(dw), _p(ws), nbytes,
         _lib.WGRAD_OVERWRITE if overwrite else (_lib.WGRAD_DST_ZERO if dst_zero else 0), _stream(),
         flops=conv_flops(g), tag=_conv_tag(g, "wgrad"), nbytes=conv_bytes(g))
    if ws is not None and WgradOverlap.streams:
        WgradOverlap.keepalive.append(ws)
    return dw


def _ws(nbytes: int, device) -> torch.Tensor:
    return torch.empty(((nbytes + 3) // 4,), dtype=torch.float32, device=device)


def bn_train_fwd(x: torch.Tensor, gamma, beta, running_mean, running_var, num_batches_tracked, momentum: float,
                 eps: float):
    C = x.shape[-1]
    rows = x.numel() // C
    stats = torch.empty((4, C), dtype=torch.float32, device=x.device)  # mean, rstd, scale, shift
    nbytes = _lib.load().vg_reduce_workspace_bytes(rows, C)
    ws = _ws(nbytes, x.device)
    call("vg_bn_train_fwd", _p(x), _DT[x.dtype], rows, C, _p(gamma), _p(beta), _p(running_mean), _p(running_var),
         _p(num_batches_tracked), float(momentum), float(eps), _p(stats[0]), _p(stats[1]), _p(stats[2]), _p(stats[3]),
         _p(ws), nbytes, _stream())
    return stats


def bn_act_train_fwd(x: torch.Tensor, gamma, beta, running_mean, running_var, num_batches_tracked, momentum: float,
                     eps: float, act: int, slope: float, out=None):
    """Fused training BatchNorm + activation (one cooperative launch).  Returns (y, stats[4, C]).
    Measured slower than the two-kernel path at the VAE-GAN's sizes (round 1: 63 vs 32 us average forward, 113 vs 65
    backward - few co-resident blocks, long per-thread chains), so ConvLayerFn does not use it yet."""
    C = x.shape[-1]
    rows = x.numel() // C
    stats = torch.empty((4, C), dtype=torch.float32, device=x.device)
    y = out if out is not None else torch.empty_like(x)
    call("vg_bn_act_train_fwd", _p(x), _DT[x.dtype], rows, C, _p(gamma), _p(beta), _p(running_mean), _p(running_var),
         _p(num_batches_tracked), float(momentum), float(eps), act, float(slope), _p(stats), _p(y), _stream())
    return y, stats


def bn_act_train_bwd(dy: torch.Tensor, x: torch.Tensor, stats: torch.Tensor, act: int, slope: float, dgamma, dbeta,
                     out=None):
    C = x.shape[-1]
    rows = x.numel() // C
    dx = out if out is not None else torch.empty_like(x)
    nbytes = _lib.load().vg_bn_bwd_workspace_bytes(rows, C)
    ws = _ws(nbytes, x.device)
    call("vg_bn_act_train_bwd", _p(dy), _p(x), _DT[x.dtype], rows, C, _p(stats), act, float(slope), _p(dgamma),
         _p(dbeta), _p(dx), _p(ws), nbytes, _stream())
    return dx


class SumsArena:
    """Zero-initialised fp32 scratch for the channel sums of the fused convolution epilogues.  Slices are handed
    out sequentially.  A fused step OWNS its arena (`activate(buffer)` at the start of every step, after zeroing
    the part it used: the addresses are baked into its CUDA graph, so the buffer lives exactly as long as the step
    object and nothing else ever hands out or replaces it).  Module calls outside a step draw from a per-device
    default buffer; when that runs out a fresh zero buffer replaces it - the old one stays alive through the slices
    still referencing it."""
    FLOATS = 1 << 20
    _buf = {}
    _off = {}
    _owned = None          # [buffer, offset, high-water mark] of the step currently running

    @classmethod
    def activate(cls, buf: torch.Tensor) -> None:
        cls._owned = [buf, 0, 0]

    @classmethod
    def deactivate(cls) -> int:
        """-> floats the owner handed out (its high-water mark: what it has to re-zero next time)."""
        used = cls._owned[2] if cls._owned is not None else 0
        cls._owned = None
        return used

    @classmethod
    def take(cls, n: int, device) -> torch.Tensor:
        n = (n + 31) // 32 * 32
        if cls._owned is not None:
            buf, off, _ = cls._owned
            if off + n > buf.numel():
                raise _lib.VaeganB200Error(f"fused-epilogue sums arena exhausted ({buf.numel()} floats)")
            cls._owned[1] = off + n
            cls._owned[2] = max(cls._owned[2], off + n)
            return buf[off:off + n]
        key = torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device()
        buf = cls._buf.get(key)
        if buf is None or cls._off[key] + n > buf.numel():
            buf = torch.zeros(max(cls.FLOATS, n), dtype=torch.float32, device=device)
            cls._buf[key], cls._off[key] = buf, 0
        off = cls._off[key]
        cls._off[key] = off + n
        return buf[off:off + n]

    @classmethod
    def reset(cls, device) -> None:
        key = torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device()
        buf = cls._buf.get(key)
        if buf is None:
            cls._buf[key] = torch.zeros(cls.FLOATS, dtype=torch.float32, device=device)
        elif cls._off[key]:
            buf.zero_()
        cls._off[key] = 0


def bn_apply_from_sums(x: torch.Tensor, sums: torch.Tensor, groups: int, gamma, beta, running_mean, running_var,
                       num_batches_tracked, momentum: float, eps: float, act: int, slope: float):
    """Finalise the BatchNorm statistics from the raw sums of a VG_EPI_BN_STATS epilogue and apply
    y = act(BN(x)) in the same launch.  Returns (y, stats[groups, 4, C])."""
    C = x.shape[-1]
    rows = x.numel() // C // groups
    stats = torch.empty((groups, 4, C), dtype=torch.float32, device=x.device)
    y = torch.empty_like(x)
    call("vg_bn_apply_from_sums", _p(x), _DT[x.dtype], rows, C, groups, _p(sums), _p(gamma), _p(beta), _p(running_mean),
         _p(running_var), _p(num_batches_tracked), float(momentum), float(eps), act, float(slope), _p(stats), _p(y),
         _stream())
    return y, stats


def bn_bwd_apply_from_sums(dz: torch.Tensor, x: torch.Tensor, stats: torch.Tensor, sums: torch.Tensor, groups: int,
                           dgamma, dbeta) -> torch.Tensor:
    """dx of BatchNorm from dz (activation derivative already applied) and the raw sums of a VG_EPI_BN_BWD epilogue."""
    C = x.shape[-1]
    rows = x.numel() // C // groups
    dx = torch.empty_like(x)
    call("vg_bn_bwd_apply_from_sums", _p(dz), _p(x), _DT[x.dtype], rows, C, groups, _p(stats), _p(sums), _p(dgamma),
         _p(dbeta), _p(dx), _stream())
    return dx


def bn_eval_coeffs(gamma, beta, running_mean, running_var, eps: float):
    C = running_mean.numel()
    stats = torch.zeros((4, C), dtype=torch.float32, device=running_mean.device)
    call("vg_bn_eval_coeffs", _p(gamma), _p(beta), _p(running_mean), _p(running_var), float(eps), C, _p(stats[2]),
         _p(stats[3]), _stream())
    return stats


def scale_shift_act(x: torch.Tensor, scale, shift, act: int, slope: float, out_dtype=None, out=None) -> torch.Tensor:
    out_dtype = out_dtype or (out.dtype if out is not None else x.dtype)
    C = x.shape[-1]
    y = out if out is not None else torch.empty(x.shape, dtype=out_dtype, device=x.device)
    call("vg_scale_shift_act", _p(x), _DT[x.dtype], x.numel() // C, C, _p(scale), _p(shift), act, float(slope), _p(y),
         _DT[out_dtype], _stream())
    return y


def bn_act_bwd(dy: torch.Tensor, x: torch.Tensor, stats: torch.Tensor, act: int, slope: float, dgamma, dbeta,
               out=None):
    C = x.shape[-1]
    rows = x.numel() // C
    dx = out if out is not None else torch.empty_like(x)
    nbytes = _lib.load().vg_bn_bwd_workspace_bytes(rows, C)
    ws = _ws(nbytes, x.device)
    call("vg_bn_act_bwd", _p(dy), _p(x), _DT[x.dtype], rows, C, _p(stats[2]), _p(stats[3]), _p(stats[0]), _p(stats[1]),
         act, float(slope), _p(dgamma), _p(dbeta), _p(dx), _p(ws), nbytes, _stream())
    return dx


def act_bwd(dy: torch.Tensor, x: torch.Tensor, act: int, slope: float, out_dtype=None) -> torch.Tensor:
    out_dtype = out_dtype or x.dtype
    dx = torch.empty(x.shape, dtype=out_dtype, device=x.device)
    call("vg_act_bwd", _p(dy), _p(x), _DT[x.dtype], x.numel(), act, float(slope), _p(dx), _DT[out_dtype], _stream())
    return dx


def colsum(x: torch.Tensor, out: torch.Tensor) -> None:
    C = x.shape[-1]
    rows = x.numel() // C
    nbytes = _lib.load().vg_reduce_workspace_bytes(rows, C)
    ws = _ws(nbytes, x.device)
    call("vg_colsum", _p(x), _DT[x.dtype], rows, C, _p(out), _p(ws), nbytes, _stream())


def image_nhwc_shape(B: int, C: int, H: int, W: int, dtype, s2d_origin: Optional[int] = None):
    """Shape / dtype of what nchw_to_nhwc produces for a [B, C, H, W] image (to pre-allocate its `out`)."""
    if s2d_origin is not None:
        o = int(s2d_origin)
        return (B, H // 2 + o, W // 2 + o, 64), torch.bfloat16
    return (B, H, W, padded_channels(C, dtype)), dtype


def nchw_to_nhwc(src: torch.Tensor, dtype, aux: Optional[torch.Tensor] = None, mode: int = 0, sigma: float = 0.0,
                 clamp: bool = False, out: Optional[torch.Tensor] = None, s2d_origin: Optional[int] = None) -> torch.Tensor:
    """fp32 NCHW -> internal NHWC (channel-padded per `padded_channels`), or, with `s2d_origin` in {0, 1}, the
    space-to-depth image form [B, H/2+o, W/2+o, 64] (include/vaegan_b200.h, vg_nchw_to_s2d)."""
    B, C, H, W = src.shape
    if s2d_origin is not None:
        o = int(s2d_origin)
        dst = out if out is not None else torch.empty((B, H // 2 + o, W // 2 + o, 64), dtype=torch.bfloat16,
                                                      device=src.device)
        call("vg_nchw_to_s2d", _p(src), _p(aux), _p(dst), B, C, H, W, o, mode, float(sigma), int(clamp), _stream())
        return dst
    Cd = padded_channels(C, dtype)
    dst = out if out is not None else torch.empty((B, H, W, Cd), dtype=dtype, device=src.device)
    call("vg_nchw_to_nhwc", _p(src), _p(aux), _p(dst), _DT[dtype], B, C, H, W, Cd, mode, float(sigma), int(clamp),
         _stream())
    return dst


def nhwc_to_nchw(src: torch.Tensor, act: int = ACT_NONE, slope: float = 0.0, channels: Optional[int] = None,
                 s2d_origin: Optional[int] = None) -> torch.Tensor:
    """internal NHWC -> fp32 NCHW, keeping the first `channels` channels of a padded tensor; `s2d_origin` marks a
    space-to-depth image tensor (vg_s2d_to_nchw)."""
    B, H, W, Cs = src.shape
    if s2d_origin is not None:
        o = int(s2d_origin)
        H, W = 2 * (H - o), 2 * (W - o)
        dst = torch.empty((B, channels, H, W), dtype=torch.float32, device=src.device)
        call("vg_s2d_to_nchw", _p(src), _p(dst), B, channels, H, W, o, act, float(slope), _stream())
        return dst
    C = channels or Cs
    dst = torch.empty((B, C, H, W), dtype=torch.float32, device=src.device)
    call("vg_nhwc_to_nchw", _p(src), _DT[src.dtype], Cs, _p(dst), B, C, H, W, act, float(slope), _stream())
    return dst


# --------------------------------------------------------------------------------------------- side stream for wgrad
class WgradOverlap:
    """Weight gradients are off the backward critical path (BN-backward -> dgrad -> BN-backward ...), and they are
    tensor/L2-bound while the BatchNorm passes are HBM-bound.  The fused step therefore issues every wgrad on side
    streams (parallel branches of the captured CUDA graph), round-robin, and joins them before the optimizer: the
    many small weight-gradient launches of the encoder / discriminator tail then overlap EACH OTHER as well.
    Tensors the side streams read are kept alive until the join, so the caching allocator cannot hand their memory
    out early."""
    streams: list = []
    keepalive: list = []
    _next = 0

    @classmethod
    def enable(cls, streams):
        cls.streams = list(streams) if isinstance(streams, (list, tuple)) else [streams]
        cls.keepalive, cls._next = [], 0

    @classmethod
    def pick(cls):
        """Next side stream (None when overlap is off)."""
        if not cls.streams:
            return None
        s = cls.streams[cls._next % len(cls.streams)]
        cls._next += 1
        return s

    @classmethod
    def join(cls):
        for s in cls.streams:
            torch.cuda.current_stream().wait_stream(s)
        cls.keepalive = []

    @classmethod
    def disable(cls):
        cls.join()
        cls.streams = []


class GradReady:
    """Data parallelism hook: the fused step registers, per parameter, a callable that is invoked once the kernels
    producing that parameter's gradient have all been ISSUED (end of its layer's backward; weight gradients may still
    be running on a WgradOverlap side stream - the callee waits for those streams).  dp.BucketedAllReduce.mark_ready
    is what gets registered: a bucket's all-reduce starts under the rest of the backward pass."""
    handlers: dict = {}

    @classmethod
    def set(cls, handlers: dict):
        cls.handlers = handlers

    @classmethod
    def clear(cls):
        cls.handlers = {}

    @classmethod
    def notify(cls, *params):
        if not cls.handlers:
            return
        for p in params:
            if p is not None:
                h = cls.handlers.get(id(p))
                if h is not None:
                    h()


# --------------------------------------------------------------------------------------------- space-to-depth layers
def is_image_s2d(t: torch.Tensor, true_channels: int) -> bool:
    """An image-side activation (true_channels < 16) is either channel-padded to 16 or in space-to-depth form (64)."""
    return true_channels < 16 and t.dtype == torch.bfloat16 and t.shape[-1] == 64


class S2DWeightMap:
    """The three image-side convolutions of the reference rewritten as 64-channel convolutions over space-to-depth
    image tensors, so that their TMA rows are 128 bytes instead of one 32-byte padded pixel:

      Conv2d(C, n, 4, 2, p) (main_vae.py:37 first ConvBlock p=0; gan_code.py:59 p=1)
          == Conv2d(64, n, 2, 1, 0) over the s2d input with origin p:   weq[n][(sy,sx,c)][by][bx] = w[n][c][2by+sy][2bx+sx]
      ConvTranspose2d(m, C, 3, 1, 1) (gan_code.py:49)
          == Conv2d(m, 64, 4, 2, 1) whose OUTPUT is the s2d image (origin 0):
                                                    weq[(sy,sx,c)][m][KY][KX] = w[m][c][sy+2-KY][sx+2-KX]  (0 outside)

    The masters keep the reference layout; `materialize` gathers the equivalent weights (one tiny kernel) and
    `scatter` folds the equivalent weight gradient back (each master element is the sum of 1 resp. 4 terms)."""

    @staticmethod
    def eligible(spec: "ConvSpec") -> bool:
        if spec.big_c >= 16:
            return False
        if spec.kind == "down":
            return spec.kernel == 4 and spec.stride == 2 and spec.pad in (0, 1)
        return spec.kernel == 3 and spec.stride == 1 and spec.pad == 1

    def __init__(self, spec: "ConvSpec"):
        import numpy as np
        self.spec = spec
        C = spec.big_c
        if spec.kind == "down":
            n = spec.small_c
            self.eq_spec = ConvSpec("down", n, 64, 2, 1, 0)
            self.origin = spec.pad
            fwd = np.full((n, 64, 2, 2), -1, dtype=np.int32)
            bwd = np.full((n, C, 4, 4, 1), -1, dtype=np.int32)
            master = np.arange(n * C * 16, dtype=np.int32).reshape(n, C, 4, 4)
            eq = np.arange(n * 64 * 4, dtype=np.int32).reshape(n, 64, 2, 2)
            for ky in range(4):
                for kx in range(4):
                    by, sy, bx, sx = ky >> 1, ky & 1, kx >> 1, kx & 1
                    slot = (sy * 2 + sx) * 16
                    fwd[:, slot:slot + C, by, bx] = master[:, :, ky, kx]
                    bwd[:, :, ky, kx, 0] = eq[:, slot:slot + C, by, bx]
        else:
            m = spec.small_c
            self.eq_spec = ConvSpec("down", 64, m, 4, 2, 1)
            self.origin = 0
            fwd = np.full((64, m, 4, 4), -1, dtype=np.int32)
            bwd = np.full((m, C, 3, 3, 4), -1, dtype=np.int32)
            master = np.arange(m * C * 9, dtype=np.int32).reshape(m, C, 3, 3)
            eq = np.arange(64 * m * 16, dtype=np.int32).reshape(64, m, 4, 4)
            for sy in range(2):
                for sx in range(2):
                    slot = (sy * 2 + sx) * 16
                    for KY in range(4):
                        for KX in range(4):
                            ky, kx = sy + 2 - KY, sx + 2 - KX
                            if 0 <= ky <= 2 and 0 <= kx <= 2:
                                fwd[slot:slot + C, :, KY, KX] = master[:, :, ky, kx].T
                                bwd[:, :, ky, kx, sy * 2 + sx] = eq[slot:slot + C, :, KY, KX].T
        self.fan = bwd.shape[-1]
        self._fwd_np, self._bwd_np = fwd.reshape(-1), bwd.reshape(-1)
        self._dev = {}

    def _tensors(self, device):
        t = self._dev.get(device)
        if t is None:
            e = self.eq_spec
            t = dict(fwd=torch.from_numpy(self._fwd_np).to(device), bwd=torch.from_numpy(self._bwd_np).to(device),
                     weq=torch.empty((e.small_c, e.big_c, e.kernel, e.kernel), dtype=torch.float32, device=device),
                     dweq=torch.empty((e.small_c, e.big_c, e.kernel, e.kernel), dtype=torch.float32, device=device))
            self._dev[device] = t
        return t

    def materialize(self, master: torch.Tensor) -> torch.Tensor:
        t = self._tensors(master.device)
        call("vg_gather_f32", _p(t["weq"]), _p(_contig(master)), _p(t["fwd"]), t["weq"].numel(), 1, 0, _stream())
        return t["weq"]

    def grad_buffer(self, device, zero: bool = True) -> torch.Tensor:
        t = self._tensors(device)
        if zero:
            t["dweq"].zero_()
        return t["dweq"]

    def scatter(self, dweq: torch.Tensor, target: torch.Tensor) -> None:
        """target (master layout, fp32, contiguous) += the equivalent-weight gradient folded back."""
        t = self._tensors(dweq.device)
        call("vg_gather_f32", _p(target), _p(dweq), _p(t["bwd"]), target.numel(), self.fan, 1, _stream())


class LinearGemmMap:
    """nn.Linear over a flattened feature map (fc_mu / fc_logvar, main_vae.py:47-48,53) as ONE dense GEMM
    [B, h*w*C] x [h*w*C, n_pad] over the NHWC activation, for the cases the full-extent-convolution form does not put
    on the tensor cores: more than 64 taps (the reference's own 256x256 encoder flattens 14x14x256) or a latent size
    that is not a multiple of 32 (its default 100).  Same interface as S2DWeightMap: the master keeps the reference
    layout [n, C*h*w]; `materialize` builds the GEMM operand ((h, w, c) order, zero rows up to n_pad), `scatter` folds
    the gradient back."""

    @staticmethod
    def needed(spec: "ConvSpec") -> bool:
        return spec.kernel * spec.kernel > 64 or spec.small_c % 32 != 0

    def __init__(self, spec: "ConvSpec"):
        self.spec = spec
        self.n, self.C, self.kk = spec.small_c, spec.big_c, spec.kernel * spec.kernel
        self.n_pad = (self.n + 63) // 64 * 64
        self.eq_spec = ConvSpec("down", self.n_pad, self.C * self.kk, 1, 1, 0)
        self.origin = 0
        self._dev = {}

    def _tensors(self, device):
        t = self._dev.get(device)
        if t is None:
            shape = (self.n_pad, self.C * self.kk, 1, 1)
            t = dict(weq=torch.empty(shape, dtype=torch.float32, device=device),
                     dweq=torch.empty(shape, dtype=torch.float32, device=device))
            self._dev[device] = t
        return t

    def materialize(self, master: torch.Tensor) -> torch.Tensor:
        t = self._tensors(master.device)
        call("vg_linear_permute", _p(_contig(master)), _p(t["weq"]), self.n, self.n_pad, self.C, self.kk, 0, _stream())
        return t["weq"]

    def grad_buffer(self, device, zero: bool = True) -> torch.Tensor:
        t = self._tensors(device)
        if zero:
            t["dweq"].zero_()
        return t["dweq"]

    def scatter(self, dweq: torch.Tensor, target: torch.Tensor) -> None:
        call("vg_linear_permute", _p(dweq), _p(target), self.n, self.n_pad, self.C, self.kk, 1, _stream())


class PadRowsMap:
    """A convolution whose small-side channel count is not a multiple of 16 (the generator's first
    ConvTranspose2d(nz, ...) with the reference's default latent size 100, gan_code.py:19 / vaegan_code.py:26) run
    with zero-padded small-side channels: equivalent weights = the master plus zero rows, the activation on that side is
    padded by the caller.  Same interface as S2DWeightMap."""

    @staticmethod
    def needed(spec: "ConvSpec") -> bool:
        return spec.small_c % 16 != 0 and spec.small_c > 16

    def __init__(self, spec: "ConvSpec"):
        self.spec = spec
        self.n, self.row = spec.small_c, spec.big_c * spec.kernel * spec.kernel
        self.n_pad = (self.n + 63) // 64 * 64
        self.eq_spec = ConvSpec(spec.kind, self.n_pad, spec.big_c, spec.kernel, spec.stride, spec.pad)
        self.origin = 0
        self._dev = {}

    def _tensors(self, device):
        t = self._dev.get(device)
        if t is None:
            e = self.eq_spec
            shape = (e.small_c, e.big_c, e.kernel, e.kernel)
            t = dict(weq=torch.empty(shape, dtype=torch.float32, device=device),
                     dweq=torch.empty(shape, dtype=torch.float32, device=device))
            self._dev[device] = t
        return t

    def materialize(self, master: torch.Tensor) -> torch.Tensor:
        t = self._tensors(master.device)      # (kk = 1: a row-wise copy with zero rows appended)
        call("vg_linear_permute", _p(_contig(master)), _p(t["weq"]), self.n, self.n_pad, self.row, 1, 0, _stream())
        return t["weq"]

    def grad_buffer(self, device, zero: bool = True) -> torch.Tensor:
        t = self._tensors(device)
        if zero:
            t["dweq"].zero_()
        return t["dweq"]

    def scatter(self, dweq: torch.Tensor, target: torch.Tensor) -> None:
        call("vg_linear_permute", _p(dweq), _p(target), self.n, self.n_pad, self.row, 1, 1, _stream())


# --------------------------------------------------------------------------------------------- weight cache
class PackedWeights:
    """bf16 K-major copies of one fp32 master weight, refreshed when the master changes.  L1 (drop-in modules):
    staleness is detected through the parameter's autograd version counter (optimizer.step() bumps it);
    L2 (fused step) invalidates explicitly after its own Adam kernel."""

    def __init__(self):
        self.version = None
        self.key = None
        self.wd = None
        self.wu = None

    def invalidate(self):
        self.version = None

    def stage(self, w: torch.Tensor, small_c: int, big_c: int, kernel: int):
        """Make sure persistent wd / wu buffers of the right shape exist and mark them current for `w` (the caller
        is about to fill them through vg_pack_weights_multi)."""
        kk = kernel * kernel
        if self.wd is None or self.wd.shape != (kk, small_c, big_c) or self.wd.device != w.device:
            self.wd = torch.empty((kk, small_c, big_c), dtype=torch.bfloat16, device=w.device)
            self.wu = torch.empty((kk, big_c, small_c), dtype=torch.bfloat16, device=w.device)
        self.version, self.key = w._version, (w.data_ptr(), small_c, big_c, kernel)

    def get(self, w: torch.Tensor, g: VgConvGeom, wmap: Optional[S2DWeightMap] = None):
        """`w` is the fp32 master; with `wmap` the copies are packed from its equivalent space-to-depth weights."""
        key = (w.data_ptr(), g.small_c, g.big_c, g.kernel)
        if self.version != w._version or self.key != key:
            src = wmap.materialize(w.detach()) if wmap is not None else w.detach()
            kk = g.kernel * g.kernel
            if self.wd is not None and self.wd.shape == (kk, g.small_c, g.big_c) and self.wd.device == w.device:
                # refresh IN PLACE: a captured CUDA graph (fused step) may hold these addresses
                call("vg_pack_weights_bf16", ctypes.byref(g), _p(src), _p(self.wd), _p(self.wu), _stream())
            else:
                self.wd, self.wu = pack_weights(src, g)
            self.version, self.key = w._version, key
        return self.wd, self.wu


def pack_layers(layers, dtype) -> None:
    """Refresh the bf16 copies of every layer in `layers` (objects with .spec, .conv, .cache) with ONE launch."""
    if dtype != torch.bfloat16 or not layers:
        return
    items = (_lib.VgPackItem * len(layers))()
    for i, layer in enumerate(layers):
        sp, w = layer.spec, layer.conv.weight
        wmap = getattr(layer, "active_wmap", None)      # the equivalent-weights form the layer last ran in, if any
        src = w
        if wmap is not None:                    # image layer running in space-to-depth form: pack its equivalent weights
            sp, src = wmap.eq_spec, wmap.materialize(w.detach())
        bc = padded_channels(sp.big_c, dtype)
        layer.cache.stage(w, sp.small_c, bc, sp.kernel)
        items[i] = _lib.VgPackItem(src.data_ptr(), layer.cache.wd.data_ptr(), layer.cache.wu.data_ptr(), sp.small_c, bc,
                                   sp.big_c if bc != sp.big_c else 0, sp.kernel * sp.kernel)
    call("vg_pack_weights_multi", items, len(layers), _stream())


# --------------------------------------------------------------------------------------------- autograd glue
def _accumulate_or_return(param: Optional[torch.Tensor], grad: Optional[torch.Tensor]):
    """Megatron-style fused gradient accumulation: a parameter carrying `.main_grad` receives its gradient there
    (the kernels already accumulated into it) and autograd sees None."""
    return None if (param is not None and getattr(param, "main_grad", None) is not None) else grad


class LayerLink:
    """Hand-shake between two CONSECUTIVE layers of a sequential chain (the output of the first feeds the second and
    nothing else).  The producer records what its backward needs (raw conv output, BatchNorm statistics,
    activation); the consumer's dgrad then applies the activation derivative and accumulates the BatchNorm-backward
    sums in its own epilogue (VG_EPI_BN_BWD / VG_EPI_ACT_BWD), and the producer's backward starts from dz."""
    __slots__ = ("raw", "stats", "act", "slope", "groups", "sums", "fused")

    def __init__(self):
        self.raw = self.stats = self.sums = None
        self.act, self.slope, self.groups, self.fused = ACT_NONE, 0.0, 1, False


class ConvLayerFn(torch.autograd.Function):
    """conv / convT (+bias) -> [BatchNorm2d, training or eval] -> activation, on NHWC tensors.

    Replaces, per layer, ConvBlock.forward (main_vae.py:27-31) and the (ConvT|Conv, BN, ReLU|LeakyReLU) triples of
    Generator.main / Discriminator.main (gan_code.py:19-51, 59-86) together with their autograd backward."""

    @staticmethod
    def forward(ctx, x, weight, bias, gamma, beta, spec: ConvSpec, act: int, slope: float, bn, training: bool,
                cache: PackedWeights, out_f32: bool, groups: int = 1, link_in: Optional[LayerLink] = None,
                link_out: Optional[LayerLink] = None, wmap: Optional[S2DWeightMap] = None, no_grad: bool = False):
        """`wmap`: the layer runs in space-to-depth form - `spec` is then the EQUIVALENT convolution (S2DWeightMap),
        `weight` still the reference-layout master."""
        """`groups` > 1: the batch holds that many independent sub-batches (e.g. the discriminator's real and fake
        batches of vaegan_code.py:96-97 run through ONE convolution launch); BatchNorm statistics, running-stat
        updates and the BN backward stay per sub-batch, in order, exactly as separate forward calls would."""
        _require_cuda(x, "ConvLayerFn")
        x = _contig(x)
        B, H, W, Cx = x.shape
        if spec.kind == "down":
            if Cx != padded_channels(spec.big_c, x.dtype):
                raise RuntimeError(f"Given groups=1, weight of size {list(weight.shape)}, expected input with "
                                   f"{spec.big_c} channels, but got {Cx} channels instead")
            g = spec.geom(B, H, W, Cx)
        else:
            if Cx != spec.small_c:
                raise RuntimeError(f"Given transposed=1, weight of size {list(weight.shape)}, expected input with "
                                   f"{spec.small_c} channels, but got {Cx} channels instead")
            g = spec.geom(B, H, W, padded_channels(spec.big_c, x.dtype))
        if wmap is not None:
            o, e = wmap.spec, wmap.eq_spec
            g.algo_scale = (o.small_c * o.big_c * o.kernel ** 2 * (4 if o.kind == "up" else 1)) / \
                           (e.small_c * e.big_c * e.kernel ** 2)
        if x.dtype == torch.bfloat16:
            wd, wu = cache.get(weight, g, wmap)
            w_fwd = wd if spec.kind == "down" else wu
        else:
            w_fwd = _contig(weight.detach())
        # training BatchNorm on the tensor-core path: the statistics ride the convolution's epilogue
        ep = sums = None
        if bn is not None and training and x.dtype == torch.bfloat16 and not out_f32:
            if B % groups:
                raise RuntimeError(f"batch {B} is not divisible into {groups} sub-batches")
            C = spec.small_c if spec.kind == "down" else g.big_c
            sums = SumsArena.take(groups * 2 * C, x.device)
            ep = make_epilogue(EPI_BN_STATS, groups, C, sums=sums)
            if not epilogue_supported(g, spec.kind == "up", ep):
                ep = sums = None
        # ReLU / LeakyReLU without BatchNorm: applied in the epilogue; only the activated tensor is kept (its sign is
        # all the backward needs)
        act_in_epilogue = False
        if bn is None and act in (ACT_RELU, ACT_LEAKY) and x.dtype == torch.bfloat16 and not out_f32:
            ep_act = make_epilogue(EPI_ACT_FWD, 1, 0, act, slope)
            if epilogue_supported(g, spec.kind == "up", ep_act):
                ep, act_in_epilogue = ep_act, True
        # eval-mode BatchNorm (+ activation) of a pass that needs no gradients (generation / validation under
        # torch.no_grad(), main_vae.py:348-374, vaegan_code.py:147-171; `no_grad` is the CALLER's grad mode - inside
        # forward() it is always off and needs_input_grad ignores it): scale / shift are known before the launch and
        # ride the epilogue - act(conv * scale + shift) straight from the fp32 accumulator, no separate pass
        eval_stats, eval_fused = None, False
        if (bn is not None and not training and x.dtype == torch.bfloat16 and not out_f32
                and act in (ACT_NONE, ACT_RELU, ACT_LEAKY) and no_grad):
            C = spec.small_c if spec.kind == "down" else g.big_c
            eval_stats = bn_eval_coeffs(gamma.detach(), beta.detach(), bn.running_mean, bn.running_var, bn.eps)
            ep_aff = make_epilogue(EPI_AFFINE_ACT_FWD, 1, C, act, slope, stats=eval_stats)
            if epilogue_supported(g, spec.kind == "up", ep_aff):
                ep, eval_fused = ep_aff, True
        bias_k = bias.detach() if bias is not None else None
        if bias_k is not None and bias_k.numel() < spec.small_c:           # padded GEMM rows (LinearGemmMap)
            padded = torch.zeros(spec.small_c, dtype=torch.float32, device=x.device)
            padded[:bias_k.numel()].copy_(bias_k)
            bias_k = padded
        if spec.kind == "down":
            raw = conv_down(x, w_fwd, g, bias_k, out_f32=out_f32, ep=ep)
        else:
            raw = conv_up(x, w_fwd, g, ep=ep)
        stats = None
        if bn is not None:
            if training:
                rm, rv, nbt = (bn.running_mean, bn.running_var, bn.num_batches_tracked) if bn.track_running_stats \
                    else (None, None, None)
                # F.batch_norm semantics: momentum=None means cumulative average - the reference never uses it
                if ep is not None:
                    y, stats = bn_apply_from_sums(raw, sums, groups, gamma.detach(), beta.detach(), rm, rv, nbt,
                                                  bn.momentum, bn.eps, act, slope)
                    if groups == 1:
                        stats = stats[0]
                elif groups == 1:
                    stats = bn_train_fwd(raw, gamma.detach(), beta.detach(), rm, rv, nbt, bn.momentum, bn.eps)
                    y = scale_shift_act(raw, stats[2], stats[3], act, slope)
                else:
                    if B % groups:
                        raise RuntimeError(f"batch {B} is not divisible into {groups} sub-batches")
                    rg, y = raw.view(groups, -1, raw.shape[-1]), torch.empty_like(raw)
                    yg = y.view(groups, -1, raw.shape[-1])
                    per = [bn_train_fwd(rg[i], gamma.detach(), beta.detach(), rm, rv, nbt, bn.momentum, bn.eps)
                           for i in range(groups)]
                    for i in range(groups):
                        scale_shift_act(rg[i], per[i][2], per[i][3], act, slope, out=yg[i])
                    stats = torch.stack(per)
            elif eval_fused:
                stats, y = eval_stats, raw
            else:
                stats = eval_stats if eval_stats is not None else \
                    bn_eval_coeffs(gamma.detach(), beta.detach(), bn.running_mean, bn.running_var, bn.eps)
                y = scale_shift_act(raw, stats[2], stats[3], act, slope)
        elif act != ACT_NONE and not act_in_epilogue:
            y = scale_shift_act(raw, None, None, act, slope)
        else:
            y = raw
        ctx.spec, ctx.act, ctx.slope, ctx.g = spec, act, slope, g
        ctx.has_bn, ctx.bn_training, ctx.cache, ctx.out_f32 = bn is not None, training, cache, out_f32
        ctx.groups = groups
        ctx.params = (weight, bias, gamma, beta)
        ctx.link_in, ctx.link_out, ctx.wmap = link_in, None, wmap
        if (link_out is not None and x.dtype == torch.bfloat16 and not out_f32 and act in (ACT_NONE, ACT_RELU, ACT_LEAKY)
                and (bn is None or training) and (bn is not None or act != ACT_NONE)):
            # (a detached alias: when `raw` is also this function's OUTPUT it acquires grad_fn -> ctx -> link -> raw,
            # a reference cycle that would keep a whole step's tensors alive until the garbage collector runs)
            link_out.raw, link_out.stats, link_out.act, link_out.slope, link_out.groups = raw.detach(), stats, act, slope, groups
            ctx.link_out = link_out
        ctx.save_for_backward(x, raw, stats)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, raw, stats = ctx.saved_tensors
        weight, bias, gamma, beta = ctx.params
        spec, g, act, slope = ctx.spec, ctx.g, ctx.act, ctx.slope
        dy = _contig(dy)
        need_x, need_w = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        dgamma = dbeta = dbias = dw = dx = None

        # ---- through activation (+ BatchNorm) to the gradient of the raw conv output
        link_out = ctx.link_out
        if link_out is not None and link_out.fused:
            # the next layer's dgrad epilogue already produced dz = dy * act'(.) and the two channel sums
            if ctx.has_bn:
                if ctx.needs_input_grad[3]:
                    dgamma = getattr(gamma, "main_grad", None)
                    dbeta = getattr(beta, "main_grad", None)
                    if dgamma is None:
                        dgamma = torch.zeros_like(gamma, dtype=torch.float32)
                        dbeta = torch.zeros_like(beta, dtype=torch.float32)
                d_raw = bn_bwd_apply_from_sums(dy, raw, stats, link_out.sums, ctx.groups, dgamma, dbeta)
            else:
                d_raw = dy
        elif ctx.has_bn:
            if ctx.bn_training:
                want_affine = ctx.needs_input_grad[3]
                if want_affine:
                    dgamma = getattr(gamma, "main_grad", None)
                    dbeta = getattr(beta, "main_grad", None)
                    if dgamma is None:
                        dgamma = torch.zeros_like(gamma, dtype=torch.float32)
                        dbeta = torch.zeros_like(beta, dtype=torch.float32)
                if ctx.groups == 1:
                    d_raw = bn_act_bwd(dy, raw, stats, act, slope, dgamma, dbeta)
                else:
                    C = raw.shape[-1]
                    d_raw = torch.empty_like(raw)
                    dyg, rg, dg = dy.view(ctx.groups, -1, C), raw.view(ctx.groups, -1, C), d_raw.view(ctx.groups, -1, C)
                    for i in range(ctx.groups):
                        bn_act_bwd(dyg[i], rg[i], stats[i], act, slope, dgamma, dbeta, out=dg[i])
            else:
                raise _lib.VaeganB200Error("backward through eval-mode BatchNorm is not part of the VAE-GAN step")
        elif act != ACT_NONE:
            d_raw = act_bwd(dy, raw, act, slope, out_dtype=x.dtype)
        else:
            d_raw = dy if dy.dtype == x.dtype else scale_shift_act(dy, None, None, ACT_NONE, 0.0, out_dtype=x.dtype)

        # ---- bias, weight and input gradients
        want_bias = bias is not None and ctx.needs_input_grad[2]
        if want_bias:
            dbias = getattr(bias, "main_grad", None)
            if dbias is None:
                dbias = torch.zeros_like(bias, dtype=torch.float32)
        bias_with_wgrad = want_bias and need_w       # the bias gradient then rides the weight-gradient stream
        def bias_sum(t, target):
            if target.numel() == t.shape[-1]:
                colsum(t, target)
            else:                                   # padded GEMM rows: reduce all columns, keep the real ones
                tmp = torch.zeros(t.shape[-1], dtype=torch.float32, device=t.device)
                colsum(t, tmp)
                target.add_(tmp[:target.numel()])

        if want_bias and not bias_with_wgrad:
            bias_sum(d_raw, dbias)
        small, big = (d_raw, x) if spec.kind == "down" else (x, d_raw)
        if need_w:
            main_grad = getattr(weight, "main_grad", None)
            side = WgradOverlap.pick() if main_grad is not None else None
            wmap = ctx.wmap

            def run_wgrad():
                if bias_with_wgrad:
                    bias_sum(d_raw, dbias)
                if wmap is None:
                    # (a flat-buffer owner that zeroes before every backward pass and runs each layer once per pass
                    # marks its gradients `first_touch`: the kernel then skips reading the (zero) destination)
                    return conv_wgrad(small, big, g, main_grad,
                                      dst_zero=main_grad is not None and getattr(weight, "grad_first_touch", False))
                # space-to-depth layer: gradient of the equivalent weights, folded back into the master layout
                dweq = conv_wgrad(small, big, g, wmap.grad_buffer(small.device, zero=False), overwrite=True)
                target = main_grad if main_grad is not None else torch.zeros(weight.shape, dtype=torch.float32,
                                                                             device=small.device)
                wmap.scatter(dweq, target)
                return target

            if side is not None and main_grad is not None:
                side.wait_stream(torch.cuda.current_stream())          # d_raw (and x) are complete
                with torch.cuda.stream(side):
                    dw = run_wgrad()
                WgradOverlap.keepalive.append((small, big))
            else:
                dw = run_wgrad()
            dw = dw.view(weight.shape)
        if need_x:
            if x.dtype == torch.bfloat16:
                wd, wu = ctx.cache.get(weight, g, ctx.wmap)
                w_bwd = wu if spec.kind == "down" else wd
            else:
                w_bwd = _contig(weight.detach())
            # the producer of x left a link: fold ITS activation / BatchNorm backward reduction into this dgrad
            link_in, ep, sums = ctx.link_in, None, None
            if link_in is not None and link_in.raw is not None and x.dtype == torch.bfloat16:
                C = link_in.raw.shape[-1]
                if link_in.stats is not None:
                    sums = SumsArena.take(link_in.groups * 2 * C, x.device)
                    ep = make_epilogue(EPI_BN_BWD, link_in.groups, C, link_in.act, link_in.slope, sums, link_in.raw,
                                       link_in.stats)
                else:
                    ep = make_epilogue(EPI_ACT_BWD, 1, C, link_in.act, link_in.slope, None, link_in.raw, None)
                if not epilogue_supported(g, spec.kind == "down", ep):
                    ep = None
            dx = conv_up(d_raw, w_bwd, g, ep=ep) if spec.kind == "down" else conv_down(d_raw, w_bwd, g, ep=ep)
            if ep is not None:
                link_in.sums, link_in.fused = sums, True
        GradReady.notify(weight, bias, gamma, beta)
        return (dx, _accumulate_or_return(weight, dw), _accumulate_or_return(bias, dbias),
                _accumulate_or_return(gamma, dgamma), _accumulate_or_return(beta, dbeta),
                None, None, None, None, None, None, None, None, None, None, None, None)


class ToNHWCFn(torch.autograd.Function):
    """fp32 NCHW module input -> internal NHWC activation (and the reverse for its gradient)."""

    @staticmethod
    def forward(ctx, x, dtype, s2d_origin=None):
        _require_cuda(x, "ToNHWCFn")
        if x.dtype != torch.float32:
            raise _lib.VaeganB200Error(f"module inputs must be float32 (got {x.dtype}), like the reference's loaders")
        ctx.channels, ctx.s2d_origin = x.shape[1], s2d_origin
        return nchw_to_nhwc(_contig(x), dtype, s2d_origin=s2d_origin)

    @staticmethod
    def backward(ctx, dy):
        return nhwc_to_nchw(_contig(dy), channels=ctx.channels, s2d_origin=ctx.s2d_origin), None, None


class ToNCHWActFn(torch.autograd.Function):
    """internal NHWC activation -> fp32 NCHW module output with the final Tanh / no-op fused in."""

    @staticmethod
    def forward(ctx, x, act, channels=None):
        # an image-side tensor (< 16 true channels) with 64 slots per pixel block is in space-to-depth form, origin 0
        ctx.s2d = 0 if (channels is not None and is_image_s2d(x, channels)) else None
        y = nhwc_to_nchw(_contig(x), act, channels=channels, s2d_origin=ctx.s2d)
        ctx.act, ctx.dtype = act, x.dtype
        ctx.save_for_backward(y if act == ACT_TANH else None)
        return y

    @staticmethod
    def backward(ctx, dy):
        (y,) = ctx.saved_tensors
        dy = _contig(dy)
        if ctx.act == ACT_TANH:
            return nchw_to_nhwc(dy, ctx.dtype, aux=y, mode=2, s2d_origin=ctx.s2d), None, None
        return nchw_to_nhwc(dy, ctx.dtype, s2d_origin=ctx.s2d), None, None


class PointwiseActFn(torch.autograd.Function):
    """Stand-alone activation on an fp32 tensor (the discriminator's final Sigmoid, gan_code.py:85)."""

    @staticmethod
    def forward(ctx, x, act, slope):
        x = _contig(x)
        ctx.act, ctx.slope = act, slope
        ctx.save_for_backward(x)
        return scale_shift_act(x.view(-1, 1), None, None, act, slope).view(x.shape)

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        return act_bwd(_contig(dy), x, ctx.act, ctx.slope), None, None
