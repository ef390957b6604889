"""torch.library registration of the C-ABI contractions: `torch.ops.vaegan_b200.*`.

The drop-in modules reach the library through ctypes + autograd.Function (functional.ConvLayerFn: layer links, packed
weight caches, fused epilogues - Python state a dispatcher schema cannot carry).  This file registers the stateless
core of the same path as real custom ops, so that torch.compile / torch.export / FakeTensor tracing see opaque ops
with shape functions and autograd formulas instead of ctypes calls:

    y  = torch.ops.vaegan_b200.conv2d_nhwc(x, weight, bias, stride, padding)            # F.conv2d,  NHWC activations
    y  = torch.ops.vaegan_b200.conv_transpose2d_nhwc(x, weight, stride, padding)        # F.conv_transpose2d
    dx = torch.ops.vaegan_b200.conv_dgrad_nhwc(...)   dw = torch.ops.vaegan_b200.conv_wgrad_nhwc(...)

`weight` is the reference-layout fp32 master (Conv2d.weight[Cout,Cin,k,k] / ConvTranspose2d.weight[Cin,Cout,k,k],
main_vae.py:24, gan_code.py:21-49); activations are NHWC in bf16 (tcgen05 path) or fp32 (CUDA-core path).  Only a CUDA
implementation is registered: a CPU tensor raises NotImplementedError from the dispatcher - there is no fallback.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import functional as F_

_NS = "vaegan_b200"


def _spec(kind: str, weight: torch.Tensor, stride: int, padding: int) -> F_.ConvSpec:
    # both reference layouts are [small_c, big_c, k, k]
    return F_.ConvSpec(kind, weight.shape[0], weight.shape[1], weight.shape[2], stride, padding)


def _geom(kind: str, weight, stride, padding, batch, in_h, in_w, big_c_tensor):
    return _spec(kind, weight, stride, padding).geom(batch, in_h, in_w, big_c_tensor)


def _operands(weight: torch.Tensor, g, dtype):
    """(down operand, up operand) of the weight for this arithmetic mode."""
    if dtype == torch.bfloat16:
        return F_.pack_weights(weight.detach().contiguous(), g)
    w = weight.detach().contiguous()
    return w, w


# ------------------------------------------------------------------------------------------------- forward ops
@torch.library.custom_op(f"{_NS}::conv2d_nhwc", mutates_args=(), device_types="cuda")
def conv2d_nhwc(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], stride: int,
                padding: int) -> torch.Tensor:
    g = _geom("down", weight, stride, padding, x.shape[0], x.shape[1], x.shape[2], x.shape[3])
    wd, _ = _operands(weight, g, x.dtype)
    return F_.conv_down(x.contiguous(), wd, g, bias)


@conv2d_nhwc.register_fake
def _(x, weight, bias, stride, padding):
    oh, ow = _spec("down", weight, stride, padding).out_hw(x.shape[1], x.shape[2])
    return x.new_empty((x.shape[0], oh, ow, weight.shape[0]))


@torch.library.custom_op(f"{_NS}::conv_transpose2d_nhwc", mutates_args=(), device_types="cuda")
def conv_transpose2d_nhwc(x: torch.Tensor, weight: torch.Tensor, stride: int, padding: int) -> torch.Tensor:
    g = _geom("up", weight, stride, padding, x.shape[0], x.shape[1], x.shape[2], None)
    _, wu = _operands(weight, g, x.dtype)
    return F_.conv_up(x.contiguous(), wu, g)


@conv_transpose2d_nhwc.register_fake
def _(x, weight, stride, padding):
    oh, ow = _spec("up", weight, stride, padding).out_hw(x.shape[1], x.shape[2])
    return x.new_empty((x.shape[0], oh, ow, weight.shape[1]))


# ------------------------------------------------------------------------------------------------- gradient ops
@torch.library.custom_op(f"{_NS}::conv_dgrad_nhwc", mutates_args=(), device_types="cuda")
def conv_dgrad_nhwc(dy: torch.Tensor, weight: torch.Tensor, transposed: bool, stride: int, padding: int, in_h: int,
                    in_w: int) -> torch.Tensor:
    """Input gradient of conv2d_nhwc (transposed=False: an `up` contraction of dy) or of conv_transpose2d_nhwc
    (transposed=True: a `down` contraction); in_h / in_w = spatial size of the forward input."""
    kind = "up" if transposed else "down"
    g = _geom(kind, weight, stride, padding, dy.shape[0], in_h, in_w, None)
    wd, wu = _operands(weight, g, dy.dtype)
    return F_.conv_down(dy.contiguous(), wd, g) if transposed else F_.conv_up(dy.contiguous(), wu, g)


@conv_dgrad_nhwc.register_fake
def _(dy, weight, transposed, stride, padding, in_h, in_w):
    return dy.new_empty((dy.shape[0], in_h, in_w, weight.shape[0] if transposed else weight.shape[1]))


@torch.library.custom_op(f"{_NS}::conv_wgrad_nhwc", mutates_args=(), device_types="cuda")
def conv_wgrad_nhwc(dy: torch.Tensor, x: torch.Tensor, transposed: bool, kernel: int, stride: int,
                    padding: int) -> torch.Tensor:
    """Weight gradient in the reference layout [small_c, big_c, k, k], fp32."""
    small, big = (x, dy) if transposed else (dy, x)
    spec = F_.ConvSpec("up" if transposed else "down", small.shape[3], big.shape[3], kernel, stride, padding)
    g = spec.geom(x.shape[0], x.shape[1], x.shape[2], None)
    return F_.conv_wgrad(small.contiguous(), big.contiguous(), g)


@conv_wgrad_nhwc.register_fake
def _(dy, x, transposed, kernel, stride, padding):
    small_c, big_c = (x.shape[3], dy.shape[3]) if transposed else (dy.shape[3], x.shape[3])
    return dy.new_empty((small_c, big_c, kernel, kernel), dtype=torch.float32)


# ------------------------------------------------------------------------------------------------- autograd
def _conv_setup(ctx, inputs, output):
    x, weight, bias, stride, padding = inputs
    ctx.save_for_backward(x, weight)
    ctx.stride, ctx.padding, ctx.has_bias = stride, padding, bias is not None


def _conv_backward(ctx, dy):
    x, weight = ctx.saved_tensors
    dy = dy.contiguous()
    dx = dw = db = None
    if ctx.needs_input_grad[0]:
        dx = conv_dgrad_nhwc(dy, weight, False, ctx.stride, ctx.padding, x.shape[1], x.shape[2])
    if ctx.needs_input_grad[1]:
        dw = conv_wgrad_nhwc(dy, x, False, weight.shape[2], ctx.stride, ctx.padding)
    if ctx.has_bias and ctx.needs_input_grad[2]:
        db = dy.float().sum(dim=(0, 1, 2))
    return dx, dw, db, None, None


conv2d_nhwc.register_autograd(_conv_backward, setup_context=_conv_setup)


def _convt_setup(ctx, inputs, output):
    x, weight, stride, padding = inputs
    ctx.save_for_backward(x, weight)
    ctx.stride, ctx.padding = stride, padding


def _convt_backward(ctx, dy):
    x, weight = ctx.saved_tensors
    dy = dy.contiguous()
    dx = dw = None
    if ctx.needs_input_grad[0]:
        dx = conv_dgrad_nhwc(dy, weight, True, ctx.stride, ctx.padding, x.shape[1], x.shape[2])
    if ctx.needs_input_grad[1]:
        dw = conv_wgrad_nhwc(dy, x, True, weight.shape[2], ctx.stride, ctx.padding)
    return dx, dw, None, None


conv_transpose2d_nhwc.register_autograd(_convt_backward, setup_context=_convt_setup)
