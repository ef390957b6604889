"""Drop-in replacements for the reference's hot-path nn.Modules, backed by the sm_100a kernels.

    from vaegan_b200 import Encoder, Generator, Decoder, Discriminator, weights_init

Same constructor signatures, forward signatures, child-module structure and state_dict keys as
  main_vae.py:20-58   ConvBlock, Encoder          (cnn.{i}.conv / cnn.{i}.bn / fc_mu / fc_logvar)
  gan_code.py:16-54   Generator  (`main.{i}`)     aliased `Decoder` like main_vae.py:10
  gan_code.py:56-89   Discriminator (`main.{i}`)
  gan_code.py:91-97   weights_init
so the loop of vaegan_code.py:29-135 (construction, .apply(weights_init), .parameters() -> Adam, .train()/.eval(),
forward, loss.backward(), state_dict()/load_state_dict()) runs unchanged.  The children are real torch.nn.Conv2d /
ConvTranspose2d / BatchNorm2d / Linear objects - they hold the fp32 master parameters and buffers - but `forward`
never calls them: it walks the same layer list and dispatches fused conv(+BN)(+activation) layers to the C-ABI.

Extra, optional keyword arguments (all default to the reference behaviour):
  hw        image side for the resolution-derived variants of SURVEY.md Appendix A.1 (256 = the reference nets)
  width     encoder channel multiplier (BASELINE.json config 4)
  precision "bf16" (tcgen05 tensor-core path, fp32 accumulate) or "fp32" (CUDA-core path, reference-accurate)
"""
from __future__ import annotations

import math
from typing import List, Optional, Tuple

import torch
import torch.nn as nn

from . import functional as F_
from ._lib import ACT_LEAKY, ACT_NONE, ACT_RELU, ACT_SIGMOID, ACT_TANH

_DEFAULT_PRECISION = "bf16"


def set_default_precision(precision: str) -> None:
    global _DEFAULT_PRECISION
    if precision not in F_.PRECISION_DTYPE:
        raise ValueError(f"precision must be one of {list(F_.PRECISION_DTYPE)}")
    _DEFAULT_PRECISION = precision


def _act_of(m: nn.Module) -> Tuple[int, float]:
    if isinstance(m, nn.LeakyReLU):
        return ACT_LEAKY, float(m.negative_slope)
    if isinstance(m, nn.ReLU):
        return ACT_RELU, 0.0
    if isinstance(m, nn.Tanh):
        return ACT_TANH, 0.0
    if isinstance(m, nn.Sigmoid):
        return ACT_SIGMOID, 0.0
    raise TypeError(f"unsupported activation {type(m).__name__}")


def _spec_of(conv: nn.Module) -> F_.ConvSpec:
    k, s, p = conv.kernel_size, conv.stride, conv.padding
    if k[0] != k[1] or s[0] != s[1] or p[0] != p[1]:
        raise ValueError("only square kernels / strides / paddings occur on the VAE-GAN path")
    if isinstance(conv, nn.ConvTranspose2d):
        return F_.ConvSpec("up", conv.in_channels, conv.out_channels, k[0], s[0], p[0])
    return F_.ConvSpec("down", conv.out_channels, conv.in_channels, k[0], s[0], p[0])


class _Layer:
    """(conv, optional BatchNorm2d, activation) group of a reference nn.Sequential."""

    def __init__(self, conv, bn, act, slope):
        self.conv, self.bn, self.act, self.slope = conv, bn, act, slope
        self.spec = _spec_of(conv)
        self.cache = F_.PackedWeights()
        # image-side layers (3 channels) can run as 64-channel convolutions over space-to-depth image tensors
        self.wmap = None
        if F_.S2DWeightMap.eligible(self.spec) and (self.spec.kind == "down" or conv.bias is None):
            self.wmap = F_.S2DWeightMap(self.spec)
        # small-side channels that are not a multiple of 16 (latent size 100 into the generator's first layer)
        self.padmap = F_.PadRowsMap(self.spec) if (self.spec.kind == "up" and F_.PadRowsMap.needed(self.spec)) else None
        self.s2d_active = False

    def input_s2d_origin(self, dtype, h: int, w: int):
        """Space-to-depth origin this layer wants for an IMAGE input of h x w (None = plain channel-padded NHWC)."""
        if self.wmap is None or self.spec.kind != "down" or dtype != torch.bfloat16 or (h | w) & 1:
            return None
        return self.wmap.origin

    def __call__(self, x, training: bool, out_f32: bool = False, fuse_act: bool = True, groups: int = 1,
                 link_in=None, link_out=None):
        bn = self.bn
        act = self.act if fuse_act else ACT_NONE
        spec, wmap = self.spec, None
        if self.padmap is not None and x.dtype == torch.bfloat16:
            spec, wmap = self.padmap.eq_spec, self.padmap
            x = torch.nn.functional.pad(x, (0, self.padmap.n_pad - self.padmap.n))      # zero channels; backward slices
        if self.wmap is not None and x.dtype == torch.bfloat16:
            if spec.kind == "down":
                s2d = x.shape[-1] == 64                      # the caller converted the image with input_s2d_origin()
            else:
                s2d = not ((x.shape[1] | x.shape[2]) & 1)    # the image this layer produces has even extents
            if s2d:
                spec, wmap = self.wmap.eq_spec, self.wmap
        self.s2d_active, self.active_wmap = wmap is not None, wmap
        return F_.ConvLayerFn.apply(x, self.conv.weight, self.conv.bias, bn.weight if bn is not None else None,
                                    bn.bias if bn is not None else None, spec, act, self.slope, bn, training,
                                    self.cache, out_f32, groups, link_in, link_out, wmap, not torch.is_grad_enabled())


def _run_chain(layers, h, training: bool, groups: int = 1, link=None):
    """Run consecutive layers whose outputs feed only the next layer, handing a LayerLink from each producer to its
    consumer so that the backward reductions fuse into the consumer's dgrad.  Returns (output, link of the last layer)."""
    for layer in layers:
        nxt = F_.LayerLink() if torch.is_grad_enabled() else None
        h = layer(h, training, groups=groups, link_in=link, link_out=nxt)
        link = nxt
    return h, link


def _group_layers(seq: nn.Sequential) -> List[_Layer]:
    mods = list(seq)
    layers, i = [], 0
    while i < len(mods):
        conv = mods[i]
        if not isinstance(conv, (nn.Conv2d, nn.ConvTranspose2d)):
            raise TypeError(f"expected a convolution at main[{i}], found {type(conv).__name__}")
        i += 1
        bn = None
        if i < len(mods) and isinstance(mods[i], nn.BatchNorm2d):
            bn = mods[i]
            i += 1
        act, slope = ACT_NONE, 0.0
        if i < len(mods) and not isinstance(mods[i], (nn.Conv2d, nn.ConvTranspose2d)):
            act, slope = _act_of(mods[i])
            i += 1
        layers.append(_Layer(conv, bn, act, slope))
    return layers


class _KernelModule(nn.Module):
    precision: str

    def _dtype(self):
        return F_.PRECISION_DTYPE[self.precision]

    def repack_weights(self) -> None:
        """Refresh every layer's packed bf16 weights from the fp32 masters in one launch (fused step: after its
        own Adam kernel, which updates the masters without touching autograd's version counters)."""
        F_.pack_layers(self._layers(), self._dtype())

    def invalidate_packed_weights(self) -> None:
        """Force a re-pack of the bf16 weight copies (needed after the masters were changed behind autograd's back)."""
        for layer in self._layers():
            layer.cache.invalidate()


# --------------------------------------------------------------------------------------------- encoder
class ConvBlock(nn.Module):
    """main_vae.py:20-31 - Conv2d(k4, s2, p0, bias) -> BatchNorm2d -> LeakyReLU(0.01)."""

    def __init__(self, in_channels, out_channels, kernel_size=4, stride=2):
        super().__init__()
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size, stride)
        self.bn = nn.BatchNorm2d(out_channels)
        self.leaky_relu = nn.LeakyReLU(inplace=True)

    def _layer(self) -> _Layer:
        layer = self.__dict__.get("_vg_layer")
        if layer is None or layer.conv is not self.conv or layer.bn is not self.bn:
            layer = _Layer(self.conv, self.bn, ACT_LEAKY, float(self.leaky_relu.negative_slope))
            self.__dict__["_vg_layer"] = layer
        return layer

    def forward(self, x):
        """Stand-alone use takes / returns fp32 NCHW like the reference; Encoder chains blocks in NHWC instead."""
        dtype = F_.PRECISION_DTYPE[_DEFAULT_PRECISION]
        y = self._layer()(F_.ToNHWCFn.apply(x, dtype), self.training)
        return F_.ToNCHWActFn.apply(y, ACT_NONE, self.conv.out_channels)


class Encoder(_KernelModule):
    """main_vae.py:34-58."""

    def __init__(self, img_size, latent_dim, width: int = 1, precision: Optional[str] = None):
        super().__init__()
        self.precision = precision or _DEFAULT_PRECISION
        channels = [img_size[0]] + [c * width for c in (32, 64, 128, 256)]
        self.cnn = nn.Sequential(*[ConvBlock(a, b) for a, b in zip(channels[:-1], channels[1:])])
        # The reference sizes its heads with a TRAIN-mode forward of zeros (main_vae.py:43-45).  Its observable side
        # effects are reproduced in closed form: a zero image gives conv output == bias everywhere, so batch mean ==
        # bias, batch variance == 0 and the block output is beta == 0 again for the next block.
        h, w = int(img_size[1]), int(img_size[2])
        for blk in self.cnn:
            h, w = _spec_of(blk.conv).out_hw(h, w)
            if h * w < 2:
                raise ValueError("Expected more than 1 value per channel when training")  # same failure as the reference
            with torch.no_grad():
                m = blk.bn.momentum
                blk.bn.running_mean.mul_(1 - m).add_(blk.conv.bias.detach(), alpha=m)
                blk.bn.running_var.mul_(1 - m)
                blk.bn.num_batches_tracked.add_(1)
        self._feat_hw = (h, w)
        self.flatten_size = channels[-1] * h * w
        self.fc_mu = nn.Linear(self.flatten_size, latent_dim)
        self.fc_logvar = nn.Linear(self.flatten_size, latent_dim)
        self._heads = None

    def _layers(self):
        return [blk._layer() for blk in self.cnn] + list(self._head_layers())

    def _head_layers(self):
        """fc_mu / fc_logvar as full-extent convolutions over the NHWC feature map: the reference flattens (c,h,w)
        (main_vae.py:53), so Linear.weight[nz, c*h*w] IS a conv weight [nz, c, h, w] - no permutation needed."""
        if self._heads is None or self._heads[0].conv is not self.fc_mu or self._heads[1].conv is not self.fc_logvar:
            self._heads = (_LinearAsConv(self.fc_mu, self._feat_hw), _LinearAsConv(self.fc_logvar, self._feat_hw))
        return self._heads

    def forward(self, x):
        if x.dim() != 4:
            raise RuntimeError(f"Expected 4D (batched) input to conv2d, but got input of size: {list(x.shape)}")
        return self.forward_nhwc(F_.ToNHWCFn.apply(x, self._dtype(), self.input_s2d_origin(x.shape[2], x.shape[3])))

    def input_s2d_origin(self, h: int, w: int):
        """How forward_nhwc wants an h x w image laid out: space-to-depth origin, or None for channel-padded NHWC."""
        return self.cnn[0]._layer().input_s2d_origin(self._dtype(), h, w)

    def forward_nhwc(self, h):
        """Same as forward for an input that is already an internal NHWC activation (used by the fused step)."""
        blocks = [blk._layer() for blk in self.cnn]
        h, link = _run_chain(blocks[:-1], h, self.training)
        h = blocks[-1](h, self.training, link_in=link)          # two heads read this output: no link beyond it
        if (h.shape[1], h.shape[2]) != self._feat_hw:
            raise RuntimeError(f"mat1 and mat2 shapes cannot be multiplied ({h.shape[0]}x"
                               f"{h.shape[1] * h.shape[2] * h.shape[3]} and {self.flatten_size}x{self.fc_mu.out_features})")
        mu_head, lv_head = self._head_layers()
        mu = mu_head(h, self.training, out_f32=True)
        logvar = lv_head(h, self.training, out_f32=True)
        return mu.view(mu.shape[0], -1), logvar.view(logvar.shape[0], -1)


class _LinearAsConv(_Layer):
    def __init__(self, linear: nn.Linear, feat_hw):
        h, w = feat_hw
        if h != w:
            raise ValueError("square feature maps only")
        self.conv, self.bn, self.act, self.slope = linear, None, ACT_NONE, 0.0
        self.spec = F_.ConvSpec("down", linear.out_features, linear.in_features // (h * w), h, 1, 0)
        self.cache = F_.PackedWeights()
        # large feature maps / latent sizes that are not a multiple of 32 run as one dense GEMM (bf16 path)
        self.wmap = F_.LinearGemmMap(self.spec) if F_.LinearGemmMap.needed(self.spec) else None
        self.s2d_active = False

    def __call__(self, x, training: bool, out_f32: bool = False, fuse_act: bool = True, groups: int = 1,
                 link_in=None, link_out=None):
        self.active_wmap = self.wmap if x.dtype == torch.bfloat16 else None
        if self.wmap is None or x.dtype != torch.bfloat16:
            self.s2d_active = False
            return F_.ConvLayerFn.apply(x, self.conv.weight, self.conv.bias, None, None, self.spec, ACT_NONE, 0.0, None,
                                        training, self.cache, out_f32, 1, None, None, None, False)
        self.s2d_active = True                       # pack_layers: pack the GEMM operand, not the conv form
        flat = x.reshape(x.shape[0], 1, 1, -1)       # NHWC flatten = (h, w, c) order, what LinearGemmMap packs for
        y = F_.ConvLayerFn.apply(flat, self.conv.weight, self.conv.bias, None, None, self.wmap.eq_spec, ACT_NONE, 0.0,
                                 None, training, self.cache, out_f32, 1, None, None, self.wmap, False)
        return y[..., :self.spec.small_c]


# --------------------------------------------------------------------------------------------- generator
def _generator_channels(ngf: int, hw: int) -> List[int]:
    full = [ngf * 16, ngf * 8, ngf * 4, ngf * 2, ngf, ngf // 2, ngf // 4]   # feature-map sizes 4 .. 256
    if hw < 8 or hw > 256 or hw & (hw - 1):
        raise ValueError("hw must be a power of two in [8, 256]")
    return full[:int(math.log2(hw)) - 1]


class Generator(_KernelModule):
    """gan_code.py:16-54 (hw=256), or its resolution-aligned truncation for hw in {8..128}."""

    def __init__(self, nz=128, ngf=64, nc=3, hw: int = 256, precision: Optional[str] = None):
        super().__init__()
        self.precision = precision or _DEFAULT_PRECISION
        ch = _generator_channels(ngf, hw)
        mods: List[nn.Module] = [nn.ConvTranspose2d(nz, ch[0], 4, 1, 0, bias=False), nn.BatchNorm2d(ch[0]), nn.ReLU(True)]
        for a, b in zip(ch[:-1], ch[1:]):
            mods += [nn.ConvTranspose2d(a, b, 4, 2, 1, bias=False), nn.BatchNorm2d(b), nn.ReLU(True)]
        mods += [nn.ConvTranspose2d(ch[-1], nc, 3, 1, 1, bias=False), nn.Tanh()]
        self.main = nn.Sequential(*mods)
        self._plan = None

    def _layers(self):
        mods = list(self.main)
        if self._plan is None or self._plan[0] != [id(m) for m in mods]:
            self._plan = ([id(m) for m in mods], _group_layers(self.main))
        return self._plan[1]

    def forward(self, input):
        last = self._layers()[-1]
        return F_.ToNCHWActFn.apply(self.forward_nhwc(F_.ToNHWCFn.apply(input, self._dtype())), last.act,
                                    last.conv.out_channels)

    def forward_nhwc(self, h):
        """z as NHWC [B,1,1,nz] -> RAW last-conv output NHWC [B,H,W,nc]; the final Tanh rides the NCHW store."""
        layers = self._layers()
        h, link = _run_chain(layers[:-1], h, self.training)
        return layers[-1](h, self.training, fuse_act=False, link_in=link)


Decoder = Generator   # main_vae.py:10


# --------------------------------------------------------------------------------------------- discriminator
def _discriminator_channels(ndf: int, hw: int) -> List[int]:
    full = [ndf // 4, ndf // 2, ndf, ndf * 2, ndf * 4, ndf * 8]   # stage input sizes 256 .. 8
    if hw < 8 or hw > 256 or hw & (hw - 1):
        raise ValueError("hw must be a power of two in [8, 256]")
    return full[len(full) - (int(math.log2(hw)) - 2):]


class Discriminator(_KernelModule):
    """gan_code.py:56-89 (hw=256), or its resolution-aligned truncation."""

    def __init__(self, ndf=64, nc=3, hw: int = 256, precision: Optional[str] = None):
        super().__init__()
        self.precision = precision or _DEFAULT_PRECISION
        ch = _discriminator_channels(ndf, hw)
        mods: List[nn.Module] = [nn.Conv2d(nc, ch[0], 4, 2, 1, bias=False), nn.LeakyReLU(0.2, inplace=True)]
        for a, b in zip(ch[:-1], ch[1:]):
            mods += [nn.Conv2d(a, b, 4, 2, 1, bias=False), nn.BatchNorm2d(b), nn.LeakyReLU(0.2, inplace=True)]
        mods += [nn.Conv2d(ch[-1], 1, 4, 1, 0, bias=False), nn.Sigmoid()]
        self.main = nn.Sequential(*mods)
        self._plan = None

    _layers = Generator._layers

    def forward(self, input):
        return self.forward_nhwc(F_.ToNHWCFn.apply(input, self._dtype(),
                                                   self.input_s2d_origin(input.shape[2], input.shape[3])))

    def input_s2d_origin(self, h: int, w: int):
        return self._layers()[0].input_s2d_origin(self._dtype(), h, w)

    def forward_nhwc(self, h, groups: int = 1, tap=None):
        """`groups` independent sub-batches stacked along the batch axis share every convolution launch while
        BatchNorm treats them separately (the fused step runs D(real) and D(fake) of vaegan_code.py:96-97 this way).
        `tap = (l, fn)`: the activation of conv group `l` (Dis_l of README.md eq. 2) is passed through `fn` before the
        next layer reads it; the fused backward chain is cut there so that `fn` sees (and may add to) its gradient."""
        layers = self._layers()
        if tap is None:
            h, link = _run_chain(layers[:-1], h, self.training, groups=groups)
        else:
            cut = tap[0] % len(layers)
            if cut >= len(layers) - 1:
                raise ValueError("the tapped layer must be a hidden layer of the discriminator")
            h = tap[1](self.features_nhwc(h, cut, groups=groups))
            h, link = _run_chain(layers[cut + 1:-1], h, self.training, groups=groups)
        last = layers[-1]
        logits = last(h, self.training, out_f32=True, fuse_act=False, link_in=link)     # [B, 1, 1, 1] fp32
        return F_.PointwiseActFn.apply(logits, last.act, last.slope).view(-1)

    def features_nhwc(self, h, layer: int, groups: int = 1):
        """Activation (NHWC, internal dtype) after conv group `layer` - Dis_l(x)."""
        layers = self._layers()
        cut = layer % len(layers)
        h, link = _run_chain(layers[:cut], h, self.training, groups=groups)
        return layers[cut](h, self.training, groups=groups, link_in=link)


def weights_init(m):
    """gan_code.py:91-97: DCGAN initialisation keyed on the CLASS NAME (so it can be passed to `.apply`):
    any module whose class name contains "Conv" gets weight ~ N(0, 0.02); any "BatchNorm" gets
    weight ~ N(1, 0.02) and bias = 0.  Applied to the decoder and the discriminator only (vaegan_code.py:37-38)."""
    kind = type(m).__name__
    if "Conv" in kind:
        nn.init.normal_(m.weight.data, mean=0.0, std=0.02)
    elif "BatchNorm" in kind:
        nn.init.normal_(m.weight.data, mean=1.0, std=0.02)
        nn.init.zeros_(m.bias.data)
