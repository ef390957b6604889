"""Import the UNMODIFIED reference classes from /root/reference (build container only).  TEST INFRASTRUCTURE.

The reference modules import matplotlib / torchmetrics at module scope (main_vae.py:6-13, gan_code.py:3-12,
vaegan_code.py:8-16); neither is installed here and neither is on the hot path, so they are stubbed in
sys.modules before the import.  Nothing in this file is used on the GPU box.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("VAEGAN_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "vaegan_code.py"))


class _Swallow:
    """Stands in for any metric / plotting object: accepts every construction, call and attribute."""

    def __init__(self, *a, **k):
        pass

    def __getattr__(self, name):
        return _Swallow()

    def __call__(self, *a, **k):
        return _Swallow()


def _stub(name, **attrs):
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    sys.modules.setdefault(name, mod)
    return sys.modules[name]


def import_reference():
    """Returns (Encoder, ConvBlock, Generator, Discriminator, weights_init, configure_seed) of the reference."""
    if not reference_available():
        raise RuntimeError(f"reference sources not found under {REFERENCE_ROOT}")
    mpl = _stub("matplotlib")
    mpl.pyplot = _stub("matplotlib.pyplot", rcParams={})
    mpl.gridspec = _stub("matplotlib.gridspec")
    _stub("torchmetrics")
    _stub("torchmetrics.image", FrechetInceptionDistance=_Swallow, StructuralSimilarityIndexMeasure=_Swallow)
    _stub("torchmetrics.image.inception", InceptionScore=_Swallow)
    _stub("torchmetrics.image.fid", FrechetInceptionDistance=_Swallow)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    from main_vae import Encoder, ConvBlock  # noqa: E402
    from gan_code import Generator, Discriminator, weights_init  # noqa: E402
    from utils import configure_seed  # noqa: E402
    return Encoder, ConvBlock, Generator, Discriminator, weights_init, configure_seed


def build_reference_nets(hw: int, nz: int, width: int = 1, seed: int = 42):
    """The reference's own networks; for hw < 256 the resolution-derived variants of SURVEY.md Appendix A.1, built by
    slicing the reference nn.Sequential objects so every surviving layer is the reference's layer."""
    import math
    import torch
    import torch.nn as nn
    Encoder, ConvBlock, Generator, Discriminator, weights_init, configure_seed = import_reference()
    configure_seed(seed)
    enc = Encoder([3, hw, hw], nz)
    if width != 1:
        raise NotImplementedError("width != 1 is validated structurally only (encoder channels are literals)")
    gen = Generator(nz=nz, ngf=64 * width)
    dis = Discriminator(ndf=64 * width)
    d = int(math.log2(256 // hw))
    if d:
        g = list(gen.main)[:21 - 3 * d]
        gen.main = nn.Sequential(*g, nn.ConvTranspose2d(g[-3].out_channels, 3, 3, 1, 1, bias=False), nn.Tanh())
        t = list(dis.main)[2 + 3 * d:]
        dis.main = nn.Sequential(nn.Conv2d(3, t[0].in_channels, 4, 2, 1, bias=False), nn.LeakyReLU(0.2, True), *t)
    gen.apply(weights_init)
    dis.apply(weights_init)
    return enc, gen, dis
