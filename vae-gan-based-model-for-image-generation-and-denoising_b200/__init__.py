"""vaegan_b200 - B200-native (sm_100a) VAE-GAN training step behind the reference's nn.Module interfaces.

Import as `vaegan_b200` (the repo-root shim maps that name onto this directory).  See DESIGN.md.
"""
from ._lib import LIB_PATH, VaeganB200Error, load as load_library  # noqa: F401
from .modules import (ConvBlock, Decoder, Discriminator, Encoder, Generator, set_default_precision,  # noqa: F401
                      weights_init)
from .generate import GraphedGenerator  # noqa: F401
from . import ops  # noqa: F401  (registers torch.ops.vaegan_b200.*)

__all__ = ["ConvBlock", "Encoder", "Generator", "Decoder", "Discriminator", "weights_init", "set_default_precision",
           "GraphedGenerator", "load_library", "VaeganB200Error", "LIB_PATH"]
