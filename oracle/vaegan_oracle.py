"""CPU oracle for the VAE-GAN training step.  TEST INFRASTRUCTURE ONLY.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may import
this module; the product package never does (it has no CPU path at all).

What it restates, and from where (reference = viniciusmenesessouza/VAE-GAN-based-model-for-image-generation-and-
denoising, mounted at /root/reference in the build container only):

  * `OracleEncoder`        <- main_vae.py:20-58   (ConvBlock: Conv2d k4 s2 p0 +bias -> BatchNorm2d -> LeakyReLU(0.01);
                                                    channels [C,32,64,128,256]; ctor dry-run in train mode :43-45;
                                                    (c,h,w) flatten :53; two Linear heads :47-48)
  * `make_generator`       <- gan_code.py:16-54   (ConvT(nz,16ngf,4,1,0) -> [BN, ReLU, ConvT(k4,s2,p1)]* -> ... ->
                                                    ConvT(.,nc,3,1,1) -> Tanh; bias=False everywhere)
  * `make_discriminator`   <- gan_code.py:56-89   (Conv(nc,.,4,2,1)+LeakyReLU(0.2) (no BN) -> [Conv(k4,s2,p1), BN,
                                                    LeakyReLU(0.2)]* -> Conv(8ndf,1,4,1,0) -> Sigmoid -> view(-1))
  * `weights_init`         <- gan_code.py:91-97
  * `reference_step`       <- vaegan_code.py:74-135, line by line, with the three torch.randn_like draws
                              (:77, :91, :92) replaced by injected tensors so CPU and GPU see the same noise.

The arithmetic itself (convolution, batch norm, BCE, MSE, Adam) lives in the third-party dependency `torch`
(unpinned by the reference - it has no requirements file; this image has torch 2.11.0+cu128, CPU = ATen/oneDNN).
The oracle calls the same torch.nn modules the reference calls.

Pinning: the reference ships no tests / golden vectors.  The restatement is pinned instead
against the reference's own classes executed in the build container: `oracle/gen_golden.py` imports
main_vae.Encoder / gan_code.Generator / gan_code.Discriminator (matplotlib / torchmetrics stubbed), checks
this file reproduces them BIT-EXACTLY (state_dict keys, forward, gradients, one full step), and writes
`tests/golden/*.npz`, which `tests/test_oracle_golden.py` replays without the reference.

Resolution variants: the reference networks are hard-wired to 256x256.  For H in
{64, 128} with d = log2(256/H): G keeps its first 7-d upsampling stages and gets a fresh ConvT(c,3,3,1,1)+Tanh;
D drops its first d downsampling stages and starts with Conv(3,c,4,2,1)+LeakyReLU(0.2).  `make_generator` /
`make_discriminator` build those directly with the reference's module indices (state_dict keys `main.<i>`).
"""
from __future__ import annotations

import math
import os
import random
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn as nn


# --------------------------------------------------------------------------------------------- seeding
def configure_seed(seed: int) -> None:
    """utils.py:6-14 (CPU part)."""
    os.environ["PYTHONHASHSEED"] = str(seed)
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)


# --------------------------------------------------------------------------------------------- networks
class OracleConvBlock(nn.Module):
    """main_vae.py:20-31."""

    def __init__(self, cin: int, cout: int, kernel_size: int = 4, stride: int = 2):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, kernel_size, stride)
        self.bn = nn.BatchNorm2d(cout)
        self.leaky_relu = nn.LeakyReLU(inplace=True)

    def forward(self, x):
        return self.leaky_relu(self.bn(self.conv(x)))


class OracleEncoder(nn.Module):
    """main_vae.py:34-58.  `width` multiplies the literal channel list (cfg 4 of BASELINE.json)."""

    def __init__(self, img_size, latent_dim: int, width: int = 1):
        super().__init__()
        chans = [img_size[0]] + [c * width for c in (32, 64, 128, 256)]
        self.cnn = nn.Sequential(*[OracleConvBlock(a, b) for a, b in zip(chans[:-1], chans[1:])])
        probe = self.cnn(torch.zeros(1, img_size[0], img_size[1], img_size[2]))  # train-mode dry run, as the reference
        self.flatten_size = probe.view(1, -1).size(1)
        self.fc_mu = nn.Linear(self.flatten_size, latent_dim)
        self.fc_logvar = nn.Linear(self.flatten_size, latent_dim)

    def forward(self, x):
        h = self.cnn(x)
        h = h.view(h.size(0), -1)
        return self.fc_mu(h), self.fc_logvar(h)


class _Seq(nn.Module):
    """Holds `main` so state_dict keys read `main.<i>.*` like the reference classes."""

    def __init__(self, layers: List[nn.Module], flatten_out: bool):
        super().__init__()
        self.main = nn.Sequential(*layers)
        self._flatten_out = flatten_out

    def forward(self, x):
        y = self.main(x)
        return y.view(-1) if self._flatten_out else y


def generator_channels(ngf: int, hw: int) -> List[int]:
    """Channel count after each upsampling stage 4,8,...,hw  (gan_code.py:21-46 at the same feature-map size)."""
    full = [ngf * 16, ngf * 8, ngf * 4, ngf * 2, ngf, ngf // 2, ngf // 4]  # sizes 4..256
    n = int(math.log2(hw)) - 1
    return full[:n]


def make_generator(nz: int = 128, ngf: int = 64, nc: int = 3, hw: int = 256) -> nn.Module:
    chans = generator_channels(ngf, hw)
    layers: List[nn.Module] = [nn.ConvTranspose2d(nz, chans[0], 4, 1, 0, bias=False), nn.BatchNorm2d(chans[0]),
                               nn.ReLU(True)]
    for a, b in zip(chans[:-1], chans[1:]):
        layers += [nn.ConvTranspose2d(a, b, 4, 2, 1, bias=False), nn.BatchNorm2d(b), nn.ReLU(True)]
    layers += [nn.ConvTranspose2d(chans[-1], nc, 3, 1, 1, bias=False), nn.Tanh()]
    return _Seq(layers, flatten_out=False)


def discriminator_channels(ndf: int, hw: int) -> List[int]:
    """Output channels of each stride-2 stage for an hw x hw input (gan_code.py:61-80 at the same map size)."""
    full = [ndf // 4, ndf // 2, ndf, ndf * 2, ndf * 4, ndf * 8]  # input sizes 256..8
    n = int(math.log2(hw)) - 2
    return full[len(full) - n:]


def make_discriminator(ndf: int = 64, nc: int = 3, hw: int = 256) -> nn.Module:
    chans = discriminator_channels(ndf, hw)
    layers: List[nn.Module] = [nn.Conv2d(nc, chans[0], 4, 2, 1, bias=False), nn.LeakyReLU(0.2, inplace=True)]
    for a, b in zip(chans[:-1], chans[1:]):
        layers += [nn.Conv2d(a, b, 4, 2, 1, bias=False), nn.BatchNorm2d(b), nn.LeakyReLU(0.2, inplace=True)]
    layers += [nn.Conv2d(chans[-1], 1, 4, 1, 0, bias=False), nn.Sigmoid()]
    return _Seq(layers, flatten_out=True)


def conv_group_end(main: nn.Sequential, group: int) -> int:
    """Index into `main` just past conv group `group` (conv [+ BatchNorm] + activation); negative = from the end."""
    starts = [i for i, m in enumerate(main) if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d))]
    group = group % len(starts)
    return starts[group + 1] if group + 1 < len(starts) else len(main)


def weights_init(m: nn.Module) -> None:
    """gan_code.py:91-97."""
    name = m.__class__.__name__
    if name.find("Conv") != -1:
        nn.init.normal_(m.weight.data, 0.0, 0.02)
    elif name.find("BatchNorm") != -1:
        nn.init.normal_(m.weight.data, 1.0, 0.02)
        nn.init.constant_(m.bias.data, 0)


@dataclass
class NetConfig:
    hw: int = 64          # image side
    nz: int = 128         # latent size
    width: int = 1        # channel multiplier (ngf = ndf = 64 * width; encoder channels * width)
    seed: int = 42


def build_nets(cfg: NetConfig):
    """vaegan_code.py:19,29-38: seed, construct E, G, D in that order, weights_init on G and D only."""
    configure_seed(cfg.seed)
    enc = OracleEncoder([3, cfg.hw, cfg.hw], cfg.nz, cfg.width)
    gen = make_generator(nz=cfg.nz, ngf=64 * cfg.width, hw=cfg.hw)
    dis = make_discriminator(ndf=64 * cfg.width, hw=cfg.hw)
    gen.apply(weights_init)
    dis.apply(weights_init)
    return enc, gen, dis


def make_optimizers(enc, gen, dis, lr: float = 2e-4):
    """vaegan_code.py:42-44 (torch.optim.Adam defaults: betas (0.9, 0.999), eps 1e-8, no weight decay)."""
    return (torch.optim.Adam(enc.parameters(), lr=lr), torch.optim.Adam(gen.parameters(), lr=lr),
            torch.optim.Adam(dis.parameters(), lr=lr))


def make_inputs(batch: int, hw: int, nz: int, seed: int = 42):
    """real ~ U[-1,1) from seed, eps / n_real / n_fake ~ N(0,1) from seed+1/+2/+3."""
    g = torch.Generator().manual_seed(seed)
    real = torch.rand(batch, 3, hw, hw, generator=g) * 2 - 1
    eps = torch.randn(batch, nz, generator=torch.Generator().manual_seed(seed + 1))
    n_real = torch.randn(batch, 3, hw, hw, generator=torch.Generator().manual_seed(seed + 2))
    n_fake = torch.randn(batch, 3, hw, hw, generator=torch.Generator().manual_seed(seed + 3))
    return real, eps, n_real, n_fake


# --------------------------------------------------------------------------------------------- the step
@dataclass
class StepResult:
    losses: Dict[str, float]
    d_grads: List[Dict[str, torch.Tensor]] = field(default_factory=list)   # per D iteration (before its Adam step)
    e_grads: Dict[str, torch.Tensor] = field(default_factory=dict)
    g_grads: Dict[str, torch.Tensor] = field(default_factory=dict)
    mu: Optional[torch.Tensor] = None
    logvar: Optional[torch.Tensor] = None
    recon: Optional[torch.Tensor] = None


def reference_step(enc, gen, dis, opt_e, opt_g, opt_d, real, epoch: int, eps, n_real, n_fake, *, n_dis: int = 2,
                   alpha_kl: float = 0.1, alpha_adv: float = 0.1, sigma_inst: float = 0.05,
                   denoise_sigma: float = 0.0, n_denoise=None, keep_grads: bool = True,
                   recon_mode: str = "pixel", dis_layer: int = -2) -> StepResult:
    """One iteration of the hot loop, vaegan_code.py:66-135.  Line numbers refer to that file.

    `recon_mode="dis_l"` replaces the pixel MSE of :113 by the feature-matching term the reference's README.md:11-14
    (eq. 2, Larsen et al.) describes but vaegan_code.py does not implement: MSE between the discriminator's
    `dis_layer`-th conv-group activations of recon_noisy and (detached) of real_noisy.  PARITY UNPINNED: there is no
    reference code for this mode, the restatement below is the definition both sides of the test share.

    `denoise_sigma` > 0 is BASELINE.json config 3: the encoder sees clamp(real + sigma*n_denoise, -1, 1)
    (pattern of main_vae.py:104-105 / vaegan_code.py:153-154); the reconstruction target stays `real`.
    """
    bce = nn.BCELoss()                                   # :46
    mse = nn.MSELoss(reduction="mean")                   # :47
    batch = real.size(0)                                 # :67
    res = StepResult(losses={})

    enc_in = real
    if denoise_sigma > 0.0:
        enc_in = torch.clamp(real + denoise_sigma * n_denoise, -1.0, 1.0)
    mu, logvar = enc(enc_in)                             # :74
    logvar = torch.clamp(logvar, min=-10, max=10)        # :75
    std = torch.exp(0.5 * logvar)                        # :76
    z = mu + std * eps                                   # :77 (eps injected for randn_like)
    z = z.unsqueeze(-1).unsqueeze(-1)                    # :78
    recon = gen(z)                                       # :83

    real_labels = torch.full((batch,), 0.9, device=real.device, dtype=real.dtype)   # :88 (dtype: float64 twins)
    fake_labels = torch.full((batch,), 0.1, device=real.device, dtype=real.dtype)   # :89
    real_noisy = real + sigma_inst * n_real              # :91
    recon_noisy = recon + sigma_inst * n_fake            # :92

    for it in range(n_dis):                              # :95
        real_out = dis(real_noisy)                       # :96
        fake_out = dis(recon_noisy.detach())             # :97
        d_loss = bce(real_out, real_labels) + bce(fake_out, fake_labels)   # :99-101
        opt_d.zero_grad()                                # :103
        d_loss.backward()                                # :104
        if keep_grads:
            res.d_grads.append({k: p.grad.detach().clone() for k, p in dis.named_parameters()})
        opt_d.step()                                     # :105
        res.losses[f"d_loss_{it}"] = float(d_loss.item())

    if recon_mode == "dis_l":
        cut = conv_group_end(dis.main, dis_layer)
        with torch.no_grad():
            feat_real = dis.main[:cut](real_noisy)       # Dis_l(x): train-mode call, its own BatchNorm statistics
        feat_fake = dis.main[:cut](recon_noisy)          # Dis_l(x~)
        fake_out = dis.main[cut:](feat_fake).view(-1)    # :110 continues from the tapped features
        recon_loss = mse(feat_fake, feat_real)           # README.md eq. (2)
    else:
        fake_out = dis(recon_noisy)                      # :110
        recon_loss = mse(recon, real)                    # :113
    kl_loss = -0.5 * torch.sum(1 + logvar - mu.pow(2) - logvar.exp()) / batch   # :114
    g_adv = bce(fake_out, real_labels)                   # :115
    total = recon_loss + alpha_kl * min(1.0, epoch / 50) * kl_loss + alpha_adv * g_adv   # :117
    opt_e.zero_grad()                                    # :131
    opt_g.zero_grad()                                    # :132
    total.backward()                                     # :133
    if keep_grads:
        res.e_grads = {k: p.grad.detach().clone() for k, p in enc.named_parameters()}
        res.g_grads = {k: p.grad.detach().clone() for k, p in gen.named_parameters()}
    opt_e.step()                                         # :134
    opt_g.step()                                         # :135

    res.losses.update(recon=float(recon_loss.item()), kl=float(kl_loss.item()), adv=float(g_adv.item()),
                      total=float(total.item()))
    res.mu, res.logvar, res.recon = mu.detach(), logvar.detach(), recon.detach()
    return res


# --------------------------------------------------------------------------------------------- bf16 emulation
def _round_bf16(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.bfloat16).to(t.dtype)       # (dtype-preserving: the emulation also runs on float64 twins)


class _RoundSTE(torch.autograd.Function):
    """Round to bf16 in forward AND backward (the GPU path stores activations and their gradients as bf16)."""

    @staticmethod
    def forward(ctx, x):
        return _round_bf16(x)

    @staticmethod
    def backward(ctx, g):
        return _round_bf16(g)


def attach_bf16_emulation(*nets: nn.Module) -> None:
    """Make fp32 CPU modules mimic the rounding points of the bf16 tensor-core path: conv / linear operands
    (activations AND weights) and their outputs are bf16-representable in forward, and the gradients that flow
    back through those points are rounded too; accumulation stays fp32 (SURVEY.md Appendix D protocol).
    Applies parametrizations in place - use on deep copies."""
    import torch.nn.utils.parametrize as parametrize

    def pre(mod, args):
        return (_RoundSTE.apply(args[0]),)

    def post(mod, args, out):
        return _RoundSTE.apply(out)

    for net in nets:
        for m in list(net.modules()):
            if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d, nn.Linear)):
                m.register_forward_pre_hook(pre)
                m.register_forward_hook(post)
                parametrize.register_parametrization(m, "weight", _RoundParam(), unsafe=True)


class _RoundParam(nn.Module):
    def forward(self, w):
        return _RoundFwd.apply(w)


class _RoundFwd(torch.autograd.Function):
    """bf16 rounding of a weight in forward; its fp32 gradient passes through untouched (fp32 wgrad output)."""

    @staticmethod
    def forward(ctx, x):
        return _round_bf16(x)

    @staticmethod
    def backward(ctx, g):
        return g


# --------------------------------------------------------------------------------------------- decoder-only
@torch.no_grad()
def generate(gen: nn.Module, z: torch.Tensor) -> torch.Tensor:
    """main_vae.py:361-366: eval-mode decoder on z ~ N(0, I) shaped [B, nz, 1, 1]."""
    was_training = gen.training
    gen.eval()
    out = gen(z)
    gen.train(was_training)
    return out
