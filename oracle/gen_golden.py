"""Generate tests/golden/*.npz by executing the UNMODIFIED reference classes.  TEST INFRASTRUCTURE ONLY.

Run in the build container (needs /root/reference):   python -m oracle.gen_golden

For every configuration it
  1. builds the oracle restatement (oracle/vaegan_oracle.py) and the reference's own Encoder / Generator /
     Discriminator (resolution-derived variants built by slicing the reference nn.Sequential objects),
  2. copies one set of weights into both and checks that forward outputs, every gradient of one full
     VAE-GAN step (vaegan_code.py:74-135) and the post-step weights / BatchNorm buffers agree BIT-EXACTLY,
  3. writes a compact fixture (losses, per-tensor gradient / weight statistics, BN buffers, latent outputs)
     taken from the REFERENCE-class run.  The fixtures are what pins the oracle on machines without the reference.
"""
import copy
import json
import os
import sys

import numpy as np
import torch

from . import ref_import
from . import vaegan_oracle as vo

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

CONFIGS = {
    # name: hw, nz, batch, epoch, denoise_sigma
    "tiny64_e50": dict(hw=64, nz=128, batch=4, epoch=50, denoise_sigma=0.0),
    "tiny64_e0": dict(hw=64, nz=128, batch=4, epoch=0, denoise_sigma=0.0),
    "denoise64_e50": dict(hw=64, nz=128, batch=4, epoch=50, denoise_sigma=0.1),
    "native256_e50": dict(hw=256, nz=100, batch=2, epoch=50, denoise_sigma=0.0),
}


def tensor_stats(t: torch.Tensor) -> np.ndarray:
    """[l2 norm, sum, first 6 values] in float64 - enough to pin a tensor without storing it."""
    f = t.detach().double().flatten()
    head = torch.zeros(6, dtype=torch.float64)
    n = min(6, f.numel())
    head[:n] = f[:n]
    return torch.cat([f.norm().reshape(1), f.sum().reshape(1), head]).numpy()


def run(nets, cfg, inputs):
    enc, gen, dis = nets
    opts = vo.make_optimizers(enc, gen, dis)
    real, eps, n_real, n_fake, n_den = inputs
    return vo.reference_step(enc, gen, dis, *opts, real, cfg["epoch"], eps, n_real, n_fake,
                             denoise_sigma=cfg["denoise_sigma"], n_denoise=n_den)


def assert_same(a: torch.Tensor, b: torch.Tensor, what: str):
    if not torch.equal(a, b):
        raise AssertionError(f"oracle restatement differs from the reference at {what}: "
                             f"max abs diff {(a - b).abs().max().item():.3e}")


def generate(name: str, cfg: dict) -> dict:
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    o_nets = vo.build_nets(vo.NetConfig(hw=cfg["hw"], nz=cfg["nz"]))
    r_nets = ref_import.build_reference_nets(cfg["hw"], cfg["nz"])
    for o, r in zip(o_nets, r_nets):
        assert list(o.state_dict().keys()) == list(r.state_dict().keys()), "state_dict keys differ"
        r.load_state_dict(copy.deepcopy(o.state_dict()))
    weights_before = {f"{n}.{k}": tensor_stats(v) for n, net in zip("EGD", o_nets)
                      for k, v in net.state_dict().items() if v.dtype.is_floating_point}

    real, eps, n_real, n_fake = vo.make_inputs(cfg["batch"], cfg["hw"], cfg["nz"], seed=42)
    n_den = torch.randn(real.shape, generator=torch.Generator().manual_seed(46))
    inputs = (real, eps, n_real, n_fake, n_den)

    res_r = run(r_nets, cfg, inputs)
    res_o = run(o_nets, cfg, inputs)

    # ---- the pin: restatement == reference, bit for bit
    for k in res_r.losses:
        assert res_r.losses[k] == res_o.losses[k], f"loss {k}: {res_r.losses[k]} vs {res_o.losses[k]}"
    assert_same(res_r.mu, res_o.mu, "mu")
    assert_same(res_r.logvar, res_o.logvar, "logvar")
    assert_same(res_r.recon, res_o.recon, "recon")
    for it in range(2):
        for k in res_r.d_grads[it]:
            assert_same(res_r.d_grads[it][k], res_o.d_grads[it][k], f"D grad iter {it} {k}")
    for k in res_r.e_grads:
        assert_same(res_r.e_grads[k], res_o.e_grads[k], f"E grad {k}")
    for k in res_r.g_grads:
        assert_same(res_r.g_grads[k], res_o.g_grads[k], f"G grad {k}")
    for n, o, r in zip("EGD", o_nets, r_nets):
        for (k, a), (_, b) in zip(r.state_dict().items(), o.state_dict().items()):
            assert_same(a, b, f"{n} post-step {k}")

    # ---- the fixture (from the reference-class run)
    out = {"meta": np.frombuffer(json.dumps(dict(cfg, name=name, torch=torch.__version__)).encode(), dtype=np.uint8)}
    for k, v in res_r.losses.items():
        out[f"loss/{k}"] = np.float64(v)
    for k, v in weights_before.items():
        out[f"w0/{k}"] = v
    for it in range(2):
        for k, v in res_r.d_grads[it].items():
            out[f"grad/D{it}.{k}"] = tensor_stats(v)
    for k, v in res_r.e_grads.items():
        out[f"grad/E.{k}"] = tensor_stats(v)
    for k, v in res_r.g_grads.items():
        out[f"grad/G.{k}"] = tensor_stats(v)
    for n, r in zip("EGD", r_nets):
        for k, v in r.state_dict().items():
            if "running_" in k or "num_batches" in k:
                out[f"bn/{n}.{k}"] = v.detach().numpy().copy()
            else:
                out[f"w1/{n}.{k}"] = tensor_stats(v)
    out["out/mu"] = res_r.mu.numpy()
    out["out/logvar"] = res_r.logvar.numpy()
    out["out/recon_sub"] = res_r.recon[:, :, ::8, ::8].numpy().copy()
    return out


def main():
    if not ref_import.reference_available():
        print("reference sources not available; nothing generated", file=sys.stderr)
        return 1
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    for name, cfg in CONFIGS.items():
        fx = generate(name, cfg)
        path = os.path.join(GOLDEN_DIR, f"{name}.npz")
        np.savez_compressed(path, **fx)
        print(f"{name}: oracle == reference bit-exact; wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB) "
              f"losses " + ", ".join(f"{k.split('/')[1]}={float(v):.6f}" for k, v in fx.items() if k.startswith("loss/")))
    return 0


if __name__ == "__main__":
    sys.exit(main())
