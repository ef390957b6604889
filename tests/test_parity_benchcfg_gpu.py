"""GPU parity of the fused step AT THE BENCHMARKED CONFIGURATIONS (BASELINE.json configs 2, 3 and 4): bf16 tensor-core
mode, CUDA graph, the batch sizes bench.py times - so the launch plans that produce the headline number (tile
narrowing, split-K, persistent item walk, the discriminator's stacked real/fake batch of 2*B) are the ones compared.

Reference = oracle.reference_step (vaegan_code.py:74-135) on identical weights, inputs and injected noise:
  * all six losses of the step within 2e-2 relative of the fp32 oracle (north_star's bf16 tolerance);
  * every parameter-gradient tensor (E, G, and D of BOTH discriminator updates) against the bf16-EMULATED oracle
    (same rounding points, SURVEY.md Appendix D protocol (i)): cosine > 0.999, generator 0.998 (two bf16 pipelines with
    identical rounding points cannot agree better through five ReLU+BN stages);
  * per-network flattened gradient against the PURE fp32 oracle (protocol (ii)), asserted at the floors SURVEY.md
    Appendix D measured for bf16 operands: E 0.995, G 0.9999, D 0.9995.
The measured values are printed (pytest -s) and copied to profiles/ by scripts/parity_report.sh.
"""
import copy
import json

import pytest
import torch

from tests.util import cosine, make_pair

pytestmark = pytest.mark.gpu

# name, hw, nz, width, batch per GPU, denoise sigma
CASES = [
    ("cfg2_b256", 64, 128, 1, 256, 0.0),
    ("cfg3_b256_denoise", 64, 128, 1, 256, 0.1),
    ("cfg4_b64", 128, 256, 2, 64, 0.0),
    ("cfg4_b128", 128, 256, 2, 128, 0.0),
]

LOSS_TOL = 2e-2
PER_TENSOR_FLOOR = {"E": 0.999, "G": 0.998, "D": 0.999}
FLAT_FP32_FLOOR = {"E": 0.995, "G": 0.9999, "D": 0.9995}


def _clean(name: str) -> str:
    return name.replace("parametrizations.weight.original", "weight")


def _flat(grads, skip_bias=True):
    keys = [k for k in sorted(grads) if not (skip_bias and "conv.bias" in k)]
    return torch.cat([grads[k].detach().double().cpu().flatten() for k in keys])


@pytest.mark.parametrize("name,hw,nz,width,batch,sigma", CASES, ids=[c[0] for c in CASES])
def test_bf16_graph_step_at_bench_config(name, hw, nz, width, batch, sigma):
    from importlib import import_module
    from oracle import vaegan_oracle as vo
    VAEGANStep = import_module("vaegan_b200.step").VAEGANStep
    epoch = 50
    o_nets, nets = make_pair(hw, nz, "bf16", width=width)
    emu_nets = copy.deepcopy(o_nets)
    vo.attach_bf16_emulation(*emu_nets)
    real, eps, n_real, n_fake = vo.make_inputs(batch, hw, nz)
    n_den = torch.randn(real.shape, generator=torch.Generator().manual_seed(46))
    kw = dict(denoise_sigma=sigma, n_denoise=n_den)
    res32 = vo.reference_step(*o_nets, *vo.make_optimizers(*o_nets), real, epoch, eps, n_real, n_fake, **kw)
    res16 = vo.reference_step(*emu_nets, *vo.make_optimizers(*emu_nets), real, epoch, eps, n_real, n_fake, **kw)

    step = VAEGANStep(*nets, use_cuda_graph=True, denoise_sigma=sigma, capture_grads=True)
    losses = step.step(real.cuda(), epoch, eps.cuda(), n_real.cuda(), n_fake.cuda(), n_den.cuda() if sigma > 0 else None)
    torch.cuda.synchronize()
    report = {"case": name, "losses": {}, "per_tensor_min_cos_vs_emulated": {}, "flat_cos_vs_fp32": {},
              "flat_cos_vs_emulated": {}}
    for k, v in res32.losses.items():
        got = float(losses[k])
        report["losses"][k] = {"gpu": got, "fp32_oracle": v, "emulated_oracle": res16.losses[k],
                               "rel_err_vs_fp32": abs(got - v) / (abs(v) + 1e-12)}
    g = step.gradients()
    mine = {"E": [g["E"]], "G": [g["G"]], "D": g["D"]}
    ref16 = {"E": [res16.e_grads], "G": [res16.g_grads], "D": res16.d_grads}
    ref32 = {"E": [res32.e_grads], "G": [res32.g_grads], "D": res32.d_grads}
    worst = []
    for net in "EGD":
        for it, (gm, g16, g32) in enumerate(zip(mine[net], ref16[net], ref32[net])):
            g16 = {_clean(k): v for k, v in g16.items()}
            tag = net if net != "D" else f"D{it}"
            cos = {k: cosine(gm[k], g16[k]) for k in g16 if "conv.bias" not in k}
            kmin = min(cos, key=cos.get)
            report["per_tensor_min_cos_vs_emulated"][tag] = {"min": cos[kmin], "tensor": kmin,
                                                             "median": sorted(cos.values())[len(cos) // 2]}
            report["flat_cos_vs_fp32"][tag] = cosine(_flat(gm), _flat(g32))
            report["flat_cos_vs_emulated"][tag] = cosine(_flat(gm), _flat(g16))
            worst += [(tag, k, c) for k, c in cos.items() if c <= PER_TENSOR_FLOOR[net]]
    print("PARITY_REPORT " + json.dumps(report))
    for k, r in report["losses"].items():
        assert r["rel_err_vs_fp32"] <= LOSS_TOL, f"{name} loss {k}: {r}"
    assert not worst, f"{name}: gradient tensors below the cosine floor vs the bf16-emulated oracle: {worst}"
    for tag, c in report["flat_cos_vs_fp32"].items():
        assert c > FLAT_FP32_FLOOR[tag[0]], f"{name}: flattened {tag} gradient cosine vs fp32 oracle {c:.6f}"


def test_eval_forward_after_step_uses_updated_weights():
    """the fused Adam updates the fp32 masters behind autograd's back; an eval / generation pass through
    the SAME modules after a step must see bf16 copies of the updated weights (not the ones packed at the start of
    the step)."""
    from importlib import import_module
    import vaegan_b200 as vb
    from oracle import vaegan_oracle as vo
    VAEGANStep = import_module("vaegan_b200.step").VAEGANStep
    hw, nz, batch = 64, 128, 16
    _, nets = make_pair(hw, nz, "bf16")
    step = VAEGANStep(*nets, use_cuda_graph=True)
    real, eps, n_real, n_fake = vo.make_inputs(batch, hw, nz)
    for _ in range(3):          # a few steps so that stale copies would differ visibly (3 x lr per weight)
        step.step(real.cuda(), 50, eps.cuda(), n_real.cuda(), n_fake.cuda())
    E, G, _ = nets
    z = torch.randn(8, nz, 1, 1, generator=torch.Generator().manual_seed(3)).cuda()
    x = real[:8].cuda()
    E.eval(), G.eval()
    with torch.no_grad():
        img, (mu, _) = G(z), E(x)
        # fresh modules holding exactly the current masters / buffers: their packed copies are made from scratch
        e2 = vb.Encoder([3, hw, hw], nz, precision="bf16").cuda()
        g2 = vb.Generator(nz=nz, hw=hw, precision="bf16").cuda()
        e2.load_state_dict(E.state_dict()), g2.load_state_dict(G.state_dict())
        e2.eval(), g2.eval()
        img2, (mu2, _) = g2(z), e2(x)
    assert torch.equal(img, img2), float((img - img2).abs().max())
    assert torch.equal(mu, mu2), float((mu - mu2).abs().max())
