"""Debug: run the same fused step from identical state several times; report run-to-run differences of losses and
gradients with the wgrad side stream on / off, eager / graph."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.util import make_pair
from oracle import vaegan_oracle as vo
from importlib import import_module
import vaegan_b200
VAEGANStep = import_module("vaegan_b200.step").VAEGANStep
hw, nz, B = 64, 128, 8
real, eps, n_real, n_fake = [t.cuda() for t in vo.make_inputs(B, hw, nz)]
for overlap in (False, True):
    for graph in (False, True):
        outs = []
        for rep in range(3):
            _, nets = make_pair(hw, nz, "bf16")
            st = VAEGANStep(*nets, use_cuda_graph=graph, overlap_wgrad=overlap)
            l = st.step(real, 50, eps, n_real, n_fake)
            torch.cuda.synchronize()
            outs.append(({k: float(v) for k, v in l.items()}, st.opt_D.params.clone(), st.opt_G.params.clone(), st.opt_E.params.clone(),
                         st.opt_D.grads.clone(), st.opt_G.grads.clone()))
        ref = outs[0]
        dl = max(abs(o[0][k] - ref[0][k]) / abs(ref[0][k]) for o in outs[1:] for k in ref[0])
        dD = max(float((o[1] - ref[1]).abs().max()) for o in outs[1:])
        dG = max(float((o[2] - ref[2]).abs().max()) for o in outs[1:])
        gG = max(float((o[5] - ref[5]).abs().max() / ref[5].abs().max()) for o in outs[1:])
        print(f"overlap={overlap} graph={graph}: max rel loss diff {dl:.2e}  max |dD params| {dD:.2e}  max |dG params| {dG:.2e}  G grad rel diff {gG:.2e}")
