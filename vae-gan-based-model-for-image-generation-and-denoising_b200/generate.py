"""Decoder-only generation (BASELINE.json config 5; main_vae.py:348-374, vaegan_code.py:147-171): the eval-mode
generator under no_grad, one CUDA graph per batch size.

    gen = Generator(...).cuda().eval()
    gg = GraphedGenerator(gen)
    images = gg(z)            # z [B, nz, 1, 1] (device, or host: copied in) -> fp32 NCHW [B, 3, H, W], Tanh range

A forward of the 64x64 generator is 12 launches of this library (5 convolutions, 4 folded-BatchNorm passes, layout
edges); at batch 1 they are ~3 us kernels behind ~5 us of Python + ctypes dispatch each, so replaying a captured graph
is what makes small batches latency-bound on the device instead of on the host.  The graphs read the packed bf16
weights and the BatchNorm running statistics in place: training steps in between need no re-capture - a stale weight
copy is re-packed (in place) before the replay.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch

from . import functional as F_


class GraphedGenerator:
    def __init__(self, generator, max_graphs: int = 32):
        self.gen = generator
        self.dev = next(generator.parameters()).device
        if self.dev.type != "cuda":
            raise F_._lib.VaeganB200Error("GraphedGenerator needs the generator on a CUDA (sm_100) device; no CPU fallback")
        self.max_graphs = max_graphs
        self._graphs: Dict[Tuple[int, ...], tuple] = {}

    def _stale(self) -> bool:
        return any(l.cache.version != l.conv.weight._version for l in self.gen._layers())

    def _forward(self, z):
        with torch.no_grad():
            return self.gen(z)

    @torch.no_grad()
    def __call__(self, z: torch.Tensor, clone: bool = False) -> torch.Tensor:
        """Returns the graph's static output buffer (overwritten by the next call with the same batch size) unless
        `clone`."""
        if self.gen.training:
            raise RuntimeError("GraphedGenerator replays the eval-mode forward: call generator.eval() first "
                               "(main_vae.py:361)")
        key = tuple(z.shape)
        entry = self._graphs.get(key)
        if entry is None:
            if len(self._graphs) >= self.max_graphs:
                self._graphs.pop(next(iter(self._graphs)))
            z_static = torch.zeros(key, dtype=torch.float32, device=self.dev)
            z_static.copy_(z)
            side = torch.cuda.Stream(device=self.dev)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):           # warm-up outside capture: packs weights, plans the layers
                self._forward(z_static)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                out = self._forward(z_static)
            entry = (g, z_static, out)
            self._graphs[key] = entry
        g, z_static, out = entry
        if self._stale():
            self.gen.repack_weights()               # in place: the captured graphs keep reading the same buffers
        z_static.copy_(z, non_blocking=True)
        g.replay()
        return out.clone() if clone else out
