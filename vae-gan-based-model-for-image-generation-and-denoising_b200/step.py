"""VAEGANStep - the fused training step (integration level 2).

Reproduces one iteration of the reference hot loop, vaegan_code.py:66-135, on the drop-in modules with
  * the fused loss-side kernels (reparameterisation + KL, BCE, MSE with gradient seeds; no ATen arithmetic),
  * gradients accumulated by the wgrad / BatchNorm kernels straight into flat fp32 buffers (`param.main_grad`),
  * one fused Adam launch per optimizer over its flat buffer (torch.optim.Adam defaults of vaegan_code.py:42-44),
  * the discriminator's weight gradients skipped in the generator step (the reference computes and discards them:
    vaegan_code.py:110-135 never steps opt_Dis after `total.backward()`),
  * data parallelism: one process per GPU, gradient all-reduce (NCCL) per optimizer update, 1/world folded into Adam,
  * the whole schedule captured in ONE CUDA graph and replayed per step; losses stay on the device.

    step = VAEGANStep(encoder, decoder, discriminator)        # same modules the reference loop would use
    losses = step.step(real_images, epoch)                    # dict of 0-d device tensors
"""
from __future__ import annotations

import ctypes
from typing import Dict, Optional

import torch
import torch.nn as nn

from . import functional as F_
from .dp import BucketedAllReduce, LocalAdam, PeerAdam, PeerShared
from ._lib import ACT_TANH, call
from .functional import _p, _stream

LOSS_KEYS = ("d_loss_0", "d_loss_1", "recon", "kl", "adv", "total")


class _FlatAdam:
    """All parameters of one network in one fp32 buffer + matching gradient / moment buffers."""

    def __init__(self, net: nn.Module, lr, betas, eps, symmetric: bool = False):
        """`symmetric`: parameters and gradients live in symmetric memory (peer-mapped across the ranks of the data-
        parallel group; dp.PeerAdam reduces / updates / broadcasts them with one kernel per bucket)."""
        params = [p for p in net.parameters()]
        self.names = [k for k, _ in net.named_parameters()]
        dev = params[0].device
        offs, total = [], 0
        for p in params:
            offs.append(total)
            total += (p.numel() + 3) // 4 * 4        # keep every view 16-byte aligned
        self.n = total
        self.offsets, self.sizes, self.shapes = offs, [p.numel() for p in params], [tuple(p.shape) for p in params]
        self.param_ids = [id(p) for p in params]
        if symmetric:
            import torch.distributed._symmetric_memory as symm_mem
            self.params = symm_mem.empty(total, dtype=torch.float32, device=dev).zero_()
            self.grads = symm_mem.empty(total, dtype=torch.float32, device=dev).zero_()
        else:
            self.params = torch.zeros(total, dtype=torch.float32, device=dev)
            self.grads = torch.zeros(total, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(total, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(total, dtype=torch.float32, device=dev)
        self.step_count = torch.zeros((), dtype=torch.int64, device=dev)
        for p, o in zip(params, offs):
            view = self.params[o:o + p.numel()].view_as(p)
            view.copy_(p.detach())
            p.data = view
            p.main_grad = self.grads[o:o + p.numel()].view_as(p)
            # the step zeroes `grads` before every backward pass and each layer runs once per pass: the weight-gradient
            # kernel may store instead of read-add-write (functional.ConvLayerFn.backward)
            p.grad_first_touch = True
            p.grad = None
        self.lr, self.betas, self.eps = lr, betas, eps

    def zero_grad(self):
        self.grads.zero_()      # cudaMemsetAsync

    def named_views(self, flat: torch.Tensor):
        """{parameter name: view of `flat` shaped like the parameter} for a flat buffer laid out like `self.grads`."""
        return {k: flat[o:o + n].view(shape) for k, o, n, shape in zip(self.names, self.offsets, self.sizes, self.shapes)}

    def state_dict(self):
        """Per-parameter moments in `net.parameters()` order and the step count - the content of
        torch.optim.Adam.state_dict()['state'] for the same network."""
        views = lambda flat: [flat[o:o + n].view(shape).clone() for o, n, shape in zip(self.offsets, self.sizes, self.shapes)]
        return {"step": int(self.step_count), "exp_avg": views(self.exp_avg), "exp_avg_sq": views(self.exp_avg_sq)}

    def load_state_dict(self, sd):
        for key, flat in (("exp_avg", self.exp_avg), ("exp_avg_sq", self.exp_avg_sq)):
            if len(sd[key]) != len(self.offsets):
                raise RuntimeError(f"optimizer state has {len(sd[key])} tensors, the network has {len(self.offsets)} parameters")
            for t, o, n, shape in zip(sd[key], self.offsets, self.sizes, self.shapes):
                if tuple(t.shape) != shape:
                    raise RuntimeError(f"size mismatch in optimizer state: {tuple(t.shape)} vs parameter {shape}")
                flat[o:o + n].view(shape).copy_(t)
        self.step_count.fill_(int(sd["step"]))

    def step(self, grad_scale: float):
        call("vg_adam_step", _p(self.params), _p(self.grads), _p(self.exp_avg), _p(self.exp_avg_sq), self.n,
             float(self.lr), float(self.betas[0]), float(self.betas[1]), float(self.eps), _p(self.step_count),
             float(grad_scale), _stream())


class _ReparamFn(torch.autograd.Function):
    """vaegan_code.py:75-78 + the KL prior of :114; backward adds w_kl * dKL to the gradients of mu / logvar."""

    @staticmethod
    def forward(ctx, mu, logvar, eps, kl_out, kl_weight_dev, dtype):
        B, nz = mu.shape
        z = torch.empty((B, 1, 1, nz), dtype=dtype, device=mu.device)
        call("vg_reparam_fwd", _p(mu), _p(logvar), _p(eps), B, nz, _p(z), F_._DT[dtype], _p(kl_out), _stream())
        ctx.save_for_backward(mu, logvar, eps, kl_weight_dev)
        return z

    @staticmethod
    def backward(ctx, dz):
        mu, logvar, eps, klw = ctx.saved_tensors
        B, nz = mu.shape
        dmu, dlv = torch.empty_like(mu), torch.empty_like(logvar)
        dz = dz.contiguous()
        call("vg_reparam_bwd", _p(dz), F_._DT[dz.dtype], _p(mu), _p(logvar), _p(eps), B, nz, _p(klw), 0.0, _p(dmu),
             _p(dlv), _stream())
        return dmu, dlv, None, None, None, None


class _ReconHubFn(torch.autograd.Function):
    """recon (fp32 NCHW) -> recon + sigma*n_fake as the discriminator's NHWC input (vaegan_code.py:92).
    Backward folds in the pixel-MSE term of :113, d_recon = d(adv path) + 2 (recon - real) / N, and the same launch
    writes the reconstruction loss and the step's total (:117; KL and the adversarial term are already on the device
    when the backward pass gets here).  `terms` = None in Dis_l mode (the feature tap owns the reconstruction term)."""

    @staticmethod
    def forward(ctx, recon, real, n_fake, sigma, dtype, terms, s2d_origin=None, out=None):
        """`out`: where to write (the fake half of the discriminator's stacked input batch - no copy afterwards)."""
        ctx.save_for_backward(recon, real)
        ctx.s2d_origin, ctx.terms = s2d_origin, terms
        return F_.nchw_to_nhwc(recon, dtype, aux=n_fake, mode=1, sigma=sigma, s2d_origin=s2d_origin, out=out)

    @staticmethod
    def backward(ctx, dy):
        recon, real = ctx.saved_tensors
        d_adv = F_.nhwc_to_nchw(dy.contiguous(), channels=recon.shape[1], s2d_origin=ctx.s2d_origin)
        if ctx.terms is None:
            return d_adv, None, None, None, None, None, None, None
        d_total = torch.empty_like(recon)
        ctx.terms.mse_total(recon, real, d_adv, d_total)
        return d_total, None, None, None, None, None, None, None


class _FeatureMatchFn(torch.autograd.Function):
    """Dis_l reconstruction term (README.md:11-14, eq. 2): identity on the tapped discriminator features of the
    reconstruction; backward adds 2 (f - f_real) / N to their gradient and writes the loss and the step's total."""

    @staticmethod
    def forward(ctx, feat, feat_real, terms):
        ctx.save_for_backward(feat, feat_real)
        ctx.terms = terms
        return feat.view_as(feat)

    @staticmethod
    def backward(ctx, dy):
        feat, feat_real = ctx.saved_tensors
        g = torch.empty_like(feat)
        ctx.terms.mse_total(feat, feat_real, dy.contiguous(), g)
        return g, None, None


class _LossTerms:
    """Device scalars the fused MSE + total launch reads / writes (slots of VAEGANStep's loss vector)."""

    def __init__(self, losses, kl_w, alpha_adv, ws):
        self.losses, self.kl_w, self.alpha_adv, self.ws = losses, kl_w, alpha_adv, ws

    def mse_total(self, a, b, grad_in, grad_out):
        L = self.losses
        call("vg_mse_total", _p(a), _p(b), F_._DT[a.dtype], a.numel(), 1.0, _p(grad_in), _p(grad_out), _p(L[2:3]),
             _p(L[3:4]), _p(L[4:5]), _p(self.kl_w), float(self.alpha_adv), _p(L[5:6]), _p(self.ws),
             self.ws.numel() * 4, _stream())


class VAEGANStep:
    def __init__(self, encoder, decoder, discriminator, *, lr: float = 2e-4, betas=(0.9, 0.999), eps: float = 1e-8,
                 alpha_kl: float = 0.1, alpha_adv: float = 0.1, kl_warmup_epochs: int = 50, sigma_inst: float = 0.05,
                 denoise_sigma: float = 0.0, n_dis: int = 2, real_label: float = 0.9, fake_label: float = 0.1,
                 process_group=None, use_cuda_graph: bool = True, seed: int = 0, overlap_wgrad: bool = True,
                 capture_grads: bool = False, bucket_bytes: Optional[Dict[str, int]] = None,
                 recon_mode: str = "pixel", dis_layer: int = -2, dp_transport: Optional[str] = None):
        self.E, self.G, self.D = encoder, decoder, discriminator
        self.dtype = encoder._dtype()
        self.dev = next(encoder.parameters()).device
        if self.dev.type != "cuda":
            raise F_._lib.VaeganB200Error("VAEGANStep needs the modules on a CUDA (sm_100) device; no CPU fallback")
        self.alpha_kl, self.alpha_adv, self.kl_warmup = alpha_kl, alpha_adv, kl_warmup_epochs
        self.sigma_inst, self.denoise_sigma, self.n_dis = sigma_inst, denoise_sigma, n_dis
        self.real_label, self.fake_label = real_label, fake_label
        # reconstruction term: "pixel" = MSE(recon, real) as vaegan_code.py:113 has it; "dis_l" = the feature-matching
        # form of README.md:11-14 (eq. 2) on the discriminator's conv group `dis_layer` (no reference code: parity
        # is pinned against the oracle's restatement only)
        if recon_mode not in ("pixel", "dis_l"):
            raise ValueError("recon_mode must be 'pixel' or 'dis_l'")
        self.recon_mode, self.dis_layer = recon_mode, dis_layer
        self.pg = process_group
        self.world, self.rank = 1, 0
        if process_group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()):
            self.world = torch.distributed.get_world_size(process_group)
            self.rank = torch.distributed.get_rank(process_group)
        # data-parallel transport: "peer" = one kernel per gradient bucket over NVLink / NVSwitch peer memory
        # (reduce-scatter + sharded Adam + all-gather, csrc/dp_comm.cu); "nccl" = bucketed ncclAllReduce + replicated
        # Adam.  Default: peer when symmetric memory can be set up for the group, NCCL otherwise (environment override
        # VG_DP_TRANSPORT=peer|peer-nomc|nccl; peer-nomc keeps to peer loads / stores, no multimem instructions)
        import os as _os
        transport = dp_transport or _os.environ.get("VG_DP_TRANSPORT", "auto")
        if transport not in ("auto", "peer", "peer-nomc", "nccl"):
            raise ValueError("dp_transport must be auto, peer, peer-nomc or nccl")
        self.peer = None
        if self.world > 1 and transport != "nccl":
            try:
                self.peer = PeerShared(self.dev, process_group, allow_multicast=(transport != "peer-nomc"))
            except Exception as exc:        # no peer access / no symmetric-memory support on this box
                if transport != "auto":
                    raise
                import warnings
                warnings.warn(f"vaegan_b200: symmetric memory unavailable ({exc}); gradients go through NCCL")
        sym = self.peer is not None
        self.opt_E = _FlatAdam(encoder, lr, betas, eps, symmetric=sym)
        self.opt_G = _FlatAdam(decoder, lr, betas, eps, symmetric=sym)
        self.opt_D = _FlatAdam(discriminator, lr, betas, eps, symmetric=sym)
        if self.world > 1:
            # replicas must start from the same state (DDP's constructor does the same): rank 0's parameters and
            # BatchNorm buffers go to everyone
            dist = torch.distributed
            src = dist.get_global_rank(process_group, 0) if process_group is not None else 0
            for opt in (self.opt_E, self.opt_G, self.opt_D):
                dist.broadcast(opt.params, src, group=process_group)
            for net in (encoder, decoder, discriminator):
                for b in net.buffers():
                    dist.broadcast(b, src, group=process_group)
        # channel-sum scratch of the fused convolution epilogues: owned by this step (its graph bakes the addresses in)
        self._sums = torch.zeros(F_.SumsArena.FLOATS, dtype=torch.float32, device=self.dev)
        self._sums_used = self._sums.numel()
        self.use_graph = use_cuda_graph
        self.seed = seed
        # parity tests: keep a copy of the discriminator's gradient of EVERY update (the flat buffer is re-used)
        self.capture_grads = capture_grads
        self._d_grad_copies = [torch.zeros_like(self.opt_D.grads) for _ in range(n_dis)] if capture_grads else []
        # side streams (parallel branches of the captured graph): weight gradients round-robin, plus the start-of-step
        # weight packing and noise generation that the encoder's forward pass does not wait for
        self.wgrad_streams = [torch.cuda.Stream(device=self.dev) for _ in range(3)] if overlap_wgrad else []
        # data parallel (world > 1): bucketed gradient all-reduce on a communication stream, launched from INSIDE the
        # backward passes - a bucket goes out as soon as the kernels producing its gradients have been issued
        # (functional.GradReady, reverse parameter order) and runs under the rest of the backward pass.  The
        # discriminator's 256->512 weight (76 % of its bytes) is ready first and travels under its remaining
        # dgrad / wgrad launches; the generator's 53 MB leave in three buckets under its own and the encoder's backward.
        self.comm_stream = torch.cuda.Stream(device=self.dev) if self.world > 1 else None
        bb = dict(E=2 << 20, G=8 << 20, D=4 << 20)
        if self.peer is not None:
            # a peer-memory bucket costs two cross-GPU barriers on the communication stream and nothing on the main
            # stream; only the LAST bucket of a backward pass is exposed, so it is kept small (D: first + second layer)
            bb["D"] = 2 << 20
        bb.update(bucket_bytes or {})
        self.buckets = {}
        self.peer_adam = {}
        # one GPU, experiment switch (off): networks whose Adam is issued per bucket from inside the backward pass
        # (dp.LocalAdam), VG_LOCAL_ADAM = comma list out of E,G,D.  Measured (40 steps each): none 3.633 ms, D 3.653,
        # D,E 3.646, D,E,G 3.631 - the 14 / 9 / 56 us of Adam it hides come back as contention and extra launches
        local = _os.environ.get("VG_LOCAL_ADAM", "")
        self.local_adam = set(k for k in local.split(",") if k in ("E", "G", "D")) \
            if (self.world == 1 and overlap_wgrad) else set()
        if self.local_adam:
            self.comm_stream = torch.cuda.Stream(device=self.dev)
            for key, opt in (("E", self.opt_E), ("G", self.opt_G), ("D", self.opt_D)):
                if key in self.local_adam:
                    padded = [(n + 3) // 4 * 4 for n in opt.sizes]
                    self.buckets[key] = BucketedAllReduce(opt.grads, opt.offsets, padded, process_group, bb[key],
                                                          self.comm_stream, self.wgrad_streams,
                                                          reducer=LocalAdam(opt).reduce_and_step)
        if self.world > 1:
            for key, opt in (("E", self.opt_E), ("G", self.opt_G), ("D", self.opt_D)):
                padded = [(n + 3) // 4 * 4 for n in opt.sizes]
                reducer = None
                if self.peer is not None:
                    pa = PeerAdam(opt.grads, opt.params, opt.exp_avg, opt.exp_avg_sq, opt.step_count,
                                  (lr, betas, eps), process_group, self.peer, write_grads=capture_grads)
                    self.peer_adam[key], reducer = pa, pa.reduce_and_step
                self.buckets[key] = BucketedAllReduce(opt.grads, opt.offsets, padded, process_group, bb[key],
                                                      self.comm_stream, self.wgrad_streams, reducer=reducer)
        self._graph = None
        self._static = None
        self._copy_stream = None            # prefetch(): host -> device copies of the next batch
        self._prefetched = None
        self.launches_per_step = None
        for net in (encoder, decoder, discriminator):
            net.train()

    # ------------------------------------------------------------------------------------------ buffers
    def _alloc_static(self, batch, hw, nz):
        dev = self.dev
        f32 = dict(dtype=torch.float32, device=dev)
        s = {
            "real": torch.zeros((batch, 3, hw, hw), **f32), "eps": torch.zeros((batch, nz), **f32),
            "n_real": torch.zeros((batch, 3, hw, hw), **f32), "n_fake": torch.zeros((batch, 3, hw, hw), **f32),
            "n_den": torch.zeros((batch, 3, hw, hw), **f32) if self.denoise_sigma > 0 else None,
            "kl_w": torch.zeros((), **f32), "losses": torch.zeros((len(LOSS_KEYS),), **f32),
            "rng_offset": torch.zeros((), dtype=torch.int64, device=dev),
            "mse_ws": torch.zeros((F_._lib.load().vg_mse_workspace_bytes() // 4,), **f32),
            "dp_a": torch.empty((batch,), **f32), "dp_pair": torch.empty((2 * batch,), **f32),
        }
        return s

    def _randn_into(self, t: torch.Tensor, stream_id: int):
        # (Philox stream = 16 * rank + tensor id: data-parallel replicas draw different noise from the same seed)
        call("vg_randn", _p(t), t.numel(), ctypes.c_ulonglong(self.seed), _p(self._static["rng_offset"]),
             ctypes.c_ulonglong(16 * self.rank + stream_id), _stream())

    # ------------------------------------------------------------------------------------------ the schedule
    def _run(self, gen_noise: bool):
        if self.wgrad_streams:
            F_.WgradOverlap.enable(self.wgrad_streams)
        self._sums[:self._sums_used].zero_()          # one memset per step (only what the previous run handed out)
        F_.SumsArena.activate(self._sums)
        try:
            self._run_body(gen_noise)
        finally:
            self._sums_used = max(32, F_.SumsArena.deactivate())
            F_.WgradOverlap.disable()

    def _run_body(self, gen_noise: bool):
        s = self._static
        E, G, D = self.E, self.G, self.D
        real, loss = s["real"], s["losses"]
        B = real.shape[0]
        # every replay starts from freshly packed bf16 weights (one launch per net): the encoder's are needed at once,
        # the generator's and the discriminator's are packed on the second stream under the encoder's forward pass
        sides = self.wgrad_streams
        cur = torch.cuda.current_stream()
        if not sides:
            E.repack_weights()

        def noise():                   # torch.randn_like of vaegan_code.py:77,91,92 -> Philox kernel
            self._randn_into(s["eps"], 1)
            self._randn_into(s["n_real"], 2)
            self._randn_into(s["n_fake"], 3)
            if s["n_den"] is not None:
                self._randn_into(s["n_den"], 4)

        if sides:
            for st in sides:                    # fork every side stream here: all of them belong to the capture
                st.wait_stream(cur)
            with torch.cuda.stream(sides[2]):   # the encoder's copies: needed first, packed next to the input conversion
                E.repack_weights()
            with torch.cuda.stream(sides[0]):
                G.repack_weights()
                D.repack_weights()
            if gen_noise and s["n_den"] is not None:
                noise()                         # the denoising noise feeds the encoder's input: nothing to hide it under
            elif gen_noise:
                with torch.cuda.stream(sides[1]):
                    noise()
        else:
            G.repack_weights()
            D.repack_weights()
            if gen_noise:
                noise()

        # ---- encode, reparameterise, decode                                         (:74-83)
        # image-side tensors go to the networks in the layout their first layer asks for (space-to-depth on the
        # bf16 path: 128-byte TMA rows instead of 32-byte padded pixels)
        e_fmt = E.input_s2d_origin(real.shape[2], real.shape[3])
        d_fmt = D.input_s2d_origin(real.shape[2], real.shape[3])
        # D(real_noisy) and D(recon_noisy.detach()) of :96-97 share every convolution launch: the two batches are
        # stacked, BatchNorm statistics / running-stat updates stay per batch, real first (ConvLayerFn groups=2).
        # The real half (:91) depends on nothing the encoder / generator compute: it is converted on a side stream
        # under the encoder's forward pass; the noisy reconstruction is later written straight into the second half
        shape, dt = F_.image_nhwc_shape(B, real.shape[1], real.shape[2], real.shape[3], self.dtype, d_fmt)
        pair = torch.empty((2 * B,) + tuple(shape[1:]), dtype=dt, device=self.dev)
        if sides:
            sides[1].wait_stream(cur)               # (the noise, when it was drawn on the main stream)
            with torch.cuda.stream(sides[1]):
                F_.nchw_to_nhwc(real, self.dtype, aux=s["n_real"], mode=1, sigma=self.sigma_inst, out=pair[:B],
                                s2d_origin=d_fmt)
        else:
            F_.nchw_to_nhwc(real, self.dtype, aux=s["n_real"], mode=1, sigma=self.sigma_inst, out=pair[:B],
                            s2d_origin=d_fmt)
        if self.denoise_sigma > 0:
            enc_in = F_.nchw_to_nhwc(real, self.dtype, aux=s["n_den"], mode=1, sigma=self.denoise_sigma, clamp=True,
                                     s2d_origin=e_fmt)
        else:
            enc_in = F_.nchw_to_nhwc(real, self.dtype, s2d_origin=e_fmt)
        if sides:
            cur.wait_stream(sides[2])                                # packed encoder weights
        mu, logvar = E.forward_nhwc(enc_in)
        z = _ReparamFn.apply(mu, logvar, s["eps"], loss[3:4], s["kl_w"], self.dtype)
        for st in sides[:2]:
            cur.wait_stream(st)                                  # packed generator / discriminator weights, noise
        recon = F_.ToNCHWActFn.apply(G.forward_nhwc(z), ACT_TANH, 3)

        # ---- instance noise                                                           (:88-92)
        terms = _LossTerms(loss, s["kl_w"], self.alpha_adv, s["mse_ws"])
        dis_l = self.recon_mode == "dis_l"
        recon_noisy = _ReconHubFn.apply(recon, real, s["n_fake"], self.sigma_inst, self.dtype,
                                        None if dis_l else terms, d_fmt, pair[B:])
        dp = s["dp_pair"]

        # ---- discriminator updates                                                    (:95-105)
        for it in range(self.n_dis):
            self.opt_D.zero_grad()
            p_pair = D.forward_nhwc(pair, groups=2)
            slot = loss[it:it + 1] if it < 2 else None
            call("vg_bce_pair", _p(p_pair), B, self.real_label, self.fake_label, 1.0, _p(slot), _p(dp), _stream())
            self._arm_buckets("D")
            try:
                torch.autograd.backward([p_pair], [dp])
            finally:
                F_.GradReady.clear()
            F_.WgradOverlap.join()
            if self.capture_grads and self.world == 1:
                self._d_grad_copies[it].copy_(self.opt_D.grads)
            self._finish_buckets("D")
            if self.capture_grads and self.world > 1:
                self._d_grad_copies[it].copy_(self.opt_D.grads)
            if self.peer is None and "D" not in self.local_adam:   # (else the bucket kernels have applied Adam)
                self.opt_D.step(1.0 / self.world)
            D.repack_weights()

        # ---- generator / encoder update                                               (:110-135)
        self.opt_E.zero_grad()
        self.opt_G.zero_grad()
        d_params = list(D.parameters())
        for p in d_params:             # weight gradients of D are never used in this phase (reference discards them)
            p.requires_grad_(False)
        self._arm_buckets("E", "G")
        try:
            tap = None
            if dis_l:
                with torch.no_grad():                   # Dis_l(x): its own train-mode call, like the oracle's
                    feat_real = D.features_nhwc(pair[:B], self.dis_layer)
                tap = (self.dis_layer, lambda f: _FeatureMatchFn.apply(f, feat_real, terms))
            p_fake = D.forward_nhwc(recon_noisy, tap=tap)
            call("vg_bce", _p(p_fake), B, self.real_label, self.alpha_adv, _p(loss[4:5]), 0, _p(s["dp_a"]), _stream())
            torch.autograd.backward([p_fake], [s["dp_a"]])      # (the MSE launch inside also writes recon and total)
        finally:
            F_.GradReady.clear()
            for p in d_params:
                p.requires_grad_(True)
        F_.WgradOverlap.join()
        if self.world > 1:
            # the generator's buckets left first and are (nearly) done; the encoder's tail bucket is launched now and
            # its latency (~40 us of ring all-reduce at 8 GPUs) hides under the generator's HBM-bound Adam
            self.buckets["G"].flush()
            self.buckets["E"].flush()
            self.buckets["G"].wait()             # (only for G's own last bucket: dp.BucketedAllReduce.last_event)
            if self.peer is None:
                self.opt_G.step(1.0 / self.world)
            self.buckets["E"].wait()
            if self.peer is None:
                self.opt_E.step(1.0 / self.world)
        else:
            self._finish_buckets("G", "E")
            if "E" not in self.local_adam:
                self.opt_E.step(1.0)
            if "G" not in self.local_adam:
                self.opt_G.step(1.0)
        self._last = dict(mu=mu.detach(), logvar=logvar.detach(), recon=recon.detach())

    def _arm_buckets(self, *keys):
        """Register dp.BucketedAllReduce.mark_ready for every parameter of the named networks: the layer backward
        that has just issued a parameter's gradient kernels calls it (functional.GradReady)."""
        handlers = {}
        for key in keys:
            if key not in self.buckets:
                continue
            opt, bk = {"E": self.opt_E, "G": self.opt_G, "D": self.opt_D}[key], self.buckets[key]
            bk.reset()
            for i, pid in enumerate(opt.param_ids):
                handlers[pid] = (lambda bk=bk, i=i: bk.mark_ready(i))
        F_.GradReady.set(handlers)

    def _finish_buckets(self, *keys):
        """Launch what the backward pass did not trigger and make the current stream wait for the communication."""
        for key in keys:
            if key in self.buckets:
                self.buckets[key].finish()

    # ------------------------------------------------------------------------------------------ public API
    def step(self, real: torch.Tensor, epoch: int, eps: Optional[torch.Tensor] = None,
             n_real: Optional[torch.Tensor] = None, n_fake: Optional[torch.Tensor] = None,
             n_denoise: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        """One training step on `real` (fp32 NCHW in [-1, 1], device or pinned host tensor).  Noise tensors may be
        injected (parity tests) - otherwise they are drawn on the device.  Returns 0-d device tensors."""
        u8 = real.dtype == torch.uint8          # decoded images [B, H, W, 3]: normalised on the device
        if u8:
            batch, hw, _, ch = real.shape
            if ch != 3:
                raise RuntimeError(f"uint8 input must be NHWC with 3 channels, got {list(real.shape)}")
            shape = (batch, 3, hw, hw)
        else:
            batch, _, hw, _ = real.shape
            shape = tuple(real.shape)
        nz = self.E.fc_mu.out_features
        injected = eps is not None
        if self._static is None or tuple(self._static["real"].shape) != shape:
            old_rng = int(self._static["rng_offset"]) if self._static is not None else getattr(self, "_pending_rng", 0)
            self._static = self._alloc_static(batch, hw, nz)
            self._static["rng_offset"].fill_(old_rng)
            self._graph = None
        s = self._static
        staged = self._take_prefetched(real)
        if u8:
            src = staged
            if src is None:
                if s.get("real_u8") is None or s["real_u8"].shape != real.shape:
                    s["real_u8"] = torch.empty(real.shape, dtype=torch.uint8, device=self.dev)
                s["real_u8"].copy_(real, non_blocking=True)
                src = s["real_u8"]
            call("vg_u8_nhwc_to_nchw", _p(src), _p(s["real"]), batch, 3, hw, hw, 0.5, 0.5, _stream())
        else:
            s["real"].copy_(real if staged is None else staged, non_blocking=True)
        if staged is not None:
            self._prefetched[3].record(torch.cuda.current_stream())    # the staging buffer may be refilled after this
        s["kl_w"].fill_(self.alpha_kl * min(1.0, epoch / self.kl_warmup) if self.kl_warmup > 0 else self.alpha_kl)
        if injected:
            s["eps"].copy_(eps, non_blocking=True)
            s["n_real"].copy_(n_real, non_blocking=True)
            s["n_fake"].copy_(n_fake, non_blocking=True)
            if s["n_den"] is not None:
                s["n_den"].copy_(n_denoise, non_blocking=True)
        if not self.use_graph:
            lib = F_._lib.load()
            n0 = lib.vg_launch_count()
            self._run(gen_noise=not injected)
            self.launches_per_step = int(lib.vg_launch_count() - n0)
        else:
            key = (not injected)
            if self._graph is None or self._graph[0] != key:
                # warm-up run outside capture (allocator, lazy module plans), then capture
                # (measured and dropped: capturing the main chain on a high-priority stream so that it wins SMs over the
                # weight-gradient side streams - 3.50-3.54 ms against 3.42-3.44 ms on the same box)
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    self._snapshot = self._save_state()
                    self._run(gen_noise=key)
                torch.cuda.current_stream().wait_stream(side)
                torch.cuda.synchronize()
                self._restore_state(self._snapshot)
                g = torch.cuda.CUDAGraph()
                lib = F_._lib.load()
                n0 = lib.vg_launch_count()
                with torch.cuda.graph(g):
                    self._run(gen_noise=key)
                self.launches_per_step = int(lib.vg_launch_count() - n0)   # kernel nodes of ours in the graph
                self._graph = (key, g)
                self._restore_state(self._snapshot)
                self._snapshot = None
            self._graph[1].replay()
        # The encoder's and the generator's masters were updated by the fused Adam at the end of the step, behind
        # autograd's version counters; their bf16 copies are re-packed at the start of the NEXT step.  Anything that
        # uses the modules in between (validation, generation: vaegan_code.py:147-171, main_vae.py:348-374) must not
        # see the stale copies: mark them so that the next module call re-packs (in place).
        self.E.invalidate_packed_weights()
        self.G.invalidate_packed_weights()
        return {k: s["losses"][i] for i, k in enumerate(LOSS_KEYS)}

    # ------------------------------------------------------------------------------------------ loss read-back
    def losses_lagged(self) -> Optional[Dict[str, float]]:
        """Host values of the losses WITHOUT stalling the device: enqueues an asynchronous device -> host copy of the
        step that was just launched (24 bytes into a pinned double buffer) and returns the PREVIOUS step's six losses
        as Python floats (None after the first step).  The `.item()` reads of vaegan_code.py:119-127 drain the queue
        every step - the host then prepares the next launch while the GPU idles; read one step late and the next
        replay is already queued when the host blocks.  `losses_flush()` returns the last step's values."""
        if self._static is None:
            return None
        if getattr(self, "_lag", None) is None:
            self._lag = {"buf": [torch.empty(len(LOSS_KEYS), dtype=torch.float32).pin_memory() for _ in range(2)],
                         "ev": [torch.cuda.Event(), torch.cuda.Event()], "cur": 0, "have": False}
        lag = self._lag
        prev = None
        if lag["have"]:
            j = lag["cur"] ^ 1
            lag["ev"][j].synchronize()
            prev = {k: float(lag["buf"][j][i]) for i, k in enumerate(LOSS_KEYS)}
        i = lag["cur"]
        lag["buf"][i].copy_(self._static["losses"], non_blocking=True)
        lag["ev"][i].record(torch.cuda.current_stream())
        lag["cur"], lag["have"] = i ^ 1, True
        return prev

    def losses_flush(self) -> Optional[Dict[str, float]]:
        lag = getattr(self, "_lag", None)
        if lag is None or not lag["have"]:
            return None
        j = lag["cur"] ^ 1
        lag["ev"][j].synchronize()
        lag["have"] = False
        return {k: float(lag["buf"][j][i]) for i, k in enumerate(LOSS_KEYS)}

    # ------------------------------------------------------------------------------------------ input prefetch
    def prefetch(self, real: torch.Tensor) -> None:
        """Start the host -> device copy of the NEXT step's batch (pinned host tensor) on a copy stream, so that it
        runs under the step that is executing now; the next `step(real)` called with this same tensor picks the
        staged copy up (a device-to-device move of ~12 MB) instead of copying from the host on the compute stream.
        The data loader's double buffering (DataLoader(pin_memory=True) + non_blocking copies, main_vae.py:204-209)."""
        if real.device.type != "cpu":
            return
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.dev)
        prev = self._prefetched
        if prev is not None and prev[1].shape == real.shape and prev[1].dtype == real.dtype:
            stage, ready, consumed = prev[1], prev[2], prev[3]
        else:
            stage = torch.empty(real.shape, dtype=real.dtype, device=self.dev)
            ready, consumed = torch.cuda.Event(), torch.cuda.Event()
            consumed.record(torch.cuda.current_stream())
        self._copy_stream.wait_event(consumed)          # the previous consumer of the staging buffer has read it
        with torch.cuda.stream(self._copy_stream):
            stage.copy_(real, non_blocking=True)
            ready.record(self._copy_stream)
        self._prefetched = ((real.data_ptr(), tuple(real.shape), real.dtype, real._version), stage, ready, consumed)

    def _take_prefetched(self, real: torch.Tensor):
        pf = self._prefetched
        if pf is None or real.device.type != "cpu":
            return None
        if pf[0] != (real.data_ptr(), tuple(real.shape), real.dtype, real._version):
            return None
        torch.cuda.current_stream().wait_event(pf[2])
        return pf[1]

    # state save / restore so that warm-up + capture do not advance training
    def _save_state(self):
        st = {"opt": [(o.params.clone(), o.exp_avg.clone(), o.exp_avg_sq.clone(), o.step_count.clone())
                      for o in (self.opt_E, self.opt_G, self.opt_D)],
              "buf": [[b.clone() for b in net.buffers()] for net in (self.E, self.G, self.D)],
              "rng": self._static["rng_offset"].clone()}
        return st

    def _restore_state(self, st):
        for o, (p, m, v, c) in zip((self.opt_E, self.opt_G, self.opt_D), st["opt"]):
            o.params.copy_(p); o.exp_avg.copy_(m); o.exp_avg_sq.copy_(v); o.step_count.copy_(c)
        for net, bufs in zip((self.E, self.G, self.D), st["buf"]):
            for b, saved in zip(net.buffers(), bufs):
                b.copy_(saved)
        self._static["rng_offset"].copy_(st["rng"])

    # checkpoint / resume (vaegan_code.py:193 saves only the decoder; a resumable run also needs the optimizer state)
    def state_dict(self):
        """State of the fused step beyond the three modules' own (reference-keyed) state_dicts: the three Adam states
        and the device noise counter.  Save it next to encoder / decoder / discriminator .state_dict()."""
        rng = int(self._static["rng_offset"]) if self._static is not None else getattr(self, "_pending_rng", 0)
        if self.peer is not None:
            self.peer.check()                      # a cross-GPU barrier that timed out would have corrupted the step
        for key, pa in self.peer_adam.items():     # sharded optimizer state: collect every slice from its owner
            pa.gather_moments(self.buckets[key].buckets)
        return {"opt_E": self.opt_E.state_dict(), "opt_G": self.opt_G.state_dict(), "opt_D": self.opt_D.state_dict(),
                "rng_offset": rng, "seed": self.seed}

    def load_state_dict(self, sd):
        """Load after the modules' own load_state_dict (their parameters are views of this step's flat buffers, so that
        copy lands in place)."""
        self.opt_E.load_state_dict(sd["opt_E"])
        self.opt_G.load_state_dict(sd["opt_G"])
        self.opt_D.load_state_dict(sd["opt_D"])
        if sd.get("seed", self.seed) != self.seed:
            self.seed = sd["seed"]
            self._graph = None          # the seed is a launch argument baked into the captured graph
        self._pending_rng = int(sd.get("rng_offset", 0))
        if self._static is not None:
            self._static["rng_offset"].fill_(self._pending_rng)

    def check_comm(self) -> None:
        """Peer transport: raise if a cross-GPU barrier of the optimizer kernels ever timed out (reads one device
        integer, i.e. synchronises - call it at checkpoints / epoch ends, not every step)."""
        if self.peer is not None:
            self.peer.check()

    def gradients(self):
        """Gradients of the most recent step as {name: tensor} per network: "E" / "G" (what their Adam consumed, i.e.
        after the all-reduce, before the 1/world scale) and "D" = a list with one dict per discriminator update (all of
        them with `capture_grads=True`, else only the last).  Views of the flat buffers: clone to keep.  With the peer
        transport the summed gradient is only written back with `capture_grads=True`; otherwise these are the rank's
        own, un-reduced gradients (the sum exists in registers of the optimizer kernel only)."""
        d = [self.opt_D.named_views(c) for c in self._d_grad_copies] or [self.opt_D.named_views(self.opt_D.grads)]
        return {"E": self.opt_E.named_views(self.opt_E.grads), "G": self.opt_G.named_views(self.opt_G.grads), "D": d}

    def last_outputs(self):
        """mu, logvar, recon of the most recent step (device tensors; static under CUDA-graph replay)."""
        return self._last
