#!/usr/bin/env python
"""bench.py - VAE-GAN training-step throughput (BASELINE.json metric) on B200.

    python bench.py --gpus N --steps K --warmup W                  # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K --warmup W # the reference's CPU path (oracle port), rank 0

Workload (config.workload = "cfg2"): the 64x64-derived VAE-GAN, latent 128, batch 256 per GPU,
bf16 tensor-core mode, synthetic CelebA-shaped data (uniform [-1,1] images), one step = one iteration of the
reference hot loop vaegan_code.py:66-135 (encode, decode, 2 discriminator updates, generator/encoder update, 3 Adam).
Prints ONE JSON line (rank 0).  `value` = images/s with inputs resident in HBM; `e2e` = the same step through the
public API from pinned host memory with the losses read back every step.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

HW, NZ, BATCH_PER_GPU = 64, 128, 256
METRIC = "VAE-GAN train imgs/sec, 64x64 b256, 1/2/4/8 B200; conv tensor-pipe % of peak"     # BASELINE.json "metric", verbatim
STEP_GFLOP_PER_IMAGE = 6.245          # algorithmic conv+linear FLOPs of one step / image


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            p = json.load(f)
        return dict(burst=p["bf16_tflops"], sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                    hbm=p["hbm_gbs"], source="measured (MEASURED_PEAKS.json)")
    return dict(burst=1590.0, sustained=1400.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


# ------------------------------------------------------------------------------------------------ CPU side
def _cpu_step_rate(batch: int, timed: int, warm: int = 1, budget_s: float = 150.0):
    """images/s of the oracle's reference_step (the reference loop on the reference's own torch CPU arithmetic).
    Stops early (at least one timed step) once `budget_s` of wall clock is spent, so a slow host cannot stall a run."""
    import torch
    from oracle import vaegan_oracle as vo
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    nets = vo.build_nets(vo.NetConfig(hw=HW, nz=NZ))
    opts = vo.make_optimizers(*nets)
    times = []
    t_begin = time.perf_counter()
    for it in range(warm + timed):
        real, eps, n_real, n_fake = vo.make_inputs(batch, HW, NZ, seed=42 + it)
        t0 = time.perf_counter()
        vo.reference_step(*nets, *opts, real, 50, eps, n_real, n_fake, keep_grads=False)
        dt = time.perf_counter() - t0
        if it >= warm:
            times.append(dt)
            if time.perf_counter() - t_begin > budget_s:
                break
    return batch / statistics.median(times), statistics.median(times), cores, len(times)


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    sample_batch = BATCH_PER_GPU
    # each "step" is one full cfg-2 step of ONE GPU's batch (256 images, same nets, fp32) on all host cores: about
    # 1.3 s per step on the box's 16 cores - exactly --steps timed steps after --warmup untimed ones (a wall-clock
    # guard of 240 s ends the loop early on a slow host; `steps` then reports what was timed)
    steps, warm = max(1, args.steps), max(0, args.warmup)
    rate, sec, cores, steps = _cpu_step_rate(sample_batch, timed=steps, warm=warm, budget_s=240.0)
    line = {
        "impl": "reference", "metric": METRIC, "value": rate,
        "unit": "images/s", "n_gpus": args.gpus, "steps": steps, "warmup": warm,
        "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "cfg2", "nets": "64x64-derived VAE-GAN (SURVEY A.1), latent 128",
                   "batch_per_gpu": BATCH_PER_GPU, "global_batch": BATCH_PER_GPU * args.gpus,
                   "parallelism": f"dp{args.gpus}", "measured_batch": sample_batch,
                   "note": "one process on the host cores steps one GPU's batch; reference CPU arithmetic (torch CPU/oneDNN) on the oracle's line-by-line restatement of "
                           "vaegan_code.py:66-135"},
        "cpu_baseline": {"value": rate, "unit": "images/s", "cores": cores, "kind": "port",
                         "sample": f"{steps} timed steps of {sample_batch} images (one GPU's full cfg-2 batch per step)"},
        "e2e": {"value": rate, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi polled every 20 ms in the background; stop(t0, t1) keeps the samples whose timestamp lies inside
    the measured window [t0, t1] (datetime.now() taken after a device synchronise at both ends)."""
    FIELDS = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self, t0=None, t1=None):
        import datetime
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, reasons, seen = [], [], set(), 0
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in out.strip().splitlines():
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 8:
                continue
            seen += 1
            try:
                ts = datetime.datetime.strptime(parts[0], "%Y/%m/%d %H:%M:%S.%f")
                if t0 is not None and not (t0 <= ts <= t1):
                    continue
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples inside the timed window"],
                    "samples": 0, "samples_total": seen}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm), "samples_total": seen,
                "window": "resident loop + e2e loop (the same step under continuous load)"}


# ------------------------------------------------------------------------------------------------ GPU arm
def run_gpu_arm(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - this repo has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("TORCH_NCCL_ASYNC_ERROR_HANDLING", "0")   # NCCL collectives are captured in the CUDA graph
        dist.init_process_group("nccl", device_id=dev)

    import vaegan_b200 as vb
    from importlib import import_module
    VAEGANStep = import_module("vaegan_b200.step").VAEGANStep
    lib = vb.load_library()

    torch.manual_seed(42)                       # identical initial weights on every rank
    enc = vb.Encoder([3, HW, HW], NZ, precision="bf16")
    gen = vb.Generator(nz=NZ, hw=HW, precision="bf16")
    dis = vb.Discriminator(hw=HW, precision="bf16")
    gen.apply(vb.weights_init)
    dis.apply(vb.weights_init)
    for m in (enc, gen, dis):
        m.to(dev)
    step = VAEGANStep(enc, gen, dis, use_cuda_graph=not args.no_graph, seed=1234 + rank)

    B = BATCH_PER_GPU
    g = torch.Generator(device="cpu").manual_seed(1000 + rank)       # a different data shard per rank
    n_pool = 4
    host_pool = [(torch.rand(B, 3, HW, HW, generator=g) * 2 - 1).pin_memory() for _ in range(n_pool)]
    dev_pool = [h.to(dev) for h in host_pool]
    epoch = 50                                   # KL weight fully warmed up (vaegan_code.py:117)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    import datetime
    sampler = ClockSampler(local_rank) if rank == 0 else None      # polling starts now; samples are windowed later
    # ---- warm-up (includes CUDA-graph capture), then count kernel launches of one step
    for i in range(max(3, args.warmup)):
        step.step(dev_pool[i % n_pool], epoch)
    torch.cuda.synchronize()
    launches_per_step = step.launches_per_step

    # ---- value: inputs resident in HBM, K steps, CUDA events, max over ranks
    barrier()
    t_window0 = datetime.datetime.now()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        losses = step.step(dev_pool[i % n_pool], epoch)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t)
    ms_per_step = ms_total / args.steps
    value = world * B * args.steps / (ms_total * 1e-3)
    final_total = float(losses["total"])

    # ---- e2e: same step through the public API from pinned host memory, losses read back every step
    # (the next batch's host -> device copy is started on the copy stream right after a step is launched - the data
    # loader's double buffering - and therefore runs under that step; every batch still crosses PCIe inside the loop)
    step.prefetch(host_pool[0])                          # warm-up through the very code path that is timed (copy
    for i in range(3):                                   # stream, staging buffer, pinned read-back buffers: one-time
        step.step(host_pool[i % n_pool], epoch)          # allocations that do not belong to a steady-state step)
        step.prefetch(host_pool[(i + 1) % n_pool])
        step.losses_lagged()
    step.losses_flush()
    barrier()
    t0 = time.perf_counter()
    step.prefetch(host_pool[0])
    host_losses = None
    for i in range(args.steps):
        losses = step.step(host_pool[i % n_pool], epoch)
        step.prefetch(host_pool[(i + 1) % n_pool])
        # device -> host read of every step's result: the six losses go to pinned memory asynchronously and are read
        # on the host one step late (VAEGANStep.losses_lagged), so the next replay is queued before the host blocks
        host_losses = step.losses_lagged() or host_losses
    host_losses = step.losses_flush() or host_losses     # the last step's values, still inside the timed region
    barrier()
    e2e_s = time.perf_counter() - t0
    assert host_losses is not None and host_losses["total"] == host_losses["total"]
    clocks = sampler.stop(t_window0, datetime.datetime.now()) if sampler is not None else None
    t = torch.tensor([e2e_s], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * B * args.steps / float(t)

    # ---- dominant kernel (igemm_fprop_kernel: every forward / dgrad contraction of the step), timed live: one more
    # step is run launch by launch (no graph) with a CUDA-event pair around each tensor-core launch on its stream.
    # Every rank runs it (the step contains the gradient all-reduce).
    dom = _dominant_kernel(torch, step, dev_pool[0], epoch, ms_per_step)

    # ---- e2e with the images shipped as decoded uint8 NHWC (1 byte per value) and normalised on the device
    u8_pool = [((h.permute(0, 2, 3, 1) * 0.5 + 0.5) * 255.0).round().to(torch.uint8).contiguous().pin_memory()
               for h in host_pool]
    step.prefetch(u8_pool[0])
    for i in range(3):
        step.step(u8_pool[i % n_pool], epoch)
        step.prefetch(u8_pool[(i + 1) % n_pool])
        step.losses_lagged()
    step.losses_flush()
    barrier()
    t0 = time.perf_counter()
    step.prefetch(u8_pool[0])
    for i in range(args.steps):
        losses = step.step(u8_pool[i % n_pool], epoch)
        step.prefetch(u8_pool[(i + 1) % n_pool])
        step.losses_lagged()
    step.losses_flush()
    barrier()
    t = torch.tensor([time.perf_counter() - t0], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_u8_value = world * B * args.steps / float(t)

    if world == 1:
        transport = None
    elif step.peer is None:
        transport = "nccl all-reduce per gradient bucket + replicated Adam"
    else:
        transport = ("peer-memory kernel per gradient bucket (reduce-scatter + sharded Adam + all-gather, "
                     + ("multimem instructions over the NVSwitch multicast mapping)" if step.peer_adam["G"].multicast
                        else "peer loads / stores)"))
        step.peer.check()
    extra = None
    del step, losses
    torch.cuda.empty_cache()
    if not args.no_extra:
        extra = _extra_configs(torch, dist, vb, VAEGANStep, dev, rank, world, barrier, min(args.steps, 10))

    if rank != 0:
        _finish(world)
        return 0

    peaks = _peaks()
    achieved_tflops = STEP_GFLOP_PER_IMAGE * B / ms_per_step          # GFLOP / ms == TFLOP/s, per GPU
    kernels = _kernel_microbench(torch, vb, dev) if not args.no_micro else None
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        rate, sec, cores, n_timed = _cpu_step_rate(BATCH_PER_GPU, timed=3, warm=1, budget_s=60.0)
        cpu = {"value": rate, "unit": "images/s", "cores": cores, "kind": "port",
               "sample": f"{n_timed} timed steps (median) of {BATCH_PER_GPU} images (the cfg-2 batch), fp32, same nets, "
                         "oracle restatement of vaegan_code.py:66-135 on torch CPU"}
    line = {
        "metric": METRIC, "value": value, "unit": "images/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "cfg2", "nets": "64x64-derived VAE-GAN (SURVEY A.1), latent 128",
                   "batch_per_gpu": B, "global_batch": B * world, "parallelism": f"dp{world}", "dp_transport": transport,
                   "cuda_graph": not args.no_graph,
                   "l2_policy": "per-step working set (~3 GB of activations/gradients) far exceeds the 126 MB L2; "
                                "4 rotating input batches"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": B * 3 * HW * HW * 4,
                "d2h_bytes_per_step": 24,
                "what": "VAEGANStep.step(pinned fp32 host batch) every step, the next batch prefetched on a copy stream "
                        "under the running step; all six losses of every step copied to pinned host memory and read on "
                        "the host one step late (VAEGANStep.losses_lagged), the last step's inside the timed region"},
        "e2e_uint8_input": {"value": e2e_u8_value, "unit": "images/s", "h2d_bytes_per_step": B * 3 * HW * HW,
                            "d2h_bytes_per_step": 24,
                            "what": "same step, host batch = decoded uint8 NHWC images, ToTensor+Normalize on the device"},
        "gpu_launches": int(launches_per_step * args.steps) if launches_per_step else None,
        "gpu_launches_per_step": launches_per_step,
        "roofline": {"bound": "tensor", "achieved": dom["tflops"], "peak": peaks["sustained"], "unit": "TFLOP/s",
                     "frac": dom["tflops"] / peaks["sustained"], "traffic": dom["traffic"],
                     "kernel": "igemm_fprop_kernel", "launches_per_step": dom["launches"],
                     "gflop_per_launch": dom["gflop_per_launch"], "us_per_launch": dom["us_per_launch"],
                     "algorithmic_bytes_per_launch": dom["algorithmic_bytes_per_launch"],
                     "share_of_step": dom["share"], "share_basis": "sum of this kernel's event-bracketed launch times / "
                     "sum over all %d C-ABI calls of the step (%.2f ms %s; the graph overlaps the wgrad "
                     "streams and replays in %.2f ms)" % (dom["timed_calls"], dom["timed_ms"],
                                                         "serialised on one stream" if dom["serialised"] else
                                                         "with the wgrad streams overlapping", dom["graph_ms_per_step"]),
                     "traffic_source": dom["traffic_source"],
                     "what": "dominant kernel = the tcgen05 implicit-GEMM kernel behind every forward / dgrad "
                             "contraction: algorithmic FLOPs (2 x MACs over the valid channels) of its launches in one "
                             "step / their summed durations (CUDA events around each launch on its stream, eager "
                             "replay of the same step"
                             + (", all kernels on one stream so that a launch is timed alone" if dom["serialised"] else
                                ", wgrad streams overlapping as in the graph")
                             + "; `in_step_overlapped` = the same launches while a weight-gradient kernel shares the "
                             "SMs); peak = sustained bf16, " + peaks["source"],
                     "in_step_overlapped": dom["overlapped"],
                     "per_launch": dom["per_launch"],
                     "wgrad_kernel": dom["wgrad"],
                     "whole_step": {"achieved": achieved_tflops, "frac": achieved_tflops / peaks["sustained"],
                                    "what": "6.245 GFLOP/image x batch / step time"},
                     "kernels": kernels},
        "cpu_baseline": cpu,
        "other_configs": extra,
        "final_total_loss": final_total,
    }
    print(json.dumps(line), flush=True)
    _finish(world)
    return 0


# ------------------------------------------------------------------------------------------------ configs 3, 4, 5
CFG4_GFLOP_PER_IMAGE = 33.49          # 128x128 nets, double width, latent 256


def _time_step(torch, dist, vb, VAEGANStep, dev, rank, world, barrier, *, hw, nz, width, batch, sigma, steps):
    """ms/step (max over ranks) of the graph-replayed bf16 step for one net / batch configuration, inputs resident."""
    torch.manual_seed(42)
    enc = vb.Encoder([3, hw, hw], nz, width=width, precision="bf16")
    gen = vb.Generator(nz=nz, ngf=64 * width, hw=hw, precision="bf16")
    dis = vb.Discriminator(ndf=64 * width, hw=hw, precision="bf16")
    gen.apply(vb.weights_init)
    dis.apply(vb.weights_init)
    for m in (enc, gen, dis):
        m.to(dev)
    step = VAEGANStep(enc, gen, dis, use_cuda_graph=True, seed=4321 + rank, denoise_sigma=sigma)
    g = torch.Generator(device="cpu").manual_seed(2000 + rank)
    pool = [(torch.rand(batch, 3, hw, hw, generator=g) * 2 - 1).to(dev) for _ in range(2)]
    for i in range(3):
        step.step(pool[i % 2], 50)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        losses = step.step(pool[i % 2], 50)
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total = float(losses["total"])
    del step, enc, gen, dis, pool
    torch.cuda.empty_cache()
    return float(t) / steps, total


def _generation_sweep(torch, vb, dev):
    """cfg 5: decoder-only generation (main_vae.py:360-366: eval-mode decoder under no_grad on z ~ N(0, I)) through
    vaegan_b200.GraphedGenerator (one CUDA graph per batch size), batch 1 ... 4096, latency and images/s per batch;
    the reference decoder on the host cores beside it at three batch sizes (oracle port, bounded)."""
    peaks = _peaks()
    torch.manual_seed(42)
    gen = vb.Generator(nz=NZ, hw=HW, precision="bf16")
    gen.apply(vb.weights_init)
    gen.to(dev).eval()
    gg = vb.GraphedGenerator(gen)
    rows = []
    gflop_per_image = 2 * 0.41805        # forward MACs of the 64x64 generator: 0.836 GFLOP / image
    for b in (1, 4, 16, 64, 256, 1024, 4096):
        z = torch.randn(b, NZ, 1, 1, device=dev)
        for _ in range(3):
            gg(z)
        torch.cuda.synchronize()
        best = None
        for _ in range(3):
            n = 50 if b <= 256 else 10
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n):
                out = gg(z)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / n
            best = ms if best is None else min(best, ms)
        rows.append({"batch": b, "ms": best, "images_per_s": b / (best * 1e-3),
                     "tflops": gflop_per_image * b / best, "frac_of_bf16_peak": gflop_per_image * b / best / peaks["sustained"]})
    cpu = []
    try:
        from oracle import vaegan_oracle as vo
        torch.set_num_threads(os.cpu_count() or 1)
        g_ref = vo.make_generator(nz=NZ, ngf=64, hw=HW).eval()
        for b in (1, 64, 1024):
            z = torch.randn(b, NZ, 1, 1)
            vo.generate(g_ref, z)
            t0 = time.perf_counter()
            reps = 3
            for _ in range(reps):
                vo.generate(g_ref, z)
            dt = (time.perf_counter() - t0) / reps
            cpu.append({"batch": b, "ms": dt * 1e3, "images_per_s": b / dt})
    except Exception as e:      # the CPU leg is a reported baseline, never a reason to lose the GPU numbers
        cpu = [{"error": repr(e)}]
    return {"workload": "cfg5", "what": "eval-mode generator forward z -> 64x64 image (fp32 NCHW out), bf16 tensor-core "
            "path, one CUDA graph per batch size, z resident, best of 3 x n replays", "sweep": rows,
            "cpu_reference_decoder": {"cores": os.cpu_count() or 1, "kind": "port", "rows": cpu}}


def _extra_configs(torch, dist, vb, VAEGANStep, dev, rank, world, barrier, steps):
    """BASELINE.json configs 3, 4 and 5 next to the headline cfg-2 line (same timing rules: warm-up 3, CUDA events,
    max over ranks, inputs resident and far larger than L2)."""
    peaks = _peaks()
    out = {}
    ms, total = _time_step(torch, dist, vb, VAEGANStep, dev, rank, world, barrier, hw=HW, nz=NZ, width=1,
                           batch=BATCH_PER_GPU, sigma=0.1, steps=steps)
    out["cfg3"] = {"workload": "cfg3", "what": "denoising mode: encoder input = clamp(real + 0.1 * noise, -1, 1), drawn "
                   "on the device every step; reconstruction target stays real", "batch_per_gpu": BATCH_PER_GPU,
                   "global_batch": BATCH_PER_GPU * world, "n_gpus": world, "ms_per_step": ms, "steps": steps,
                   "images_per_s": world * BATCH_PER_GPU / (ms * 1e-3), "final_total_loss": total,
                   "whole_step_frac_of_bf16_peak": STEP_GFLOP_PER_IMAGE * BATCH_PER_GPU / ms / peaks["sustained"]}
    # cfg 4: global batch 512 over 2 / 4 / 8 GPUs; a single GPU runs the 8-GPU shard (64 images)
    b4 = 512 // world if world >= 2 else 64
    ms, total = _time_step(torch, dist, vb, VAEGANStep, dev, rank, world, barrier, hw=128, nz=256, width=2, batch=b4,
                           sigma=0.0, steps=steps)
    out["cfg4"] = {"workload": "cfg4", "what": "128x128 nets, 2x channel width, latent 256" +
                   ("" if world >= 2 else "; ONE GPU running the per-GPU shard of the 8-GPU case (64 images)"),
                   "batch_per_gpu": b4, "global_batch": b4 * world, "n_gpus": world, "ms_per_step": ms, "steps": steps,
                   "images_per_s": world * b4 / (ms * 1e-3), "final_total_loss": total,
                   "whole_step_frac_of_bf16_peak": CFG4_GFLOP_PER_IMAGE * b4 / ms / peaks["sustained"]}
    if rank == 0 and world == 1:
        out["cfg5"] = _generation_sweep(torch, vb, dev)
    return out


def _finish(world: int):
    """Tear the NCCL communicator down.  destroy_process_group() hangs while a CUDA graph that holds captured NCCL
    kernels is alive (scripts/nccl_teardown_probe.py: > 30 s with the graph, 0.7 s once it is released), so every
    VAEGANStep - and with it its graph - is dropped first.  A watchdog still ends the process if teardown stalls."""
    if world <= 1:
        return
    import gc
    import threading
    import torch
    import torch.distributed as dist
    sys.stdout.flush()
    sys.stderr.flush()

    def bail():
        time.sleep(60)
        sys.stderr.write("bench.py: NCCL teardown stalled for 60 s, exiting without it\n")
        sys.stderr.flush()
        os._exit(0)

    threading.Thread(target=bail, daemon=True).start()
    gc.collect()
    torch.cuda.synchronize()
    dist.destroy_process_group()


def _dominant_kernel(torch, step, batch, epoch, graph_ms_per_step):
    """Every C-ABI call of one eagerly replayed step bracketed by CUDA events on its launching stream
    (vaegan_b200._lib.LaunchTimer); the tcgen05 fprop-type kernel's launches are summed for the roofline.
    Two passes: (a) every kernel on ONE stream - a launch's event pair then brackets that kernel alone, back to back
    with its neighbours of the step (what the roofline fraction of a KERNEL means; only at world == 1, the bucketed
    all-reduce needs its streams), (b) the product schedule, weight gradients on their side streams - the same launches
    while they share the SMs with a concurrent wgrad kernel (what each launch costs on the step's critical path)."""
    from importlib import import_module
    lib = import_module("vaegan_b200._lib")

    def one_pass():
        step.step(batch, epoch)                      # eager warm-up of this code path
        torch.cuda.synchronize()
        # the host needs ~10-15 us per C-ABI call + event pair, more than the small kernels run: park the stream behind
        # a ~10 ms spin so that every launch of the step is queued before the first one executes and an event pair
        # brackets GPU time only
        torch.cuda._sleep(20_000_000)
        lib.LaunchTimer.start()
        step.step(batch, epoch)
        recs = lib.LaunchTimer.stop()
        all_ms = sum(r[-1] for r in recs)
        out = {"timed_calls": len(recs), "timed_ms": all_ms}
        for kind in ("fprop", "wgrad"):
            sel = [(f, nb, ms) for _, tag, f, nb, ms in recs if tag == kind]
            n, flops, nb, ms = len(sel), sum(r[0] for r in sel), sum(r[1] for r in sel), sum(r[2] for r in sel)
            out[kind] = {"launches": n, "gflop_per_launch": flops / n / 1e9, "us_per_launch": ms / n * 1e3,
                         "tflops": flops / (ms * 1e-3) / 1e12, "ms_per_step": ms,
                         "algorithmic_bytes_per_launch": nb / n, "share": ms / all_ms}
        out["per_launch"] = [{"call": name, "kind": tag, "gflop": round(f / 1e9, 3), "us": round(ms * 1e3, 2),
                              "tflops": round(f / (ms * 1e-3) / 1e12, 1)}
                             for name, tag, f, nb, ms in recs if tag in ("fprop", "wgrad")]
        return out

    was_graph, sides, local, buckets = step.use_graph, step.wgrad_streams, step.local_adam, step.buckets
    step.use_graph = False
    try:
        overlapped = one_pass()
        alone = None
        if step.world == 1:
            # one stream for everything: no weight-gradient side streams, Adam as one launch after the backward pass
            step.wgrad_streams, step.local_adam, step.buckets = [], set(), {}
            alone = one_pass()
    finally:
        step.use_graph, step.wgrad_streams, step.local_adam, step.buckets = was_graph, sides, local, buckets
    traffic, src = None, "no ncu capture committed"
    path = os.path.join(ROOT, "profiles", "fprop_traffic.json")
    if os.path.isfile(path):
        with open(path) as f:
            t = json.load(f)
        traffic, src = t["dram_bytes_per_launch"], t["source"]
    main = alone or overlapped
    d = main["fprop"]
    return {"tflops": d["tflops"], "launches": d["launches"], "gflop_per_launch": d["gflop_per_launch"],
            "us_per_launch": d["us_per_launch"], "share": d["share"],
            "algorithmic_bytes_per_launch": d["algorithmic_bytes_per_launch"],
            "timed_calls": main["timed_calls"], "timed_ms": main["timed_ms"], "graph_ms_per_step": graph_ms_per_step,
            "serialised": alone is not None, "per_launch": main["per_launch"],
            "traffic": traffic, "traffic_source": src,
            "wgrad": {k: main["wgrad"][k] for k in ("launches", "gflop_per_launch", "us_per_launch", "tflops")},
            "overlapped": {"fprop": {k: overlapped["fprop"][k] for k in ("launches", "us_per_launch", "tflops")},
                           "wgrad": {k: overlapped["wgrad"][k] for k in ("launches", "us_per_launch", "tflops")},
                           "what": "the same launches in the product schedule: weight gradients run on side streams "
                                   "and share the SMs with the forward / dgrad launch that is timed"}}


def _kernel_microbench(torch, vb, dev):
    """Per-kernel evidence, timed live with CUDA events on the launching stream: the tcgen05 implicit-GEMM kernels on
    the fattest cfg-2 layer shapes (B=256) with random data, and the HBM-bound BatchNorm-apply / Adam kernels."""
    from importlib import import_module
    fn = import_module("vaegan_b200.functional")
    lib = vb.load_library()
    peaks = _peaks()
    out = []
    B = BATCH_PER_GPU

    def timeit(f, iters=20, repeats=3):
        for _ in range(3):
            f()
        torch.cuda.synchronize()
        best = None
        for _ in range(repeats):           # min of 3 x 20: one slow repeat (clock ramp, a stray host stall) does not stick
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(iters):
                f()
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / iters
            best = ms if best is None else min(best, ms)
        return best

    shapes = [("G ConvT 512->256 8^2->16^2", fn.ConvSpec("up", 512, 256, 4, 2, 1), 8),
              ("G ConvT 256->128 16^2->32^2", fn.ConvSpec("up", 256, 128, 4, 2, 1), 16),
              ("D Conv 128->256 16^2->8^2", fn.ConvSpec("down", 256, 128, 4, 2, 1), 16)]
    for name, spec, h in shapes:
        g = spec.geom(B, h, h)
        small = torch.randn(B, g.small_h, g.small_w, g.small_c, device=dev).bfloat16()
        big = torch.randn(B, g.big_h, g.big_w, g.big_c, device=dev).bfloat16()
        w = torch.randn(g.small_c, g.big_c, 4, 4, device=dev) * 0.02
        wd, wu = fn.pack_weights(w, g)
        dw = torch.zeros_like(w)
        flops = 2.0 * B * g.small_h * g.small_w * g.small_c * g.big_c * 16
        for op, f in (("down", lambda: fn.conv_down(big, wd, g)), ("up", lambda: fn.conv_up(small, wu, g)),
                      ("wgrad", lambda: fn.conv_wgrad(small, big, g, dw))):
            ms = timeit(f)
            tf = flops / (ms * 1e-3) / 1e12
            out.append({"kernel": f"igemm_{'wgrad' if op == 'wgrad' else 'fprop'}_kernel", "op": op, "layer": name,
                        "us": ms * 1e3, "tflops": tf, "frac_of_burst_peak": tf / peaks["burst"]})
    # HBM-bound: BN apply (bf16 in + out) on the largest BatchNorm tensor, and fused Adam (28 B / parameter)
    x = torch.randn(B, 64, 64, 64, device=dev).bfloat16()
    sc, sh = torch.rand(64, device=dev), torch.rand(64, device=dev)
    ms = timeit(lambda: fn.scale_shift_act(x, sc, sh, 1, 0.0))
    gbs = x.numel() * 4 / (ms * 1e-3) / 1e9
    out.append({"kernel": "scale_shift_act_vec_kernel", "op": "bn_apply+relu", "layer": "G 64@64^2", "us": ms * 1e3,
                "gbs": gbs, "frac_of_hbm_peak": gbs / peaks["hbm"]})
    n = 13_243_968
    import ctypes
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    p, gr, m, v = (torch.zeros(n, device=dev) for _ in range(4))
    stp = torch.zeros((), dtype=torch.int64, device=dev)
    ms = timeit(lambda: lib.vg_adam_step(P(p), P(gr), P(m), P(v), n, 2e-4, 0.9, 0.999, 1e-8, P(stp), 1.0,
                                         ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    gbs = n * 28 / (ms * 1e-3) / 1e9
    out.append({"kernel": "adam_kernel", "op": "adam", "layer": "G (13.24 M params)", "us": ms * 1e3, "gbs": gbs,
                "frac_of_hbm_peak": gbs / peaks["hbm"]})
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-graph", action="store_true", help="run the step eagerly instead of replaying a CUDA graph")
    ap.add_argument("--no-micro", action="store_true", help="skip the per-kernel micro-benchmarks")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the cfg 3 / 4 / 5 measurements (other_configs)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
