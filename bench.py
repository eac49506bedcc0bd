#!/usr/bin/env python
"""Benchmark of the VQ bottleneck hot path (BASELINE.json metric: VQ lookup tokens/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload all|c2|c3|c1|c4|c5|n1]
    python bench.py --impl reference ...      # the UNMODIFIED reference module on the host CPU

A step = one quantizer forward + backward over one batch of synthetic latents
(per GPU: 1M tokens for c2/c3).  Weak scaling: every rank has its own batch, the
codebook is replicated, and the per-step codebook statistics (dE, histogram,
squared error) are all-reduced over NCCL.  Prints ONE JSON line on rank 0.

The default workload ("all") keeps BASELINE.json's configs[1] (c2, FMA-bound) as the
top-level `value` and adds equally complete blocks for the tensor-bound configs[2]
(`"c3"`) and, on more than one GPU, the codebook-sharded bulk encode of configs[4]
(`"c5"`), plus a `"parity_check"` block: on-device / oracle self-checks of the very
tensors the timed steps produced (histogram mass, all-reduced dE == sum of per-rank dE,
sharded winners == unsharded search, sampled indices == CPU oracle).
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (images per GPU, D, H, W, K, description)
    "c2": (1024, 4, 32, 32, 16384, "low-dim f=4 LDM latent: 1M tokens x d=4 x codebook 16384, fwd+bwd"),
    "c3": (1024, 256, 32, 32, 16384, "high-dim tokenizer: 1M tokens x d=256 x codebook 16384, fwd+bwd"),
    "c1": (4, 256, 32, 32, 128, "train_vqgan.py default: 4096 tokens x d=256 x codebook 128, fwd+bwd"),
    "c4": (64, 256, 32, 32, 128, "quantizer of the 256x256 VQ-GAN step, global batch 64 split over the ranks, "
                                 "K=128 d=256, fwd+bwd + stats all-reduce (encoder/decoder out of scope)"),
    "c5": (56, 256, 32, 32, 65536, "bulk encode: 56 images (57344 tokens) per rank per step, codebook 65536 x d=256 "
                                   "sharded over the ranks, all-gather + local search + MIN reduce-scatter"),
}
WORKLOADS["n1"] = (1024, 256, 32, 32, 256, "next row N1: pre_quant_conv-style 1x1 convolution 256 -> 256 channels on 1M tokens "
                   "(NCHW fp32, 3xTF32 tcgen05), forward only; K here = output channels")
CPU_CHUNK_TOKENS = {"c2": 32768, "c3": 16384, "c1": 4096, "c4": 4096, "c5": 4096, "n1": 65536}
BETA = 0.25
N_ROTATE = 8  # distinct input sets cycled through so the working set exceeds the 126 MB L2


def ref_quantizer_class():
    """The UNMODIFIED reference VectorQuantizer staged in oracle/_ref (None if the copy is absent)."""
    try:
        from oracle import ref_loader
        return ref_loader.reference_quantizer_class() if ref_loader.available() else None
    except Exception:
        return None


def make_codebook(kind, K, D, device=None, tokens=None):
    """normal: N(0,1) seed 1 (SURVEY 8d primary).  refinit: the reference constructor's U(+-1/K)
    (quantizer.py:48) under torch.manual_seed(42).  trained: K distinct tokens + 0.01 N(0,1) (SURVEY 8d C1
    variant).  clustered: 16 centres, codes 1e-4 apart (collapsed codebook, adversarial for any low-precision
    first pass).  copied: K/4 N(0,1) codes, each present four times bit for bit, shuffled (a codebook restarted by copying
    live codes onto dead ones)."""
    if kind == "normal":
        E = torch.randn(K, D, generator=torch.Generator().manual_seed(1))
    elif kind == "refinit":
        state = torch.get_rng_state()
        torch.manual_seed(42)
        emb = torch.nn.Embedding(K, D)
        emb.weight.data.uniform_(-1.0 / K, 1.0 / K)
        E = emb.weight.detach().clone()
        torch.set_rng_state(state)
    elif kind == "trained":
        rows = tokens.permute(0, 2, 3, 1).reshape(-1, D)
        pick = torch.randperm(rows.shape[0], generator=torch.Generator().manual_seed(1))[:K].to(rows.device)
        E = rows[pick].cpu() + 0.01 * torch.randn(K, D, generator=torch.Generator().manual_seed(11))
    elif kind == "copied":
        g = torch.Generator().manual_seed(1)
        base = torch.randn(max(K // 4, 1), D, generator=g)
        E = base.repeat_interleave(4, dim=0)[:K]
        if E.shape[0] < K:
            E = torch.cat([E, torch.randn(K - E.shape[0], D, generator=g)])
        E = E[torch.randperm(K, generator=g)]
    elif kind == "clustered":
        g = torch.Generator().manual_seed(1)
        centres = torch.randn(16, D, generator=g)
        E = centres[torch.randint(0, 16, (K,), generator=g)] + 1e-4 * torch.randn(K, D, generator=g)
    else:
        raise ValueError(kind)
    return E.contiguous() if device is None else E.contiguous().to(device)


def workload_config(workload, world, codebook="normal"):
    """The `config` object of a line: the workload's definition only, identical for the B200 arm and the
    reference arm (run-dependent facts -- kernel chosen, L2 rotation, chunking of the CPU arm -- go to `details`)."""
    B, D, H, W, K, desc = WORKLOADS[workload]
    if workload == "c4":
        B = max(B // world, 1)
    return {"workload": f"{workload}: {desc}", "tokens_per_gpu_per_step": B * H * W, "D": D, "K": K, "beta": BETA,
            "codebook": codebook, "parallelism": f"dp{world}"}


def load_traffic(kernel, workload):
    """DRAM bytes per launch of `kernel` from the committed ncu --set full capture (None if not captured
    for this workload)."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        with open(path) as f:
            ent = json.load(f).get(kernel)
        return ent["bytes"] if ent and ent.get("workload") == workload else None
    except Exception:
        return None


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": p.get("hbm_gbs"), "bf16_tflops": p.get("bf16_tflops"),
                "bf16_tflops_sustained": p.get("bf16_tflops_sustained"), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


# ---------------------------------------------------------------------------
# clocks (pynvml sampling thread during the timed region)
# ---------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
               0x4: "sw_power_cap", 0x80: "hw_power_brake"}

    def __init__(self, index, period=0.02):
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def __enter__(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                    "samples": 0}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ---------------------------------------------------------------------------
# CPU arm: the reference op sequence (oracle port) on the host cores
# ---------------------------------------------------------------------------
def cpu_step_fn(workload):
    """Returns (fn, tokens_per_call, kind): one fwd+bwd of the reference quantizer on a bounded token chunk of
    the workload (the reference cannot hold [1M, 16384]).  kind "reference": the UNMODIFIED module
    (quantizer.py:17-149) staged in oracle/_ref; "port": the oracle's restatement when the copy is absent."""
    _, D, H, W, K, _ = WORKLOADS[workload]
    chunk = CPU_CHUNK_TOKENS[workload]
    B = max(chunk // (H * W), 1)
    if workload == "n1":  # the reference layer itself: nn.Conv2d(cin, cout, 1) (vq_vae.py:75), forward
        conv = torch.nn.Conv2d(D, K, 1)
        xc = torch.randn(B, D, H, W, generator=torch.Generator().manual_seed(0))

        def fn_conv():
            with torch.no_grad():
                return conv(xc)
        return fn_conv, B * H * W, "reference"
    z = torch.randn(B, D, H, W, generator=torch.Generator().manual_seed(0))
    E = torch.randn(K, D, generator=torch.Generator().manual_seed(1))
    g = torch.randn(B, D, H, W, generator=torch.Generator().manual_seed(2))
    Ref = ref_quantizer_class()
    if Ref is not None:
        vq = Ref(K, D, BETA)
        with torch.no_grad():
            vq.embedding.weight.copy_(E)

        def fn_ref():
            zr = z.detach().requires_grad_(True)
            vq.embedding.weight.grad = None
            z_q, loss_dict, idx = vq(zr)
            if workload == "c5":  # encode_to_indices (vq_vae.py:162-175): forward only
                return idx
            (loss_dict["vq_loss"] + (z_q * g).sum()).backward()
            return idx
        return fn_ref, B * H * W, "reference"
    from oracle import vq_oracle as orc

    def fn():
        return orc.autograd_step(z, E, BETA, g)
    return fn, B * H * W, "port"


def run_cpu_baseline(workload, budget_s=12.0, max_calls=6):
    torch.set_num_threads(os.cpu_count() or 1)
    fn, tokens, kind = cpu_step_fn(workload)
    fn()  # warm-up
    times = []
    t_all = time.perf_counter()
    while len(times) < max_calls and (time.perf_counter() - t_all) < budget_s:
        t0 = time.perf_counter()
        fn()
        times.append(time.perf_counter() - t0)
    best = min(times)
    what = ("the unmodified reference VectorQuantizer (oracle/_ref)" if kind == "reference"
            else "reference op sequence (oracle port)")
    return {"value": tokens / best, "unit": "tokens/s", "cores": torch.get_num_threads(),
            "kind": kind,
            "sample": f"{tokens}-token chunk of {workload} ({'forward' if workload in ('c5', 'n1') else 'fwd+bwd'}, "
                      f"{what}, torch CPU fp32, best of {len(times)})"}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    workload = "c2" if args.workload == "all" else args.workload
    fn, tokens, kind = cpu_step_fn(workload)
    for _ in range(max(args.warmup, 1) if args.warmup else 0):
        fn()
    steps = max(args.steps, 1)
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    dt = time.perf_counter() - t0
    value = tokens * steps / dt
    desc = WORKLOADS[workload][5]
    line = {
        "impl": "reference", "metric": "vq_lookup_tokens_per_sec", "value": value, "unit": "tokens/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup, "ms_per_step": dt / steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": workload_config(workload, args.gpus, args.codebook),
        "details": {"tokens_per_step": tokens,
                   "note": ("the UNMODIFIED reference VectorQuantizer.forward + autograd backward (quantizer.py:50-110, "
                            "staged in oracle/_ref)" if kind == "reference" else
                            "reference op sequence (quantizer.py:63-98 + autograd), oracle port") +
                           " on the host CPU; each step is a bounded token chunk because the reference "
                           "materialises [N, K]"},
        "cpu_baseline": {"value": value, "unit": "tokens/s", "cores": torch.get_num_threads(),
                         "kind": kind, "sample": f"{tokens}-token chunk x {steps} steps"},
        "e2e": {"value": value, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------
def gpu_strawman_tokens_per_s(workload, device):
    """The reference op sequence with stock torch CUDA ops, on one chunk (context only)."""
    _, D, H, W, K, _ = WORKLOADS[workload]
    chunk = CPU_CHUNK_TOKENS[workload]
    B = max(chunk // (H * W), 1)
    z = torch.randn(B, D, H, W, device=device, requires_grad=True)
    E = torch.randn(K, D, device=device, requires_grad=True)
    g = torch.randn(B, D, H, W, device=device)

    def fn():
        rows = z.permute(0, 2, 3, 1).reshape(-1, D)
        d = (rows ** 2).sum(1, keepdim=True) + (E ** 2).sum(1) - 2 * rows @ E.t()
        idx = torch.argmin(d, 1)
        e = torch.nn.functional.embedding(idx, E).view(B, H, W, D).permute(0, 3, 1, 2).contiguous()
        loss = torch.nn.functional.mse_loss(e.detach(), z) + BETA * torch.nn.functional.mse_loss(e, z.detach())
        zq = z + (e - z).detach()
        (loss + (zq * g).sum()).backward()
        z.grad = None
        E.grad = None
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        fn()
    b.record()
    b.synchronize()
    return B * H * W * 5 / (a.elapsed_time(b) * 1e-3)


def run_conv(args, world, rank, local_rank, device, B, Cin, H, W, Cout, desc, peaks):
    """n1: the 1x1 convolution next to the quantizer (vq_vae.py:74-79,115), forward, HBM-bound."""
    import torch.distributed as dist
    from vq_gan_b200 import ops
    tokens = B * H * W
    g = torch.Generator(device=device).manual_seed(100 + rank)
    xs = [torch.randn(B, Cin, H, W, device=device, generator=g) for _ in range(2)]
    wt = torch.randn(Cout, Cin, device=device, generator=g) / Cin ** 0.5
    bs = torch.randn(Cout, device=device, generator=g)

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local_rank])
        torch.cuda.synchronize()

    warmup = max(args.warmup, 3)
    for i in range(warmup):
        ops.conv1x1(xs[i % 2], wt, bs)
    barrier()
    launches0 = ops.LAUNCHES["total"]
    ops.PROFILE_CONV = []
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clk:
        barrier()
        ev0.record()
        for i in range(args.steps):
            ops.conv1x1(xs[i % 2], wt, bs)
        ev1.record()
        barrier()
    ms_total = ev0.elapsed_time(ev1)
    k_ms = statistics.mean(a.elapsed_time(b) for a, b in ops.PROFILE_CONV)
    ops.PROFILE_CONV = None
    if world > 1:
        t = torch.tensor([ms_total], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    # end to end: pinned host activations in, result's checksum out, every step
    x_host = torch.randn(B, Cin, H, W).pin_memory()
    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(2):
        float(ops.conv1x1(x_host.to(device, non_blocking=True), wt, bs).sum())
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(e2e_steps):
        float(ops.conv1x1(x_host.to(device, non_blocking=True), wt, bs).sum())
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1)
    line = None
    if rank == 0:
        nbytes = tokens * (Cin + Cout) * 4 + 4 * Cin * Cout
        flops = 2.0 * tokens * Cin * Cout
        line = {
            "metric": "vq_lookup_tokens_per_sec", "value": tokens * world * args.steps / (ms_total * 1e-3),
            "unit": "tokens/s", "n_gpus": world, "steps": args.steps, "warmup": warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"n1: {desc}", "tokens_per_gpu_per_step": tokens, "Cin": Cin, "Cout": Cout,
                       "l2": "2 rotating input sets (2147 MB) > 126 MB L2"},
            "roofline": {"bound": "hbm", "kernel": "conv1x1_tc_kernel", "achieved": nbytes / (k_ms * 1e-3) / 1e9,
                         "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": nbytes / (k_ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                         "traffic": load_traffic("conv1x1_tc_kernel", "n1"), "kernel_ms": k_ms,
                         "algorithmic_bytes_per_launch": nbytes,
                         "tensor_tflops_algorithmic": flops / (k_ms * 1e-3) / 1e12, "executed_flops_factor": 3,
                         "tensor_frac_executed": 3 * flops / (k_ms * 1e-3) / 1e12 / (peaks["bf16_tflops_sustained"] / 2)},
            "e2e": {"value": tokens * world * e2e_steps / (e2e_ms * 1e-3), "unit": "tokens/s",
                    "h2d_bytes_per_step": tokens * Cin * 4, "d2h_bytes_per_step": 4, "steps": e2e_steps,
                    "ms_per_step": e2e_ms / e2e_steps},
            "gpu_launches": ops.LAUNCHES["total"] - launches0, "clocks": clk.summary(),
        }
        if world == 1:
            torch.set_num_threads(os.cpu_count() or 1)
            xc = torch.randn(64, Cin, H, W)
            conv = torch.nn.Conv2d(Cin, Cout, 1)
            conv(xc)
            t0 = time.perf_counter()
            for _ in range(3):
                conv(xc)
            dt = (time.perf_counter() - t0) / 3
            line["cpu_baseline"] = {"value": 64 * H * W / dt, "unit": "tokens/s", "cores": torch.get_num_threads(),
                                    "kind": "reference", "sample": "nn.Conv2d(256, 256, 1) forward on 65536 tokens, torch CPU fp32"}
    return line


def run_full_step(args, world, rank, local_rank, device, peaks, steps=None, warmup=3):
    """c4 as SURVEY 8(d) specifies it: the reference's own VQVAE (staged in oracle/_ref, default VQGANConfig
    architecture, 67.5 M parameters) with `vqvae.quantizer` swapped for the drop-in (vq_vae.py:82-86), wrapped
    in DistributedDataParallel like accelerate does (train_vqgan.py:197-209), global batch 64 of 256x256 images
    split over the ranks (split_batches=True, :112), one step = forward, L1 + vq_loss, backward (DDP NCCL
    all-reduce of all gradients incl. the codebook's), clip_grad_norm_(1.0), Adam (:267-271).  LPIPS and the
    discriminator are out of scope and lpips is not installed: reconstruction term = L1 only.  The conv stacks
    are the reference's stock cuDNN layers -- this measures the drop-in INSIDE the caller, and the
    quantizer's share of the step."""
    import torch.distributed as dist
    from torch.nn.parallel import DistributedDataParallel as DDP
    from oracle import ref_loader
    from vq_gan_b200 import VectorQuantizer, ops
    if not ref_loader.available():
        return {"workload": "c4full", "unavailable": "oracle/_ref not staged (python oracle/make_ref.py)"} if rank == 0 else None
    VQVAE = ref_loader.reference_vqvae_class()
    kw = ref_loader.default_vqvae_kwargs()
    torch.manual_seed(42)
    model = VQVAE(**kw)
    model.quantizer = VectorQuantizer(kw["num_embeddings"], kw["embedding_dim"], kw["commitment_cost"])
    model = model.to(device).train()
    global_batch = 64
    Bl = max(global_batch // world, 1)
    net = DDP(model, device_ids=[local_rank]) if world > 1 else model
    opt = torch.optim.Adam(net.parameters(), lr=4.5e-5, betas=(0.5, 0.9), weight_decay=0.0)
    imgs = [torch.rand(Bl, 3, 256, 256, device=device, generator=torch.Generator(device=device).manual_seed(100 + rank + 7 * j))
            for j in range(2)]
    tokens = Bl * 32 * 32

    def step(i):
        x = imgs[i % 2]
        rec, ld = net(x)
        total = (rec - x).abs().mean() + ld["vq_loss"]
        opt.zero_grad(set_to_none=True)
        total.backward()
        torch.nn.utils.clip_grad_norm_(net.parameters(), 1.0)
        opt.step()
        return ld

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local_rank])
        torch.cuda.synchronize()

    steps = steps or max(3, min(args.steps, 10))
    for i in range(warmup):
        step(i)
    barrier()
    ops.PROFILE, ops.PROFILE_TAIL, ops.PROFILE_BWD = [], [], []
    launches0 = ops.LAUNCHES["total"]
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clk:
        barrier()
        ev0.record()
        for i in range(steps):
            ld = step(i)
        ev1.record()
        barrier()
    ms_total = ev0.elapsed_time(ev1)
    q_ms = sum(a.elapsed_time(b) for lst in (ops.PROFILE, ops.PROFILE_TAIL, ops.PROFILE_BWD) for a, b in lst)
    ops.PROFILE = ops.PROFILE_TAIL = ops.PROFILE_BWD = None
    if world > 1:
        t = torch.tensor([ms_total], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    if rank != 0:
        return None
    return {
        "metric": "vq_lookup_tokens_per_sec", "value": tokens * world * steps / (ms_total * 1e-3), "unit": "tokens/s",
        "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": ms_total / steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "c4full: full VQ-GAN training step at 256x256 (reference VQVAE 67.5M params + drop-in "
                               "quantizer K=128 d=256), global batch 64, DDP, L1 + vq_loss (LPIPS / discriminator off), "
                               "clip 1.0 + Adam(4.5e-5, (0.5, 0.9))",
                   "images_per_gpu": Bl, "tokens_per_gpu_per_step": tokens, "parallelism": f"ddp{world}",
                   "images_per_s": Bl * world * steps / (ms_total * 1e-3)},
        "quantizer_share_of_step": q_ms / ms_total,
        "quantizer_ms_per_step": q_ms / steps,
        "loss_dict_keys": sorted(ld.keys()),
        "gpu_launches": ops.LAUNCHES["total"] - launches0, "clocks": clk.summary(),
    }


def run_sharded_encode(args, world, rank, local_rank, device, B, D, H, W, K, desc, peaks):
    """c5: search only (encode_to_indices), codebook rows sharded over the ranks.  Returns the line (rank 0)."""
    import torch.distributed as dist
    from vq_gan_b200 import ops
    from vq_gan_b200 import distributed as vdist
    tokens = B * H * W
    klo, khi = vdist.shard_range(K, world, rank)
    E_full = torch.randn(K, D, generator=torch.Generator().manual_seed(1))
    E = E_full[klo:khi].contiguous().to(device)
    gen = torch.Generator(device=device).manual_seed(100 + rank)
    zs = [torch.randn(B, D, H, W, device=device, generator=gen) for _ in range(4)]

    enc = vdist.ShardedEncoder(E, klo) if world > 1 else None

    def step(i):
        # steady state of the bulk-encode loop: the all-gather of batch i+1 overlaps the search of batch i
        # (every step still does one gather, one search and one key reduce-scatter)
        if world > 1:
            enc.submit(zs[(i + 1) % 4])
            return enc.collect()
        idx, dmin, _ = ops.search(zs[i % 4], E)
        return idx, dmin

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local_rank])
        torch.cuda.synchronize()

    steps = max(min(args.steps, 30), 1)
    warmup = max(args.warmup, 3)
    if world > 1:
        enc.submit(zs[0])  # prime the one-batch look-ahead
    for i in range(warmup):
        step(i)
    barrier()
    launches0 = ops.LAUNCHES["total"]
    ops.PROFILE = []
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clk:
        barrier()
        ev0.record()
        for i in range(steps):
            step(i)
        ev1.record()
        barrier()
    ms_total = ev0.elapsed_time(ev1)
    search_ms = [a.elapsed_time(b) for a, b in ops.PROFILE]
    ops.PROFILE = None
    gpu_launches = ops.LAUNCHES["total"] - launches0
    if world > 1:
        t = torch.tensor([ms_total], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())

    # ---- end to end: pinned host latents in, compact index map (vq_gan_b200.indexio, 2 bytes/token) out
    z_host = [torch.randn(B, D, H, W, generator=torch.Generator().manual_seed(300 + rank + j)).pin_memory()
              for j in range(2)]
    codes_host = torch.empty(B, H, W, dtype=torch.uint16).pin_memory()
    copy_stream = torch.cuda.Stream(device=device)

    def prefetch(i):
        with torch.cuda.stream(copy_stream):
            zt = z_host[i % 2].to(device, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return zt, ev

    def e2e_step(i, staged):
        zt, ev = staged
        torch.cuda.current_stream().wait_event(ev)
        zt.record_stream(torch.cuda.current_stream())
        nxt = prefetch(i + 1)
        if world > 1:
            enc.submit(zt)                # this step's batch: gather starts now ...
            idx, _ = enc.collect()        # ... while the previous step's batch is searched (one batch of look-ahead)
        else:
            idx, _, _ = ops.search(zt, E)
        codes, _ = ops.indices_narrow(idx, K)
        codes_host.copy_(codes, non_blocking=True)
        torch.cuda.current_stream().synchronize()  # the step's result is on the host
        return int(codes_host[0, 0, 0]), nxt

    e2e_steps = max(3, min(steps, 10))
    staged = prefetch(0)
    for i in range(2):
        _, staged = e2e_step(i, staged)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(e2e_steps):
        _, staged = e2e_step(2 + i, staged)
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([e2e_ms], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())

    # ---- self-check (outside the timed regions): the sharded winners of rank 0's tokens equal an UNSHARDED
    # search of the full codebook on rank 0, and a sample of them equals the CPU oracle
    parity = {}
    if world > 1:
        while enc._pending:      # drain the look-ahead so the check sees exactly batch 0
            enc.collect()
        enc.submit(zs[0])
        idx_s, dmin_s = enc.collect()
    else:
        idx_s, dmin_s = step(0)
    idx_u, dmin_u, _ = ops.search(zs[0], E_full.to(device))
    if world > 1:
        parity["sharded_equals_unsharded_search"] = bool(torch.equal(idx_s, idx_u))
        parity["sharded_min_score_max_abs_diff"] = float((dmin_s - dmin_u).abs().max())
    if rank == 0:
        from oracle import vq_oracle as orc
        rows = orc.tokens_of(zs[0][:2].cpu())
        rep = orc.compare_indices(idx_s[:2].reshape(-1).cpu(), orc.search_with_gap(rows, E_full))
        parity["oracle_sample"] = rep
    ok = parity.get("sharded_equals_unsharded_search", True) and (rank != 0 or parity["oracle_sample"]["outside"] == 0)
    if world > 1:
        t = torch.tensor([0 if ok else 1], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        ok = int(t.item()) == 0
    parity["status"] = "ok" if ok else "FAILED"

    line = None
    if rank == 0:
        s_ms_step = sum(search_ms) / steps           # search time per step (all slices)
        flops = 2.0 * tokens * world * (khi - klo) * D  # per-rank search: all tokens x local codes
        achieved = flops / (s_ms_step * 1e-3) / 1e12
        line = {
            "metric": "vq_lookup_tokens_per_sec", "value": tokens * world * steps / (ms_total * 1e-3),
            "unit": "tokens/s", "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": ms_total / steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"c5: {desc}", "tokens_per_gpu_per_step": tokens, "D": D, "K": K,
                       "codes_per_rank": khi - klo, "parallelism": f"codebook-sharded x{world}",
                       "l2": "4 rotating latent sets (235 MB) exceed the 126 MB L2",
                       "pipeline": "ShardedEncoder: the all-gather of batch i+1 overlaps the search of batch i (one batch of "
                                   "look-ahead); one gather, one search, one key reduce-scatter per step"},
            "roofline": {"bound": "tensor", "kernel": "search_tc16_kernel", "achieved": achieved,
                         "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                         "frac": achieved / peaks["bf16_tflops_sustained"], "executed_flops_factor": 1,
                         "frac_executed": achieved / peaks["bf16_tflops_sustained"], "traffic": None,
                         "kernel_ms": s_ms_step, "algorithmic_flops_per_launch": flops,
                         "step_share": s_ms_step * steps / ms_total},
            "exchange_bytes_per_step": {"all_gather_latents": tokens * D * 4 * max(world - 1, 0),
                                        "reduce_scatter_keys": tokens * world * 8 if world > 1 else 0},
            "e2e": {"value": tokens * world * e2e_steps / (e2e_ms * 1e-3), "unit": "tokens/s",
                    "h2d_bytes_per_step": tokens * D * 4, "d2h_bytes_per_step": tokens * 2, "steps": e2e_steps,
                    "ms_per_step": e2e_ms / e2e_steps,
                    "api": "distributed.sharded_search_dp on pinned host latents (H2D of step i+1 overlaps step i), "
                           "compact uint16 index map read back every step"},
            "parity_check": parity,
            "gpu_launches": gpu_launches, "clocks": clk.summary(),
        }
        if world == 1:
            line["cpu_baseline"] = run_cpu_baseline("c5")
    del zs, z_host
    torch.cuda.empty_cache()
    return line


def measure_quantizer(args, workload, world, rank, local_rank, device, peaks, fma, codebook="normal",
                      steps=None, with_cpu=True):
    """Forward + backward of the drop-in on `workload`; returns the complete line (a dict) on rank 0."""
    import torch.distributed as dist
    from vq_gan_b200 import VectorQuantizer, ops
    from vq_gan_b200 import distributed as vdist

    B, D, H, W, K, desc = WORKLOADS[workload]
    if workload == "c4":
        B = max(B // world, 1)  # global batch 64 split over the ranks (strong scaling of one step)
    tokens = B * H * W
    steps = steps or args.steps
    fma_scalar, fma_packed = fma
    fma_peak = max(fma)

    n_rot = N_ROTATE if workload == "c2" else (2 if workload == "c3" else 16)
    gen = torch.Generator(device=device).manual_seed(100 + rank)
    zs = [torch.randn(B, D, H, W, device=device, generator=gen).requires_grad_(True) for _ in range(n_rot)]
    gs = [torch.randn(B, D, H, W, device=device, generator=gen) for _ in range(min(n_rot, 2))]
    vq = VectorQuantizer(K, D, BETA, lazy_stats=True, algo=args.algo).to(device)
    with torch.no_grad():
        vq.embedding.weight.copy_(make_codebook(codebook, K, D, tokens=zs[0].detach()))
    weight = vq.embedding.weight
    bytes_per_set = zs[0].numel() * 4 * 4 + tokens * 8  # z, g, z_q, dz + idx
    one = torch.ones((), device=device)
    last = {}

    def step(i):
        z = zs[i % n_rot]
        z.grad = None
        weight.grad = None
        z_q, loss_dict, idx = vq(z)
        # the decoder's gradient w.r.t. z_q and d(total loss)/d(vq_loss) = 1 are handed to autograd
        # directly, so the timed region holds the quantizer's own kernels only
        torch.autograd.backward((z_q, loss_dict["vq_loss"]), (gs[i % len(gs)], one))
        if world > 1:
            # data parallel: one packed all-reduce of dE + histogram + squared-error sum
            usage, _, _ = ops.codebook_usage(idx, K)
            sq = (loss_dict["codebook_loss"] * float(z.numel())).reshape(1)
            last["local_dE"] = weight.grad
            dE, hist, s = vdist.allreduce_stats(weight.grad, usage, sq, average_dE=True)
            weight.grad = dE
            last["hist"], last["sq"] = hist, s
        last["idx"], last["i"] = idx, i
        z.grad = None  # hand dz back to the caching allocator: no cudaMalloc in the timed region
        return loss_dict["vq_loss"]

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local_rank])
        torch.cuda.synchronize()

    warmup = max(args.warmup, 3)
    for i in range(warmup):
        step(i)
    barrier()

    # ---- device-resident timing: exactly K steps between two events --------
    ops.PROFILE, ops.PROFILE_TAIL, ops.PROFILE_BWD = [], [], []
    launches0 = ops.LAUNCHES["total"]
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clk:
        barrier()
        ev0.record()
        for i in range(steps):
            step(i)
        ev1.record()
        barrier()
    ms_total = ev0.elapsed_time(ev1)
    gpu_launches = ops.LAUNCHES["total"] - launches0
    search_ms = [a.elapsed_time(b) for a, b in ops.PROFILE]
    tail_ms = [a.elapsed_time(b) for a, b in ops.PROFILE_TAIL]
    bwd_ms = [a.elapsed_time(b) for a, b in ops.PROFILE_BWD]
    ops.PROFILE = ops.PROFILE_TAIL = ops.PROFILE_BWD = None
    if world > 1:
        t = torch.tensor([ms_total], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    stats = vq.last_search_stats.tolist()

    # ---- self-check of what the timed steps produced (outside the timed region) ----
    parity = {}
    idx_last = last["idx"]
    usage, _, _ = ops.codebook_usage(idx_last, K)
    if world > 1:
        parity["hist_sum_equals_tokens_x_world"] = int(last["hist"].sum()) == tokens * world
        parts = [torch.empty_like(last["local_dE"]) for _ in range(world)]
        dist.all_gather(parts, last["local_dE"].contiguous())
        want = torch.stack(parts).double().sum(0) / world
        got = weight.grad.double()
        err = float((got - want).abs().max() / want.abs().max().clamp_min(1e-30))
        parity["allreduced_dE_vs_sum_of_rank_dE_rel_err"] = err
        parity["allreduced_dE_ok"] = err < 1e-5
    else:
        parity["hist_sum_equals_tokens"] = int(usage.sum()) == tokens
    if rank == 0:
        # a seeded sample of the LAST timed step's tokens against the CPU oracle (reference op order, fp32)
        from oracle import vq_oracle as orc
        nb = max(1, min(B, 4096 // (H * W)))
        zb = zs[last["i"] % n_rot][:nb].detach().cpu()
        rep = orc.compare_indices(idx_last[:nb].reshape(-1).cpu(),
                                  orc.search_with_gap(orc.tokens_of(zb), weight.detach().cpu()))
        parity["oracle_sample"] = rep
    flags = [v for k, v in parity.items() if isinstance(v, bool)]
    ok = all(flags) and (rank != 0 or parity["oracle_sample"]["outside"] == 0)
    if world > 1:
        t = torch.tensor([0 if ok else 1], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        ok = int(t.item()) == 0
    parity["status"] = "ok" if ok else "FAILED"

    # ---- optional: the same K steps as CUDA-graph replays (launch-bound shapes) ----
    eager_ms_total = None
    if args.graph:
        from vq_gan_b200.graphs import GraphedVectorQuantizer
        gvq = GraphedVectorQuantizer(vq, zs[0])

        def gstep(i):
            z = zs[i % n_rot]
            z.grad = None
            weight.grad = None
            z_q, loss_dict, idx = gvq(z)
            torch.autograd.backward((z_q, loss_dict["vq_loss"]), (gs[i % len(gs)], one))
            if world > 1:
                usage, _, _ = ops.codebook_usage(idx, K)
                sq = (loss_dict["codebook_loss"] * float(z.numel())).reshape(1)
                dE, hist, s = vdist.allreduce_stats(weight.grad, usage, sq, average_dE=True)
                weight.grad = dE
            z.grad = None
        for i in range(warmup):
            gstep(i)
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for i in range(steps):
            gstep(i)
        g1.record()
        barrier()
        eager_ms_total = ms_total
        ms_total = g0.elapsed_time(g1)
        if world > 1:
            t = torch.tensor([ms_total], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_total = float(t.item())

    # ---- end to end through the public module API with host buffers --------
    z_host = [torch.randn(B, D, H, W, generator=torch.Generator().manual_seed(200 + rank + j)).pin_memory()
              for j in range(2)]
    idx_host = torch.empty(B, H, W, dtype=torch.int64).pin_memory()
    vq_sync = VectorQuantizer(K, D, BETA, algo=args.algo).to(device)
    vq_sync.embedding.weight = weight

    # input pipeline like a training loop's data loader: the NEXT step's latents are copied
    # host->device on a side stream while the current step computes (every step still pays
    # its own H2D copy and D2H read-back inside the timed region)
    copy_stream = torch.cuda.Stream(device=device)
    d2h_stream = torch.cuda.Stream(device=device)

    def prefetch(i):
        with torch.cuda.stream(copy_stream):
            zt = z_host[i % 2].to(device, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return zt, ev

    RING = 4                                                          # result slots: the host consumes step i - (RING - 1)
    loss_host = torch.zeros(RING, dtype=torch.float32).pin_memory()
    loss_events = [None] * RING
    losses_read = []

    def e2e_step(i, staged):
        zt, ev = staged
        torch.cuda.current_stream().wait_event(ev)
        zt.record_stream(torch.cuda.current_stream())
        nxt = prefetch(i + 1)
        z = zt.requires_grad_(True)
        weight.grad = None
        # the module with the REFERENCE contract (loss_dict holds Python floats; they are fetched on first access,
        # vq_gan_b200.LossDict -- this loop, like train_vqgan.py:303-315, never looks at them)
        z_q, loss_dict, idx = vq_sync(z)
        torch.autograd.backward((z_q, loss_dict["vq_loss"]), (gs[i % len(gs)], one))
        if world > 1:
            usage, _, _ = ops.codebook_usage(idx, K)
            sq = (vq_sync.last_mse * float(z.numel())).reshape(1)   # device value: no pageable H2D copy
            dE, hist, s = vdist.allreduce_stats(weight.grad, usage, sq, average_dE=True)
            weight.grad = dE
        # the step's results go back to pinned host memory inside the timed region: the loss into a small ring on the
        # compute stream, the indices on their own stream; the HOST reads the loss of step i-3 here, so that it never
        # waits for the GPU while there is nothing queued behind and a late rank's jitter is absorbed by the queue
        # instead of stalling the all-reduce of every other rank (every step's loss is still read)
        slot = i % RING
        loss_host[slot:slot + 1].copy_(loss_dict["vq_loss"].detach().reshape(1), non_blocking=True)
        done = torch.cuda.Event()
        done.record()
        loss_events[slot] = done
        with torch.cuda.stream(d2h_stream):
            d2h_stream.wait_event(done)
            idx_host.copy_(idx, non_blocking=True)
            idx.record_stream(d2h_stream)
        old = (i + 1) % RING                       # the oldest slot = the one the NEXT step will overwrite
        prev = loss_events[old]
        if prev is not None:
            prev.synchronize()
            losses_read.append(float(loss_host[old]))
            loss_events[old] = None
        return None, nxt

    e2e_steps = max(3, min(steps, 20 if workload != "c3" else 10))
    staged = prefetch(0)
    for i in range(3):
        _, staged = e2e_step(i, staged)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n_read0 = len(losses_read)
    for i in range(e2e_steps):
        _, staged = e2e_step(3 + i, staged)
    torch.cuda.current_stream().wait_stream(d2h_stream)
    e1.record()
    barrier()
    for k in range(RING):                                            # drain: the last RING-1 steps' losses
        if loss_events[k] is not None:
            loss_events[k].synchronize()
            losses_read.append(float(loss_host[k]))
            loss_events[k] = None
    assert len(losses_read) == 3 + e2e_steps and all(v == v for v in losses_read)  # every step's loss reached the host
    e2e_ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([e2e_ms], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())

    line = None
    if rank == 0:
        algo = int(stats[1])
        flops = 2.0 * tokens * K * D
        s_ms = statistics.mean(search_ms) if search_ms else float("nan")
        factor = {3: 3, 4: 1}.get(algo, 1)
        if algo == 5:
            # tf32x3 on the tensor cores: the executed contraction is 8*ceil((3D+3)/8) tf32 slots per
            # (token, code) instead of D multiply-adds; tf32 dense peak = half the bf16 peak
            slots = 8 * ((3 * D + 3 + 7) // 8)
            factor = slots / D
            bound, peak, unit = "tensor", peaks["bf16_tflops_sustained"] / 2, "TFLOP/s"
            peak_note = (f"tf32 dense = 1/2 of cuBLAS bf16 sustained ({peaks['source']}); algorithmic flops 2NKD, the "
                         f"kernel executes {slots} tf32 slots per pair (tf32x3 split + half norm), x{factor:.0f}; "
                         f"for reference the FP32 FMA peak measured in this run is {fma_peak:.1f} TFLOP/s")
        elif algo in (3, 4):
            bound, peak, unit = "tensor", peaks["bf16_tflops_sustained"], "TFLOP/s"
            peak_note = (f"cuBLAS bf16 sustained, {peaks['source']} (algorithmic flops 2NKD; the kernel executes "
                         f"x{factor}: " + ("bf16x3 split" if algo == 3 else "one fp16 pass, exact fp32 re-score "
                                            "of the certified candidates") + ")")
        else:
            bound, peak, unit = "fma", fma_peak, "TFLOP/s"
            peak_note = ("FP32 FMA peak measured in this run by vqb_fma_peak_launch (libvqb200_bench.so) "
                         f"(scalar FFMA {fma_scalar:.1f}, packed FFMA2 {fma_packed:.1f} TFLOP/s); "
                         "MEASURED_PEAKS.json has no FMA figure")
        achieved = flops / (s_ms * 1e-3) / 1e12
        if algo == 6:
            peak_note += ("; algo 6 runs the CUDA-core kernel (FP32 FMA pipe) and the tf32x3 tensor kernel side by side on "
                          f"disjoint images ({int(stats[3])} of {tokens} tokens on the tensor engine): the fraction is the "
                          "step's algorithmic flops over the FMA-pipe peak alone, i.e. it can exceed what one pipe delivers")
        roofline = {"bound": bound, "kernel": {1: "search_lowd_kernel", 2: "search_fp32_kernel",
                                               3: "search_tc_kernel", 4: "search_tc16_kernel",
                                               5: "search_tclow_kernel",
                                               6: "search_dual_kernel"}.get(algo, str(algo)),
                    "achieved": achieved, "peak": peak, "unit": unit, "frac": achieved / peak,
                    "traffic": None, "kernel_ms": s_ms, "algorithmic_flops_per_launch": flops,
                    "algorithmic_bytes_per_launch": tokens * (4 * D + 8) + 4 * K * D,
                    "executed_flops_factor": factor,
                    "frac_executed": achieved * factor / peak,
                    "peak_source": peak_note,
                    "step_share": s_ms * len(search_ms) / (eager_ms_total or ms_total) if search_ms else None}
        roofline["traffic"] = load_traffic(roofline["kernel"], workload)
        # HBM-side kernels: algorithmic bytes per token 8D+8 (tail) and 12D+8 (+ dE once) (backward)
        hbm = {}
        for name, ms_list, nbytes in (("gather_loss_st_kernel", tail_ms, tokens * (8 * D + 8)),
                                      ("backward_kernel", bwd_ms, tokens * (12 * D + 8) + 4 * K * D)):
            if ms_list:
                t = statistics.mean(ms_list)
                hbm[name] = {"kernel_ms": t, "algorithmic_bytes": nbytes, "achieved_gbs": nbytes / (t * 1e-3) / 1e9,
                             "peak_gbs": peaks["hbm_gbs"], "frac": nbytes / (t * 1e-3) / 1e9 / peaks["hbm_gbs"],
                             "traffic": load_traffic(name, workload),
                             "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({peaks['source']})"}
                if t < 0.05:  # a few tens of microseconds: two launches' latency, not bandwidth, sets the time
                    hbm[name]["note"] = (f"launch-latency bound: {nbytes / 1e6:.0f} MB per call would take "
                                         f"{nbytes / peaks['hbm_gbs'] / 1e3:.1f} us at the HBM peak; the bandwidth "
                                         "figure of these kernels is the c3 workload's")
        line = {
            "metric": "vq_lookup_tokens_per_sec", "value": tokens * world * steps / (ms_total * 1e-3),
            "unit": "tokens/s", "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": ms_total / steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(workload, world, codebook),
            "details": {"l2": f"{n_rot} rotating input sets ({n_rot * bytes_per_set / 1e6:.0f} MB) > 126 MB L2",
                       "search_algo": {1: "lowd_fma", 2: "fp32_tile", 3: "tcgen05_bf16x3",
                                       4: "tcgen05_f16_certified", 5: "tcgen05_tf32x3_certified",
                                       6: "dual: lowd_fma + tcgen05_tf32x3 side by side"}.get(algo, str(algo)),
                       "cuda_graph": bool(args.graph),
                       "eager_ms_per_step": (eager_ms_total / steps) if eager_ms_total is not None else None,
                       "rescored_tokens_last_step": int(stats[0]),
                       "multi_group_tokens_last_step": int(stats[2])},
            "roofline": roofline,
            "e2e": {"value": tokens * world * e2e_steps / (e2e_ms * 1e-3), "unit": "tokens/s",
                    "h2d_bytes_per_step": zs[0].numel() * 4, "d2h_bytes_per_step": tokens * 8 + 8,
                    "steps": e2e_steps, "ms_per_step": e2e_ms / e2e_steps,
                    "api": "VectorQuantizer.forward (reference contract) + backward on pinned host latents; H2D of step i+1 "
                           "overlaps step i on a copy stream; every step's loss and indices are copied to pinned host memory "
                           "inside the timed region, the host consumes each loss three steps late (no stall with an empty queue)"},
            "gpu_launches": gpu_launches,
            "gpu_launches_note": "host-side count of libvqb200 kernels per entry point (mirrors the dispatch)",
            "hbm_kernels": hbm,
            "parity_check": parity,
            "clocks": clk.summary(),
            "fma_peak_tflops": {"scalar": fma_scalar, "packed": fma_packed},
        }
        if world == 1 and with_cpu:
            line["cpu_baseline"] = run_cpu_baseline(workload)
            if not args.no_strawman:
                try:
                    line["gpu_strawman_tokens_per_s"] = gpu_strawman_tokens_per_s(workload, device)
                except Exception as e:  # context only, never fatal
                    line["gpu_strawman_tokens_per_s"] = f"failed: {type(e).__name__}"
    del zs, gs, z_host, idx_host, last
    torch.cuda.empty_cache()
    return line


def measure_codebook_variants(args, workload, world, rank, device, kinds=("refinit", "trained", "copied", "clustered")):
    """Search-only timing of `workload` on the secondary codebooks of SURVEY 8(d) and on a collapsed one:
    how many tokens fall through the certified tensor pass to the exact tiers, and what that costs."""
    from vq_gan_b200 import ops
    B, D, H, W, K, _ = WORKLOADS[workload]
    z = torch.randn(B, D, H, W, device=device, generator=torch.Generator(device=device).manual_seed(100 + rank))
    out = {}
    for kind in kinds:
        if kind == "clustered":  # tokens near the 16 centres (tests/test_gpu_adversarial.py)
            g = torch.Generator().manual_seed(1)
            centres = torch.randn(16, D, generator=g).to(device)
            pick = torch.randint(0, 16, (B * H * W,), device=device)
            zz = (centres[pick] + 0.05 * torch.randn(B * H * W, D, device=device)).view(B, H, W, D).permute(0, 3, 1, 2)
            zz = zz.contiguous()
        else:
            zz = z
        E = make_codebook(kind, K, D, device=device, tokens=zz)
        for _ in range(2):
            idx, dmin, st = ops.search(zz, E, args.algo)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 3
        a.record()
        for _ in range(n):
            idx, dmin, st = ops.search(zz, E, args.algo)
        b.record()
        b.synchronize()
        st = st.tolist()
        out[kind] = {"search_ms": a.elapsed_time(b) / n, "algo": int(st[1]), "exact_full_search_tokens": int(st[0]),
                     "multi_group_tokens": int(st[2]), "codes_used": int(torch.unique(idx).numel())}
        if int(st[1]) == 6:
            out[kind]["tensor_role_tokens"] = int(st[3])
        if int(st[1]) == 4:  # of the uncertified tokens, how many the pruned exact tier took (the rest: full fp32 re-search)
            out[kind]["pruned_tier_tokens"] = int(st[3])
        del E, idx, dmin
    del z
    torch.cuda.empty_cache()
    return out


def measure_next_rows(device, peaks):
    """Rows N1 / N2 of SURVEY 8(f) in the driver-run line: the 1x1 convolution either side of the quantizer (forward and
    parameter gradients) and the GroupNorm+SiLU in front of `conv_out`, each timed with CUDA events on inputs larger than
    L2, as a fraction of the HBM peak for its algorithmic bytes, next to the stock torch op, with a parity self-check."""
    import torch.nn.functional as F
    from vq_gan_b200 import ops

    def timed(fn, n=8):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            fn()
        b.record()
        b.synchronize()
        return a.elapsed_time(b) / n

    hbm = peaks["hbm_gbs"]
    out = {"note": "builder-side extras of the same line; inputs 0.5-1 GB each (> 126 MB L2); frac = algorithmic bytes / time / "
                   f"{hbm:.0f} GB/s"}
    g = torch.Generator(device=device).manual_seed(7)
    B, C, H, W = 512, 256, 32, 32
    tokens = B * H * W
    x = torch.randn(B, C, H, W, device=device, generator=g)
    gy = torch.randn(B, C, H, W, device=device, generator=g)
    wt = torch.randn(C, C, device=device, generator=g) / C ** 0.5
    bs = torch.randn(C, device=device, generator=g)
    t_f = timed(lambda: ops.conv1x1(x, wt, bs))
    t_w = timed(lambda: ops.conv1x1_param_grads(gy, x))
    conv = torch.nn.Conv2d(C, C, 1).to(device)
    t_f_ref = timed(lambda: conv(x))
    t_w_ref = timed(lambda: (torch.einsum("bot,bct->oc", gy.reshape(B, C, -1), x.reshape(B, C, -1)), gy.sum(dim=(0, 2, 3))))
    nb = 4.0 * tokens * 2 * C
    y = ops.conv1x1(x[:8], wt, bs)
    ref = torch.einsum("oc,bchw->bohw", wt.double(), x[:8].double()) + bs.double().view(1, -1, 1, 1)
    bound = torch.einsum("oc,bchw->bohw", wt.abs().double(), x[:8].abs().double()) + bs.abs().double().view(1, -1, 1, 1)
    e_f = float(((y.double() - ref).abs() / bound).max())
    gw, gb = ops.conv1x1_param_grads(gy[:32], x[:32])
    wref = torch.einsum("bot,bct->oc", gy[:32].double().reshape(32, C, -1), x[:32].double().reshape(32, C, -1))
    wbound = torch.einsum("bot,bct->oc", gy[:32].double().abs().reshape(32, C, -1), x[:32].double().abs().reshape(32, C, -1))
    e_w = float(((gw.double() - wref).abs() / wbound).max())
    out["n1_conv1x1_256_to_256"] = {
        "tokens": tokens,
        "forward": {"kernel": "conv1x1_tc_kernel", "ms": t_f, "frac_hbm": nb / (t_f * 1e-3) / 1e9 / hbm,
                    "tflops_executed_tf32": 6.0 * tokens * C * C / (t_f * 1e-3) / 1e12, "torch_cudnn_tf32_ms": t_f_ref,
                    "max_err_over_bound": e_f},
        "param_grads": {"kernel": "conv1x1_dw_kernel", "ms": t_w, "frac_hbm": nb / (t_w * 1e-3) / 1e9 / hbm,
                        "tflops_executed_tf32": 6.0 * tokens * C * C / (t_w * 1e-3) / 1e12, "torch_fp32_gemm_ms": t_w_ref,
                        "max_err_over_bound": e_w},
        "parity_check": "ok" if (e_f < 3e-6 and e_w < 4e-6) else "FAILED"}
    del x, gy, y, ref, bound, conv
    torch.cuda.empty_cache()
    B, C = 512, 512
    x = torch.randn(B, C, H, W, device=device, generator=g)
    gy = torch.randn(B, C, H, W, device=device, generator=g)
    w = 1 + 0.3 * torch.randn(C, device=device, generator=g)
    b = 0.2 * torch.randn(C, device=device, generator=g)
    t_f = timed(lambda: ops.groupnorm_silu(x, w, b, 32, 1e-6))
    y, mean, rstd = ops.groupnorm_silu(x, w, b, 32, 1e-6)
    t_b = timed(lambda: ops.groupnorm_silu_backward(gy, x, w, b, mean, rstd, 32))
    t_f_ref = timed(lambda: F.silu(F.group_norm(x, 32, w, b, 1e-6)))
    xs = x[:4].clone().requires_grad_(True)
    yr = F.silu(F.group_norm(xs, 32, w, b, 1e-6))
    yr.backward(gy[:4])
    dx, _, _ = ops.groupnorm_silu_backward(gy[:4], x[:4], w, b, mean[:4 * 32], rstd[:4 * 32], 32)
    e_y = float((y[:4] - yr.detach()).abs().max())
    e_dx = float((dx - xs.grad).abs().max() / xs.grad.abs().max())
    nbx = 4.0 * x.numel()
    out["n2_groupnorm_silu_512ch_32x32"] = {
        "elements": x.numel(),
        "forward": {"kernel": "groupnorm_silu_fwd_reg_kernel", "ms": t_f, "frac_hbm": 2 * nbx / (t_f * 1e-3) / 1e9 / hbm,
                    "torch_ms": t_f_ref, "max_abs_err": e_y},
        "backward": {"kernel": "groupnorm_silu_bwd_reg2_kernel", "ms": t_b, "frac_hbm": 3 * nbx / (t_b * 1e-3) / 1e9 / hbm,
                     "max_rel_err_dx": e_dx},
        "parity_check": "ok" if (e_y < 2e-5 and e_dx < 1e-4) else "FAILED"}
    del x, gy, y
    torch.cuda.empty_cache()
    return out


def run_gpu_arm(args):
    import torch.distributed as dist
    from vq_gan_b200 import ops
    from vq_gan_b200 import distributed as vdist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    cores = vdist.pin_rank_to_gpu_cores(local_rank, int(os.environ.get("LOCAL_WORLD_SIZE", str(world))))
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    peaks = load_peaks()

    def finish(line):
        if rank == 0 and line is not None:
            line.setdefault("details", {})["host_cores_pinned_rank0"] = len(cores) if cores else None
            print(json.dumps(line), flush=True)
        if world > 1:
            dist.barrier(device_ids=[local_rank])
            dist.destroy_process_group()

    if args.workload == "n1":
        B, D, H, W, K, desc = WORKLOADS["n1"]
        return finish(run_conv(args, world, rank, local_rank, device, B, D, H, W, K, desc, peaks))
    if args.workload == "c5":
        B, D, H, W, K, desc = WORKLOADS["c5"]
        return finish(run_sharded_encode(args, world, rank, local_rank, device, B, D, H, W, K, desc, peaks))
    if args.workload == "c4full":
        return finish(run_full_step(args, world, rank, local_rank, device, peaks))

    # FMA peak of this device (roofline denominator for the CUDA-core search)
    fma = (ops.fma_peak_tflops(False), ops.fma_peak_tflops(True))
    if args.workload != "all":
        line = measure_quantizer(args, args.workload, world, rank, local_rank, device, peaks, fma, args.codebook)
        if rank == 0 and args.variants and args.workload in ("c2", "c3"):
            line["codebooks"] = measure_codebook_variants(args, args.workload, world, rank, device)
        return finish(line)

    # default: c2 on top, c3 (and c5 on more than one GPU) as blocks of the same line
    line = measure_quantizer(args, "c2", world, rank, local_rank, device, peaks, fma, args.codebook)
    c3_steps = max(3, min(args.steps, 20))
    c3 = measure_quantizer(args, "c3", world, rank, local_rank, device, peaks, fma, args.codebook, steps=c3_steps)
    variants = measure_codebook_variants(args, "c3", world, rank, device) if rank == 0 else None
    variants_c2 = measure_codebook_variants(args, "c2", world, rank, device) if rank == 0 else None
    c5 = None
    if world > 1:
        B, D, H, W, K, desc = WORKLOADS["c5"]
        c5 = run_sharded_encode(args, world, rank, local_rank, device, B, D, H, W, K, desc, peaks)
    # configs[3]: the reference's own VQVAE training step with the drop-in inside it (short: 3 timed steps)
    c4 = None
    if not args.no_c4:
        try:
            c4 = run_full_step(args, world, rank, local_rank, device, peaks, steps=3, warmup=2)
        except Exception as e:  # the block is context; it must never take the headline down with it
            c4 = {"workload": "c4full", "failed": f"{type(e).__name__}: {e}"[:300]} if rank == 0 else None
    nxt = None
    if world == 1 and not args.no_c4:
        try:
            nxt = measure_next_rows(device, peaks)
        except Exception as e:  # context only, never fatal
            nxt = {"failed": f"{type(e).__name__}: {e}"[:300]}
    if rank == 0:
        if nxt is not None:
            line["next_rows"] = nxt
        c3["codebooks"] = variants
        line["codebooks"] = variants_c2
        line["c3"] = c3
        if c4 is not None:
            line["c4"] = c4
        if c5 is not None:
            line["c5"] = c5
        blocks = [line["parity_check"], c3["parity_check"]] + ([c5["parity_check"]] if c5 else [])
        line["parity_check"] = {"status": "ok" if all(b["status"] == "ok" for b in blocks) else "FAILED",
                                "c2": line["parity_check"], "c3": c3["parity_check"],
                                **({"c5": c5["parity_check"]} if c5 else {})}
    finish(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="all", choices=sorted(WORKLOADS) + ["all", "c4full"],
                    help="all (default): c2 on top + c3 block (+ c5 block on > 1 GPU); c4full: the reference VQVAE "
                         "training step with the drop-in under DDP")
    ap.add_argument("--codebook", default="normal", choices=["normal", "refinit", "trained", "copied", "clustered"],
                    help="codebook law of the quantizer workloads (SURVEY 8d: normal is primary)")
    ap.add_argument("--variants", action="store_true", help="single c2/c3 run: also time the secondary codebooks")
    ap.add_argument("--no-strawman", action="store_true")
    ap.add_argument("--no-c4", action="store_true", help="default workload: skip the full VQVAE training-step block")
    ap.add_argument("--algo", type=int, default=0, help="search kernel override (0 = auto)")
    ap.add_argument("--graph", action="store_true",
                    help="time the step as a CUDA-graph replay (vq_gan_b200.graphs); the per-kernel roofline "
                         "figures still come from an eager pass")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 arm has no CPU fallback")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus != world and world == 1 and args.gpus > 1:
        raise SystemExit("launch multi-GPU runs with torch.distributed.run --nproc-per-node N")
    run_gpu_arm(args)


if __name__ == "__main__":
    main()
