"""Two-process tests of the multi-GPU modes that run on ANY GPU box.

With >= 2 visible GPUs the ranks use one GPU each over NCCL (the product configuration).  On a 1-GPU box
(the driver's GPU-test box) both ranks share cuda:0 and the collectives go through gloo -- NCCL refuses two
ranks on one device -- so the exchange logic of every mode is still exercised against this library's CUDA
kernels: data-parallel statistics, codebook-sharded search (incl. a NaN code on the last shard), the EMA
extension, and a DistributedDataParallel training step of the reference VQVAE with the drop-in quantizer
(train_vqgan.py:197-209 wraps the model the same way through accelerate).
"""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _setup(rank, world, port):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    ngpu = torch.cuda.device_count()
    dev = torch.device("cuda", rank if ngpu >= world else 0)
    torch.cuda.set_device(dev)
    if ngpu >= world:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    return dev


def _worker(rank, world, port, q):
    dev = _setup(rank, world, port)
    out = {"backend": dist.get_backend()}
    try:
        from oracle import ref_loader
        from oracle import vq_oracle as orc
        from vq_gan_b200 import EMAVectorQuantizer, VectorQuantizer, ops
        from vq_gan_b200 import distributed as vdist

        # ---- (1) data parallel: packed stats all-reduce == global-batch gradient / histogram / loss
        g = torch.Generator().manual_seed(5)
        z = torch.randn(4, 8, 16, 16, generator=g)
        E = torch.randn(300, 8, generator=g)
        lo, hi = vdist.shard_range(z.shape[0], world, rank)
        vq = VectorQuantizer(300, 8, 0.25, lazy_stats=True).to(dev)
        with torch.no_grad():
            vq.embedding.weight.copy_(E)
        zl = z[lo:hi].to(dev).requires_grad_(True)
        z_q, ld, idx = vq(zl)
        ld["vq_loss"].backward()
        usage, _, _ = ops.codebook_usage(idx, 300)
        sq = (ld["codebook_loss"] * float(zl.numel())).reshape(1)
        dE, hist, s = vdist.allreduce_stats(vq.embedding.weight.grad, usage, sq)
        fg = orc.forward(z, E, 0.25)
        bg = orc.backward(z, E, fg["indices"], 0.25, None, 1.0)
        hg, _ = orc.codebook_usage(fg["indices"], 300)
        out["dp_dE"] = bool(torch.allclose(dE.cpu(), bg["dE"], rtol=1e-5, atol=1e-6 * float(bg["dE"].abs().max())))
        out["dp_hist"] = bool(torch.equal(hist.cpu(), hg))
        out["dp_loss"] = bool(torch.allclose(s.cpu() / z.numel(), fg["mse"].reshape(1), rtol=1e-6))

        # ---- (2) codebook sharded: winners equal the unsharded search
        for D, K in ((4, 512), (64, 512)):
            g = torch.Generator().manual_seed(9)
            z = torch.randn(2, D, 16, 16, generator=g)
            E = torch.randn(K, D, generator=g)
            E[K // 2:] = E[:K // 2]  # duplicates across shards: ties must go to the lower global index
            klo, khi = vdist.shard_range(K, world, rank)
            idx, dmin = vdist.sharded_search(z.to(dev), E[klo:khi].contiguous().to(dev), klo)
            want, wmin, _ = ops.search(z.to(dev), E.to(dev))
            out[f"sharded_d{D}"] = bool(torch.equal(idx, want)) and int(idx.max()) < K // 2 \
                and bool(torch.allclose(dmin, wmin, rtol=1e-5, atol=1e-5))
            zr = torch.randn(2, D, 16, 16, generator=torch.Generator().manual_seed(20 + rank))
            idx2, _ = vdist.sharded_search_dp(zr.to(dev), E[klo:khi].contiguous().to(dev), klo)
            want2, _, _ = ops.search(zr.to(dev), E.to(dev))
            out[f"sharded_dp_d{D}"] = bool(torch.equal(idx2, want2))
            # the bulk-encode pipeline (all-gather of batch i+1 under the search of batch i) returns the same winners
            batches = [torch.randn(2, D, 16, 16, generator=torch.Generator().manual_seed(40 + 10 * j + rank)).to(dev)
                       for j in range(3)]
            enc = vdist.ShardedEncoder(E[klo:khi].contiguous().to(dev), klo)
            got = [i for i, _ in enc.encode_all(batches)]
            want = [ops.search(b, E.to(dev))[0] for b in batches]
            out[f"sharded_encoder_d{D}"] = len(got) == 3 and all(torch.equal(a, b) for a, b in zip(got, want))
            idx_c, _ = vdist.sharded_search_dp(zr.to(dev), E[klo:khi].contiguous().to(dev), klo, chunks=2)
            out[f"sharded_dp_chunked_d{D}"] = bool(torch.equal(idx_c, want2))
            # a NaN code on the LAST shard wins everywhere, like the unsharded ATen argmin (quantizer.py:76)
            En = torch.randn(K, D, generator=torch.Generator().manual_seed(11))
            En[K - 5, 1] = float("nan")
            idx3, _ = vdist.sharded_search(z.to(dev), En[klo:khi].contiguous().to(dev), klo)
            want3, _, _ = ops.search(z.to(dev), En.to(dev))
            out[f"sharded_nan_d{D}"] = bool(torch.equal(idx3, want3)) and int(want3.min()) == K - 5

        # ---- (3) EMA extension: counts / sums summed over the ranks, identical replicas
        torch.manual_seed(7)
        ema = EMAVectorQuantizer(64, 8, 0.25, decay=0.9).to(dev)
        ema.train()
        zg = torch.randn(4, 8, 8, 8, generator=torch.Generator().manual_seed(8))
        w0 = ema.embedding.weight.detach().cpu().clone()
        lo, hi = vdist.shard_range(4, world, rank)
        _, _, idx_l = ema(zg[lo:hi].to(dev))
        gathered = [torch.empty_like(ema.embedding.weight) for _ in range(world)]
        dist.all_gather(gathered, ema.embedding.weight.detach().clone())
        out["ema_replicas_bitwise_equal"] = all(torch.equal(gathered[0], t) for t in gathered)
        rows = orc.tokens_of(zg)
        idx_g = orc.nearest_code(rows, w0)
        want_w, _, _ = orc.ema_update(w0, torch.zeros(64), w0.clone(), rows, idx_g, 0.9, 1e-5)
        out["ema_matches_global_batch"] = bool(torch.allclose(ema.embedding.weight.detach().cpu(), want_w,
                                                              rtol=1e-5, atol=1e-7))

        # ---- (4) DDP training step of the reference VQVAE with the drop-in == the global-batch step
        if ref_loader.available():
            import copy
            from torch.nn.parallel import DistributedDataParallel as DDP
            torch.backends.cudnn.deterministic, torch.backends.cudnn.benchmark = True, False
            torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = False
            VQVAE = ref_loader.reference_vqvae_class()
            kw = ref_loader.default_vqvae_kwargs()
            kw.update(ch=64)
            torch.manual_seed(42)
            base = VQVAE(**kw)
            base.quantizer = VectorQuantizer(kw["num_embeddings"], kw["embedding_dim"], kw["commitment_cost"])
            base = base.to(dev).train()
            images = torch.rand(4, 3, 256, 256, generator=torch.Generator().manual_seed(100)).to(dev)

            def step(model, x):
                opt = torch.optim.Adam(model.parameters(), lr=4.5e-5, betas=(0.5, 0.9))
                rec, ldd = model(x)
                total = (rec - x).abs().mean() + ldd["vq_loss"]
                opt.zero_grad()
                total.backward()
                torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
                opt.step()

            single = copy.deepcopy(base)
            step(single, images)                                   # global batch, no DDP
            ddp = DDP(copy.deepcopy(base), device_ids=[dev.index])
            lo, hi = vdist.shard_range(4, world, rank)
            step(ddp, images[lo:hi])                               # split_batches=True (train_vqgan.py:112)
            w_s = single.quantizer.embedding.weight.detach()
            w_d = ddp.module.quantizer.embedding.weight.detach()
            moved = float((w_s - base.quantizer.embedding.weight.detach()).abs().max())
            out["ddp_codebook_moved"] = moved > 1e-6
            out["ddp_codebook_equals_global_batch"] = bool(torch.allclose(w_d, w_s, rtol=1e-5, atol=1e-7))
            # Adam's first step moves every weight by ~lr * sign(g): where a gradient is pure rounding noise its sign
            # (and so 2 * lr) may differ between the two batch splits; everything else must agree closely
            with torch.no_grad():
                diffs = torch.cat([(a - b).abs().reshape(-1) for a, b in zip(single.parameters(), ddp.module.parameters())])
            out["ddp_param_max_diff"] = float(diffs.max())
            out["ddp_param_frac_above_1e-6"] = float((diffs > 1e-6).float().mean())
            out["ddp_all_params_equal_global_batch"] = bool(diffs.max() <= 2.05 * 4.5e-5) and \
                float((diffs > 1e-6).float().mean()) < 1e-3
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


def test_two_rank_modes_match_single_process_results():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29800 + (os.getpid() % 1000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=600) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    print("two-rank results:", res[0])
    for r in range(world):
        bad = [k for k, v in res[r].items() if isinstance(v, bool) and v is not True]
        assert not bad, (r, bad, res[r])
