"""Multi-GPU paths over NCCL (needs >= 2 visible GPUs; skipped otherwise).  The host logic of
the same paths is covered on CPU by tests/test_distributed_gloo.py."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from oracle import vq_oracle as orc
        from vq_gan_b200 import VectorQuantizer, ops
        from vq_gan_b200 import distributed as vdist
        out = {}
        # ---- data parallel: per-rank batch, packed stats all-reduce == global-batch gradient
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
        # ---- codebook sharded, same tokens everywhere
        for D, K in ((4, 512), (64, 512)):
            g = torch.Generator().manual_seed(9)
            z = torch.randn(2, D, 16, 16, generator=g)
            E = torch.randn(K, D, generator=g)
            E[K // 2:] = E[:K // 2]  # duplicates across shards: ties must go to the lower global index
            klo, khi = vdist.shard_range(K, world, rank)
            idx, dmin = vdist.sharded_search(z.to(dev), E[klo:khi].contiguous().to(dev), klo)
            want, _, _ = ops.search(z.to(dev), E.to(dev))
            out[f"sharded_d{D}"] = bool(torch.equal(idx, want)) and int(idx.max()) < K // 2
            # ---- codebook sharded, different tokens per rank (all-gather + MIN reduce-scatter)
            zr = torch.randn(2, D, 16, 16, generator=torch.Generator().manual_seed(20 + rank))
            idx2, _ = vdist.sharded_search_dp(zr.to(dev), E[klo:khi].contiguous().to(dev), klo)
            want2, _, _ = ops.search(zr.to(dev), E.to(dev))
            out[f"sharded_dp_d{D}"] = bool(torch.equal(idx2, want2))
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_nccl_dp_and_sharded_modes():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 1000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=240) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in range(world):
        assert all(res[r].values()), res
