"""world_size-2 gloo tests of the multi-GPU host logic (run on CPU).

The local compute either side of the collectives is CUDA-only in the product, so
here the oracle plays the local kernel and the tests check the exchange step:
the packed all-reduce of codebook statistics (data parallel) and the MIN
reduction of (distance, index) keys (codebook sharded)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, fn_name, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        q.put((rank, globals()[fn_name](rank, world)))
    finally:
        dist.destroy_process_group()


def _run(fn_name, world=2):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, fn_name, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return out


def dp_stats(rank, world):
    from cases import make_case
    from oracle import vq_oracle as orc
    from vq_gan_b200 import distributed as vdist
    c = make_case("small_d8")
    z, E, beta = c["z"], c["E"], c["beta"]
    lo, hi = vdist.shard_range(z.shape[0], world, rank)  # shard the batch
    zl = z[lo:hi]
    f = orc.forward(zl, E, beta)
    b = orc.backward(zl, E, f["indices"], beta, None, 1.0)
    hist, _ = orc.codebook_usage(f["indices"], E.shape[0])
    sq = (f["mse"].double() * zl.numel()).float().reshape(1)
    dE, h, s = vdist.allreduce_stats(b["dE"], hist, sq, average_dE=True)
    # global-batch reference
    fg = orc.forward(z, E, beta)
    bg = orc.backward(z, E, fg["indices"], beta, None, 1.0)
    hg, _ = orc.codebook_usage(fg["indices"], E.shape[0])
    ok_dE = torch.allclose(dE, bg["dE"], rtol=1e-5, atol=1e-8)
    ok_h = torch.equal(h, hg)
    ok_s = torch.allclose(s / z.numel(), fg["mse"].reshape(1), rtol=1e-6)
    return bool(ok_dE and ok_h and ok_s)


def sharded_keys(rank, world):
    from cases import make_case
    from oracle import vq_oracle as orc
    from vq_gan_b200 import distributed as vdist
    c = make_case("ties_d4")  # duplicate codes across the two shards: ties must go to the lower index
    rows, E = orc.tokens_of(c["z"]), c["E"]
    lo, hi = vdist.shard_range(E.shape[0], world, rank)
    d = orc.half_distance(rows, E[lo:hi]).float()
    dmin, idx = d.min(dim=1)
    keys = orc.argmin_key(dmin, idx + lo)
    vdist.reduce_argmin_keys(keys)
    got = keys & 0xFFFFFFFF
    want = torch.argmin(orc.half_distance(rows, E).float(), dim=1)
    return bool(torch.equal(got, want) and int(got.max()) < E.shape[0] // 2)


def sharded_keys_nan_code(rank, world):
    """A NaN code on the LAST shard must win for every finite token, exactly like the unsharded argmin
    (quantizer.py:76: NaN is minimal), whatever finite scores the other shards report."""
    from cases import make_case
    from oracle import vq_oracle as orc
    from vq_gan_b200 import distributed as vdist
    c = make_case("small_d8")
    rows, E = orc.tokens_of(c["z"]), c["E"].clone()
    nan_code = E.shape[0] - 3
    E[nan_code, 2] = float("nan")
    lo, hi = vdist.shard_range(E.shape[0], world, rank)
    d = orc.distance_matrix(rows, E[lo:hi])
    idx = torch.argmin(d, dim=1)                       # ATen: first NaN of the shard, else the minimum
    dmin = d.gather(1, idx[:, None]).squeeze(1)         # NaN where the shard's winner is the NaN code
    keys = orc.argmin_key(dmin, idx + lo)
    vdist.reduce_argmin_keys(keys)
    got = keys & 0xFFFFFFFF
    want = torch.argmin(orc.distance_matrix(rows, E), dim=1)
    return bool(torch.equal(got, want) and int(want[0]) == nan_code)


def test_dp_stats_allreduce_matches_global_batch():
    assert _run("dp_stats") == {0: True, 1: True}


def test_codebook_sharded_min_reduce_matches_global_argmin():
    assert _run("sharded_keys") == {0: True, 1: True}


def test_codebook_sharded_nan_code_on_last_shard():
    assert _run("sharded_keys_nan_code") == {0: True, 1: True}
