"""How fast can the codebook-gradient scatter go?  red.global.add.v4.f32 of N x D floats into a [K, D] table, by
access pattern (lanes per row piece) and index distribution.  No other traffic: the floor of the dE part of the backward."""
import os
os.environ.setdefault("VQB200_EXPERIMENTAL", "1")
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vq_gan_b200 import _cabi, ops
lib = _cabi.lib()
N, D = 1 << 20, 256
for K, dist in ((16384, "uniform"), (16384, "search"), (128, "uniform"), (16384, "sorted")):
    if dist == "uniform":
        idx = torch.randint(0, K, (N,), device="cuda")
    elif dist == "sorted":
        idx = torch.sort(torch.randint(0, K, (N,), device="cuda")).values
    else:  # what the C3 search returns on Gaussian data
        z = torch.randn(1024, D, 32, 32, device="cuda")
        E = torch.randn(K, D, generator=torch.Generator().manual_seed(1)).cuda()
        idx = ops.search(z, E)[0].reshape(-1)
        del z
        h = torch.bincount(idx, minlength=K)
        print(f"   search histogram: max {int(h.max())} mean {float(h.float().mean()):.1f} unused {int((h == 0).sum())}")
    table = torch.zeros(K, D, device="cuda")
    for lpr in (1, 4, 16):
        for _ in range(2):
            _cabi.check(lib.vqb_ubench_red(ops._p(table), ops._p(idx), N, D, lpr, ops._stream()), "red")
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            _cabi.check(lib.vqb_ubench_red(ops._p(table), ops._p(idx), N, D, lpr, ops._stream()), "red")
        b.record()
        b.synchronize()
        ms = a.elapsed_time(b) / 5
        print(f"K={K} {dist:8s} lanes_per_row={lpr:2d}: {ms:.3f} ms  ({N * D * 4 / ms / 1e6:.0f} GB/s of reductions)", flush=True)
