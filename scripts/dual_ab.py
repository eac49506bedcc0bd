"""Two-engine low-D search (vqb_search_dual_f32): sweep of the image split, against the single engines (algo 1 / 5)."""
import os
os.environ.setdefault("VQB200_EXPERIMENTAL", "1")
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vq_gan_b200 import ops


def timed(fn, n=10):
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


for D in ((4,) if os.environ.get("VQB_DUAL_TENSOR_FIRST") else (4, 2, 3)):
    B, K = 1024, 16384
    zs = [torch.randn(B, D, 32, 32, device="cuda") for _ in range(4)]
    E = torch.randn(K, D, generator=torch.Generator().manual_seed(1)).cuda()
    it = [0]

    def run(algo):
        it[0] = (it[0] + 1) % 4
        return ops.search(zs[it[0]], E, algo)
    t1 = timed(lambda: run(1))
    t5 = timed(lambda: run(5))
    i1, d1, _ = ops.search(zs[0], E, 1)
    print(f"D={D}: algo1 {t1:.3f} ms, algo5 {t5:.3f} ms", flush=True)
    for frac in (-1, 0.40, 0.45, 0.50, 0.55, 0.60, 0.65, 0.70):
        ops.DUAL_TENSOR_IMAGES = -1 if frac < 0 else int(B * frac)
        t6 = timed(lambda: run(6))
        i6, d6, st = ops.search(zs[0], E, 6)
        same = torch.equal(i1, i6) and torch.equal(d1, d6)
        print(f"   dual frac={frac}: {t6:.3f} ms  ({B * 1024 / t6 / 1e3:.0f} M tok/s) bit-identical to algo 1: {same} stats={st.tolist()}",
              flush=True)
    ops.DUAL_TENSOR_IMAGES = -1
