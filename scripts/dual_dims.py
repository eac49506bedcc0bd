"""Two-engine low-D search over the embedding dimension: algo 6 against the single engines (1 = CUDA cores,
5 = tensor cores), 1M tokens x 16384 codes; results must be bit-identical to algo 1.  Optional split sweep."""
import os
os.environ.setdefault("VQB200_EXPERIMENTAL", "1")
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vq_gan_b200 import _cabi, ops
lib = _cabi.lib()


def timed(fn, n=6):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    b.synchronize()
    return a.elapsed_time(b) / n


B, K = 1024, 16384
for D in (3, 4, 5, 6, 8, 12, 16):
    zs = [torch.randn(B, D, 32, 32, device="cuda") for _ in range(3)]
    E = torch.randn(K, D, generator=torch.Generator().manual_seed(1)).cuda()
    it = [0]

    def run(algo):
        it[0] = (it[0] + 1) % 3
        return ops.search(zs[it[0]], E, algo)
    _cabi.check(lib.vqb_tune(b"dual_permille", 0), "t")   # (0 is rejected by the range check: see below)
    i1, d1, _ = ops.search(zs[0], E, 1)
    t1, t5 = timed(lambda: run(1)), timed(lambda: run(5))
    t6 = timed(lambda: run(6))
    i6, d6, st = ops.search(zs[0], E, 6)
    same = torch.equal(i1, i6) and torch.equal(d1, d6)
    line = (f"D={D:2d}: cuda cores {t1:.3f} ms, tensor {t5:.3f} ms, both {t6:.3f} ms = {B * 1024 / t6 / 1e3:.0f} M tok/s "
            f"({min(t1, t5) / t6:.2f}x the better single engine), tensor share {st[3].item() / (B * 1024):.2f}, identical {same}")
    if D in (8, 16):
        sweep = []
        for pm in (450, 550, 650, 750):
            _cabi.check(lib.vqb_tune(b"dual_permille", pm), "t")
            sweep.append(f"{pm / 10:.0f}%: {timed(lambda: run(6)):.3f}")
        line += " | split sweep " + ", ".join(sweep)
    print(line, flush=True)
