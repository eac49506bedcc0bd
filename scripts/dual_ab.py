"""Two-engine low-D search (algo 6: CUDA-core role + tensor role in one CTA): sweep of the image split
(vqb_tune dual_permille), against the single engines (algo 1 / 5).  Results must be bit-identical to algo 1."""
import os
os.environ.setdefault("VQB200_EXPERIMENTAL", "1")
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vq_gan_b200 import _cabi, ops
lib = _cabi.lib()


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


D, B, K = 4, 1024, 16384
zs = [torch.randn(B, D, 32, 32, device="cuda") for _ in range(4)]
E = torch.randn(K, D, generator=torch.Generator().manual_seed(1)).cuda()
it = [0]


def run(algo):
    it[0] = (it[0] + 1) % 4
    return ops.search(zs[it[0]], E, algo)


i1, d1, _ = ops.search(zs[0], E, 1)
torch.cuda.synchronize()
i6, d6, st = ops.search(zs[0], E, 6)
torch.cuda.synchronize()
print("first call ok:", torch.equal(i1, i6), torch.equal(d1, d6), st.tolist(), flush=True)
t1 = timed(lambda: run(1))
t5 = timed(lambda: run(5))
print(f"D={D}: algo1 {t1:.3f} ms, algo5 {t5:.3f} ms", flush=True)
for pm in (350, 400, 450, 500, 550, 600, 650, 700):
    _cabi.check(lib.vqb_tune(b"dual_permille", pm), "tune")
    t6 = timed(lambda: run(6))
    i6, d6, st = ops.search(zs[0], E, 6)
    same = torch.equal(i1, i6) and torch.equal(d1, d6)
    print(f"   dual tensor share {pm / 10:.0f}%: {t6:.3f} ms  ({B * 1024 / t6 / 1e3:.0f} M tok/s) bit-identical to algo 1: {same} "
          f"stats={st.tolist()}", flush=True)
# ragged / small shapes and the tie-heavy reference init
_cabi.check(lib.vqb_tune(b"dual_permille", 550), "tune")
for Bx, HW, Kx in ((2, 7, 100), (3, 1000, 777), (64, 1024, 16384), (17, 333, 5000)):
    zz = torch.randn(Bx, D, HW, device="cuda")
    for name, EE in (("normal", torch.randn(Kx, D, device="cuda")), ("refinit", (torch.rand(Kx, D, device="cuda") * 2 - 1) / Kx)):
        a1, b1, _ = ops.search(zz, EE, 1)
        a6, b6, s6 = ops.search(zz, EE, 6)
        print(f"   B={Bx} HW={HW} K={Kx} {name}: identical {torch.equal(a1, a6) and torch.equal(b1, b6)} stats={s6.tolist()}", flush=True)
