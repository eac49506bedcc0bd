"""A/B: CTA-wide tiled forward tail vs the experimental warp-private variant (vqb_tune tail_warp)."""
import os
os.environ.setdefault("VQB200_EXPERIMENTAL", "1")  # route the ops through libvqb200_bench.so (vqb_tune, microbenchmarks)
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vq_gan_b200 import _cabi, ops
lib = _cabi.lib()
for D, K, B in ((256, 16384, 1024), (128, 4096, 1024), (64, 4096, 1024)):
    z = torch.randn(B, D, 32, 32, device="cuda")
    E = torch.randn(K, D, device="cuda")
    idx = torch.randint(0, K, (B, 32, 32), device="cuda")
    n = B * 1024
    outs = {}
    for mode in (0, 2):
        _cabi.check(lib.vqb_tune(b"tail_warp", mode), "t")
        best = 1e9
        for _ in range(4):
            zq = torch.empty_like(z); loss = torch.empty(2, device="cuda")
            pb = lib.vqb_tail_partials_bytes(n); part = torch.empty(pb, dtype=torch.uint8, device="cuda")
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _cabi.check(lib.vqb_gather_loss_st_f32(ops._p(z), ops._p(E), ops._p(idx), B, D, 1024, K, 0.25, ops._p(zq), ops._p(loss),
                                                   ops._p(part), pb, None, ops._stream()), "tail")
            e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        outs[mode] = (zq, loss)
        print(f"D={D} tail_warp={mode}: {best:.3f} ms ({n*(8*D+8)/best/1e6:.0f} GB/s)", flush=True)
    print("   same z_q:", torch.equal(outs[0][0], outs[1][0]), " loss:", outs[0][1].tolist(), outs[1][1].tolist())
_cabi.check(lib.vqb_tune(b"tail_warp", 1), "t")
