"""A/B: CTA-wide tiled backward vs the warp-private variant (vqb_tune bwd_warp)."""
import os
os.environ.setdefault("VQB200_EXPERIMENTAL", "1")  # route the ops through libvqb200_bench.so (vqb_tune, microbenchmarks)
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vq_gan_b200 import _cabi, ops
lib = _cabi.lib()
for D, K, B in ((256, 16384, 1024), (128, 4096, 1024), (256, 128, 64)):
    z = torch.randn(B, D, 32, 32, device="cuda"); g = torch.randn(B, D, 32, 32, device="cuda")
    E = torch.randn(K, D, device="cuda")
    idx = torch.randint(0, K, (B, 32, 32), device="cuda")
    gv = torch.ones(1, device="cuda")
    n = B * 1024
    outs = {}
    for mode in (0, 2):
        _cabi.check(lib.vqb_tune(b"bwd_warp", mode), "t")
        best = 1e9
        for _ in range(4):
            dz = torch.empty_like(z); dE = torch.zeros_like(E); hist = torch.zeros(K, dtype=torch.int64, device="cuda")
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _cabi.check(lib.vqb_backward_f32(ops._p(z), ops._p(E), ops._p(idx), ops._p(g), ops._p(gv), 0.25, B, D, 1024, K,
                                             ops._p(dz), ops._p(dE), ops._p(hist), ops._stream()), "bwd")
            e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        outs[mode] = (dz, dE, hist)
        print(f"D={D} K={K} bwd_warp={mode}: {best:.3f} ms ({(n*(12*D+8)+4*K*D)/best/1e6:.0f} GB/s)", flush=True)
    a, b = outs[0], outs[2]
    print("   same dz:", torch.equal(a[0], b[0]), " dE max rel diff:", float((a[1] - b[1]).abs().max() / a[1].abs().max()),
          " same hist:", torch.equal(a[2], b[2]))
_cabi.check(lib.vqb_tune(b"bwd_warp", 0), "t")
