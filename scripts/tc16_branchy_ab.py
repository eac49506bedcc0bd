"""A/B of the fp16 tensor search epilogue: unconditional 7-op sorted insert vs insert-on-improvement (vqb_tune tc16_branchy)."""
import os
os.environ.setdefault("VQB200_EXPERIMENTAL", "1")
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vq_gan_b200 import _cabi, ops
lib = _cabi.lib()
torch.manual_seed(0)
for D in (64, 128, 192, 256, 100):
    z = torch.randn(1024, D, 32, 32, device="cuda")
    E = torch.randn(16384, D, device="cuda")
    ref = None
    for mode in (0, 1, 0, 1):
        _cabi.check(lib.vqb_tune(b"tc16_branchy", mode), "t")
        best = 1e9
        for _ in range(4):
            ops.PROFILE = []
            idx, dmin, st = ops.search(z, E, 4)
            torch.cuda.synchronize()
            (s, e), = ops.PROFILE
            best = min(best, s.elapsed_time(e))
        ops.PROFILE = None
        if ref is None:
            ref = idx
        n = 1 << 20
        print(f"D={D:3d} branchy={mode}: {best:7.3f} ms {2.0 * n * 16384 * D / best / 1e9:7.1f} TFLOP/s  same idx: {torch.equal(idx, ref)} stats={st.tolist()}",
              flush=True)
_cabi.check(lib.vqb_tune(b"tc16_branchy", 0), "t")
