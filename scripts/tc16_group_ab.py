"""A/B of the fp16 tensor search's candidate-group size (vqb_tune tc16_group): groups of 4 codes (2.5 ALU operations per
score in the epilogue, 4 / 8 / 12 exact candidates) vs groups of 8 (1.5 per score, 8 / 16 / 24 candidates)."""
import os
os.environ.setdefault("VQB200_EXPERIMENTAL", "1")
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vq_gan_b200 import _cabi, ops
lib = _cabi.lib()
torch.manual_seed(0)
dims = tuple(int(a) for a in sys.argv[1:]) or (32, 64, 100, 128, 192, 256)
for D in dims:
    z = torch.randn(1024, D, 32, 32, device="cuda")
    for law in ("normal", "refinit"):
        E = torch.randn(16384, D, device="cuda") if law == "normal" else (torch.rand(16384, D, device="cuda") * 2 - 1) / 16384
        ref = None
        for mode in (4, 8, 4, 8):
            _cabi.check(lib.vqb_tune(b"tc16_group", mode), "t")
            best = 1e9
            for _ in range(4):
                ops.PROFILE = []
                idx, dmin, st = ops.search(z, E, 4)
                torch.cuda.synchronize()
                (s, e), = ops.PROFILE
                best = min(best, s.elapsed_time(e))
            ops.PROFILE = None
            if ref is None:
                ref = (idx, dmin)
            n = 1 << 20
            print(f"D={D:3d} {law:8s} group={mode}: {best:7.3f} ms {2.0 * n * 16384 * D / best / 1e9:7.1f} TFLOP/s  same idx: "
                  f"{torch.equal(idx, ref[0])} same dmin: {torch.equal(dmin, ref[1])} stats={st.tolist()}", flush=True)
    # exactness of the wide groups against the CUDA-core kernel on a slice
    _cabi.check(lib.vqb_tune(b"tc16_group", 8), "t")
    zs = z[:64].contiguous()
    E = torch.randn(16384, D, device="cuda")
    i8, d8, _ = ops.search(zs, E, 4)
    i2, d2, _ = ops.search(zs, E, 2)
    print(f"D={D:3d} group=8 vs fp32 tile kernel on 65536 tokens: idx mismatches {(i8 != i2).sum().item()}, "
          f"max |dmin diff| {(d8 - d2).abs().max().item():.3e}", flush=True)
_cabi.check(lib.vqb_tune(b"tc16_group", 0), "t")
