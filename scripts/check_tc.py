"""Quick check of the tcgen05 search against the fp32 tile kernel (debug aid)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vq_gan_b200 import ops  # noqa: E402

torch.manual_seed(0)
for (B, D, K) in ((1, 64, 256), (2, 64, 300), (3, 128, 1000), (4, 192, 512), (8, 256, 4096), (64, 256, 16384),
                  (1024, 256, 16384)):
    z = torch.randn(B, D, 32, 32, device="cuda")
    E = torch.randn(K, D, device="cuda")
    if B <= 64:
        i2, d2, _ = ops.search(z, E, 2)
    else:
        i2, d2, _ = ops.search(z, E, 3)  # the fp32 tile kernel would take ~0.25 s here
    torch.cuda.synchronize()
    for algo in (3, 4):
        ops.search(z, E, algo)
        torch.cuda.synchronize()
        ops.PROFILE = []
        i3, d3, st = ops.search(z, E, algo)
        torch.cuda.synchronize()
        (a, b), = ops.PROFILE
        ops.PROFILE = None
        ms = a.elapsed_time(b)
        n = B * 1024
        bad = int((i2 != i3).sum())
        err = float((d2 - d3).abs().max())
        print(f"algo={algo} B={B} D={D} K={K}: mismatches={bad}/{n} max|dmin diff|={err:.3e} stats={st.tolist()} "
              f"{ms:.3f} ms {2.0 * n * K * D / ms / 1e9:.1f} TFLOP/s algorithmic", flush=True)
