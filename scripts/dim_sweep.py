"""Search time per kernel over the embedding dimension (1M tokens x 16384 codes)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vq_gan_b200 import ops
torch.manual_seed(0)
for D, algos in ((2, (1, 5)), (4, (1, 5)), (8, (1, 5)), (16, (1, 5)), (32, (2,)), (64, (3, 4)), (128, (3, 4)), (192, (4,)), (256, (3, 4)), (320, (4,)), (384, (4,)), (512, (2, 4))):
    B = 1024 if D not in (32, 512) else (128 if D == 32 else 256)
    z = torch.randn(B, D, 32, 32, device="cuda")
    E = torch.randn(16384, D, device="cuda")
    for a in algos:
        best = 1e9
        for _ in range(3):
            ops.PROFILE = []
            _, _, st = ops.search(z, E, a)
            torch.cuda.synchronize()
            (s, e), = ops.PROFILE
            best = min(best, s.elapsed_time(e))
        ops.PROFILE = None
        n = B * 1024
        print(f"D={D:3d} algo={a} tokens={n}: {best:8.3f} ms  {n / best / 1e3:8.1f} M tok/s  {2.0 * n * 16384 * D / best / 1e9:7.1f} TFLOP/s algorithmic  stats={st.tolist()}", flush=True)
