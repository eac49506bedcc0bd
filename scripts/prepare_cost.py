"""Time of the per-forward codebook pre-pass (vqb_codebook_prepare_f32) for the headline shapes."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vq_gan_b200 import ops
for K, D in ((16384, 4), (16384, 256), (8192, 256), (128, 256)):
    E = torch.randn(K, D, device="cuda")
    for _ in range(10):
        ops.prepare_codebook(E)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(200):
        ops.prepare_codebook(E)
    b.record()
    b.synchronize()
    print(f"K={K} D={D}: prepare {a.elapsed_time(b) / 200 * 1e3:.1f} us per call (back to back, includes host launch cost)")
