"""Checks the low-D tensor search (algo 5) against the CUDA-core kernel (algo 1): indices and dmin must be identical."""
import os
os.environ.setdefault("VQB200_EXPERIMENTAL", "1")  # route the ops through libvqb200_bench.so (vqb_tune, microbenchmarks)
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vq_gan_b200 import _cabi, ops  # noqa: E402

torch.manual_seed(0)
cases = [(1, 4, 256, "n"), (2, 4, 300, "n"), (3, 3, 1000, "n"), (4, 8, 512, "n"), (8, 16, 4096, "n"), (5, 1, 700, "n"),
         (64, 4, 16384, "n"), (64, 4, 16384, "u"), (1024, 4, 16384, "n"), (256, 8, 16384, "n"), (128, 16, 8192, "n")]
if len(sys.argv) > 1:
    cases = cases[:int(sys.argv[1])]
for cl in (2, 1, 4):
    _cabi.check(_cabi.lib().vqb_tune(b"tclow_cluster", cl), "vqb_tune")
    for (B, D, K, law) in cases:
        if cl != 2 and B < 1024:
            continue
        z = torch.randn(B, D, 32, 32, device="cuda")
        E = torch.randn(K, D, device="cuda") if law == "n" else (torch.rand(K, D, device="cuda") * 2 - 1) / K
        i1, d1, _ = ops.search(z, E, 1)
        torch.cuda.synchronize()
        ms = {}
        for algo in (1, 5):
            ops.search(z, E, algo)
            torch.cuda.synchronize()
            ops.PROFILE = []
            i5, d5, st = ops.search(z, E, algo)
            torch.cuda.synchronize()
            (a, b), = ops.PROFILE
            ops.PROFILE = None
            ms[algo] = a.elapsed_time(b)
        n = B * 1024
        bad = int((i1 != i5).sum())
        err = float((d1 - d5).abs().max())
        print(f"cluster={cl} B={B} D={D} K={K} {law}: mismatches={bad}/{n} max|dmin diff|={err:.3e} stats={st.tolist()} "
              f"algo1 {ms[1]:.3f} ms, algo5 {ms[5]:.3f} ms ({2.0 * n * K * D / ms[5] / 1e9:.1f} TFLOP/s algorithmic)", flush=True)
