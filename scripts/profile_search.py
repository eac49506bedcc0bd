"""Runs only the search op a few times on one workload (short target for `ncu --set full`)."""
import os
os.environ.setdefault("VQB200_EXPERIMENTAL", "1")  # route the ops through libvqb200_bench.so (vqb_tune, microbenchmarks)
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vq_gan_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--tokens", type=int, default=1 << 20)
ap.add_argument("--dim", type=int, default=4)
ap.add_argument("--codes", type=int, default=16384)
ap.add_argument("--algo", type=int, default=0)
ap.add_argument("--iters", type=int, default=4)
ap.add_argument("--variant", type=int, default=0)
ap.add_argument("--cluster", type=int, default=2)
a = ap.parse_args()
from vq_gan_b200 import _cabi
_cabi.check(_cabi.lib().vqb_tune(b"lowd_variant", a.variant), "vqb_tune")
_cabi.check(_cabi.lib().vqb_tune(b"tc16_cluster", a.cluster), "vqb_tune")
B = a.tokens // 1024
z = torch.randn(B, a.dim, 32, 32, device="cuda")
E = torch.randn(a.codes, a.dim, device="cuda")
for _ in range(a.iters):
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ops.PROFILE = []
    idx, dmin, st = ops.search(z, E, a.algo)
    torch.cuda.synchronize()
    (s, e), = ops.PROFILE
    ms = s.elapsed_time(e)
    print(f"variant={a.variant} cluster={a.cluster} search algo={st.tolist()[1]} tokens={B * 1024} D={a.dim} K={a.codes}: {ms:.3f} ms "
          f"{2.0 * B * 1024 * a.codes * a.dim / ms / 1e9:.2f} TFLOP/s algorithmic, rescored={st.tolist()[0]}")
