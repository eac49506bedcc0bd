"""Idle cost of the pruned exact tier: healthy data (the tier's kernels are launched and return at once), tier off / on."""
import os
os.environ.setdefault("VQB200_EXPERIMENTAL", "1")
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vq_gan_b200 import _cabi, ops
lib = _cabi.lib()
g = torch.Generator().manual_seed(9)
for (B, D, K) in ((1024, 256, 16384), (1024, 64, 16384), (448, 256, 8192), (64, 256, 16384)):
    z = torch.randn(B, D, 32, 32, generator=g).cuda()
    E = torch.randn(K, D, generator=g).cuda()
    res = {}
    for rep in range(2):
        for name, pr in (("off", 0), ("on", 1)):
            _cabi.check(lib.vqb_tune(b"tc16_pruned", pr), "t")
            ts = []
            for _ in range(12):
                ops.PROFILE = []
                ops.search(z, E, 4)
                torch.cuda.synchronize()
                (s, e), = ops.PROFILE
                ts.append(s.elapsed_time(e))
            ops.PROFILE = None
            ts.sort()
            res.setdefault(name, []).append(sum(ts[:6]) / 6)
    print(f"tokens={B * 1024} D={D} K={K}: " + "  ".join(f"{k} {min(v):.4f} ms" for k, v in res.items()), flush=True)
_cabi.check(lib.vqb_tune(b"tc16_pruned", 1), "t")
