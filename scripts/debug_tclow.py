import os
os.environ.setdefault("VQB200_EXPERIMENTAL", "1")  # route the ops through libvqb200_bench.so (vqb_tune, microbenchmarks)
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vq_gan_b200 import _cabi, ops
cl = int(sys.argv[1]) if len(sys.argv) > 1 else 2
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
K = int(sys.argv[3]) if len(sys.argv) > 3 else 256
_cabi.check(_cabi.lib().vqb_tune(b"tclow_cluster", cl), "vqb_tune")
dbg = int(os.environ.get("TCLOW_DEBUG", "0"))
if dbg:
    _cabi.check(_cabi.lib().vqb_tune(b"tclow_skip_stages", dbg), "vqb_tune")
torch.manual_seed(0)
z = torch.randn(B, 4, 32, 32, device="cuda")
E = torch.randn(K, 4, device="cuda")
i1, d1, _ = ops.search(z, E, 1)
torch.cuda.synchronize()
i5, d5, st = ops.search(z, E, 5)
torch.cuda.synchronize()
print("cluster", cl, "mismatch", int((i1 != i5).sum()), "of", i1.numel(), "stats", st.tolist(), "dmin err", float((d1 - d5).abs().max()))
