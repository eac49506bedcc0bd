import os
os.environ.setdefault("VQB200_EXPERIMENTAL", "1")  # route the ops through libvqb200_bench.so (vqb_tune, microbenchmarks)
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vq_gan_b200 import ops, _cabi
_cabi.check(_cabi.lib().vqb_tune(b'conv_debug', int(os.environ.get('CONV_DEBUG', '0'))), 'tune')
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
x = torch.randn(B, 256, 32, 32, device="cuda"); w = torch.randn(256, 256, device="cuda") / 16; b = torch.randn(256, device="cuda")
for _ in range(3):
    ops.PROFILE_CONV = []
    ops.conv1x1(x, w, b, 1); torch.cuda.synchronize()
    print("ms", ops.PROFILE_CONV[0][0].elapsed_time(ops.PROFILE_CONV[0][1]))
