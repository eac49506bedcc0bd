"""Times the instruction-mix microbenchmarks of the low-D inner loop (see csrc/vqb_ubench.cu)."""
import os
os.environ.setdefault("VQB200_EXPERIMENTAL", "1")  # route the ops through libvqb200_bench.so (vqb_tune, microbenchmarks)
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vq_gan_b200 import _cabi, ops  # noqa: E402

lib = _cabi.lib()
src = torch.randn(1 << 16, device="cuda")
sink = torch.zeros(1, device="cuda")
names = {0: "FFMA2 only", 1: "FFMA2 + FMNMX3 (shipped mix)", 2: "FFMA only", 3: "FFMA + FMNMX3",
         4: "FFMA2 + 2x FMNMX", 5: "FFMA2 + FMNMX3 one pair late"}
print(f"fma peak: scalar {ops.fma_peak_tflops(False):.1f} packed {ops.fma_peak_tflops(True):.1f} TFLOP/s")
for mode in range(6):
    best = 0.0
    for _ in range(4):
        flops = ctypes.c_double(0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        _cabi.check(lib.vqb_ubench_launch(mode, 8, src.data_ptr(), sink.data_ptr(), ctypes.byref(flops),
                                          ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "ubench")
        b.record()
        b.synchronize()
        best = max(best, flops.value / (a.elapsed_time(b) * 1e-3) / 1e12)
    print(f"mode {mode} {names[mode]:30s}: {best:6.2f} TFLOP/s")
