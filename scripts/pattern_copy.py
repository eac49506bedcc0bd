"""Bandwidth ceiling of the NCHW 32-token x strided-channel access pattern used by the tail kernels."""
import os
os.environ.setdefault("VQB200_EXPERIMENTAL", "1")  # route the ops through libvqb200_bench.so (vqb_tune, microbenchmarks)
import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vq_gan_b200 import _cabi
lib = _cabi.lib()
B, D, HW = 1024, 256, 1024
a = torch.randn(B, D, HW, device="cuda"); b = torch.randn(B, D, HW, device="cuda"); out = torch.empty_like(a)
s = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
P = lambda t: ctypes.c_void_p(t.data_ptr())
for mode, name, nbytes in ((0, "strided copy (1 read + 1 write)", 2), (1, "strided a+b (2 reads + 1 write)", 3), (2, "linear float4 copy", 2),
                           (3, "strided copy, 128 tokens/CTA float4", 2), (4, "strided a+b, 128 tokens/CTA float4", 3)):
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _cabi.check(lib.vqb_ubench_copy(P(a), P(b), P(out), B, D, HW, mode, s), "copy")
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(f"{name}: {best:.3f} ms  {nbytes * a.numel() * 4 / best / 1e6:.0f} GB/s", flush=True)
t = 1e9
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out.copy_(a); e1.record(); torch.cuda.synchronize(); t = min(t, e0.elapsed_time(e1))
print(f"torch copy_: {t:.3f} ms  {2 * a.numel() * 4 / t / 1e6:.0f} GB/s")

t = 1e9
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); torch.add(a, b, out=out); e1.record(); torch.cuda.synchronize(); t = min(t, e0.elapsed_time(e1))
print(f"torch add (linear, 2 reads + 1 write): {t:.3f} ms  {3 * a.numel() * 4 / t / 1e6:.0f} GB/s")
