"""Quick check + timing of the 1x1 convolution kernels against torch (fp64 reference on the GPU)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vq_gan_b200 import ops
torch.manual_seed(0)
torch.backends.cuda.matmul.allow_tf32 = False
cases = [(2, 32, 16, 8, 8), (2, 64, 64, 16, 16), (3, 256, 256, 32, 32), (2, 128, 48, 9, 11), (64, 256, 256, 32, 32),
         (1024, 256, 256, 32, 32), (1024, 4, 256, 32, 32), (1024, 256, 4, 32, 32)]
if len(sys.argv) > 1:
    cases = cases[:int(sys.argv[1])]
for (B, Cin, Cout, H, W) in cases:
    x = torch.randn(B, Cin, H, W, device="cuda")
    w = torch.randn(Cout, Cin, device="cuda") / Cin ** 0.5
    b = torch.randn(Cout, device="cuda")
    for algo in (1, 2):
        if algo == 1 and (Cin % 32 or Cout % 16 or Cout > 256):
            continue
        if algo == 2 and B * H * W * Cin * Cout > 3e12:
            pass
        y = ops.conv1x1(x, w, b, algo)
        torch.cuda.synchronize()
        nb = min(B, 8)
        ref = torch.einsum("oc,bchw->bohw", w.double(), x[:nb].double()) + b.double().view(1, -1, 1, 1)
        bound = torch.einsum("oc,bchw->bohw", w.abs().double(), x[:nb].abs().double()) + b.abs().double().view(1, -1, 1, 1)
        rel = float(((y[:nb].double() - ref).abs() / bound).max())
        ops.PROFILE_CONV = []
        for _ in range(3):
            ops.conv1x1(x, w, b, algo)
        torch.cuda.synchronize()
        ms = min(a.elapsed_time(e) for a, e in ops.PROFILE_CONV)
        ops.PROFILE_CONV = None
        n = B * H * W
        gb = n * (Cin + Cout) * 4 / 1e9
        print(f"B={B} Cin={Cin} Cout={Cout} HW={H*W} algo={algo}: max err/bound {rel:.2e}  {ms:.3f} ms  "
              f"{2.0 * n * Cin * Cout / ms / 1e9:.1f} TFLOP/s  {gb / ms * 1e3:.0f} GB/s", flush=True)
    if B >= 64:
        t = []
        for _ in range(3):
            a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            torch.nn.functional.conv2d(x, w.view(Cout, Cin, 1, 1), b)
            e.record(); torch.cuda.synchronize()
            t.append(a.elapsed_time(e))
        print(f"    torch conv2d (cudnn, default flags): {min(t):.3f} ms", flush=True)
