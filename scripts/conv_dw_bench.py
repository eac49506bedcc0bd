"""Row N1: the 1x1 convolution's parameter gradients -- this library's 3xTF32 tcgen05 kernel vs the fp32 library GEMM
(torch.einsum, what the backward used before) and cuDNN's own backward-filter (TF32 allowed, the reference's default)."""
import os
os.environ.setdefault("VQB200_EXPERIMENTAL", "1")
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vq_gan_b200 import _cabi, ops
lib = _cabi.lib()
PEAK = 6530.0


def timed(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    b.synchronize()
    return a.elapsed_time(b) / n


for B, Cin, Cout in ((1024, 256, 256), (1024, 128, 64), (1024, 256, 64), (1024, 64, 256), (64, 256, 256)):
    x = torch.randn(B, Cin, 32, 32, device="cuda")
    gy = torch.randn(B, Cout, 32, 32, device="cuda")
    n = B * 1024
    ref = torch.einsum("bot,bct->oc", gy.double().reshape(B, Cout, -1), x.double().reshape(B, Cin, -1))
    bound = torch.einsum("bot,bct->oc", gy.double().abs().reshape(B, Cout, -1), x.double().abs().reshape(B, Cin, -1))
    _cabi.check(lib.vqb_tune(b"dw_hw_trunc", 1), "t")
    t_trunc = timed(lambda: ops.conv1x1_param_grads(gy, x))
    gw, gb = ops.conv1x1_param_grads(gy, x)
    print(f"[{B},{Cin}->{Cout},32,32] dw_hw_trunc=1: {t_trunc:.3f} ms, max err / sum|dy||x| = "
          f"{float(((gw.double() - ref).abs() / bound).max()):.2e}", flush=True)
    _cabi.check(lib.vqb_tune(b"dw_hw_trunc", 0), "t")
    t_mine = timed(lambda: ops.conv1x1_param_grads(gy, x))
    t_lib = timed(lambda: (torch.einsum("bot,bct->oc", gy.reshape(B, Cout, -1), x.reshape(B, Cin, -1)), gy.sum(dim=(0, 2, 3))))
    conv = torch.nn.Conv2d(Cin, Cout, 1).cuda()

    def cudnn_bwd():
        conv.weight.grad = conv.bias.grad = None
        torch.autograd.backward(conv(x), gy, inputs=[conv.weight, conv.bias])
    t_cudnn = timed(cudnn_bwd)
    nbytes = 4.0 * n * (Cin + Cout)
    gw, gb = ops.conv1x1_param_grads(gy, x)
    print(f"[{B},{Cin}->{Cout},32,32] dW+dbias: vqb {t_mine:.3f} ms ({nbytes / t_mine / 1e6 / PEAK:.2f} of HBM peak, "
          f"{2.0 * n * Cin * Cout / t_mine / 1e9:.0f} TFLOP/s algorithmic, {6.0 * n * Cin * Cout / t_mine / 1e9:.0f} executed tf32); "
          f"fp32 library GEMM + sum {t_lib:.3f} ms; conv forward + cuDNN backward-filter {t_cudnn:.3f} ms; "
          f"max err / sum|dy||x| = {float(((gw.double() - ref).abs() / bound).max()):.2e}", flush=True)
