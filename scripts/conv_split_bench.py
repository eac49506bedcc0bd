"""pre_quant_conv + quantizer forward, with and without the fused token split (1M tokens, 256 -> 256 channels, K = 16384)."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vq_gan_b200 import QuantConv1x1, VectorQuantizer


def timed(fn, n=8):
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


for Cin, Cout, K in ((256, 256, 16384), (128, 64, 16384)):
    x = torch.randn(1024, Cin, 32, 32, device="cuda")
    vq = VectorQuantizer(K, Cout, 0.25, lazy_stats=True).cuda()
    with torch.no_grad():
        vq.embedding.weight.copy_(torch.randn(K, Cout))
    conv = QuantConv1x1(Cin, Cout, 1).cuda()
    with torch.no_grad():
        t_plain = timed(lambda: vq(conv(x)))
        t_conv = timed(lambda: conv(x))
        conv.feed(vq)
        t_fused = timed(lambda: vq(conv(x)))
        t_conv_f = timed(lambda: conv(x))
    print(f"{Cin}->{Cout}, K={K}, 1M tokens: conv + quantizer forward {t_plain:.3f} ms -> fused split {t_fused:.3f} ms "
          f"(conv alone {t_conv:.3f} -> {t_conv_f:.3f} ms with the split in its epilogue)", flush=True)
