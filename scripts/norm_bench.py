"""GroupNorm + SiLU (row N2): fused kernel vs stock torch, forward and backward, encoder-tail and decoder-tail shapes."""
import os
os.environ.setdefault("VQB200_EXPERIMENTAL", "1")
import sys
import torch
import torch.nn.functional as F
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


for B, C, H, W, mode in [(b, c, h, w, m) for (b, c, h, w) in ((64, 512, 32, 32), (1024, 512, 32, 32), (256, 512, 16, 16),
                                                             (16, 128, 256, 256)) for m in (0, 1)]:
    _cabi.check(lib.vqb_tune(b"norm_cluster", mode), "t")
    x = torch.randn(B, C, H, W, device="cuda")
    w = torch.randn(C, device="cuda")
    b = torch.randn(C, device="cuda")
    gy = torch.randn_like(x)
    nbytes = x.numel() * 4
    tf = timed(lambda: ops.groupnorm_silu(x, w, b, 32, 1e-6))
    tt = timed(lambda: F.silu(F.group_norm(x, 32, w, b, 1e-6)))
    y, mean, rstd = ops.groupnorm_silu(x, w, b, 32, 1e-6)
    tb = timed(lambda: ops.groupnorm_silu_backward(gy, x, w, b, mean, rstd, 32))
    if mode == 1:  # A/B of the forward: register-resident (default) vs staged in shared memory (norm_fwd_reg 0)
        y1 = ops.groupnorm_silu(x, w, b, 32, 1e-6)[0]
        _cabi.check(lib.vqb_tune(b"norm_fwd_reg", 0), "t")
        tf0 = timed(lambda: ops.groupnorm_silu(x, w, b, 32, 1e-6))
        y0 = ops.groupnorm_silu(x, w, b, 32, 1e-6)[0]
        _cabi.check(lib.vqb_tune(b"norm_fwd_reg", 1), "t")
        print(f"[{B},{C},{H},{W}] forward: register-resident (product) {tf:.3f} ms vs staged {tf0:.3f} ms; max |dy| {float((y0 - y1).abs().max()):.2e}", flush=True)
    if mode == 1:  # A/B of the register-resident backward: one CTA per SM (norm_bwd2 0) vs two (default)
        d1 = ops.groupnorm_silu_backward(gy, x, w, b, mean, rstd, 32)[0]
        _cabi.check(lib.vqb_tune(b"norm_bwd2", 0), "t")
        tb1 = timed(lambda: ops.groupnorm_silu_backward(gy, x, w, b, mean, rstd, 32))
        d0 = ops.groupnorm_silu_backward(gy, x, w, b, mean, rstd, 32)[0]
        _cabi.check(lib.vqb_tune(b"norm_bwd2", 1), "t")
        print(f"[{B},{C},{H},{W}] backward: two CTAs per SM, xhat in shared memory (product) {tb:.3f} ms vs one CTA per SM, all in "
              f"registers {tb1:.3f} ms; same dx: {torch.equal(d0, d1)}", flush=True)
    xr = x.clone().requires_grad_(True)

    def torch_fb():
        xr.grad = None
        F.silu(F.group_norm(xr, 32, w, b, 1e-6)).backward(gy)
    ttb = timed(torch_fb)
    print(f"[{B},{C},{H},{W}] cluster={mode} fwd fused {tf:.3f} ms ({2 * nbytes / tf / 1e6 / PEAK:.2f} of HBM peak for 2 passes) vs torch {tt:.3f} ms; "
          f"bwd fused {tb:.3f} ms ({3 * nbytes / tb / 1e6 / PEAK:.2f} for 3 passes); torch fwd+bwd {ttb:.3f} ms", flush=True)
