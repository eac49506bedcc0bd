"""Row N1, parameter gradients of the 1x1 convolutions either side of the quantizer (autograd of vq_vae.py:115,121):
dW[o, c] = sum over tokens dy[b, o, t] x[b, c, t], dbias[o] = sum over tokens dy[b, o, t] -- the 3xTF32 tcgen05 kernel
(vqb_conv1x1_dw_f32) against an fp64 evaluation.  Tolerance: 3xTF32 keeps ~2^-21 relative error per product and the
tokens-long fp32 accumulation adds its own rounding, so |err| <= 4e-6 * sum_t |dy||x| elementwise; the fp32 library
GEMM it replaces is held to the same bound in the same test (so the bound is not vacuous)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

SHAPES = [
    (4, 256, 256, 32, 32),   # the default VQGAN bottleneck: two 128-row tiles x one 256-channel chunk
    (3, 128, 64, 16, 16),    # one tile, one chunk
    (5, 32, 16, 6, 6),       # HW = 36: a partial last token block (TMA zero fill)
    (2, 192, 200, 8, 8),     # Cout not a multiple of 128 (the second tile's box runs into the next image), chunk 192
    (2, 320, 48, 8, 8),      # chunk 160 (a multiple of 32); see (2, 80, 48, 8, 8) for a partly unused last TMEM read
    (2, 512, 256, 16, 16),   # two chunks x two tiles
    (2, 80, 48, 8, 8),       # chunk 80: the last 32-column TMEM read is half unused
    (2, 768, 32, 8, 8),      # three chunks
    (3, 16, 1, 8, 8),        # one output channel
    (1, 64, 4, 128, 128),    # one image, many token blocks per CTA
]


@pytest.mark.parametrize("B,Cin,Cout,H,W", SHAPES)
def test_param_grads_match_fp64(B, Cin, Cout, H, W):
    from vq_gan_b200 import lib, ops
    assert lib().vqb_conv1x1_dw_supported(Cin, Cout, H * W)
    g = torch.Generator().manual_seed(Cin * 7 + Cout)
    x = (torch.randn(B, Cin, H, W, generator=g) + 0.3).cuda()
    gy = (torch.randn(B, Cout, H, W, generator=g) * torch.rand(1, Cout, 1, 1, generator=g) - 0.1).cuda()
    gw, gb = ops.conv1x1_param_grads(gy, x)
    x64, g64 = x.double().reshape(B, Cin, -1), gy.double().reshape(B, Cout, -1)
    ref = torch.einsum("bot,bct->oc", g64, x64)
    bound = torch.einsum("bot,bct->oc", g64.abs(), x64.abs())
    err = (gw.double() - ref).abs()
    assert gw.shape == (Cout, Cin) and gb.shape == (Cout,)
    assert float((err / bound).max()) < 4e-6, float((err / bound).max())
    lib32 = torch.einsum("bot,bct->oc", gy.reshape(B, Cout, -1), x.reshape(B, Cin, -1))
    assert float(((lib32.double() - ref).abs() / bound).max()) < 4e-6
    bref = g64.sum(dim=(0, 2))
    bbound = g64.abs().sum(dim=(0, 2))
    assert float(((gb.double() - bref).abs() / bbound).max()) < 2e-6


def test_accumulates_and_rejects_other_shapes():
    from vq_gan_b200 import lib, ops
    import ctypes
    assert not lib().vqb_conv1x1_dw_supported(24, 64, 64)      # Cin % 16
    assert not lib().vqb_conv1x1_dw_supported(64, 300, 64)     # Cout > 256
    assert not lib().vqb_conv1x1_dw_supported(64, 64, 35)      # HW % 4
    assert not lib().vqb_conv1x1_dw_supported(64, 64, 16)      # HW < 32
    x = torch.randn(2, 24, 8, 8, device="cuda")
    gy = torch.randn(2, 64, 8, 8, device="cuda")
    gw = torch.zeros(64, 24, device="cuda")
    rc = lib().vqb_conv1x1_dw_f32(ctypes.c_void_p(gy.data_ptr()), ctypes.c_void_p(x.data_ptr()), 2, 24, 64, 64,
                                  ctypes.c_void_p(gw.data_ptr()), None, None)
    assert rc == -3  # VQB_ERR_UNSUPPORTED
    # unsupported shapes keep the library GEMM through the Python wrapper
    w2, b2 = ops.conv1x1_param_grads(gy, x)
    torch.testing.assert_close(w2, torch.einsum("bot,bct->oc", gy.reshape(2, 64, -1), x.reshape(2, 24, -1)))
    # the C entry point accumulates: calling it twice on the same buffers doubles the result
    x = torch.randn(2, 32, 8, 8, device="cuda")
    gw = torch.zeros(64, 32, device="cuda")
    gb = torch.zeros(64, device="cuda")
    for _ in range(2):
        rc = lib().vqb_conv1x1_dw_f32(ctypes.c_void_p(gy.data_ptr()), ctypes.c_void_p(x.data_ptr()), 2, 32, 64, 64,
                                      ctypes.c_void_p(gw.data_ptr()), ctypes.c_void_p(gb.data_ptr()),
                                      ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        assert rc == 0
    once_w, once_b = ops.conv1x1_param_grads(gy, x)
    torch.testing.assert_close(gw, 2 * once_w, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(gb, 2 * once_b, rtol=1e-5, atol=1e-5)


def test_module_backward_uses_the_kernel():
    """QuantConv1x1's weight / bias gradients come from the library kernel and match nn.Conv2d's."""
    from vq_gan_b200 import QuantConv1x1, ops
    torch.manual_seed(3)
    ref = torch.nn.Conv2d(128, 64, 1).cuda()
    mine = QuantConv1x1(128, 64).cuda()
    mine.load_state_dict(ref.state_dict(), strict=True)
    x = torch.randn(8, 128, 32, 32, device="cuda")
    gy = torch.randn(8, 64, 32, 32, device="cuda")
    ops.PROFILE_CONV_DW = []
    mine(x).backward(gy)
    torch.cuda.synchronize()
    n = len(ops.PROFILE_CONV_DW)
    ops.PROFILE_CONV_DW = None
    assert n == 1
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        ref(x).backward(gy)
    finally:
        torch.backends.cudnn.allow_tf32 = old
    torch.testing.assert_close(mine.weight.grad, ref.weight.grad, rtol=1e-4, atol=1e-4 * float(ref.weight.grad.abs().max()))
    torch.testing.assert_close(mine.bias.grad, ref.bias.grad, rtol=1e-4, atol=1e-4 * float(ref.bias.grad.abs().max()))
