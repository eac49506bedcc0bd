"""Next row N1, second half: `pre_quant_conv` fused with the quantizer's token split (vq_vae.py:115 feeding :118).
`QuantConv1x1.feed(vq)` makes the convolution's epilogue emit the fp16 token rows / scales / norms of the tensor search
from the TMEM accumulator; the quantizer then skips its own pass over z.  Everything downstream must be unchanged."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("Cin,Cout,K,B,HW", [(128, 256, 2048, 16, 1024), (64, 64, 4096, 64, 1024), (256, 128, 1000, 33, 400),
                                              (96, 192, 3000, 8, 1024)])
def test_fused_conv_split_changes_nothing_downstream(Cin, Cout, K, B, HW):
    from vq_gan_b200 import QuantConv1x1, VectorQuantizer, ops
    torch.manual_seed(Cin + Cout)
    H = 32 if HW % 32 == 0 else 20
    W = HW // H
    vq = VectorQuantizer(K, Cout, 0.25).cuda()
    with torch.no_grad():
        vq.embedding.weight.copy_(torch.randn(K, Cout))
    conv = QuantConv1x1(Cin, Cout, 1).cuda()
    x = torch.randn(B, Cin, H, W, device="cuda")
    auto = ops.split_eligible(Cin, Cout, B * H * W, K)   # Cout <= 128: `feed` switches by itself

    # plain: convolution, then the quantizer splits z itself
    xa = x.clone().requires_grad_(True)
    ya = conv(xa)
    assert not hasattr(ya, "_vqb_presplit")
    zq_a, ld_a, idx_a = vq(ya)
    (ld_a["vq_loss"] + (zq_a * 0.01).sum()).backward()
    stats_a = vq.last_search_stats.tolist()
    ga = (xa.grad.clone(), conv.weight.grad.clone(), vq.embedding.weight.grad.clone())
    conv.zero_grad()
    vq.zero_grad()

    # fused
    conv.feed(vq)
    xb = x.clone().requires_grad_(True)
    if auto:
        yb = conv(xb)
    else:  # the op takes Cout up to 256; the module only switches where it pays (ops.split_eligible)
        assert not hasattr(conv(xb.detach()), "_vqb_presplit")
        w = vq.embedding.weight
        pack = ops.prepare_codebook(w.detach())
        yb, sws = ops.conv1x1_split(xb, conv.weight, conv.bias, w.detach(), pack)
        yb._vqb_presplit = (sws, pack, (w.data_ptr(), w._version, tuple(yb.shape)))
    assert hasattr(yb, "_vqb_presplit")
    assert torch.equal(yb, ya)                      # same convolution output, bit for bit
    zq_b, ld_b, idx_b = vq(yb)
    (ld_b["vq_loss"] + (zq_b * 0.01).sum()).backward()
    stats_b = vq.last_search_stats.tolist()
    assert stats_a[1] == stats_b[1] == 4            # both took the fp16 tensor search
    assert torch.equal(idx_a, idx_b) and torch.equal(zq_a, zq_b)
    assert ld_a["codebook_loss"] == ld_b["codebook_loss"]
    assert torch.equal(xa.grad, xb.grad) or torch.allclose(ga[0], xb.grad, rtol=1e-6, atol=1e-9)
    # dW is a tokens-long reduction whose partial sums meet in atomics (conv1x1_dw_kernel): the order, hence the last bits,
    # differ from run to run -- tolerance relative to the largest entry, like dE
    assert torch.allclose(ga[1], conv.weight.grad, rtol=1e-5, atol=1e-5 * float(ga[1].abs().max()))
    assert torch.allclose(ga[2], vq.embedding.weight.grad, rtol=1e-5, atol=1e-9)
    print(f"Cin={Cin} Cout={Cout} K={K}: stats plain {stats_a} fused {stats_b}")

    # the attribute is only honoured for this very tensor and codebook version
    with torch.no_grad():
        vq.embedding.weight.add_(0.0)               # bumps the version: the stale split must be ignored, not trusted
    zq_c, _, idx_c = vq(yb.detach())
    assert torch.equal(idx_c, idx_a)


def test_conv_split_falls_back_when_shapes_do_not_qualify():
    from vq_gan_b200 import QuantConv1x1, VectorQuantizer
    vq = VectorQuantizer(128, 64).cuda()            # tiny problem: the quantizer takes the fp32 tile kernel
    conv = QuantConv1x1(32, 64, 1).cuda().feed(vq)
    y = conv(torch.randn(2, 32, 8, 8, device="cuda"))
    assert not hasattr(y, "_vqb_presplit")
    z_q, ld, idx = vq(y)
    assert idx.shape == (2, 8, 8)
