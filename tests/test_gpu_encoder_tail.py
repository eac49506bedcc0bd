"""Next row N2: `norm_out -> SiLU` of the reference Encoder / Decoder (encoder_decoder.py:166-167, 249-250) as one
kernel, against the same ops in stock torch fp32 (forward and backward).  Tolerances: rtol 1e-5 / atol 1e-5 forward
(different but equally exact summation order for mean / variance; __expf in the sigmoid), 1e-4 relative to the largest
gradient backward (long reductions over H*W and the batch)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _ref(x, gn):
    return F.silu(F.group_norm(x, gn.num_groups, gn.weight, gn.bias, gn.eps))


@pytest.mark.parametrize("B,C,H,W,G", [(4, 512, 32, 32, 32),     # encoder tail (staged in shared memory)
                                        (2, 128, 256, 256, 32),   # decoder tail (1 MB per group: global passes)
                                        (3, 48, 7, 9, 8),         # ragged: HW % 4 != 0
                                        (1, 64, 16, 16, 64),      # one channel per group
                                        (2, 64, 64, 64, 16),      # cluster of 4 / 2 CTAs per group, 4 segments per channel
                                        (2, 64, 64, 64, 8),       # cluster of 8 / 4
                                        (3, 6, 50, 50, 2),        # a channel split between two CTAs of a cluster
                                        (2, 12, 35, 35, 2),       # cluster of 2 with HW % 4 != 0 (scalar path)
                                        (5, 512, 16, 16, 32)])    # 16 x 16 latents (f = 16 encoders)
def test_groupnorm_silu_matches_torch(B, C, H, W, G):
    from vq_gan_b200 import GroupNormSiLU
    torch.manual_seed(C + H)
    gn = GroupNormSiLU(G, C, eps=1e-6).cuda()
    with torch.no_grad():
        gn.weight.copy_(1 + 0.3 * torch.randn(C))
        gn.bias.copy_(0.2 * torch.randn(C))
    x = (2.0 * torch.randn(B, C, H, W, device="cuda") + 0.5).requires_grad_(True)
    gy = torch.randn(B, C, H, W, device="cuda")
    y = gn(x)
    assert y.shape == x.shape and y.is_contiguous()
    y.backward(gy)
    got = (y.detach(), x.grad.clone(), gn.weight.grad.clone(), gn.bias.grad.clone())
    x.grad = None
    gn.weight.grad = gn.bias.grad = None
    yr = _ref(x, gn)
    yr.backward(gy)
    torch.testing.assert_close(got[0], yr.detach(), rtol=1e-5, atol=1e-5)
    for a, b, name in ((got[1], x.grad, "dx"), (got[2], gn.weight.grad, "dweight"), (got[3], gn.bias.grad, "dbias")):
        scale = float(b.abs().max())
        assert float((a - b).abs().max()) <= 1e-4 * scale + 1e-6, (name, float((a - b).abs().max()), scale)


def test_state_dict_is_groupnorm_compatible_and_cpu_is_rejected():
    from vq_gan_b200 import GroupNormSiLU
    ref = torch.nn.GroupNorm(32, 512, eps=1e-6)
    mine = GroupNormSiLU(32, 512, eps=1e-6)
    assert list(mine.state_dict()) == list(ref.state_dict())
    mine.load_state_dict(ref.state_dict(), strict=True)
    with pytest.raises(RuntimeError):
        mine(torch.randn(1, 512, 4, 4))


def test_encoder_tail_of_the_reference_model():
    """The last three lines of the reference Encoder.forward on its own modules: conv_out(silu(norm_out(h)))."""
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("oracle/_ref not staged")
    from vq_gan_b200 import encoder_tail
    torch.manual_seed(0)
    VQVAE = ref_loader.reference_vqvae_class()
    m = VQVAE(**ref_loader.default_vqvae_kwargs()).cuda()
    h = torch.randn(2, 512, 32, 32, device="cuda")
    with torch.no_grad():
        want = m.encoder.conv_out(F.silu(m.encoder.norm_out(h)))
        got = encoder_tail(m.encoder, h)
    torch.testing.assert_close(got, want, rtol=1e-4, atol=1e-4)
