"""The oracle against the UNMODIFIED reference module staged in oracle/_ref (see oracle/make_ref.py), on
CPU.  The golden fixtures pin the oracle to outputs the reference produced once; this pins it to the module
itself every time the suite runs where the staged copy exists (build container and GPU box)."""
import numpy as np
import pytest
import torch

from cases import CASES, make_case
from oracle import ref_loader
from oracle import vq_oracle as orc

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="oracle/_ref not staged (run oracle/make_ref.py)")

SMALL = [n for n, c in CASES.items() if c["B"] * c["H"] * c["W"] * c["K"] <= 1 << 22]


@pytest.mark.parametrize("name", SMALL)
def test_oracle_equals_reference_module(name):
    c = make_case(name)
    K, D = c["E"].shape
    Ref = ref_loader.reference_quantizer_class()
    ref = Ref(K, D, c["beta"])
    with torch.no_grad():
        ref.embedding.weight.copy_(c["E"])
    z = c["z"].clone().requires_grad_(True)
    z_q, loss_dict, idx = ref(z)
    (loss_dict["vq_loss"] + (z_q * c["g_zq"]).sum()).backward()
    o = orc.autograd_step(c["z"], c["E"], c["beta"], c["g_zq"])
    assert torch.equal(idx, o["indices"])
    assert torch.equal(z_q.detach(), o["z_q"])
    assert loss_dict["vq_loss"].item() == o["vq_loss"].item()
    assert loss_dict["codebook_loss"] == o["mse"].item() == loss_dict["commitment_loss"]
    assert torch.equal(z.grad, o["dz"])
    np.testing.assert_allclose(ref.embedding.weight.grad.numpy(), o["dE"].numpy(), rtol=1e-6, atol=1e-12)
    # closed-form backward of the oracle == autograd of the reference
    b = orc.backward(c["z"], c["E"], idx.reshape(-1), c["beta"], c["g_zq"], 1.0)
    np.testing.assert_allclose(z.grad.numpy(), b["dz"].numpy(), rtol=1e-6, atol=1e-9)
    usage, ratio = ref.get_codebook_usage(idx)
    u2, r2 = orc.codebook_usage(idx, K)
    assert torch.equal(usage, u2) and ratio == r2
    assert torch.equal(ref.get_codebook_entry(idx), orc.codebook_entry(c["E"], idx))


def test_reference_vqvae_loads_without_its_package_init():
    VQVAE = ref_loader.reference_vqvae_class()
    kw = ref_loader.default_vqvae_kwargs()
    kw.update(ch=32, ch_mult=(1, 2), num_res_blocks=1, attn_resolutions=(), z_channels=16, embedding_dim=16,
              num_embeddings=32)
    torch.manual_seed(0)
    m = VQVAE(**kw)
    assert "quantizer.embedding.weight" in m.state_dict()   # the checkpoint key the drop-in must keep
    x = torch.rand(1, 3, 32, 32)
    rec, ld = m(x)
    assert rec.shape == x.shape and set(ld) == {"vq_loss", "codebook_loss", "commitment_loss", "codebook_usage_ratio"}
