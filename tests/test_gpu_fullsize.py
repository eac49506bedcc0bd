"""Size-independent properties at BASELINE.json's full sizes (1M tokens)."""
import numpy as np
import pytest
import torch

from oracle import vq_oracle as orc

pytestmark = pytest.mark.gpu


def _c2():
    g = torch.Generator().manual_seed(0)
    z = torch.randn(1024, 4, 32, 32, generator=g)
    E = torch.randn(16384, 4, generator=torch.Generator().manual_seed(1))
    return z, E


def test_c2_full_size_properties():
    from vq_gan_b200 import VectorQuantizer, ops
    z, E = _c2()
    vq = VectorQuantizer(16384, 4).cuda()
    with torch.no_grad():
        vq.embedding.weight.copy_(E)
    zc = z.cuda().requires_grad_(True)
    z_q, ld, idx = vq(zc)
    ld["vq_loss"].backward()
    N = idx.numel()
    assert N == 1 << 20 and int(idx.min()) >= 0 and int(idx.max()) < 16384
    # (1) histogram is a partition of the tokens
    usage, ratio = vq.get_codebook_usage(idx)
    assert int(usage.sum()) == N and 0 < ratio <= 1
    # (2) oracle parity on a seeded 8192-token sample, with the near-tie band
    pick = torch.randperm(N, generator=torch.Generator().manual_seed(7))[:8192]
    rows = orc.tokens_of(z)[pick]
    ref = orc.search_with_gap(rows, E)
    rep = orc.compare_indices(idx.reshape(-1).cpu()[pick], ref)
    print("c2 sample:", rep)
    assert rep["outside"] == 0
    # (3) the fp32 tile kernel agrees (different summation order -> only near-ties may differ)
    idx2, dmin2, _ = ops.search(zc.detach(), vq.embedding.weight.detach(), 2)
    idx1, dmin1, _ = ops.search(zc.detach(), vq.embedding.weight.detach(), 1)
    assert torch.equal(idx1, idx)
    differ = idx1 != idx2
    print("lowd vs fp32 differing tokens:", int(differ.sum()))
    assert float((dmin1 - dmin2).abs().max()) < 1e-4
    assert int(differ.sum()) < 64
    # (4) idempotence: quantising the chosen code vectors returns the same codes, zero loss
    e = vq.get_codebook_entry(idx)
    z_q2, ld2, idxe = vq(e)
    assert torch.equal(idxe, idx) and ld2["codebook_loss"] == 0.0
    # (5) loss equals the mean squared distance recomputed from the indices in float64
    rows_all = orc.tokens_of(z)
    sq = ((E[idx.reshape(-1).cpu()].double() - rows_all.double()) ** 2).sum().item() / z.numel()
    np.testing.assert_allclose(ld["codebook_loss"], sq, rtol=2e-6)
    # (6) gradient mass: sum_k dE[k] == -beta * sum_i dz_loss_i (linearity of the scatter)
    dE = vq.embedding.weight.grad.double().sum(0).cpu()
    dz = orc.tokens_of(zc.grad.cpu()).double().sum(0)
    np.testing.assert_allclose(dE.numpy(), (-0.25 * dz).numpy(), rtol=1e-3, atol=1e-7)


def test_c2_tensor_path_is_bit_identical_to_fma_kernel():
    """algo 5 (tf32x3 tcgen05 + certified chunk + exact FMA re-score) must return exactly the indices
    and minimum scores of algo 1 at full size, for a realistic and for the tie-heavy reference init."""
    from vq_gan_b200 import ops
    z, E = _c2()
    zc = z.cuda()
    K = E.shape[0]
    books = {"normal": E.cuda(),
             "reference_init": ((torch.rand(K, 4, generator=torch.Generator().manual_seed(3)) * 2 - 1) / K).cuda()}
    for name, Ec in books.items():
        i1, d1, _ = ops.search(zc, Ec, 1)
        i5, d5, st = ops.search(zc, Ec, 5)
        print(f"c2 {name}: unsure tokens re-searched exactly = {int(st[0])} of {i1.numel()}")
        assert int(st[1]) == 5
        assert torch.equal(i1, i5) and torch.equal(d1, d5)
        # both engines in one CTA (the default for this shape): same bits again
        i6, d6, st6 = ops.search(zc, Ec, 6)
        i0, d0, st0 = ops.search(zc, Ec, 0)
        assert int(st6[1]) == 6 and int(st0[1]) == 6 and 0 < int(st6[3]) < i1.numel()
        assert torch.equal(i1, i6) and torch.equal(d1, d6) and torch.equal(i1, i0)
    # other low dimensions (different numbers of k-steps), ragged token count, K not a multiple of 128
    for D, Kx, B, HW in ((1, 700, 3, 331), (3, 1000, 5, 1024), (8, 5000, 9, 777), (16, 4099, 7, 1024)):
        g = torch.Generator().manual_seed(D)
        zz = torch.randn(B, D, HW, generator=g).cuda()
        EE = torch.randn(Kx, D, generator=g).cuda()
        i1, d1, _ = ops.search(zz, EE, 1)
        i5, d5, _ = ops.search(zz, EE, 5)
        assert torch.equal(i1, i5) and torch.equal(d1, d5), (D, Kx)
    # NaN / inf tokens and a NaN code follow the same ATen rules as algo 1
    zz = torch.randn(4, 4, 16, 16, generator=torch.Generator().manual_seed(9))
    zz[0, 1, 2, 3] = float("nan")
    zz[1, 0, 0, 0] = float("inf")
    EE = torch.randn(777, 4, generator=torch.Generator().manual_seed(10))
    for nan_code in (False, True):
        if nan_code:
            EE[123, 2] = float("nan")
        i1, _, _ = ops.search(zz.cuda(), EE.cuda(), 1)
        i5, _, _ = ops.search(zz.cuda(), EE.cuda(), 5)
        assert torch.equal(i1, i5)


def test_c3_slice_tensor_path_against_fp32_kernel():
    from vq_gan_b200 import ops
    g = torch.Generator().manual_seed(0)
    z = torch.randn(64, 256, 32, 32, generator=g).cuda()      # 65536 tokens
    E = torch.randn(16384, 256, generator=torch.Generator().manual_seed(1)).cuda()
    idx3, dmin3, st = ops.search(z, E, 3)
    idx2, dmin2, _ = ops.search(z, E, 2)
    print("tensor-path stats:", st.tolist(), "differ:", int((idx3 != idx2).sum()))
    assert torch.equal(idx3, idx2)
    # single-pass fp16 path: certified candidates re-scored in fp32 with a different summation
    # order, so it may differ from the fp32 tile kernel only on fp32-rounding-level ties
    idx4, dmin4, st4 = ops.search(z, E, 4)
    differ = (idx4 != idx2).reshape(-1)
    print("fp16-path stats:", st4.tolist(), "differ:", int(differ.sum()))
    assert int(differ.sum()) <= 8
    if int(differ.sum()):
        rows = orc.tokens_of(z.cpu())[differ.cpu()]
        d64 = orc.half_distance(rows, E.cpu())
        a = d64.gather(1, idx4.reshape(-1)[differ].cpu().unsqueeze(1))
        b = d64.gather(1, idx2.reshape(-1)[differ].cpu().unsqueeze(1))
        assert float((a - b).abs().max()) < 1e-4
    assert float((dmin4 - dmin2).abs().max()) < 1e-3
    # idempotence on the tensor path
    from vq_gan_b200 import VectorQuantizer
    vq = VectorQuantizer(16384, 256).cuda()
    with torch.no_grad():
        vq.embedding.weight.copy_(E)
    e = vq.get_codebook_entry(idx3)
    assert torch.equal(vq.encode_indices(e), idx3)
