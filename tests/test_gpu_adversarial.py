"""Adversarial inputs for the certified tensor paths: exact ties on a coarse grid, clustered codebooks with
tiny gaps, and badly scaled data.  The low-D tensor kernel (algo 5) must still equal the FMA kernel (algo 1)
bit for bit; the fp16 tensor kernel (algo 4) may differ from the fp32 tile kernel (algo 2) only on
float32-rounding-level ties (checked in float64)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _make(kind, D, K, N, seed):
    g = torch.Generator().manual_seed(seed)
    if kind == "grid":          # coordinates are multiples of 1/8: thousands of exact fp32 ties
        z = torch.randint(-16, 17, (N, D), generator=g).float() / 8
        E = torch.randint(-16, 17, (K, D), generator=g).float() / 8
    elif kind == "clustered":   # 16 cluster centres, codes differ by 1e-4: top-2 gaps ~1e-5
        centres = torch.randn(16, D, generator=g)
        E = centres[torch.randint(0, 16, (K,), generator=g)] + 1e-4 * torch.randn(K, D, generator=g)
        z = centres[torch.randint(0, 16, (N,), generator=g)] + 0.05 * torch.randn(N, D, generator=g)
    elif kind == "scaled":      # |z| ~ 1e4, |e| ~ 1e-3: the half norms are negligible next to z.e
        z = 1e4 * torch.randn(N, D, generator=g)
        E = 1e-3 * torch.randn(K, D, generator=g)
    elif kind == "tiny_codes":  # the reference's own init U(+-1/K)
        z = torch.randn(N, D, generator=g)
        E = (torch.rand(K, D, generator=g) * 2 - 1) / K
    else:                       # one huge outlier code and one huge outlier token
        z = torch.randn(N, D, generator=g)
        E = torch.randn(K, D, generator=g)
        E[K // 2] *= 1e3
        z[N // 3] *= 1e3
    return z.t().contiguous().view(1, D, N), E   # [B=1, D, HW=N]


@pytest.mark.parametrize("kind", ["grid", "clustered", "scaled", "tiny_codes", "outliers"])
@pytest.mark.parametrize("D,K", [(4, 4096), (8, 1000), (16, 2048)])
def test_lowd_tensor_path_equals_fma_kernel_on_adversarial_data(kind, D, K):
    from vq_gan_b200 import ops
    z, E = _make(kind, D, K, 20000, D * 7 + K)
    i1, d1, _ = ops.search(z.cuda(), E.cuda(), 1)
    i5, d5, st = ops.search(z.cuda(), E.cuda(), 5)
    print(f"{kind} D={D} K={K}: {int(st[0])} of {i1.numel()} tokens re-searched exactly")
    assert torch.equal(i1, i5) and torch.equal(d1, d5)
    if D == 4:  # the two-engine kernel splits by images: view the same tokens as a batch of 8 "images"
        zb = z.view(D, 8, -1).permute(1, 0, 2).contiguous().cuda()          # [8, D, N/8]
        j1, e1, _ = ops.search(zb, E.cuda(), 1)
        j6, e6, st6 = ops.search(zb, E.cuda(), 6)
        assert int(st6[1]) == 6 and torch.equal(j1, j6) and torch.equal(e1, e6)


@pytest.mark.parametrize("kind", ["grid", "clustered", "scaled", "tiny_codes", "outliers"])
@pytest.mark.parametrize("D,K", [(32, 1000), (64, 4096), (256, 2048), (100, 777), (320, 1500), (512, 2048), (450, 900)])
def test_fp16_tensor_path_on_adversarial_data(kind, D, K):
    from vq_gan_b200 import ops
    z, E = _make(kind, D, K, 12000, D * 3 + K)
    zc, Ec = z.cuda(), E.cuda()
    i2, d2, _ = ops.search(zc, Ec, 2)
    i4, d4, st = ops.search(zc, Ec, 4)
    rows = z.view(D, -1).t().double()
    d64 = 0.5 * (E.double() ** 2).sum(1)[None, :] - rows @ E.double().t()
    best = d64.min(1).values
    scale = (rows.norm(dim=1) * E.double().norm(dim=1).max() + 0.5 * (E.double() ** 2).sum(1).max()).clamp_min(1e-300)
    for name, idx in (("fp32 tile", i2), ("fp16 tensor", i4)):
        chosen = d64.gather(1, idx.reshape(-1, 1).cpu()).squeeze(1)
        worst = float(((chosen - best) / scale).max())
        assert worst < 1e-6, (name, kind, worst)
    differ = (i2 != i4).reshape(-1)
    print(f"{kind} D={D} K={K}: full re-search {int(st[0])}, multi-group {int(st[2])}, differ from fp32 tile {int(differ.sum())}")
    if kind == "grid":   # exact arithmetic on the grid: both kernels must pick the lowest index of every tie
        first = (d64 == best[:, None]).double().argmax(1)
        assert torch.equal(i4.reshape(-1).cpu(), first) and torch.equal(i2.reshape(-1).cpu(), first)
