"""Seeded random shapes through every search kernel: ragged token counts, token-major inputs (HW = 1),
tiny and odd codebooks, every low dimension.  The kernels must agree with the fp32 tile kernel except on
fp32-rounding-level ties (checked in float64), and the low-D tensor path must equal the FMA kernel exactly."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _cases():
    rng = np.random.default_rng(20261018)
    out = []
    for _ in range(28):
        D = int(rng.choice([1, 2, 3, 4, 5, 6, 7, 8, 9, 12, 16]))
        K = int(rng.choice([1, 2, 17, 64, 127, 128, 129, 255, 257, 1000, 4097, 20000]))
        B = int(rng.integers(1, 6))
        HW = int(rng.choice([1, 7, 64, 129, 1000, 1024]))
        out.append((D, K, B, HW, int(rng.integers(0, 1 << 30))))
    for _ in range(14):
        D = int(rng.choice([64, 128, 192, 256]))
        K = int(rng.choice([1, 3, 100, 256, 300, 1025, 5000]))
        B = int(rng.integers(1, 5))
        HW = int(rng.choice([1, 5, 128, 200, 1024]))
        out.append((D, K, B, HW, int(rng.integers(0, 1 << 30))))
    for _ in range(6):
        D = int(rng.choice([17, 24, 33, 48, 100, 200, 255, 320]))
        out.append((D, int(rng.choice([5, 200, 1500])), int(rng.integers(1, 4)), int(rng.choice([1, 37, 256])),
                    int(rng.integers(0, 1 << 30))))
    # persistent-kernel edge cases: one more tile than SMs (second round with a lone CTA whose cluster
    # partner only pads), ragged last tile, exactly one full wave
    # 256 < D <= 512: the fp16 tensor kernel with a 80-128 KB resident token tile
    out += [(320, 1000, 2, 700, 21), (384, 513, 3, 256, 22), (512, 2000, 2, 1024, 23), (450, 300, 1, 999, 24),
            (512, 256, 149, 128, 25)]
    out += [(64, 300, 149, 128, 11), (4, 300, 149, 128, 12), (8, 257, 149, 128, 13), (256, 256, 297, 64, 14),
            (128, 130, 148, 128, 15), (16, 1000, 2, 9500, 16), (32, 700, 3, 6400, 17)]
    return out


@pytest.mark.parametrize("D,K,B,HW,seed", _cases())
def test_random_shape_all_kernels_agree(D, K, B, HW, seed):
    from vq_gan_b200 import ops
    g = torch.Generator().manual_seed(seed)
    z = torch.randn(B, D, HW, generator=g)
    E = torch.randn(K, D, generator=g)
    if seed % 3 == 0 and K > 1:           # duplicate codes: exact ties must resolve to the lowest index
        E[K - 1] = E[0]
    zc, Ec = z.cuda(), E.cuda()
    ref_idx, ref_d, _ = ops.search(zc, Ec, 2)
    rows = z.permute(0, 2, 1).reshape(-1, D).double()
    d64 = 0.5 * (E.double() ** 2).sum(1)[None, :] - rows @ E.double().t()
    # the fp32 tile kernel itself against float64: only rounding-level ties may differ
    best64 = d64.min(1).values
    got64 = d64.gather(1, ref_idx.reshape(-1, 1).cpu()).squeeze(1)
    scale = (rows.norm(dim=1) * E.double().norm(dim=1).max() + 0.5 * (E.double() ** 2).sum(1).max()).clamp_min(1e-30)
    assert float(((got64 - best64) / scale).max()) < 1e-6
    algos = [0]
    if D <= 16:
        algos += [1, 5]
    if 3 <= D <= 16 and B >= 2:
        algos.append(6)
    if 16 < D <= 512:
        algos.append(4)
    if D % 64 == 0 and D <= 256:
        algos.append(3)
    results = {}
    for a in algos:
        idx, dmin, st = ops.search(zc, Ec, a)
        results[a] = (idx, dmin)
        assert idx.shape == (B, HW) and int(idx.min()) >= 0 and int(idx.max()) < K
        chosen = d64.gather(1, idx.reshape(-1, 1).cpu()).squeeze(1)
        assert float(((chosen - best64) / scale).max()) < 1e-6, (a, float(((chosen - best64) / scale).max()))
        differ = (idx != ref_idx).reshape(-1).cpu()
        if int(differ.sum()):      # a different index is only acceptable as a float32-level tie
            assert float(((chosen - got64).abs() / scale)[differ].max()) < 1e-6
    if 1 in results:
        assert torch.equal(results[1][0], results[5][0]) and torch.equal(results[1][1], results[5][1])
    if 6 in results:   # both engines in one CTA: bit-identical to the CUDA-core kernel
        assert torch.equal(results[1][0], results[6][0]) and torch.equal(results[1][1], results[6][1])
    if 3 in results:
        assert torch.equal(results[3][0], results[4][0])
    if seed % 3 == 0 and K > 1:
        for a, (idx, _) in results.items():
            assert int((idx == K - 1).sum()) == 0, f"algo {a} picked the duplicate with the higher index"
