"""CPU oracle for the VQ bottleneck hot path.  TEST INFRASTRUCTURE ONLY.

This file restates, with stock torch CPU ops, what the reference module
`vqgan_ldm_baseline/models/quantizer.py` computes.  It is a checker: only
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / reference
arm may import it.  The product path (`vq_gan_b200`) never does and fails
loudly when the CUDA library is missing.

Pinning: the reference ships no tests or golden vectors, so
the oracle is pinned against the reference module itself, imported by file path
in the build container by `tests/golden/make_golden.py`; the outputs are
committed under `tests/golden/*.npz` and `tests/test_oracle_golden.py` checks
this restatement against them.  The extensions at the bottom (taming-style
perplexity, EMA codebook update, sharded argmin keys) have no reference code:
they are "parity unpinned" and say so.

Every function cites the reference lines it follows (paths relative to
/root/reference/).
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch

FP32_ULP = 2.0 ** -23


# --------------------------------------------------------------------------
# layout helpers
# --------------------------------------------------------------------------
def tokens_of(z: torch.Tensor) -> torch.Tensor:
    """[B,D,H,W] -> [B*H*W, D] token rows (quantizer.py:63-64: channels-last
    contiguous copy then a flat view)."""
    b, d, h, w = z.shape
    return z.permute(0, 2, 3, 1).reshape(b * h * w, d)


def image_of(rows: torch.Tensor, b: int, h: int, w: int) -> torch.Tensor:
    """[B*H*W, D] -> contiguous [B,D,H,W] (quantizer.py:83-84)."""
    return rows.reshape(b, h, w, -1).permute(0, 3, 1, 2).contiguous()


def reference_init(num_embeddings: int, embedding_dim: int) -> torch.Tensor:
    """Codebook exactly as the reference constructor draws it from the global
    torch RNG (quantizer.py:47-48): nn.Embedding's own normal_ draw happens
    first, then uniform_(-1/K, 1/K) overwrites it."""
    emb = torch.nn.Embedding(num_embeddings, embedding_dim)
    emb.weight.data.uniform_(-1.0 / num_embeddings, 1.0 / num_embeddings)
    return emb.weight.detach().clone()


# --------------------------------------------------------------------------
# search (quantizer.py:68-76)
# --------------------------------------------------------------------------
def distance_matrix(rows: torch.Tensor, codebook: torch.Tensor) -> torch.Tensor:
    """fp32 distances in the reference's evaluation order (quantizer.py:68-72):
    fl( fl(|z|^2 + |e|^2) - fl(2 * fl(z.e)) )."""
    zn = (rows * rows).sum(dim=1, keepdim=True)
    en = (codebook * codebook).sum(dim=1)
    return (zn + en) - 2 * (rows @ codebook.t())


def nearest_code(rows: torch.Tensor, codebook: torch.Tensor) -> torch.Tensor:
    """argmin over codes, lowest index on ties, NaN minimal (quantizer.py:76)."""
    return torch.argmin(distance_matrix(rows, codebook), dim=1)


def search_with_gap(rows: torch.Tensor, codebook: torch.Tensor, chunk: int = 8192
                    ) -> Dict[str, torch.Tensor]:
    """Chunked search that also returns the reference's own fp32 top-2 gap and
    the magnitude s_i = |z_i|^2 + max_k |e_k|^2 used by the parity band
    eps_i = 4 * 2^-23 * s_i.  Chunking over tokens does
    not change any per-token value (each row of quantizer.py:68-76 is
    independent)."""
    n = rows.shape[0]
    idx = torch.empty(n, dtype=torch.int64)
    gap = torch.empty(n, dtype=torch.float32)
    en_max = (codebook * codebook).sum(dim=1).max()
    for lo in range(0, n, chunk):
        d = distance_matrix(rows[lo:lo + chunk], codebook)
        if d.shape[1] >= 2:
            top2 = torch.topk(d, 2, dim=1, largest=False).values
            gap[lo:lo + chunk] = top2[:, 1] - top2[:, 0]
        else:
            gap[lo:lo + chunk] = float("inf")
        idx[lo:lo + chunk] = torch.argmin(d, dim=1)
    s = (rows * rows).sum(dim=1) + en_max
    return {"idx": idx, "gap": gap, "s": s}


def exact_distances_f64(rows: torch.Tensor, codebook: torch.Tensor) -> torch.Tensor:
    """float64 |z-e|^2, used only to classify fp32 near-ties in reports."""
    r = rows.double()
    c = codebook.double()
    return (r * r).sum(1, keepdim=True) + (c * c).sum(1) - 2 * (r @ c.t())


# --------------------------------------------------------------------------
# forward tail (quantizer.py:80-110)
# --------------------------------------------------------------------------
def forward(z: torch.Tensor, codebook: torch.Tensor, beta: float,
            idx: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
    """Whole forward.  `idx` may be forced (flat [N] int64) so the tail can be
    checked on tokens whose nearest code is a near-tie.

    Returns z_q (straight-through value, quantizer.py:98), mse (the value of
    both codebook_loss and commitment_loss, quantizer.py:89,92), vq_loss
    (quantizer.py:95) and indices [B,H,W] (quantizer.py:101)."""
    b, d, h, w = z.shape
    rows = tokens_of(z)
    if idx is None:
        idx = nearest_code(rows, codebook)
    e = image_of(codebook[idx], b, h, w)
    zc = z.contiguous()
    mse = torch.nn.functional.mse_loss(e, zc)
    vq_loss = mse + beta * mse
    z_q = zc + (e - zc)
    return {"z_q": z_q, "mse": mse, "vq_loss": vq_loss, "indices": idx.view(b, h, w)}


def backward(z: torch.Tensor, codebook: torch.Tensor, idx: torch.Tensor, beta: float,
             g_zq: Optional[torch.Tensor], g_vq: float = 1.0) -> Dict[str, torch.Tensor]:
    """Closed-form gradients that autograd derives from quantizer.py:89-98
    (SURVEY.md row a9): the straight-through path passes g_zq to z unchanged,
    codebook_loss sends (2/n)(z-e) to z, commitment_loss sends beta*(2/n)(e-z)
    to the selected codebook rows (dense embedding backward = index_add)."""
    b, d, h, w = z.shape
    n = z.numel()
    rows = tokens_of(z)
    flat = idx.reshape(-1)
    e_rows = codebook[flat]
    norm = torch.tensor(2.0 / n, dtype=torch.float32)
    gv = torch.tensor(g_vq, dtype=torch.float32)
    dz_rows = (norm * (rows - e_rows)) * gv
    dz = image_of(dz_rows, b, h, w)
    if g_zq is not None:
        dz = g_zq + dz
    contrib = (norm * (e_rows - rows)) * (gv * torch.tensor(beta, dtype=torch.float32))
    d_codebook = torch.zeros_like(codebook).index_add_(0, flat, contrib)
    return {"dz": dz, "dE": d_codebook}


def autograd_step(z: torch.Tensor, codebook: torch.Tensor, beta: float,
                  g_zq: torch.Tensor) -> Dict[str, torch.Tensor]:
    """Forward + autograd backward of `vq_loss + sum(z_q * g_zq)` built from the
    same op sequence as the reference, for pinning `backward` above."""
    zr = z.detach().clone().requires_grad_(True)
    cb = codebook.detach().clone().requires_grad_(True)
    b, d, h, w = zr.shape
    rows = tokens_of(zr)
    idx = nearest_code(rows.detach(), cb.detach())
    e = image_of(torch.nn.functional.embedding(idx, cb), b, h, w)
    l_codebook = torch.nn.functional.mse_loss(e.detach(), zr)
    l_commit = torch.nn.functional.mse_loss(e, zr.detach())
    vq_loss = l_codebook + beta * l_commit
    z_q = zr + (e - zr).detach()
    (vq_loss + (z_q * g_zq).sum()).backward()
    return {"z_q": z_q.detach(), "vq_loss": vq_loss.detach(), "mse": l_codebook.detach(),
            "indices": idx.view(b, h, w), "dz": zr.grad, "dE": cb.grad}


# --------------------------------------------------------------------------
# the two helper methods
# --------------------------------------------------------------------------
def codebook_entry(codebook: torch.Tensor, indices: torch.Tensor) -> torch.Tensor:
    """indices [B,H,W] -> [B,D,H,W] contiguous (quantizer.py:112-132)."""
    b, h, w = indices.shape
    return image_of(codebook[indices.reshape(-1)], b, h, w)


def codebook_usage(indices: torch.Tensor, num_embeddings: int) -> Tuple[torch.Tensor, float]:
    """int64 histogram with minlength K and the used fraction as a Python float
    (quantizer.py:134-149)."""
    usage = torch.bincount(indices.reshape(-1), minlength=num_embeddings)
    return usage, (usage > 0).float().mean().item()


# --------------------------------------------------------------------------
# parity band bookkeeping used by the GPU tests
# --------------------------------------------------------------------------
def parity_band(s: torch.Tensor, ulps: float = 4.0) -> torch.Tensor:
    """eps_i = ulps * 2^-23 * s_i."""
    return ulps * FP32_ULP * s


def compare_indices(got: torch.Tensor, ref: Dict[str, torch.Tensor], ulps: float = 4.0
                    ) -> Dict[str, int]:
    """Counts index disagreements and splits them by whether the reference's
    own top-2 gap lies inside the band.  `outside` must be 0."""
    got = got.reshape(-1).cpu()
    diff = got != ref["idx"]
    inside = ref["gap"] <= parity_band(ref["s"], ulps)
    return {"tokens": int(got.numel()), "mismatch": int(diff.sum()),
            "in_band_tokens": int(inside.sum()),
            "mismatch_in_band": int((diff & inside).sum()),
            "outside": int((diff & ~inside).sum())}


# --------------------------------------------------------------------------
# extensions -- PARITY UNPINNED (no reference code computes these)
# --------------------------------------------------------------------------
def perplexity(usage: torch.Tensor) -> torch.Tensor:
    """taming-transformers style exp(-sum p log(p + 1e-10)), p = usage / N."""
    p = usage.double() / usage.sum().clamp(min=1).double()
    return torch.exp(-(p * torch.log(p + 1e-10)).sum()).float()


def ema_update(codebook: torch.Tensor, cluster_size: torch.Tensor, embed_sum: torch.Tensor,
               rows: torch.Tensor, idx: torch.Tensor, decay: float, eps: float
               ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """VQ-VAE paper (van den Oord 2017, appendix A.1) EMA update with Laplace
    smoothing: N <- g N + (1-g) n ; m <- g m + (1-g) sum z ; e = m / N_smooth."""
    k = codebook.shape[0]
    counts = torch.bincount(idx.reshape(-1), minlength=k).float()
    sums = torch.zeros_like(codebook).index_add_(0, idx.reshape(-1), rows)
    new_size = cluster_size * decay + counts * (1.0 - decay)
    new_sum = embed_sum * decay + sums * (1.0 - decay)
    total = new_size.sum()
    smoothed = (new_size + eps) / (total + k * eps) * total
    return new_sum / smoothed.unsqueeze(1), new_size, new_sum


def argmin_key(dist: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """Order-preserving int64 key (monotone float bits in the high word, code
    index in the low word) whose signed minimum across codebook shards selects
    the smallest distance, lowest index on ties.  A NaN distance (a NaN code:
    ATen's argmin treats NaN as minimal, quantizer.py:76) packs as the smallest
    possible key so the lowest NaN index wins across shards."""
    nan = torch.isnan(dist)
    bits = (dist + 0.0).contiguous().view(torch.int32).to(torch.int64)  # -0.0 -> +0.0: equal scores must tie
    mono = torch.where(bits < 0, bits ^ 0x7FFFFFFF, bits)
    mono = torch.where(nan, torch.full_like(mono, -(1 << 31)), mono)    # NaN is minimal (ATen argmin)
    return (mono << 32) | idx.to(torch.int64)


def half_distance(rows: torch.Tensor, codebook: torch.Tensor) -> torch.Tensor:
    """0.5|e|^2 - z.e in float64: the argmin-equivalent score the CUDA kernels
    minimise (|z|^2 is constant per row)."""
    r = rows.double()
    c = codebook.double()
    return 0.5 * (c * c).sum(1) - r @ c.t()
