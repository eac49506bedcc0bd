"""Drop-in `VectorQuantizer` running on libvqb200 (sm_100a).

Mirrors the reference class in vqgan_ldm_baseline/models/quantizer.py:17-149:
same constructor, attributes, `forward` outputs `(z_q, loss_dict, indices)`,
`get_codebook_entry`, `get_codebook_usage`, and a `state_dict()` holding exactly
`embedding.weight`.  Inputs must be CUDA float32 tensors; there is no CPU path.
"""
from typing import Dict, Tuple

import torch
import torch.nn as nn

from . import ops
from ._cabi import ALGO_AUTO, ALGO_NAMES


_LOGGED = ("codebook_loss", "commitment_loss")


class LossDict(dict):
    """The `loss_dict` of quantizer.py:104-108 -- `{'vq_loss': Tensor, 'codebook_loss': float, 'commitment_loss':
    float}`, a real `dict` -- whose two logged floats are read from the device WHEN THEY ARE FIRST LOOKED AT instead
    of inside `forward`.  The reference calls `.item()` twice per forward (quantizer.py:106-107), which stalls the
    host until the search has finished and leaves the GPU idle while Python launches the backward pass; a training
    loop that only logs every N steps (or never reads these two keys, like train_vqgan.py:303-315, which shows
    `vq_loss` and the usage ratio) then never pays that stall.  Every read path -- `d[k]`, `get`, `items`, `values`,
    `copy`, `{**d}`, `dict(d)`, `repr`, `==`, pickling -- sees plain Python floats."""

    __slots__ = ("_pending",)

    def __init__(self, vq_loss, mse_device):
        super().__init__(vq_loss=vq_loss, codebook_loss=None, commitment_loss=None)
        self._pending = mse_device  # 0-dim device tensor, or None once the floats are in place

    def _materialize(self):
        p = self._pending
        if p is not None:
            self._pending = None
            v = p.item()  # the one device->host sync, at first use
            for k in _LOGGED:
                if dict.__getitem__(self, k) is None:
                    dict.__setitem__(self, k, v)

    def __getitem__(self, key):
        if key in _LOGGED:
            self._materialize()
        return dict.__getitem__(self, key)

    def __setitem__(self, key, value):
        if key in _LOGGED and self._pending is not None:
            self._materialize()
        dict.__setitem__(self, key, value)

    def __iter__(self):  # a Python-level __iter__ also keeps dict.update / dict(d) / {**d} off CPython's raw-storage fast path
        return dict.__iter__(self)

    def get(self, key, default=None):
        return self[key] if key in self else default

    def pop(self, key, *default):
        self._materialize()
        return dict.pop(self, key, *default)

    def popitem(self):
        self._materialize()
        return dict.popitem(self)

    def setdefault(self, key, default=None):
        self._materialize()
        return dict.setdefault(self, key, default)

    def items(self):
        self._materialize()
        return dict.items(self)

    def values(self):
        self._materialize()
        return dict.values(self)

    def copy(self):
        self._materialize()
        return dict(dict.items(self))

    def __eq__(self, other):
        self._materialize()
        return dict.__eq__(self, other)

    def __ne__(self, other):
        return not self.__eq__(other)

    __hash__ = None

    def __repr__(self):
        self._materialize()
        return dict.__repr__(self)

    def __reduce__(self):
        self._materialize()
        return (dict, (dict(dict.items(self)),))


class VectorQuantizer(nn.Module):
    """Nearest-codebook bottleneck (reference: quantizer.py:17-149).

    Args (quantizer.py:33-38):
        num_embeddings: codebook size K
        embedding_dim: code dimension D
        commitment_cost: beta, weight of the commitment term
    Keyword-only extensions (defaults reproduce the reference exactly):
        return_format: "reference" -> (z_q, loss_dict, indices) as quantizer.py:110;
            "taming" -> (z_q, vq_loss, (perplexity, None, indices)) -- values parity-unpinned
        lazy_stats: keep the two logged losses as 0-dim device tensors instead of
            Python floats (never syncs; the default returns a `LossDict` whose floats sync on first access)
        algo: search kernel override (0 auto, 1 low-D FMA, 2 fp32 tile, 3 tcgen05 bf16x3,
            4 tcgen05 single fp16 pass + exact fp32 re-score, 5 tcgen05 tf32x3 for D <= 16,
            6 CUDA-core + tf32x3 tensor roles in one CTA for D == 4)
    """

    def __init__(self, num_embeddings: int, embedding_dim: int, commitment_cost: float = 0.25, *,
                 return_format: str = "reference", lazy_stats: bool = False, algo: int = ALGO_AUTO):
        super().__init__()
        if return_format not in ("reference", "taming"):
            raise ValueError(f"unknown return_format {return_format!r}")
        if algo not in ALGO_NAMES:
            raise ValueError(f"unknown algo {algo}")
        self.num_embeddings = num_embeddings
        self.embedding_dim = embedding_dim
        self.commitment_cost = commitment_cost
        self.return_format = return_format
        self.lazy_stats = lazy_stats
        self.algo = algo
        # same two RNG draws, in the same order, as quantizer.py:47-48, so that
        # torch.manual_seed(42) (train_vqgan.py:117) yields the same codebook
        self.embedding = nn.Embedding(num_embeddings, embedding_dim)
        self.embedding.weight.data.uniform_(-1.0 / num_embeddings, 1.0 / num_embeddings)
        self.last_search_stats = None  # int64[4] device tensor of the last forward
        self.last_mse = None           # 0-dim device tensor: codebook_loss of the last forward

    # -- forward (quantizer.py:50-110) ---------------------------------------
    def forward(self, z: torch.Tensor):
        if z.dim() != 4:
            raise RuntimeError(f"expected z of shape [B, C, H, W], got {tuple(z.shape)}")
        if z.shape[1] != self.embedding_dim:
            raise RuntimeError(
                f"z has {z.shape[1]} channels but embedding_dim is {self.embedding_dim}")
        if z.dtype != torch.float32:
            z = z.float()
        pack = split_ws = None
        pre = getattr(z, "_vqb_presplit", None)
        if pre is not None:
            # z comes from a QuantConv1x1 that feeds this quantizer: the convolution already wrote the token split of the
            # fp16 tensor search into a workspace (valid only for exactly this tensor and this codebook version)
            w = self.embedding.weight
            if pre[2] == (w.data_ptr(), w._version, tuple(z.shape)) and z.is_contiguous() and self.algo in (0, 4):
                split_ws, pack = pre[0], pre[1]
        z_q, vq_loss, mse, indices, stats = ops.quantize(z, self.embedding.weight,
                                                         float(self.commitment_cost), self.algo, pack, split_ws)
        self.last_search_stats = stats
        self.last_mse = mse.detach()  # device copy of the logged loss (multi-GPU statistics need no host round trip)
        if self.return_format == "taming":
            usage, _, _ = ops.codebook_usage(indices, self.num_embeddings)
            p = usage.double() / max(indices.numel(), 1)
            perplexity = torch.exp(-(p * torch.log(p + 1e-10)).sum()).float()
            return z_q, vq_loss, (perplexity, None, indices)
        if self.lazy_stats:
            m = mse.detach()
            loss_dict = {"vq_loss": vq_loss, "codebook_loss": m, "commitment_loss": m}
        else:
            # Python floats like quantizer.py:106-107, read from the device (ONE sync for both) on first access
            loss_dict = LossDict(vq_loss, mse.detach())
        return z_q, loss_dict, indices

    # -- quantizer.py:112-132 --------------------------------------------------
    def get_codebook_entry(self, indices: torch.Tensor, strict: bool = True) -> torch.Tensor:
        if indices.dim() != 3:
            raise RuntimeError(f"expected indices of shape [B, H, W], got {tuple(indices.shape)}")
        out, err = ops.codebook_entry(self.embedding.weight.detach(), indices.long())
        if strict and int(err.item()) != 0:
            raise IndexError("index out of range in get_codebook_entry")
        return out

    # -- quantizer.py:134-149 --------------------------------------------------
    def get_codebook_usage(self, indices: torch.Tensor) -> Tuple[torch.Tensor, float]:
        usage, used, err = ops.codebook_usage(indices.long(), self.num_embeddings)
        host = torch.stack([used[0], err[0].long()]).cpu()  # one sync
        if int(host[1]) != 0:
            raise RuntimeError("get_codebook_usage: index outside [0, num_embeddings)")
        # the reference computes float32 mean of (usage > 0): count / K rounded to fp32
        ratio = (torch.tensor(float(host[0]), dtype=torch.float32) / self.num_embeddings).item()
        return usage, ratio

    # -- bulk encode (VQVAE.encode_to_indices, vq_vae.py:162-175) -------------
    @torch.no_grad()
    def encode_indices(self, z: torch.Tensor) -> torch.Tensor:
        """Search only: indices [B, H, W] without z_q / loss."""
        idx, _, stats = ops.search(z.float(), self.embedding.weight.detach(), self.algo)
        self.last_search_stats = stats
        return idx

    def extra_repr(self) -> str:
        return (f"num_embeddings={self.num_embeddings}, embedding_dim={self.embedding_dim}, "
                f"commitment_cost={self.commitment_cost}, algo={ALGO_NAMES[self.algo]}")


def usage_ratio_like_reference(usage: torch.Tensor) -> float:
    """(usage > 0).float().mean().item() -- quantizer.py:147, for host-side checks."""
    return (usage > 0).float().mean().item()


class QuantConv1x1(nn.Conv2d):
    """`pre_quant_conv` / `post_quant_conv` of the reference VQVAE (`nn.Conv2d(cin, cout, kernel_size=1)`,
    vq_vae.py:74-79) on libvqb200: same constructor arguments, parameters and state_dict keys
    (`weight [cout, cin, 1, 1]`, `bias [cout]`), so `VQVAE.pre_quant_conv = QuantConv1x1(zc, D)` loads
    reference checkpoints strictly.  Forward and the input gradient run on the hand-written kernel
    (3xTF32 tcgen05 when cin % 32 == 0, cout % 16 == 0, cout <= 256; CUDA cores otherwise); the weight
    gradient is a library GEMM."""

    def __init__(self, in_channels: int, out_channels: int, kernel_size=1, stride=1, padding=0, bias: bool = True, *,
                 algo: int = 0):
        # same positional order as nn.Conv2d, so the reference's calls `nn.Conv2d(z_channels, embedding_dim,
        # kernel_size=1)` (vq_vae.py:75-76) and `nn.Conv2d(cin, cout, 1)` become a rename
        def _one(v, want):
            return all(int(x) == want for x in (v if isinstance(v, (tuple, list)) else (v,)))
        if not (_one(kernel_size, 1) and _one(stride, 1) and _one(padding, 0)):
            raise ValueError(f"QuantConv1x1 is a 1x1 convolution: kernel_size=1, stride=1, padding=0 "
                             f"(got {kernel_size}, {stride}, {padding})")
        super().__init__(in_channels, out_channels, kernel_size=1, bias=bool(bias))
        self.algo = algo
        self._feeds = None

    def feed(self, quantizer: "VectorQuantizer") -> "QuantConv1x1":
        """Declares that this layer is the `pre_quant_conv` in front of `quantizer` (vq_vae.py:115 -> :118).  When the
        shapes qualify (`ops.split_eligible`) the forward then runs `vqb_conv1x1_split_f32`: the convolution's epilogue
        also emits the fp16 token rows / scales / norms the quantizer's tensor search needs, straight from the TMEM
        accumulator, and the search skips its own pass over z.  Results are unchanged."""
        import weakref
        self._feeds = weakref.ref(quantizer)
        return self

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if x.dim() != 4:
            raise RuntimeError(f"expected x of shape [B, C, H, W], got {tuple(x.shape)}")
        if x.dtype != torch.float32:
            x = x.float()
        vq = self._feeds() if self._feeds is not None else None
        if vq is not None and x.is_cuda and self.algo in (0, 1) and vq.algo in (0, 4) and vq.embedding_dim == self.out_channels:
            tokens = x.numel() // max(int(x.shape[1]), 1)
            if ops.split_eligible(self.in_channels, self.out_channels, tokens, vq.num_embeddings):
                w = vq.embedding.weight
                pack = ops.prepare_codebook(w.detach())
                y, sws = ops.conv1x1_split(x, self.weight, self.bias, w.detach(), pack)
                y._vqb_presplit = (sws, pack, (w.data_ptr(), w._version, tuple(y.shape)))
                return y
        return ops.conv1x1(x, self.weight, self.bias, self.algo)


class GroupNormSiLU(nn.GroupNorm):
    """`h = norm_out(h); h = F.silu(h)` of the reference Encoder / Decoder (encoder_decoder.py:166-167, 249-250) as ONE
    kernel (next row N2): same constructor, parameters and state_dict keys as the `nn.GroupNorm(32, C, eps=1e-6)` it
    replaces, but `forward` returns the ACTIVATED tensor, so the `F.silu` line that follows it in the reference
    `forward` must go.  Forward and backward run on `vqb_groupnorm_silu[_backward]_f32`; the
    3x3 convolution that follows stays a cuDNN call."""

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if not self.affine:
            raise RuntimeError("GroupNormSiLU needs affine=True (the reference's norm_out is affine)")
        if x.dtype != torch.float32:
            x = x.float()
        y, _, _ = ops.groupnorm_silu(x, self.weight, self.bias, self.num_groups, float(self.eps))
        return y


def encoder_tail(encoder: nn.Module, h: torch.Tensor) -> torch.Tensor:
    """The last three lines of the reference `Encoder.forward` / `Decoder.forward` (encoder_decoder.py:166-168,
    249-251 before the sigmoid): `conv_out(silu(norm_out(h)))` with the normalisation + activation fused.  Uses the
    module's own `norm_out` / `conv_out` parameters (any `nn.GroupNorm` / `nn.Conv2d`)."""
    gn = encoder.norm_out
    y, _, _ = ops.groupnorm_silu(h.float(), gn.weight, gn.bias, gn.num_groups, float(gn.eps))
    return encoder.conv_out(y)


class EMAVectorQuantizer(VectorQuantizer):
    """Extension (no reference code; semantics of the VQ-VAE paper, appendix A.1 -- PARITY UNPINNED):
    the codebook is updated by exponential moving averages of the per-code token counts and sums instead
    of by its gradient.  Forward outputs keep the reference contract; `vq_loss` is the commitment term
    beta * mse(z, sg(e)) only and `embedding.weight` receives no gradient.  In training mode every forward
    runs `vqb_code_sums_f32` + `vqb_ema_update_f32`; under torch.distributed the counts and sums are
    summed over the ranks first (one all-reduce), so every replica applies the same update.

    `cluster_size` / `embed_sum` are persistent buffers (extra state_dict keys: load reference checkpoints
    with strict=False)."""

    def __init__(self, num_embeddings: int, embedding_dim: int, commitment_cost: float = 0.25, *,
                 decay: float = 0.99, eps: float = 1e-5, **kw):
        super().__init__(num_embeddings, embedding_dim, commitment_cost, **kw)
        self.decay, self.eps = decay, eps
        self.embedding.weight.requires_grad_(False)
        self.register_buffer("cluster_size", torch.zeros(num_embeddings))
        self.register_buffer("embed_sum", self.embedding.weight.detach().clone())

    def forward(self, z: torch.Tensor):
        if z.dtype != torch.float32:
            z = z.float()
        # training mode updates the codebook in place below: the backward pass must see the codebook this
        # forward searched, so it gets a snapshot (K*D*4 bytes, negligible next to the latents)
        w = self.embedding.weight.detach()
        if self.training:
            w = w.clone()
        z_q, vq_loss, mse, indices, stats = ops.quantize(z, w, float(self.commitment_cost), self.algo)
        self.last_search_stats = stats
        # ops.quantize returns vq_loss = (1 + beta) * mse whose gradient w.r.t. z is (2/n)(z - e) (weight 1, the
        # reference's "codebook_loss" branch).  The EMA variant keeps only beta * mse(z, sg(e)):
        #   value    beta * (1 + beta) * mse - beta^2 * mse = beta * mse
        #   gradient beta * (2/n)(z - e)
        beta = float(self.commitment_cost)
        commit = beta * vq_loss - (beta * beta) * mse.detach()
        if self.training:
            with torch.no_grad():
                counts, sums = ops.code_sums(z.detach(), indices, self.num_embeddings)
                if torch.distributed.is_available() and torch.distributed.is_initialized() \
                        and torch.distributed.get_world_size() > 1:
                    # ONE message [sums | counts] through the packed-statistics kernels of the data-parallel path
                    flat = ops.stats_pack(sums, None, counts)
                    torch.distributed.all_reduce(flat)
                    sums, counts, _ = ops.stats_unpack(flat, sums.shape, self.num_embeddings, 0, 1.0)
                ops.ema_update(self.embedding.weight.data, self.cluster_size, self.embed_sum, counts.contiguous(),
                               sums.contiguous(), float(self.decay), float(self.eps))
        if self.lazy_stats:
            m = mse.detach()
            return z_q, {"vq_loss": commit, "codebook_loss": m, "commitment_loss": m}, indices
        return z_q, LossDict(commit, mse.detach()), indices
