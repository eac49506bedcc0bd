"""Multi-GPU modes of the VQ bottleneck (one process per GPU, torch.distributed).

Data parallel (the reference's mode: accelerate -> DDP, train_vqgan.py:111-114,
197-209): tokens are sharded, the codebook is replicated, and the per-step
codebook statistics -- dE[K,D], the usage histogram and the squared-error sum --
are summed over ranks in ONE fp32 all-reduce (NCCL over NVLink on B200).  The
histogram rides in the same message as two exactly-summable fp32 planes
(count mod 4096, count div 4096).

Codebook sharded (extension for very large K): every rank searches its slice
of the codebook for ALL tokens, packs (distance, global index) into an ordered
int64 key and a MIN all-reduce picks the nearest code, lowest index on ties.

The collectives are backend-agnostic torch.distributed calls, so the host logic
is testable with gloo on CPU; the kernels either side are CUDA-only.
"""
from typing import Optional, Tuple

import torch
import torch.distributed as dist

_HIST_RADIX = 4096


def shard_range(total: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous near-equal split of `total` items; first `total % world` ranks get one more."""
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def pack_stats(dE: torch.Tensor, hist: Optional[torch.Tensor], scalars: Optional[torch.Tensor]
               ) -> torch.Tensor:
    """[dE | scalars | hist mod R | hist div R] as one flat fp32 buffer."""
    parts = [dE.reshape(-1).float()]
    if scalars is not None:
        parts.append(scalars.reshape(-1).float())
    if hist is not None:
        h = hist.reshape(-1).long()
        parts.append((h % _HIST_RADIX).float())
        parts.append(torch.div(h, _HIST_RADIX, rounding_mode="floor").float())
    return torch.cat(parts)


def unpack_stats(flat: torch.Tensor, dE_shape, n_scalars: int, n_hist: int):
    k = 1
    for s in dE_shape:
        k *= int(s)
    dE = flat[:k].reshape(dE_shape)
    scalars = flat[k:k + n_scalars] if n_scalars else None
    hist = None
    if n_hist:
        lo = flat[k + n_scalars:k + n_scalars + n_hist]
        hi = flat[k + n_scalars + n_hist:k + n_scalars + 2 * n_hist]
        hist = lo.round().long() + hi.round().long() * _HIST_RADIX
    return dE, scalars, hist


def allreduce_stats(dE: torch.Tensor, hist: Optional[torch.Tensor] = None,
                    scalars: Optional[torch.Tensor] = None, group=None, average_dE: bool = True):
    """Sum-all-reduce of the codebook statistics in one message.

    `average_dE=True` reproduces DDP, which AVERAGES gradients over ranks while
    every rank normalises its loss by its local n; the histogram
    and scalars are always summed.  Exact for world sizes up to 4096."""
    world = dist.get_world_size(group)
    n_s = 0 if scalars is None else scalars.numel()
    n_h = 0 if hist is None else hist.numel()
    if dE.is_cuda and dE.dtype == torch.float32:
        # one libvqb200 launch each way instead of a dozen elementwise torch kernels
        from . import ops
        flat = ops.stats_pack(dE, hist, scalars)
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        out_dE, out_scalars, out_hist = ops.stats_unpack(flat, dE.shape, n_s, n_h, 1.0 / world if average_dE else 1.0)
        return out_dE, out_hist, out_scalars
    flat = pack_stats(dE, hist, scalars)  # host-side mirror (gloo tests)
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    out_dE, out_scalars, out_hist = unpack_stats(flat, dE.shape, n_s, n_h)
    if average_dE:
        out_dE = out_dE / world
    return out_dE, out_hist, out_scalars


def reduce_argmin_keys(keys: torch.Tensor, group=None) -> torch.Tensor:
    """MIN all-reduce of packed (distance, index) keys (in place)."""
    dist.all_reduce(keys, op=dist.ReduceOp.MIN, group=group)
    return keys


def sharded_search(z: torch.Tensor, weight_shard: torch.Tensor, index_offset: int, group=None,
                   algo: int = 0) -> Tuple[torch.Tensor, torch.Tensor]:
    """Codebook-sharded nearest-code search.  `z` is the SAME token batch on every
    rank of `group`; `weight_shard` is this rank's rows [index_offset, index_offset+K_r)
    of the global codebook.  Returns (global indices, min score) on every rank."""
    from . import ops
    idx, dmin, _ = ops.search(z, weight_shard, algo)
    keys = ops.pack_argmin_keys(dmin, idx, int(index_offset))
    reduce_argmin_keys(keys, group)
    return ops.unpack_argmin_keys(keys)


def sharded_search_dp(z_local: torch.Tensor, weight_shard: torch.Tensor, index_offset: int, group=None,
                      algo: int = 0) -> Tuple[torch.Tensor, torch.Tensor]:
    """Codebook-sharded search when every rank holds DIFFERENT tokens (bulk encode,
    `preprocess_latents.py`-style loop with `VQVAE.encode_to_indices`):

      1. all-gather the latents  [B_r, D, ...] -> [R*B_r, D, ...]   (the one real exchange)
      2. local search of this rank's codebook rows for all R*B_r*HW tokens
      3. pack (score, global index) keys, MIN reduce-scatter so each rank receives the
         winners of its own tokens only
    Every rank must pass the same B_r.  Returns (global indices, min score) of the local tokens."""
    from . import ops
    world = dist.get_world_size(group)
    z_local = z_local.contiguous()
    gathered = torch.empty((world * z_local.shape[0],) + tuple(z_local.shape[1:]), dtype=z_local.dtype,
                           device=z_local.device)
    dist.all_gather_into_tensor(gathered, z_local, group=group)
    idx, dmin, _ = ops.search(gathered, weight_shard, algo)
    keys = ops.pack_argmin_keys(dmin, idx, int(index_offset))
    mine = torch.empty((z_local.shape[0],) + tuple(keys.shape[1:]), dtype=torch.int64, device=keys.device)
    dist.reduce_scatter_tensor(mine, keys, op=dist.ReduceOp.MIN, group=group)
    return ops.unpack_argmin_keys(mine)
