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


def pin_rank_to_gpu_cores(local_rank: int, ranks_on_node: int, device_index: Optional[int] = None) -> Optional[list]:
    """Pins this process to its own slice of the CPU cores NVML reports as local to its GPU (same NUMA
    node / PCIe root).  One process per GPU on an 8-GPU box otherwise lets the scheduler stack several
    ranks on the same cores and allocate pinned staging buffers on the remote socket; the host side of a
    step (kernel launches, the per-forward loss read-back of quantizer.py:106-107, H2D of the next batch)
    then dominates the end-to-end time (SCALE_r01: device-timed efficiency 0.97, end to end 0.65 at N=4).
    Returns the core list, or None when NVML / sched_setaffinity is unavailable (nothing is changed then)."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        ncpu = os.cpu_count() or 1
        words = (ncpu + 63) // 64
        handle = pynvml.nvmlDeviceGetHandleByIndex(local_rank if device_index is None else device_index)
        mask = pynvml.nvmlDeviceGetCpuAffinity(handle, words)
        local = [64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1 and 64 * w + b < ncpu]
        allowed = sorted(set(local) & set(os.sched_getaffinity(0))) or sorted(os.sched_getaffinity(0))
        # GPUs that share this core set split it evenly, in local-rank order
        peers = []
        for r in range(ranks_on_node):
            try:
                m = pynvml.nvmlDeviceGetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(r), words)
            except Exception:
                continue
            if list(m) == list(mask):
                peers.append(r)
        if local_rank not in peers:
            peers = [local_rank]
        per = max(len(allowed) // len(peers), 1)
        k = peers.index(local_rank)
        mine = allowed[k * per:(k + 1) * per] or allowed
        os.sched_setaffinity(0, mine)
        return mine
    except Exception:
        return None


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


def _gloo(group) -> bool:
    return dist.get_backend(group) == "gloo"


def _all_gather_rows(out: torch.Tensor, inp: torch.Tensor, group, async_op: bool):
    """all_gather_into_tensor; on gloo (CPU tests, two ranks sharing one GPU) CUDA tensors go through the
    list form, which that backend implements for both device types."""
    if _gloo(group) and inp.is_cuda:
        dist.all_gather(list(out.chunk(dist.get_world_size(group), dim=0)), inp, group=group)
        return None
    return dist.all_gather_into_tensor(out, inp, group=group, async_op=async_op)


def _reduce_scatter_min(out: torch.Tensor, keys: torch.Tensor, group, async_op: bool):
    """MIN reduce-scatter of packed keys; gloo has no CUDA reduce-scatter: MIN all-reduce + own slice."""
    if _gloo(group) and keys.is_cuda:
        dist.all_reduce(keys, op=dist.ReduceOp.MIN, group=group)
        n = out.shape[0]
        out.copy_(keys[dist.get_rank(group) * n:(dist.get_rank(group) + 1) * n])
        return None
    return dist.reduce_scatter_tensor(out, keys, op=dist.ReduceOp.MIN, group=group, async_op=async_op)


def sharded_search_dp(z_local: torch.Tensor, weight_shard: torch.Tensor, index_offset: int, group=None,
                      algo: int = 0, chunks: int = 0) -> Tuple[torch.Tensor, torch.Tensor]:
    """Codebook-sharded search when every rank holds DIFFERENT tokens (bulk encode,
    `preprocess_latents.py`-style loop with `VQVAE.encode_to_indices`):

      1. all-gather the latents  [B_r, D, ...] -> [R*B_r, D, ...]   (the one real exchange)
      2. local search of this rank's codebook rows for all R*B_r*HW tokens
      3. pack (score, global index) keys, MIN reduce-scatter so each rank receives the
         winners of its own tokens only

    `chunks` > 1 cuts the batch into slices of images and pipelines them inside the call (every slice's
    all-gather is issued up front; the search of slice c starts when ITS gather has landed; each slice's key
    reduce-scatter overlaps the next search).  Measured at C5 on 8 B200 this is SLOWER than one slice (3.03 vs
    2.64 ms per step): every search call has a fixed cost (five launches and a latency-bound exact re-search
    of a few dozen tokens), so the default is 1 and the exchange is hidden ACROSS steps instead, by
    `ShardedEncoder` below.  The codebook shard is packed once per call.  Every rank must pass the same B_r.
    Returns (global indices, min score) of the local tokens."""
    from . import ops
    world = dist.get_world_size(group)
    z_local = z_local.contiguous()
    Bl = int(z_local.shape[0])
    tok_per_img = 1
    for d in z_local.shape[2:]:
        tok_per_img *= int(d)
    chunks = max(1, min(chunks if chunks > 0 else 1, Bl))
    bounds = [shard_range(Bl, chunks, c) for c in range(chunks)]
    pack = ops.prepare_codebook(weight_shard) if z_local.is_cuda else None
    rest = tuple(z_local.shape[1:])
    gathered, gworks = [], []
    for lo, hi in bounds:
        buf = torch.empty((world * (hi - lo),) + rest, dtype=z_local.dtype, device=z_local.device)
        gworks.append(_all_gather_rows(buf, z_local[lo:hi], group, async_op=True))
        gathered.append(buf)
    outs, rworks = [], []
    for c, (lo, hi) in enumerate(bounds):
        if gworks[c] is not None:
            gworks[c].wait()  # the compute stream waits for THIS slice only
        idx, dmin, _ = ops.search(gathered[c], weight_shard, algo, pack)
        keys = ops.pack_argmin_keys(dmin, idx, int(index_offset))
        mine = torch.empty((hi - lo,) + tuple(keys.shape[1:]), dtype=torch.int64, device=keys.device)
        rworks.append(_reduce_scatter_min(mine, keys, group, async_op=True))
        outs.append(mine)
    for w in rworks:
        if w is not None:
            w.wait()
    mine = outs[0] if chunks == 1 else torch.cat(outs, dim=0)
    return ops.unpack_argmin_keys(mine)


class ShardedEncoder:
    """Codebook-sharded bulk encode with the exchange hidden across steps (the `preprocess_latents.py`-style loop
    of config C5: one `VQVAE.encode_to_indices` batch after another against a FIXED codebook).

        enc = ShardedEncoder(weight_shard, index_offset)
        enc.submit(z0)                      # all-gather of batch 0 starts on the communicator's stream
        for z_next in batches[1:]:
            enc.submit(z_next)              # all-gather of batch i+1 ...
            idx, dmin = enc.collect()       # ... overlaps the search of batch i
        idx, dmin = enc.collect()

    `collect()` returns the winners of the OLDEST submitted batch: it waits for that batch's gather only,
    searches this rank's codebook rows for all R*B_r*HW tokens, and MIN-reduce-scatters the packed
    (score, global index) keys so each rank receives its own tokens' winners.  The codebook shard is packed
    once, in the constructor (call `refresh()` if the weights change).  Every rank must submit the same
    batch shapes in the same order."""

    def __init__(self, weight_shard: torch.Tensor, index_offset: int, group=None, algo: int = 0):
        from . import ops
        self._ops = ops
        self.weight = weight_shard.contiguous()
        self.index_offset = int(index_offset)
        self.group = group
        self.algo = algo
        self.world = dist.get_world_size(group)
        self._pending = []
        self.refresh()

    def refresh(self) -> None:
        self.pack = self._ops.prepare_codebook(self.weight) if self.weight.is_cuda else None

    def submit(self, z_local: torch.Tensor) -> None:
        z_local = z_local.contiguous()
        buf = torch.empty((self.world * z_local.shape[0],) + tuple(z_local.shape[1:]), dtype=z_local.dtype,
                          device=z_local.device)
        work = _all_gather_rows(buf, z_local, self.group, async_op=True)
        self._pending.append((buf, work, int(z_local.shape[0]), z_local))  # z_local stays alive until the gather ran

    def collect(self) -> Tuple[torch.Tensor, torch.Tensor]:
        if not self._pending:
            raise RuntimeError("ShardedEncoder.collect() without a submitted batch")
        buf, work, b_local, _ = self._pending.pop(0)
        if work is not None:
            work.wait()
        idx, dmin, _ = self._ops.search(buf, self.weight, self.algo, self.pack)
        keys = self._ops.pack_argmin_keys(dmin, idx, self.index_offset)
        mine = torch.empty((b_local,) + tuple(keys.shape[1:]), dtype=torch.int64, device=keys.device)
        _reduce_scatter_min(mine, keys, self.group, async_op=False)
        return self._ops.unpack_argmin_keys(mine)

    def encode_all(self, batches):
        """Generator over (indices, min score) for an iterable of local batches, one batch of look-ahead."""
        it = iter(batches)
        try:
            self.submit(next(it))
        except StopIteration:
            return
        for z in it:
            self.submit(z)
            yield self.collect()
        yield self.collect()
