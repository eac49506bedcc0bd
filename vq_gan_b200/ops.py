"""torch.library custom ops over the C ABI (`vqb200::*`).

PyTorch is the tensor / stream / autograd carrier only: every op validates its
arguments, allocates outputs with torch, and enqueues libvqb200 kernels on the
current CUDA stream.  CPU tensors are rejected -- there is no fallback path.
"""
import contextlib
import ctypes
from typing import Optional, Tuple

import torch
from torch import Tensor

from . import _cabi
from ._cabi import check, lib


# measurement hooks used by bench.py: LAUNCHES counts libvqb200 kernel launches by entry
# point; when PROFILE is a list, (start, stop) CUDA events bracketing the search launch on
# the current stream are appended to it; PROFILE_TAIL / PROFILE_BWD do the same for the
# forward-tail and backward launches.
LAUNCHES = {"total": 0}
PROFILE = None
PROFILE_TAIL = None
PROFILE_BWD = None
_KERNELS_PER_CALL = {"prepare": 3, "search": 2, "tail": 2, "backward": 1, "gather": 1, "hist": 2,
                     "code_sums": 1, "ema": 2, "keys": 1, "conv": 2}


def _count(kind: str, kernels: int = 0) -> None:
    LAUNCHES["total"] += kernels or _KERNELS_PER_CALL[kind]
    LAUNCHES[kind] = LAUNCHES.get(kind, 0) + 1


def _search_kernels(D: int, algo: int, N: int = 0, K: int = 0, B: int = 0) -> int:
    """libvqb200 kernels one vqb_search_f32 call launches (mirrors resolve_algo in vqb_api.cu)."""
    if algo == _cabi.ALGO_AUTO:
        algo = (_cabi.ALGO_DUAL_LOWD if (3 <= D <= 8 and B >= 16 and N * K >= 1 << 30) else
                (_cabi.ALGO_TCGEN05_TF32X3 if (D >= 5 and N * K >= 1 << 28) else _cabi.ALGO_LOWD_FMA) if D <= 16 else
                _cabi.ALGO_TCGEN05_F16 if (16 < D <= 512 and (N == 0 or N * K * D >= 1 << 29))
                else _cabi.ALGO_FP32_TILE)
    # lowd: search + stats; fp32: search (+ finalize) + stats; tcgen05: split, mma, re-score, finalize, stats
    # tf32x3: split, mma, chunk re-score, list search, stats
    return {_cabi.ALGO_LOWD_FMA: 2, _cabi.ALGO_FP32_TILE: 3, _cabi.ALGO_TCGEN05: 5,
            _cabi.ALGO_TCGEN05_F16: 6, _cabi.ALGO_TCGEN05_TF32X3: 5, _cabi.ALGO_DUAL_LOWD: 5}[algo]


def _p(t: Optional[Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _need_cuda_f32(t: Tensor, name: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"vq_gan_b200: `{name}` must be a CUDA tensor (got {t.device}); "
                           "this package has no CPU path")
    if t.dtype != torch.float32:
        raise RuntimeError(f"vq_gan_b200: `{name}` must be float32, got {t.dtype}")


def _shape_bdhw(z: Tensor, weight: Tensor) -> Tuple[int, int, int, int]:
    if z.dim() < 2:
        raise RuntimeError(f"latent must be [B, D, ...], got {tuple(z.shape)}")
    B, D = int(z.shape[0]), int(z.shape[1])
    if weight.dim() != 2 or int(weight.shape[1]) != D:
        # the reference fails in `z.view(-1, embedding_dim)` / matmul for the same input
        raise RuntimeError(f"latent channels ({D}) do not match the codebook {tuple(weight.shape)}")
    HW = 1
    for s in z.shape[2:]:
        HW *= int(s)
    return B, D, HW, int(weight.shape[0])


def _on(device):
    """Device guard that costs nothing when `device` is already current (the common case)."""
    if device.index is None or device.index == torch.cuda.current_device():
        return contextlib.nullcontext()
    return torch.cuda.device(device)


def _bytes(n: int, device) -> Tensor:
    return torch.empty(max(int(n), 1), dtype=torch.uint8, device=device)


def _prepare(weight: Tensor) -> Tensor:
    K, D = int(weight.shape[0]), int(weight.shape[1])
    nbytes = lib().vqb_codebook_pack_bytes(K, D)
    pack = _bytes(nbytes, weight.device)
    check(lib().vqb_codebook_prepare_f32(_p(weight), K, D, _p(pack), nbytes, _stream()),
          "vqb_codebook_prepare_f32")
    _count("prepare")
    return pack


def prepare_codebook(weight: Tensor) -> Tensor:
    """The packed codebook (`vqb_codebook_prepare_f32`) as an opaque uint8 tensor, for callers that search
    the SAME codebook several times (bulk encode: `search(z, weight, algo, pack)`).  It must be rebuilt
    whenever `weight` changes; `search` re-checks shape and device only."""
    _need_cuda_f32(weight, "weight")
    with _on(weight.device):
        return _prepare(weight.contiguous())


def _search_into(z: Tensor, weight: Tensor, algo: int, want_dmin: bool, pack: Optional[Tensor] = None,
                 split_ws: Optional[Tensor] = None):
    B, D, HW, K = _shape_bdhw(z, weight)
    dev = z.device
    if pack is None:
        pack = _prepare(weight)
    elif pack.device != dev or pack.dtype != torch.uint8 or pack.numel() < lib().vqb_codebook_pack_bytes(K, D):
        raise RuntimeError("search: `pack` does not belong to this codebook (wrong device, dtype or size)")
    idx = torch.empty((B,) + tuple(z.shape[2:]), dtype=torch.int64, device=dev)
    dmin = torch.empty(idx.shape, dtype=torch.float32, device=dev) if want_dmin else None
    stats = torch.empty(4, dtype=torch.int64, device=dev)  # every search path writes all four entries
    if split_ws is not None:
        # the producer of z (conv1x1_split) already wrote the token split into this workspace
        if pack is None:
            raise RuntimeError("search: a pre-split workspace needs the codebook pack it was built against")
        algo = _cabi.ALGO_TCGEN05_F16 | _cabi.SEARCH_PRESPLIT
        ws_bytes = lib().vqb_search_workspace_bytes(B, D, HW, K, _cabi.ALGO_TCGEN05_F16)
        if split_ws.device != dev or split_ws.dtype != torch.uint8 or split_ws.numel() < ws_bytes:
            raise RuntimeError("search: the pre-split workspace does not fit this problem")
        ws = split_ws
    else:
        ws_bytes = lib().vqb_search_workspace_bytes(B, D, HW, K, algo)
        ws = _bytes(ws_bytes, dev)
    prof = PROFILE
    if prof is not None:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
    check(lib().vqb_search_f32(_p(z), B, D, HW, _p(weight), K, _p(pack), _p(idx), _p(dmin), _p(ws),
                               ws_bytes, algo, _p(stats), _stream()), "vqb_search_f32")
    _count("search", _search_kernels(D, algo & 0xff, B * HW, K, B) - (1 if split_ws is not None else 0))
    if prof is not None:
        ev1.record()
        prof.append((ev0, ev1))
    return idx, dmin, stats


# ---------------------------------------------------------------------------
# search only (encode_to_indices-style bulk use; quantizer.py:68-76)
# ---------------------------------------------------------------------------
@torch.library.custom_op("vqb200::search", mutates_args=())
def search(z: Tensor, weight: Tensor, algo: int = 0, pack: Optional[Tensor] = None) -> Tuple[Tensor, Tensor, Tensor]:
    """(indices[B,*spatial] int64, dmin[B,*spatial] f32, stats int64[4]).  `pack`: optional result of
    `prepare_codebook(weight)` to skip the per-call codebook pre-pass."""
    _need_cuda_f32(z, "z")
    _need_cuda_f32(weight, "weight")
    z = z.contiguous()
    weight = weight.contiguous()
    if z.numel() == 0:
        shape = (z.shape[0],) + tuple(z.shape[2:])
        _shape_bdhw(z, weight)
        return (torch.empty(shape, dtype=torch.int64, device=z.device),
                torch.empty(shape, dtype=torch.float32, device=z.device),
                torch.zeros(4, dtype=torch.int64, device=z.device))
    with _on(z.device):
        idx, dmin, stats = _search_into(z, weight, algo, True, pack)
    return idx, dmin, stats


@search.register_fake
def _(z, weight, algo=0, pack=None):
    shape = (z.shape[0],) + tuple(z.shape[2:])
    return (z.new_empty(shape, dtype=torch.int64), z.new_empty(shape),
            z.new_empty((4,), dtype=torch.int64))


# ---------------------------------------------------------------------------
# full forward (quantizer.py:63-101)
# ---------------------------------------------------------------------------
@torch.library.custom_op("vqb200::quantize", mutates_args=())
def quantize(z: Tensor, weight: Tensor, beta: float, algo: int = 0, pack: Optional[Tensor] = None,
             split_ws: Optional[Tensor] = None) -> Tuple[Tensor, Tensor, Tensor, Tensor, Tensor]:
    """(z_q, vq_loss, mse, indices, stats).  z_q is the straight-through VALUE
    z + (e - z); vq_loss = mse + beta*mse; mse is the value of both
    codebook_loss and commitment_loss."""
    _need_cuda_f32(z, "z")
    _need_cuda_f32(weight, "weight")
    z = z.contiguous()
    weight = weight.contiguous()
    B, D, HW, K = _shape_bdhw(z, weight)
    if B * HW == 0:  # nothing to launch; mean over zero elements is NaN like F.mse_loss
        nan = torch.full((), float("nan"), device=z.device)
        return (torch.empty_like(z), nan, nan.clone(),
                torch.empty((B,) + tuple(z.shape[2:]), dtype=torch.int64, device=z.device),
                torch.zeros(4, dtype=torch.int64, device=z.device))
    with _on(z.device):
        idx, _, stats = _search_into(z, weight, algo, False, pack, split_ws)
        z_q = torch.empty_like(z)
        loss = torch.empty(2, dtype=torch.float32, device=z.device)
        pbytes = lib().vqb_tail_partials_bytes(B * HW)
        partials = _bytes(pbytes, z.device)
        prof = PROFILE_TAIL
        if prof is not None:
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
        check(lib().vqb_gather_loss_st_f32(_p(z), _p(weight), _p(idx), B, D, HW, K, float(beta),
                                           _p(z_q), _p(loss), _p(partials), pbytes, None, _stream()),
              "vqb_gather_loss_st_f32")
        if prof is not None:
            ev1.record()
            prof.append((ev0, ev1))
        _count("tail")
    return z_q, loss[1].clone(), loss[0].clone(), idx, stats


@quantize.register_fake
def _(z, weight, beta, algo=0, pack=None, split_ws=None):
    shape = (z.shape[0],) + tuple(z.shape[2:])
    return (torch.empty_like(z), z.new_empty(()), z.new_empty(()),
            z.new_empty(shape, dtype=torch.int64), z.new_empty((4,), dtype=torch.int64))


@torch.library.custom_op("vqb200::quantize_backward", mutates_args=())
def quantize_backward(z: Tensor, weight: Tensor, indices: Tensor, g_zq: Optional[Tensor],
                      g_vq: Optional[Tensor], beta: float, need_dE: bool) -> Tuple[Tensor, Tensor]:
    """(dz, dE): dz = g_zq + g_vq*(2/n)(z-e); dE = g_vq*beta*(2/n) * index_add(e - z)."""
    _need_cuda_f32(z, "z")
    z = z.contiguous()
    weight = weight.contiguous()
    B, D, HW, K = _shape_bdhw(z, weight)
    if g_zq is not None:
        g_zq = g_zq.contiguous().to(torch.float32)
    if g_vq is not None:
        g_vq = g_vq.reshape(1).to(torch.float32).contiguous()
    if z.numel() == 0:
        return torch.empty_like(z), (torch.zeros_like(weight) if need_dE else weight.new_empty((0,)))
    with _on(z.device):
        dz = torch.empty_like(z)
        dE = torch.zeros_like(weight) if need_dE else None
        prof = PROFILE_BWD
        if prof is not None:
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
        check(lib().vqb_backward_f32(_p(z), _p(weight), _p(indices), _p(g_zq), _p(g_vq), float(beta),
                                     B, D, HW, K, _p(dz), _p(dE), None, _stream()),
              "vqb_backward_f32")
        if prof is not None:
            ev1.record()
            prof.append((ev0, ev1))
        _count("backward")
    if dE is None:
        dE = weight.new_empty((0,))
    return dz, dE


@quantize_backward.register_fake
def _(z, weight, indices, g_zq, g_vq, beta, need_dE):
    return torch.empty_like(z), (torch.empty_like(weight) if need_dE else weight.new_empty((0,)))


def _quantize_setup(ctx, inputs, output):
    z, weight, beta = inputs[0], inputs[1], inputs[2]
    ctx.save_for_backward(z, weight, output[3])
    ctx.beta = beta
    # only z_q and vq_loss carry gradient.  `mse` is the LOGGED value of codebook_loss / commitment_loss (the
    # reference returns them as Python floats, quantizer.py:106-107): a loss built from it must raise instead
    # of silently receiving zero gradients
    ctx.mark_non_differentiable(output[2], output[3], output[4])


def _quantize_bwd(ctx, g_zq, g_vq, g_mse, g_idx, g_stats):
    z, weight, idx = ctx.saved_tensors
    need_dz, need_dE = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
    dz, dE = quantize_backward(z, weight, idx, g_zq, g_vq, ctx.beta, bool(need_dE))
    return (dz if need_dz else None), (dE if need_dE else None), None, None, None, None


torch.library.register_autograd("vqb200::quantize", _quantize_bwd, setup_context=_quantize_setup)


# ---------------------------------------------------------------------------
# helper methods (quantizer.py:112-149)
# ---------------------------------------------------------------------------
@torch.library.custom_op("vqb200::codebook_entry", mutates_args=())
def codebook_entry(weight: Tensor, indices: Tensor) -> Tuple[Tensor, Tensor]:
    """(out[B,D,*spatial], err int32[1]); err != 0 iff an index is out of range."""
    _need_cuda_f32(weight, "weight")
    if not indices.is_cuda or indices.dtype != torch.int64:
        raise RuntimeError("indices must be a CUDA int64 tensor")
    weight = weight.contiguous()
    indices = indices.contiguous()
    K, D = int(weight.shape[0]), int(weight.shape[1])
    B = int(indices.shape[0]) if indices.dim() > 0 else 1
    HW = indices.numel() // max(B, 1) if B > 0 else 0
    out = torch.empty((B, D) + tuple(indices.shape[1:]), dtype=torch.float32, device=weight.device)
    err = torch.zeros(1, dtype=torch.int32, device=weight.device)
    if indices.numel() == 0:
        return out, err
    with _on(weight.device):
        check(lib().vqb_gather_f32(_p(weight), _p(indices), B, D, HW, K, _p(out), _p(err), _stream()),
              "vqb_gather_f32")
        _count("gather")
    return out, err


@codebook_entry.register_fake
def _(weight, indices):
    return (weight.new_empty((indices.shape[0], weight.shape[1]) + tuple(indices.shape[1:])),
            weight.new_empty((1,), dtype=torch.int32))


@torch.library.custom_op("vqb200::codebook_usage", mutates_args=())
def codebook_usage(indices: Tensor, num_embeddings: int) -> Tuple[Tensor, Tensor, Tensor]:
    """(usage int64[K], used int64[1], err int32[1])."""
    if not indices.is_cuda or indices.dtype != torch.int64:
        raise RuntimeError("indices must be a CUDA int64 tensor")
    indices = indices.contiguous()
    hist = torch.empty(num_embeddings, dtype=torch.int64, device=indices.device)
    used = torch.empty(1, dtype=torch.int64, device=indices.device)
    err = torch.zeros(1, dtype=torch.int32, device=indices.device)
    with _on(indices.device):
        check(lib().vqb_hist_i64(_p(indices), indices.numel(), num_embeddings, _p(hist), _p(used),
                                 _p(err), _stream()), "vqb_hist_i64")
        _count("hist")
    return hist, used, err


@codebook_usage.register_fake
def _(indices, num_embeddings):
    return (indices.new_empty((num_embeddings,)), indices.new_empty((1,)),
            indices.new_empty((1,), dtype=torch.int32))


# ---------------------------------------------------------------------------
# extensions
# ---------------------------------------------------------------------------
@torch.library.custom_op("vqb200::code_sums", mutates_args=())
def code_sums(z: Tensor, indices: Tensor, num_embeddings: int) -> Tuple[Tensor, Tensor]:
    """(counts f32[K], sums f32[K,D]) = per-code token count and token sum."""
    _need_cuda_f32(z, "z")
    z = z.contiguous()
    indices = indices.contiguous()
    B, D = int(z.shape[0]), int(z.shape[1])
    HW = z.numel() // max(B * D, 1)
    counts = torch.zeros(num_embeddings, dtype=torch.float32, device=z.device)
    sums = torch.zeros(num_embeddings, D, dtype=torch.float32, device=z.device)
    with _on(z.device):
        check(lib().vqb_code_sums_f32(_p(z), _p(indices), B, D, HW, num_embeddings, _p(counts),
                                      _p(sums), _stream()), "vqb_code_sums_f32")
    return counts, sums


@code_sums.register_fake
def _(z, indices, num_embeddings):
    return z.new_empty((num_embeddings,)), z.new_empty((num_embeddings, z.shape[1]))


@torch.library.custom_op("vqb200::ema_update",
                         mutates_args=("weight", "cluster_size", "embed_sum"))
def ema_update(weight: Tensor, cluster_size: Tensor, embed_sum: Tensor, counts: Tensor,
               sums: Tensor, decay: float, eps: float) -> None:
    for t, n in ((weight, "weight"), (cluster_size, "cluster_size"), (embed_sum, "embed_sum"),
                 (counts, "counts"), (sums, "sums")):
        _need_cuda_f32(t, n)
        if not t.is_contiguous():
            raise RuntimeError(f"{n} must be contiguous")
    K, D = int(weight.shape[0]), int(weight.shape[1])
    scratch = torch.empty(1, dtype=torch.float32, device=weight.device)
    with _on(weight.device):
        check(lib().vqb_ema_update_f32(_p(weight), _p(cluster_size), _p(embed_sum), _p(counts),
                                       _p(sums), K, D, float(decay), float(eps), _p(scratch),
                                       _stream()), "vqb_ema_update_f32")


@torch.library.custom_op("vqb200::pack_argmin_keys", mutates_args=())
def pack_argmin_keys(dmin: Tensor, indices: Tensor, index_offset: int) -> Tensor:
    _need_cuda_f32(dmin, "dmin")
    dmin = dmin.contiguous()
    indices = indices.contiguous()
    keys = torch.empty(indices.shape, dtype=torch.int64, device=dmin.device)
    with _on(dmin.device):
        check(lib().vqb_pack_argmin_keys(_p(dmin), _p(indices), indices.numel(), int(index_offset),
                                         _p(keys), _stream()), "vqb_pack_argmin_keys")
    return keys


@pack_argmin_keys.register_fake
def _(dmin, indices, index_offset):
    return torch.empty_like(indices)


@torch.library.custom_op("vqb200::unpack_argmin_keys", mutates_args=())
def unpack_argmin_keys(keys: Tensor) -> Tuple[Tensor, Tensor]:
    if not keys.is_cuda or keys.dtype != torch.int64:
        raise RuntimeError("keys must be a CUDA int64 tensor")
    keys = keys.contiguous()
    idx = torch.empty_like(keys)
    dmin = torch.empty(keys.shape, dtype=torch.float32, device=keys.device)
    with _on(keys.device):
        check(lib().vqb_unpack_argmin_keys(_p(keys), keys.numel(), _p(idx), _p(dmin), _stream()),
              "vqb_unpack_argmin_keys")
    return idx, dmin


@unpack_argmin_keys.register_fake
def _(keys):
    return torch.empty_like(keys), keys.new_empty(keys.shape, dtype=torch.float32)


def stats_pack(dE: Tensor, hist: Optional[Tensor], scalars: Optional[Tensor]) -> Tensor:
    """One flat fp32 message [dE | scalars | hist mod 4096 | hist div 4096] (CUDA tensors, one launch)."""
    _need_cuda_f32(dE, "dE")
    dE = dE.contiguous()
    n_s = 0 if scalars is None else scalars.numel()
    n_h = 0 if hist is None else hist.numel()
    if scalars is not None:
        scalars = scalars.reshape(-1).float().contiguous()
    if hist is not None:
        hist = hist.reshape(-1).long().contiguous()
    flat = torch.empty(dE.numel() + n_s + 2 * n_h, dtype=torch.float32, device=dE.device)
    with _on(dE.device):
        check(lib().vqb_stats_pack(_p(dE), dE.numel(), _p(scalars), n_s, _p(hist), n_h, _p(flat), _stream()),
              "vqb_stats_pack")
        _count("keys")
    return flat


def stats_unpack(flat: Tensor, dE_shape, n_scalars: int, n_hist: int, dE_scale: float):
    """Inverse of stats_pack: (dE * dE_scale, scalars or None, int64 hist or None)."""
    n_dE = 1
    for d in dE_shape:
        n_dE *= int(d)
    dE = torch.empty(tuple(dE_shape), dtype=torch.float32, device=flat.device)
    scalars = torch.empty(n_scalars, dtype=torch.float32, device=flat.device) if n_scalars else None
    hist = torch.empty(n_hist, dtype=torch.int64, device=flat.device) if n_hist else None
    with _on(flat.device):
        check(lib().vqb_stats_unpack(_p(flat), n_dE, n_scalars, n_hist, float(dE_scale), _p(dE), _p(scalars), _p(hist),
                                     _stream()), "vqb_stats_unpack")
        _count("keys")
    return dE, scalars, hist


_NARROW_DTYPES = {1: torch.uint8, 2: torch.uint16, 4: torch.int32}


@torch.library.custom_op("vqb200::indices_narrow", mutates_args=())
def indices_narrow(indices: Tensor, num_embeddings: int) -> Tuple[Tensor, Tensor]:
    """(codes uint8/uint16/int32 of the same shape, err int32[1]) -- compact index map (row N3)."""
    if not indices.is_cuda or indices.dtype != torch.int64:
        raise RuntimeError("indices must be a CUDA int64 tensor")
    indices = indices.contiguous()
    w = lib().vqb_index_bytes(int(num_embeddings))
    if w == 0:
        raise RuntimeError("num_embeddings must be positive")
    codes = torch.empty(indices.shape, dtype=_NARROW_DTYPES[w], device=indices.device)
    err = torch.zeros(1, dtype=torch.int32, device=indices.device)
    with _on(indices.device):
        check(lib().vqb_indices_narrow(_p(indices), indices.numel(), int(num_embeddings), _p(codes), w,
                                       _p(err), _stream()), "vqb_indices_narrow")
        _count("keys")
    return codes, err


@indices_narrow.register_fake
def _(indices, num_embeddings):
    w = 1 if num_embeddings <= 256 else (2 if num_embeddings <= 65536 else 4)
    return (indices.new_empty(indices.shape, dtype=_NARROW_DTYPES[w]),
            indices.new_empty((1,), dtype=torch.int32))


@torch.library.custom_op("vqb200::indices_widen", mutates_args=())
def indices_widen(codes: Tensor) -> Tensor:
    """int64 indices from a compact index map (uint8 / uint16 / int32)."""
    if not codes.is_cuda or codes.dtype not in (torch.uint8, torch.uint16, torch.int32):
        raise RuntimeError("codes must be a CUDA uint8 / uint16 / int32 tensor")
    codes = codes.contiguous()
    out = torch.empty(codes.shape, dtype=torch.int64, device=codes.device)
    with _on(codes.device):
        check(lib().vqb_indices_widen(_p(codes), codes.numel(), codes.element_size(), _p(out), _stream()),
              "vqb_indices_widen")
        _count("keys")
    return out


@indices_widen.register_fake
def _(codes):
    return codes.new_empty(codes.shape, dtype=torch.int64)


# ---------------------------------------------------------------------------
# 1x1 convolution either side of the quantizer (row N1; vq_vae.py:74-79,115,121)
# ---------------------------------------------------------------------------
PROFILE_CONV = None


@torch.library.custom_op("vqb200::conv1x1", mutates_args=())
def conv1x1(x: Tensor, weight: Tensor, bias: Optional[Tensor], algo: int = 0) -> Tensor:
    """y[B,Cout,*spatial] = weight[Cout,Cin] . x[B,Cin,*spatial] + bias (fp32, NCHW in and out)."""
    _need_cuda_f32(x, "x")
    _need_cuda_f32(weight, "weight")
    x = x.contiguous()
    weight = weight.reshape(weight.shape[0], -1).contiguous()
    if x.dim() < 2 or int(x.shape[1]) != int(weight.shape[1]):
        raise RuntimeError(f"conv1x1: input channels {tuple(x.shape)} do not match the weight {tuple(weight.shape)}")
    if bias is not None:
        _need_cuda_f32(bias, "bias")
        bias = bias.contiguous()
    B, Cin, Cout = int(x.shape[0]), int(x.shape[1]), int(weight.shape[0])
    HW = x.numel() // max(B * Cin, 1)
    y = torch.empty((B, Cout) + tuple(x.shape[2:]), dtype=torch.float32, device=x.device)
    if y.numel() == 0:
        return y
    with _on(x.device):
        wb = lib().vqb_conv1x1_workspace_bytes(Cin, Cout)
        ws = _bytes(wb, x.device)
        prof = PROFILE_CONV
        if prof is not None:
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
        check(lib().vqb_conv1x1_f32(_p(x), B, Cin, HW, _p(weight), _p(bias), Cout, _p(y), _p(ws), wb, int(algo),
                                    _stream()), "vqb_conv1x1_f32")
        if prof is not None:
            ev1.record()
            prof.append((ev0, ev1))
        _count("conv", 2 if wb else 1)
    return y


@conv1x1.register_fake
def _(x, weight, bias, algo=0):
    return x.new_empty((x.shape[0], weight.shape[0]) + tuple(x.shape[2:]))


def split_eligible(cin: int, cout: int, tokens: int, K: int) -> bool:
    """Shapes for which `QuantConv1x1.feed` switches to `conv1x1_split`: the convolution's tensor path, feeding a
    quantizer whose automatic choice is the fp16 tensor search (the consumer of the split), and Cout <= 128 -- measured
    (scripts/conv_split_bench.py): 128 -> 64 channels 5.27 -> 4.64 ms for conv + quantizer forward on 1M tokens, but at
    256 -> 256 the extra epilogue work of the tensor-bound convolution (+0.32 ms) outweighs the split pass it replaces
    (0.27 ms).  The op itself accepts Cout up to 256."""
    return (cin % 32 == 0 and cout % 16 == 0 and 16 < cout <= 128 and tokens * K * cout >= (1 << 29))


@torch.library.custom_op("vqb200::conv1x1_split", mutates_args=())
def conv1x1_split(x: Tensor, weight: Tensor, bias: Optional[Tensor], codebook: Tensor, pack: Tensor) -> Tuple[Tensor, Tensor]:
    """(y, search_workspace): the 1x1 convolution of `conv1x1` (tensor path) that ALSO writes, from the accumulator, the
    token split the fp16 tensor search of the quantizer with `codebook` [K, Cout] needs (`pack` = prepare_codebook(codebook)).
    Hand both to `quantize(y, codebook, beta, 0, pack, search_workspace)`: the search then skips its own split pass."""
    _need_cuda_f32(x, "x")
    _need_cuda_f32(weight, "weight")
    _need_cuda_f32(codebook, "codebook")
    x = x.contiguous()
    weight = weight.reshape(weight.shape[0], -1).contiguous()
    if bias is not None:
        _need_cuda_f32(bias, "bias")
        bias = bias.contiguous()
    B, Cin, Cout = int(x.shape[0]), int(x.shape[1]), int(weight.shape[0])
    K = int(codebook.shape[0])
    if int(weight.shape[1]) != Cin or int(codebook.shape[1]) != Cout:
        raise RuntimeError("conv1x1_split: channel mismatch between x, the weight and the codebook")
    HW = x.numel() // max(B * Cin, 1)
    y = torch.empty((B, Cout) + tuple(x.shape[2:]), dtype=torch.float32, device=x.device)
    with _on(x.device):
        sb = lib().vqb_search_workspace_bytes(B, Cout, HW, K, _cabi.ALGO_TCGEN05_F16)
        sws = _bytes(sb, x.device)
        if y.numel() == 0:
            return y, sws
        wb = lib().vqb_conv1x1_workspace_bytes(Cin, Cout)
        ws = _bytes(wb, x.device)
        prof = PROFILE_CONV
        if prof is not None:
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
        check(lib().vqb_conv1x1_split_f32(_p(x), B, Cin, HW, _p(weight), _p(bias), Cout, _p(y), _p(ws), wb, _p(pack), K,
                                          _p(sws), sb, _stream()), "vqb_conv1x1_split_f32")
        if prof is not None:
            ev1.record()
            prof.append((ev0, ev1))
        _count("conv", 2)
    return y, sws


@conv1x1_split.register_fake
def _(x, weight, bias, codebook, pack):
    return (x.new_empty((x.shape[0], weight.shape[0]) + tuple(x.shape[2:])), x.new_empty((1,), dtype=torch.uint8))


def _conv1x1_split_setup(ctx, inputs, output):
    x, weight, bias, _codebook, _pack = inputs
    ctx.save_for_backward(x, weight)
    ctx.has_bias = bias is not None
    ctx.algo = 0
    ctx.mark_non_differentiable(output[1])


def _conv1x1_split_bwd(ctx, gy, g_ws):
    gx, gw, gb, _ = _conv1x1_bwd(ctx, gy)
    return gx, gw, gb, None, None


def _conv1x1_setup(ctx, inputs, output):
    x, weight, bias, algo = inputs
    ctx.save_for_backward(x, weight)
    ctx.has_bias = bias is not None
    ctx.algo = algo


PROFILE_CONV_DW = None


def conv1x1_param_grads(gy: Tensor, x: Tensor, want_w: bool = True, want_b: bool = True):
    """(dW[Cout, Cin], dbias[Cout]) of the 1x1 convolution: dW[o, c] = sum over tokens gy[b, o, t] x[b, c, t], a
    tokens-long reduction.  This library's 3xTF32 tcgen05 kernel (vqb_conv1x1_dw_f32: both operands by TMA straight from
    NCHW, dbias summed on the way) for Cin % 16 == 0, Cout <= 256, HW >= 32, HW % 4 == 0; other shapes keep the plain
    library GEMM (torch) -- there is no slow path inside the library."""
    _need_cuda_f32(gy, "gy")
    _need_cuda_f32(x, "x")
    gy = gy.contiguous()
    x = x.contiguous()
    B, Cout, Cin = int(gy.shape[0]), int(gy.shape[1]), int(x.shape[1])
    HW = x.numel() // max(B * Cin, 1)
    if want_w and B * HW > 0 and lib().vqb_conv1x1_dw_supported(Cin, Cout, HW):
        with _on(x.device):
            gw = torch.zeros((Cout, Cin), dtype=torch.float32, device=x.device)
            gb = torch.zeros((Cout,), dtype=torch.float32, device=x.device) if want_b else None
            prof = PROFILE_CONV_DW
            if prof is not None:
                ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ev0.record()
            check(lib().vqb_conv1x1_dw_f32(_p(gy), _p(x), B, Cin, Cout, HW, _p(gw), _p(gb), _stream()), "vqb_conv1x1_dw_f32")
            if prof is not None:
                ev1.record()
                prof.append((ev0, ev1))
            _count("conv", 1)
        return gw, gb
    gw = gb = None
    if want_w:
        gw = torch.einsum("bot,bct->oc", gy.reshape(B, Cout, -1), x.reshape(B, Cin, -1))
    if want_b:
        gb = gy.reshape(B, Cout, -1).sum(dim=(0, 2))
    return gw, gb


def _conv1x1_bwd(ctx, gy):
    x, weight = ctx.saved_tensors
    w2 = weight.reshape(weight.shape[0], -1)
    gx = gw = gb = None
    gy = gy.contiguous()
    if ctx.needs_input_grad[0]:
        # dx = W^T . dy is the same 1x1 convolution with the transposed weight: this library's kernel
        # (a forced tensor path only applies when the transposed shape qualifies as well)
        cin_t, cout_t = int(w2.shape[0]), int(w2.shape[1])
        algo_t = ctx.algo if (ctx.algo != 1 or (cin_t % 32 == 0 and cout_t % 16 == 0 and cout_t <= 256)) else 0
        gx = conv1x1(gy, w2.t().contiguous(), None, algo_t)
    want_w = ctx.needs_input_grad[1]
    want_b = ctx.has_bias and ctx.needs_input_grad[2]
    if want_w or want_b:
        gw, gb = conv1x1_param_grads(gy, x, want_w, want_b)
        gw = gw.reshape(weight.shape) if want_w else None
        gb = gb if want_b else None
    return gx, gw, gb, None


torch.library.register_autograd("vqb200::conv1x1", _conv1x1_bwd, setup_context=_conv1x1_setup)
torch.library.register_autograd("vqb200::conv1x1_split", _conv1x1_split_bwd, setup_context=_conv1x1_split_setup)


# ---------------------------------------------------------------------------
# GroupNorm + SiLU in front of the convolutions either side of the path (row N2; encoder_decoder.py:166-167, 249-250)
# ---------------------------------------------------------------------------
@torch.library.custom_op("vqb200::groupnorm_silu", mutates_args=())
def groupnorm_silu(x: Tensor, weight: Tensor, bias: Tensor, num_groups: int, eps: float) -> Tuple[Tensor, Tensor, Tensor]:
    """(y, mean[B*G], rstd[B*G]) with y = silu(group_norm(x, num_groups, weight, bias, eps)); x is [B, C, *spatial]."""
    _need_cuda_f32(x, "x")
    _need_cuda_f32(weight, "weight")
    _need_cuda_f32(bias, "bias")
    x = x.contiguous()
    if x.dim() < 2 or int(x.shape[1]) % int(num_groups) != 0 or weight.numel() != x.shape[1] or bias.numel() != x.shape[1]:
        raise RuntimeError(f"groupnorm_silu: {tuple(x.shape)} channels do not match {num_groups} groups / the affine parameters")
    B, C = int(x.shape[0]), int(x.shape[1])
    HW = x.numel() // max(B * C, 1)
    y = torch.empty_like(x)
    mean = torch.empty(B * num_groups, dtype=torch.float32, device=x.device)
    rstd = torch.empty(B * num_groups, dtype=torch.float32, device=x.device)
    if x.numel() == 0:
        return y, mean, rstd
    with _on(x.device):
        check(lib().vqb_groupnorm_silu_f32(_p(x), B, C, HW, _p(weight.contiguous()), _p(bias.contiguous()), int(num_groups),
                                           float(eps), _p(y), _p(mean), _p(rstd), _stream()), "vqb_groupnorm_silu_f32")
        _count("keys")
    return y, mean, rstd


@groupnorm_silu.register_fake
def _(x, weight, bias, num_groups, eps):
    return torch.empty_like(x), x.new_empty((x.shape[0] * num_groups,)), x.new_empty((x.shape[0] * num_groups,))


@torch.library.custom_op("vqb200::groupnorm_silu_backward", mutates_args=())
def groupnorm_silu_backward(gy: Tensor, x: Tensor, weight: Tensor, bias: Tensor, mean: Tensor, rstd: Tensor,
                            num_groups: int) -> Tuple[Tensor, Tensor, Tensor]:
    """(dx, dweight, dbias)."""
    gy = gy.contiguous().to(torch.float32)
    x = x.contiguous()
    B, C = int(x.shape[0]), int(x.shape[1])
    HW = x.numel() // max(B * C, 1)
    dx = torch.empty_like(x)
    dw = torch.zeros(C, dtype=torch.float32, device=x.device)
    db = torch.zeros(C, dtype=torch.float32, device=x.device)
    if x.numel() == 0:
        return dx, dw, db
    with _on(x.device):
        check(lib().vqb_groupnorm_silu_backward_f32(_p(gy), _p(x), B, C, HW, _p(weight.contiguous()), _p(bias.contiguous()),
                                                    int(num_groups), _p(mean), _p(rstd), _p(dx), _p(dw), _p(db), _stream()),
              "vqb_groupnorm_silu_backward_f32")
        _count("keys")
    return dx, dw, db


@groupnorm_silu_backward.register_fake
def _(gy, x, weight, bias, mean, rstd, num_groups):
    return torch.empty_like(x), torch.empty_like(weight), torch.empty_like(bias)


def _gn_setup(ctx, inputs, output):
    x, weight, bias, num_groups, _eps = inputs
    ctx.save_for_backward(x, weight, bias, output[1], output[2])
    ctx.num_groups = num_groups
    ctx.mark_non_differentiable(output[1], output[2])


def _gn_bwd(ctx, gy, g_mean, g_rstd):
    x, weight, bias, mean, rstd = ctx.saved_tensors
    dx, dw, db = groupnorm_silu_backward(gy, x, weight, bias, mean, rstd, ctx.num_groups)
    return (dx if ctx.needs_input_grad[0] else None, dw if ctx.needs_input_grad[1] else None,
            db if ctx.needs_input_grad[2] else None, None, None)


torch.library.register_autograd("vqb200::groupnorm_silu", _gn_bwd, setup_context=_gn_setup)


def fma_peak_tflops(packed: bool, iters: int = 4096, repeats: int = 5) -> float:
    """Measured FP32 FMA peak of the current device (roofline denominator for the
    low-D search): best of `repeats`, CUDA events on the current stream."""
    sink = torch.zeros(1, dtype=torch.float32, device="cuda")
    flops = ctypes.c_double(0.0)
    best = 0.0
    blib = _cabi.bench_lib()  # measurement build; the product library carries no microbenchmarks
    for _ in range(repeats + 1):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        check(blib.vqb_fma_peak_launch(int(packed), iters, _p(sink), ctypes.byref(flops), _stream()),
              "vqb_fma_peak_launch")
        b.record()
        b.synchronize()
        ms = a.elapsed_time(b)
        best = max(best, flops.value / (ms * 1e-3) / 1e12)
    return best
