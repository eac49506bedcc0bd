"""Compact storage of quantizer indices.

The reference caches per-image fp32 latents with `torch.save` (`preprocess_latents.py:236-238`);
for a VQ model the natural cache is the index map produced by `VQVAE.encode_to_indices`
(`vq_vae.py:162-175`).  Indices are stored in the narrowest unsigned type that holds K-1
(uint8 / uint16 / int32), i.e. 2 bytes per token for K <= 65536 instead of 4*D bytes of latent.
CUDA tensors are narrowed / widened on the device (csrc/vqb_indexio.cu); CPU tensors on the host.
"""
from typing import Dict

import torch


def index_dtype(num_embeddings: int) -> torch.dtype:
    if num_embeddings <= 0:
        raise ValueError("num_embeddings must be positive")
    if num_embeddings <= 1 << 8:
        return torch.uint8
    if num_embeddings <= 1 << 16:
        return torch.uint16
    if num_embeddings <= 1 << 31:
        return torch.int32
    return torch.int64


def pack_indices(indices: torch.Tensor, num_embeddings: int) -> Dict[str, torch.Tensor]:
    """int64 [B,H,W] -> {'codes': narrow tensor, 'num_embeddings': K} (CPU tensors, ready for torch.save).

    CUDA indices are narrowed on the device by libvqb200 (`vqb_indices_narrow`) so that only
    1-4 bytes per token cross PCIe; CPU indices (e.g. loaded from a cache) are narrowed on the host."""
    if indices.dtype != torch.int64:
        raise TypeError("indices must be int64 (as returned by the quantizer)")
    if indices.is_cuda:
        from . import ops
        codes, err = ops.indices_narrow(indices, int(num_embeddings))
        codes = codes.cpu()  # the copy also orders the read of `err` below
        if int(err.item()) != 0:
            raise ValueError("index outside [0, num_embeddings)")
        return {"codes": codes, "num_embeddings": torch.tensor(num_embeddings)}
    if indices.numel() and (int(indices.min()) < 0 or int(indices.max()) >= num_embeddings):
        raise ValueError("index outside [0, num_embeddings)")
    dt = index_dtype(num_embeddings)
    if dt == torch.uint16:  # torch has no int64->uint16 cast kernel on every backend: go through int32
        codes = indices.to(torch.int32).numpy().astype("uint16")
        codes = torch.from_numpy(codes)
    else:
        codes = indices.to(dt)
    return {"codes": codes, "num_embeddings": torch.tensor(num_embeddings)}


def unpack_indices(blob: Dict[str, torch.Tensor], device=None) -> torch.Tensor:
    """Inverse of pack_indices: int64 indices on `device` (for get_codebook_entry / decode_from_indices).
    For a CUDA device the compact codes are copied first and widened there (`vqb_indices_widen`)."""
    codes = blob["codes"]
    if device is not None and torch.device(device).type == "cuda":
        from . import ops
        return ops.indices_widen(codes.to(device))
    if codes.dtype == torch.uint16:
        out = torch.from_numpy(codes.cpu().numpy().astype("int64"))
    else:
        out = codes.to(torch.int64)
    return out.to(device) if device is not None else out


def bytes_per_token(num_embeddings: int) -> int:
    return torch.empty(0, dtype=index_dtype(num_embeddings)).element_size()
