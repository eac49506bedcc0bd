"""CUDA-graph replay of the quantizer step for launch-bound shapes.

At the reference's own default shape (4096 tokens, K=128, D=256; `vqgan_config.py:49-52`) one forward +
backward is ~15 kernel launches worth ~0.1 ms of GPU time behind ~0.5 ms of Python / dispatcher / launch
overhead.  `GraphedVectorQuantizer` captures forward and backward once (static shapes, no host sync:
the module runs with `lazy_stats=True`) and replays them with two `cudaGraphLaunch` calls per step.
Capture is plain CUDA-graph stream capture of the libvqb200 launches (kernels, memsets, TMA descriptors
passed by value) -- no tracing compiler involved.
"""
from typing import Dict, Tuple

import torch
import torch.nn as nn

from .quantizer import VectorQuantizer


class _FlatQuantizer(nn.Module):
    """(z_q, vq_loss, mse) as a flat tuple of differentiable tensors + indices (graph capture wants tensors)."""

    def __init__(self, vq: VectorQuantizer):
        super().__init__()
        self.vq = vq

    def forward(self, z: torch.Tensor):
        from . import ops
        z_q, vq_loss, mse, indices, _ = ops.quantize(z, self.vq.embedding.weight, float(self.vq.commitment_cost),
                                                     self.vq.algo)
        return z_q, vq_loss, mse.detach(), indices


class GraphedVectorQuantizer(nn.Module):
    """Drop-in for a `VectorQuantizer` whose input shape is fixed: same `forward(z) -> (z_q, loss_dict,
    indices)` contract (loss values stay 0-dim device tensors, like `lazy_stats=True`), same parameter
    (`.vq.embedding.weight` is the wrapped module's), CUDA-graph replay underneath.

    `sample_z` fixes shape / device; it must require grad if the real inputs will."""

    def __init__(self, vq: VectorQuantizer, sample_z: torch.Tensor, num_warmup_iters: int = 3):
        super().__init__()
        if not sample_z.is_cuda:
            raise RuntimeError("GraphedVectorQuantizer needs a CUDA sample input (there is no CPU path)")
        self.vq = vq
        self._shape = tuple(sample_z.shape)
        flat = _FlatQuantizer(vq)
        sample = sample_z.detach().clone().requires_grad_(sample_z.requires_grad)
        self._graphed = torch.cuda.make_graphed_callables(flat, (sample,), num_warmup_iters=num_warmup_iters)

    @property
    def embedding(self):
        return self.vq.embedding

    def forward(self, z: torch.Tensor) -> Tuple[torch.Tensor, Dict[str, torch.Tensor], torch.Tensor]:
        if tuple(z.shape) != self._shape:
            raise RuntimeError(f"graph captured for input shape {self._shape}, got {tuple(z.shape)}")
        z_q, vq_loss, mse, indices = self._graphed(z)
        return z_q, {"vq_loss": vq_loss, "codebook_loss": mse, "commitment_loss": mse}, indices
