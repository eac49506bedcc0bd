"""vq_gan_b200 -- B200-native drop-in for the VQ-GAN vector-quantizer bottleneck.

Public surface (mirrors vqgan_ldm_baseline/models/quantizer.py of heimaoqqq/vq-gan):
    VectorQuantizer(num_embeddings, embedding_dim, commitment_cost=0.25)
plus the `vqb200::*` torch.library ops in `vq_gan_b200.ops` and the multi-GPU
helpers in `vq_gan_b200.distributed`.  The compute lives in lib/libvqb200.so
(hand-written sm_100a CUDA behind the C ABI of include/vqb200.h).
"""
from ._cabi import (ALGO_AUTO, ALGO_FP32_TILE, ALGO_LOWD_FMA, ALGO_TCGEN05, ALGO_TCGEN05_F16, ALGO_TCGEN05_TF32X3,
                    LIB_PATH,
                    VqbError, lib)
from . import ops
from .quantizer import EMAVectorQuantizer, GroupNormSiLU, LossDict, QuantConv1x1, VectorQuantizer, encoder_tail

__all__ = ["VectorQuantizer", "EMAVectorQuantizer", "QuantConv1x1", "GroupNormSiLU", "encoder_tail", "LossDict", "ops", "lib", "LIB_PATH", "VqbError", "ALGO_AUTO", "ALGO_LOWD_FMA",
           "ALGO_FP32_TILE", "ALGO_TCGEN05", "ALGO_TCGEN05_F16", "ALGO_TCGEN05_TF32X3"]
__version__ = "0.1.0"
