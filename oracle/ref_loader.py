"""Loader for the staged reference modules (`oracle/_ref/refmodels`, see make_ref.py).
TEST INFRASTRUCTURE ONLY -- the product package never imports this.

`reference_quantizer_class()` is the UNMODIFIED `VectorQuantizer` of
vqgan_ldm_baseline/models/quantizer.py:17-149; `reference_vqvae_class()` the UNMODIFIED
`VQVAE` of models/vq_vae.py:18-226 (whose relative imports resolve inside the staged
package, so the reference's own `models/__init__.py` -- which needs lpips -- never runs).
"""
import importlib
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_PKG_PARENT = os.path.join(HERE, "_ref")
_PKG = "refmodels"


def available() -> bool:
    d = os.path.join(REF_PKG_PARENT, _PKG)
    return all(os.path.exists(os.path.join(d, f)) for f in ("quantizer.py", "vq_vae.py", "encoder_decoder.py"))


def _import(mod: str):
    if not available():
        raise ImportError("oracle/_ref/refmodels is absent: run `python oracle/make_ref.py` where "
                          "/root/reference exists (the build container)")
    if REF_PKG_PARENT not in sys.path:
        sys.path.insert(0, REF_PKG_PARENT)
    return importlib.import_module(f"{_PKG}.{mod}")


def reference_quantizer_class():
    return _import("quantizer").VectorQuantizer


def reference_vqvae_class():
    return _import("vq_vae").VQVAE


def default_vqvae_kwargs():
    """The architecture of `VQGANConfig` (configs/vqgan_config.py:36-52): what
    train_vqgan.py:138-150 passes to VQVAE."""
    return dict(in_channels=3, out_channels=3, ch=128, ch_mult=(1, 2, 2, 4), num_res_blocks=2,
                attn_resolutions=(16,), dropout=0.0, z_channels=256, num_embeddings=128,
                embedding_dim=256, commitment_cost=0.25)
