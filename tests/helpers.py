import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    with np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False) as f:
        return {k: f[k] for k in f.files}


def digest_matches(t: torch.Tensor, gold, key, rtol=1e-6, atol=1e-7):
    """Compares a tensor with the stored digest (strided sample + sums)."""
    flat = t.detach().cpu().reshape(-1)
    sample = flat[::97].numpy()
    np.testing.assert_allclose(sample, gold[key + "_sample"], rtol=rtol, atol=atol)
    abssum = flat.double().abs().sum().item()
    assert abs(abssum - float(gold[key + "_abssum"])) <= 1e-6 * max(1.0, float(gold[key + "_abssum"]))
