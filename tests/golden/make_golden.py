"""Generates tests/golden/*.npz by running the UNMODIFIED reference module.

Run in the build container only (it reads /root/reference, which does not
exist on the GPU box):

    python tests/golden/make_golden.py

The reference package cannot be imported as shipped (models/__init__.py needs
lpips, configs/__init__.py imports a missing file), so quantizer.py is loaded by
file path.  Inputs are regenerated from the seeds recorded in each fixture by
`tests/cases.py` (torch CPU generators are deterministic for a fixed torch
build); small cases also store their inputs so a generator change is caught.
"""
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from cases import CASES, make_case  # noqa: E402
from oracle import vq_oracle as orc  # noqa: E402

REF_FILE = "/root/reference/vqgan_ldm_baseline/models/quantizer.py"


def load_reference():
    spec = importlib.util.spec_from_file_location("ref_quantizer", REF_FILE)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def digest(t: torch.Tensor):
    f = t.detach().reshape(-1)
    return {"sum": np.float64(f.double().sum().item()),
            "abssum": np.float64(f.double().abs().sum().item()),
            "sample": f[::97].numpy().copy()}


def run_case(ref, name):
    spec = CASES[name]
    c = make_case(name)
    z, E, g = c["z"], c["E"], c["g_zq"]
    K, D = E.shape
    vq = ref.VectorQuantizer(K, D, spec["beta"])
    with torch.no_grad():
        vq.embedding.weight.copy_(E)
    zr = z.clone().requires_grad_(True)
    z_q, loss_dict, indices = vq(zr)
    (loss_dict["vq_loss"] + (z_q * g).sum()).backward()
    usage, ratio = vq.get_codebook_usage(indices)
    entry = vq.get_codebook_entry(indices)

    rows = orc.tokens_of(z)
    gapinfo = orc.search_with_gap(rows, E)
    assert torch.equal(gapinfo["idx"], indices.reshape(-1)), name

    out = {
        "indices": indices.numpy().astype(np.int32),
        "vq_loss": np.float32(loss_dict["vq_loss"].item()),
        "codebook_loss": np.float32(loss_dict["codebook_loss"]),
        "commitment_loss": np.float32(loss_dict["commitment_loss"]),
        "usage": usage.numpy().astype(np.int64),
        "usage_ratio": np.float64(ratio),
        "gap": gapinfo["gap"].numpy(),
        "s": gapinfo["s"].numpy(),
        "dE": vq.embedding.weight.grad.numpy().copy(),
        "zq_is_contiguous": np.bool_(z_q.is_contiguous()),
        "entry_equals_gather": np.bool_(torch.equal(entry, orc.codebook_entry(E, indices))),
    }
    full = spec.get("store_full", False)
    for key, t in (("z_q", z_q), ("dz", zr.grad), ("entry", entry)):
        dg = digest(t)
        out[key + "_sum"] = dg["sum"]
        out[key + "_abssum"] = dg["abssum"]
        out[key + "_sample"] = dg["sample"]
        if full:
            out[key] = t.detach().numpy().copy()
    if full:
        out["in_z"] = z.numpy().copy()
        out["in_E"] = E.numpy().copy()
        out["in_g"] = g.numpy().copy()
    return out


def main():
    ref = load_reference()
    # constructor RNG contract (quantizer.py:47-48)
    torch.manual_seed(42)
    vq = ref.VectorQuantizer(128, 256, 0.25)
    np.savez_compressed(os.path.join(HERE, "init_seed42_k128_d256.npz"),
                        weight=vq.embedding.weight.detach().numpy(),
                        state_keys=np.array(sorted(vq.state_dict().keys())),
                        next_rand=torch.rand(4).numpy())
    # argmin NaN / tie behaviour of the installed ATen (quantizer.py:76)
    np.savez_compressed(os.path.join(HERE, "argmin_semantics.npz"),
                        nan_case=np.int64(torch.argmin(torch.tensor([3.0, float("nan"), 1.0])).item()),
                        tie_case=np.int64(torch.argmin(torch.tensor([2.0, 1.0, 1.0, 5.0])).item()))
    for name in CASES:
        out = run_case(ref, name)
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print(f"{name}: N={out['indices'].size} loss={out['vq_loss']:.6g} "
              f"used={out['usage_ratio']:.3f} -> {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
