#!/usr/bin/env python
"""Recipe for `oracle/_ref/` -- TEST INFRASTRUCTURE ONLY.

The reference (heimaoqqq/vq-gan) is 100 % Python, so there is nothing to compile; what
"building the reference" means for this path is staging the UNMODIFIED module files the hot
path lives in where the GPU box can import them (`/root/reference` does not exist there):

    vqgan_ldm_baseline/models/quantizer.py        the hot path itself (VectorQuantizer)
    vqgan_ldm_baseline/models/vq_vae.py           its only caller (VQVAE; swap point :82-86)
    vqgan_ldm_baseline/models/encoder_decoder.py  the conv stacks either side (needed by VQVAE)

They are copied byte for byte into `oracle/_ref/refmodels/`, which is git-ignored (never in
history) but NOT gpurun-ignored (it travels with the snapshot like a built .so).  The
package `__init__.py` written next to them is empty on purpose: the reference's own
`models/__init__.py:9` imports `lpips`, which is not installed.  A SHA-256 manifest is
written so tests can tell that the copies are the files the goldens were generated from.

    python oracle/make_ref.py            # no-op when /root/reference is absent

Only `tests/`, `__graft_entry__` (build / smoke) and `bench.py`'s CPU arms use this.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("VQ_REFERENCE_ROOT", "/root/reference")
SRC_DIR = os.path.join(REF_ROOT, "vqgan_ldm_baseline", "models")
DST_DIR = os.path.join(HERE, "_ref", "refmodels")
FILES = ("quantizer.py", "vq_vae.py", "encoder_decoder.py")


def make(verbose: bool = True) -> bool:
    """Returns True when oracle/_ref/refmodels holds the three reference files."""
    if not os.path.isdir(SRC_DIR):
        ok = all(os.path.exists(os.path.join(DST_DIR, f)) for f in FILES)
        if verbose:
            print(f"make_ref: {SRC_DIR} not present; existing copy {'found' if ok else 'absent'}")
        return ok
    os.makedirs(DST_DIR, exist_ok=True)
    manifest = {}
    for name in FILES:
        src, dst = os.path.join(SRC_DIR, name), os.path.join(DST_DIR, name)
        shutil.copyfile(src, dst)
        with open(dst, "rb") as f:
            manifest[name] = hashlib.sha256(f.read()).hexdigest()
    with open(os.path.join(DST_DIR, "__init__.py"), "w") as f:
        f.write("# empty on purpose: see oracle/make_ref.py\n")
    with open(os.path.join(DST_DIR, "MANIFEST.json"), "w") as f:
        json.dump({"source": SRC_DIR, "sha256": manifest}, f, indent=1)
    if verbose:
        print(f"make_ref: staged {', '.join(FILES)} in {DST_DIR}")
    return True


if __name__ == "__main__":
    sys.exit(0 if make() else 1)
