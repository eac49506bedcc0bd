"""Full-size parity SWEEP: every token of BASELINE.json's configs against the chunked
CPU oracle -- not a sample.

For each (config, codebook, kernel) the report lists tokens / mismatches / tokens inside the reference's own
near-tie band eps_i = 4 * 2^-23 * (|z_i|^2 + max|e|^2) / mismatches inside / OUTSIDE (must be 0).  The report
is printed and, when $VQB_REPORT_DIR (default gpurun_out/) exists, written to r02_parity_fullsize.txt there;
the committed copy lives in profiles/.  The oracle evaluates quantizer.py:68-76 literally (fp32, the
reference's evaluation order) in token chunks; chunking over tokens changes no per-token value.
"""
import os
import time

import pytest
import torch

from oracle import vq_oracle as orc

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LINES = []


def _emit(line):
    print(line)
    _LINES.append(line)
    out_dir = os.environ.get("VQB_REPORT_DIR", os.path.join(ROOT, "gpurun_out"))
    if os.path.isdir(out_dir):
        with open(os.path.join(out_dir, "r02_parity_fullsize.txt"), "w") as f:
            f.write("# full-size parity sweep: CUDA kernels vs the chunked CPU oracle (reference op order, fp32)\n")
            f.write("# columns: config codebook kernel | tokens mismatch in_band_tokens mismatch_in_band OUTSIDE | "
                    "exact-tier tokens | oracle seconds\n")
            f.write("\n".join(_LINES) + "\n")


def _oracle_sweep(z, E, chunk):
    """idx / gap / s of the reference evaluation for ALL tokens ([B,D,H,W] on the CPU)."""
    torch.set_num_threads(os.cpu_count() or 1)
    rows = orc.tokens_of(z)
    n = rows.shape[0]
    idx = torch.empty(n, dtype=torch.int64)
    gap = torch.empty(n, dtype=torch.float32)
    en = (E * E).sum(dim=1)
    Et = E.t().contiguous()
    for lo in range(0, n, chunk):
        r = rows[lo:lo + chunk]
        d = ((r * r).sum(dim=1, keepdim=True) + en) - 2 * (r @ Et)          # quantizer.py:68-72
        i = torch.argmin(d, dim=1)                                             # :76
        best = d.gather(1, i[:, None])
        d.scatter_(1, i[:, None], float("inf"))
        gap[lo:lo + chunk] = (d.min(dim=1).values - best.squeeze(1))
        idx[lo:lo + chunk] = i
    s = (rows * rows).sum(dim=1) + en.max()
    return {"idx": idx, "gap": gap, "s": s}


def _refinit(K, D):
    state = torch.get_rng_state()
    torch.manual_seed(42)
    E = orc.reference_init(K, D)
    torch.set_rng_state(state)
    return E


def _sweep(tag, z, books, algos, chunk):
    from vq_gan_b200 import ops
    zc = z.cuda()
    for bname, E in books.items():
        t0 = time.time()
        ref = _oracle_sweep(z, E, chunk)
        dt = time.time() - t0
        Ec = E.cuda()
        for algo in algos:
            idx, _, st = ops.search(zc, Ec, algo)
            rep = orc.compare_indices(idx, ref)
            st = st.tolist()
            _emit(f"{tag} {bname} algo={algo}(ran {st[1]}) | tokens={rep['tokens']} mismatch={rep['mismatch']} "
                  f"in_band_tokens={rep['in_band_tokens']} mismatch_in_band={rep['mismatch_in_band']} "
                  f"OUTSIDE={rep['outside']} | full_exact={st[0]} multi_group={st[2]} filtered={st[3]} | oracle {dt:.0f}s")
            assert rep["outside"] == 0, (tag, bname, algo, rep)
            del idx
        del Ec
    del zc
    torch.cuda.empty_cache()


def test_c2_every_token_against_the_oracle():
    z = torch.randn(1024, 4, 32, 32, generator=torch.Generator().manual_seed(0))
    books = {"normal": torch.randn(16384, 4, generator=torch.Generator().manual_seed(1)),
             "refinit": _refinit(16384, 4)}
    _sweep("C2(N=2^20,D=4,K=16384)", z, books, (0, 1, 5), 16384)


def test_c3_every_token_against_the_oracle():
    z = torch.randn(1024, 256, 32, 32, generator=torch.Generator().manual_seed(0))
    books = {"normal": torch.randn(16384, 256, generator=torch.Generator().manual_seed(1)),
             "refinit": _refinit(16384, 256)}
    _sweep("C3(N=2^20,D=256,K=16384)", z, books, (0,), 8192)


def test_c5_one_step_against_the_oracle():
    z = torch.randn(56, 256, 32, 32, generator=torch.Generator().manual_seed(0))
    books = {"normal": torch.randn(65536, 256, generator=torch.Generator().manual_seed(1))}
    _sweep("C5(N=57344,D=256,K=65536)", z, books, (0,), 2048)
