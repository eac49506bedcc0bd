"""Seeded parity cases shared by the golden generator, the CPU tests and the
GPU tests.  Shapes follow SURVEY.md section 8(d): C1 is the reference's own
default (`train_vqgan.py` config: K=128, D=256, 32x32 latents, batch 4); the
C2/C3 slices are token slices of the 1M-token configurations that the oracle
finishes in seconds; the rest are edge cases (ragged sizes, exact ties, D=1).
"""
import torch

# codebook kinds: "normal" N(0,1); "refinit" = the reference constructor's
# U(-1/K, 1/K) under torch.manual_seed(seed_E); "trained" = tokens + noise;
# "dup" = second half duplicates the first half (exact distance ties)
CASES = {
    "small_d4":     dict(B=2, D=4,   H=8,  W=8,  K=64,    code="normal",  beta=0.25, store_full=True),
    "small_d1":     dict(B=2, D=1,   H=4,  W=6,  K=16,    code="normal",  beta=0.25, store_full=True),
    "ragged_d3":    dict(B=3, D=3,   H=5,  W=7,  K=37,    code="normal",  beta=0.5,  store_full=True),
    "small_d8":     dict(B=2, D=8,   H=9,  W=9,  K=200,   code="normal",  beta=0.25, store_full=True),
    "small_d16":    dict(B=1, D=16,  H=16, W=16, K=300,   code="normal",  beta=0.25, store_full=True),
    "mid_d32":      dict(B=2, D=32,  H=8,  W=8,  K=150,   code="normal",  beta=0.25, store_full=True),
    "mid_d48":      dict(B=1, D=48,  H=7,  W=9,  K=130,   code="normal",  beta=1.0,  store_full=True),
    "tc_d64":       dict(B=2, D=64,  H=16, W=16, K=256,   code="normal",  beta=0.25, store_full=True),
    "tc_d128_rag":  dict(B=3, D=128, H=10, W=10, K=1000,  code="normal",  beta=0.25, store_full=False),
    "ties_d4":      dict(B=1, D=4,   H=8,  W=8,  K=128,   code="dup",     beta=0.25, store_full=True),
    "ties_d64":     dict(B=1, D=64,  H=8,  W=8,  K=256,   code="dup",     beta=0.25, store_full=False),
    "c1_refinit":   dict(B=4, D=256, H=32, W=32, K=128,   code="refinit", beta=0.25, store_full=False),
    "c1_trained":   dict(B=4, D=256, H=32, W=32, K=128,   code="trained", beta=0.25, store_full=False),
    "c2_slice":     dict(B=16, D=4,  H=32, W=32, K=16384, code="normal",  beta=0.25, store_full=False),
    "c2_refinit":   dict(B=4, D=4,   H=32, W=32, K=16384, code="refinit", beta=0.25, store_full=False),
    "c3_slice":     dict(B=4, D=256, H=32, W=32, K=16384, code="normal",  beta=0.25, store_full=False),
}

SEED_Z, SEED_E, SEED_G = 0, 1, 2


def _randn(shape, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g, dtype=torch.float32)


def make_case(name):
    c = CASES[name]
    B, D, H, W, K = c["B"], c["D"], c["H"], c["W"], c["K"]
    z = _randn((B, D, H, W), SEED_Z)
    g_zq = _randn((B, D, H, W), SEED_G)
    kind = c["code"]
    if kind == "normal":
        E = _randn((K, D), SEED_E)
    elif kind == "refinit":
        # same draws as the reference constructor under manual_seed(42)
        state = torch.get_rng_state()
        torch.manual_seed(42)
        emb = torch.nn.Embedding(K, D)
        emb.weight.data.uniform_(-1.0 / K, 1.0 / K)
        E = emb.weight.detach().clone()
        torch.set_rng_state(state)
    elif kind == "trained":
        rows = z.permute(0, 2, 3, 1).reshape(-1, D)
        perm = torch.randperm(rows.shape[0], generator=torch.Generator().manual_seed(SEED_E))[:K]
        E = rows[perm] + 0.01 * _randn((K, D), SEED_E + 10)
    elif kind == "dup":
        half = _randn((K // 2, D), SEED_E)
        E = torch.cat([half, half], dim=0)
    else:
        raise ValueError(kind)
    return {"z": z, "E": E.contiguous(), "g_zq": g_zq, "beta": c["beta"], "spec": c}
