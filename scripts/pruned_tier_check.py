"""Pruned exact tier of the fp16 tensor search (vqb_search_pruned.cu): collapsed / duplicated codebooks with the tier on
and off (vqb_tune tc16_pruned) -- indices and minimum scores must be identical, the time should not be."""
import os
os.environ.setdefault("VQB200_EXPERIMENTAL", "1")
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vq_gan_b200 import _cabi, ops
lib = _cabi.lib()


def clustered(K, D, n_centres, spread, tokens, tok_sigma, seed=1):
    g = torch.Generator().manual_seed(seed)
    centres = torch.randn(n_centres, D, generator=g)
    E = centres[torch.randint(0, n_centres, (K,), generator=g)] + spread * torch.randn(K, D, generator=g)
    pick = torch.randint(0, n_centres, (tokens,), generator=g)
    z = centres[pick] + tok_sigma * torch.randn(tokens, D, generator=g)
    return z, E


def run(name, z_rows, E, HW=1024):
    N, D = z_rows.shape
    B = N // HW
    z = z_rows[:B * HW].view(B, HW, D).permute(0, 2, 1).contiguous().view(B, D, HW, 1).cuda()
    E = E.cuda()
    out = {}
    for mode in (1, 0):
        _cabi.check(lib.vqb_tune(b"tc16_pruned", mode), "t")
        best = 1e9
        for _ in range(3):
            ops.PROFILE = []
            idx, dmin, st = ops.search(z, E, 4)
            torch.cuda.synchronize()
            (s, e), = ops.PROFILE
            best = min(best, s.elapsed_time(e))
        ops.PROFILE = None
        out[mode] = (idx, dmin, st.tolist(), best)
    same_i = torch.equal(out[0][0], out[1][0])
    same_d = torch.equal(out[0][1], out[1][1])
    print(f"{name}: tokens={B * HW} K={E.shape[0]} D={D}  tier on {out[1][3]:8.3f} ms stats={out[1][2]}  |  tier off {out[0][3]:8.3f} ms "
          f"stats={out[0][2]}  | same idx {same_i} same dmin {same_d}", flush=True)
    _cabi.check(lib.vqb_tune(b"tc16_pruned", 1), "t")
    return same_i and same_d


ok = True
n = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
z, E = clustered(16384, 256, 16, 1e-4, n, 0.05)
ok &= run("collapsed 16 centres", z, E)
z, E = clustered(5000, 100, 7, 1e-4, n, 0.05, seed=3)
ok &= run("collapsed 7 centres, K=5000 D=100", z, E)
z, E = clustered(16384, 256, 1, 1e-4, n, 0.05)
ok &= run("ONE centre (nothing to prune: the tier declines)", z, E)
# a healthy codebook whose every code appears four times (1e-5 apart)
g = torch.Generator().manual_seed(5)
base = torch.randn(4096, 256, generator=g)
E = base.repeat_interleave(4, dim=0)[torch.randperm(16384, generator=g)] + 1e-5 * torch.randn(16384, 256, generator=g)
z = torch.randn(n, 256, generator=g)
ok &= run("every code duplicated 4x", z, E)
# the same with EXACT copies (a codebook restarted by copying): the prepare pass hides the later copies, the tokens certify
Ex = base.repeat_interleave(4, dim=0)[torch.randperm(16384, generator=g)].contiguous()
ok &= run("every code copied 4x exactly", z, Ex)
Ex = base[:256].repeat_interleave(64, dim=0)[torch.randperm(16384, generator=g)].contiguous()
ok &= run("256 distinct codes copied 64x exactly", z, Ex)
z = torch.randn(n, 256, generator=g)
z[::97] = float("nan")
z[5::1013] = float("inf")
zc, Ec = clustered(16384, 256, 16, 1e-4, n, 0.05)
zc[::97] = float("nan")
zc[5::1013] = float("inf")
ok &= run("collapsed + NaN / inf tokens", zc, Ec)
# healthy data: the tier's kernels are launched and return at once -- the difference is their launch overhead
g = torch.Generator().manual_seed(9)
ok &= run("N(0,1) codebook (tier idle)", torch.randn(n, 256, generator=g), torch.randn(16384, 256, generator=g))
ok &= run("N(0,1) codebook D=64 (tier idle)", torch.randn(n, 64, generator=g), torch.randn(16384, 64, generator=g))
if n >= 1 << 20:
    z, E = clustered(16384, 256, 16, 1e-4, n, 0.05)
    ok &= run("collapsed 16 centres, full size", z, E)
print("pruned tier check:", "ok" if ok else "FAILED")
sys.exit(0 if ok else 1)
