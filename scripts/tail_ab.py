"""A/B of the forward-tail / backward kernels: 32-token tiles vs 128-token float4 tiles (vqb_tune tail_tok128)."""
import os
os.environ.setdefault("VQB200_EXPERIMENTAL", "1")  # route the ops through libvqb200_bench.so (vqb_tune, microbenchmarks)
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vq_gan_b200 import VectorQuantizer, ops, _cabi
lib = _cabi.lib()
for D, K, B in ((4, 16384, 1024), (8, 4096, 1024), (64, 4096, 1024), (32, 1024, 512)):
    vq = VectorQuantizer(K, D, lazy_stats=True).cuda()
    with torch.no_grad():
        vq.embedding.weight.copy_(torch.randn(K, D))
    z = torch.randn(B, D, 32, 32, device="cuda", requires_grad=True)
    g = torch.randn(B, D, 32, 32, device="cuda")
    one = torch.ones((), device="cuda")
    n = B * 1024
    outs = {}
    for mode in (0, 1):
        _cabi.check(lib.vqb_tune(b"tail_tok128", mode), "t")
        bt, bb = 1e9, 1e9
        for it in range(3):
            ops.PROFILE_TAIL, ops.PROFILE_BWD = [], []
            z.grad = None
            vq.embedding.weight.grad = None
            zq, ld, idx = vq(z)
            torch.autograd.backward((zq, ld["vq_loss"]), (g, one))
            torch.cuda.synchronize()
            bt = min(bt, ops.PROFILE_TAIL[0][0].elapsed_time(ops.PROFILE_TAIL[0][1]))
            bb = min(bb, ops.PROFILE_BWD[0][0].elapsed_time(ops.PROFILE_BWD[0][1]))
        outs[mode] = (zq.detach().clone(), z.grad.clone(), vq.embedding.weight.grad.clone(), ld["vq_loss"].detach().clone())
        print(f"D={D} tok128={mode}: tail {bt:.3f} ms ({n*(8*D+8)/bt/1e6:.0f} GB/s)  backward {bb:.3f} ms "
              f"({(n*(12*D+8)+4*K*D)/bb/1e6:.0f} GB/s)", flush=True)
    a, b = outs[0], outs[1]
    print("   same z_q:", torch.equal(a[0], b[0]), " same dz:", torch.equal(a[1], b[1]), " dE close:",
          torch.allclose(a[2], b[2], rtol=1e-4, atol=1e-7), " loss rel diff:", float((a[3] - b[3]).abs() / a[3].abs()))
