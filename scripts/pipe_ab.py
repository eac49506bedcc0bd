"""A/B: pipelined 128-token forward tail / backward (vqb_tune tail_pipe / bwd_pipe) against the 32-token kernels.
Checks that the variants agree (z_q, dz bit-exact; loss, dE to summation-order tolerance) and prints times."""
import os
os.environ.setdefault("VQB200_EXPERIMENTAL", "1")  # route the ops through libvqb200_bench.so (vqb_tune, microbenchmarks)
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vq_gan_b200 import _cabi, ops
lib = _cabi.lib()
PEAK = 6530.0


def timed(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    b.synchronize()
    return a.elapsed_time(b) / n


for D, K, B in ((256, 16384, 1024), (128, 16384, 1024), (256, 128, 64), (192, 1000, 256)):
    torch.manual_seed(0)
    zs = [torch.randn(B, D, 32, 32, device="cuda") for _ in range(2)]
    gs = [torch.randn(B, D, 32, 32, device="cuda") for _ in range(2)]
    E = torch.randn(K, D, device="cuda")
    idx = torch.randint(0, K, (B, 32, 32), device="cuda") if K < 16384 else ops.search(zs[0], E)[0]
    one = torch.ones((), device="cuda")
    N = B * 1024
    ref = None
    for mode in (0, 1, 2):
        _cabi.check(lib.vqb_tune(b"tail_pipe", mode), "t")
        _cabi.check(lib.vqb_tune(b"bwd_pipe", mode), "t")
        loss = torch.empty(2, device="cuda")
        pb = lib.vqb_tail_partials_bytes(N)
        parts = torch.empty(pb, dtype=torch.uint8, device="cuda")
        zq = torch.empty_like(zs[0])
        it = [0]

        def fwd():
            i = it[0] = it[0] ^ 1
            _cabi.check(lib.vqb_gather_loss_st_f32(ops._p(zs[i]), ops._p(E), ops._p(idx), B, D, 1024, K, 0.25, ops._p(zq),
                                                   ops._p(loss), ops._p(parts), pb, None, ops._stream()), "tail")
        dz = torch.empty_like(zs[0])
        dE = torch.zeros_like(E)

        def bwd():
            i = it[0] = it[0] ^ 1
            _cabi.check(lib.vqb_backward_f32(ops._p(zs[i]), ops._p(E), ops._p(idx), ops._p(gs[i]), ops._p(one), 0.25, B, D, 1024,
                                             K, ops._p(dz), ops._p(dE), None, ops._stream()), "bwd")
        def bwd_nodE():
            i = it[0] = it[0] ^ 1
            _cabi.check(lib.vqb_backward_f32(ops._p(zs[i]), ops._p(E), ops._p(idx), ops._p(gs[i]), ops._p(one), 0.25, B, D, 1024,
                                             K, ops._p(dz), None, None, ops._stream()), "bwd")
        tf, tb, tn = timed(fwd), timed(bwd), timed(bwd_nodE)
        # deterministic comparison run on input set 0
        it[0] = 1
        fwd()
        dE.zero_()
        it[0] = 1
        bwd()
        cur = (zq.clone(), loss.clone(), dz.clone(), dE.clone())
        msg = ""
        if ref is None:
            ref = cur
        else:
            ok = torch.equal(cur[0], ref[0]) and torch.equal(cur[2], ref[2])
            lerr = float((cur[1] - ref[1]).abs().max() / ref[1].abs().max())
            derr = float((cur[3] - ref[3]).abs().max() / ref[3].abs().max())
            msg = f" zq/dz bit-equal={ok} loss rel {lerr:.1e} dE rel {derr:.1e}"
        fb, bb = N * (8 * D + 8), N * (12 * D + 8) + 4 * K * D
        print(f"D={D} K={K} N={N} mode={mode}: tail {tf:.3f} ms ({fb / tf / 1e6 / PEAK:.2f} of HBM peak)  "
              f"bwd {tb:.3f} ms ({bb / tb / 1e6 / PEAK:.2f}), without dE {tn:.3f} ms{msg}", flush=True)
_cabi.check(lib.vqb_tune(b"tail_pipe", 1), "t")
_cabi.check(lib.vqb_tune(b"bwd_pipe", 1), "t")
