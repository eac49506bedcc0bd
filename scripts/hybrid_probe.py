"""Feasibility probe: run the CUDA-core search (1 CTA/SM) and the tensor search concurrently on two streams."""
import os
os.environ.setdefault("VQB200_EXPERIMENTAL", "1")  # route the ops through libvqb200_bench.so (vqb_tune, microbenchmarks)
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vq_gan_b200 import _cabi, ops
lib = _cabi.lib()
torch.manual_seed(0)
B, D, K = 1024, 4, 16384
z = torch.randn(B, D, 32, 32, device="cuda")
E = torch.randn(K, D, device="cuda")
i_ref, _, _ = ops.search(z, E, 1)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
_cabi.check(lib.vqb_tune(b"lowd_ctas_per_sm", 1), "tune")
f = float(os.environ.get('FRAC', '0.5'))
Ba = int(B * f)
za, zb = z[:Ba].contiguous(), z[Ba:].contiguous()
for cl in (1, 2):
    _cabi.check(lib.vqb_tune(b"tclow_cluster", cl), "tune")
    for mode in ("fma", "tensor", "both", "both"):
        torch.cuda.synchronize()
        cur = torch.cuda.current_stream()
        ev = {k: torch.cuda.Event(enable_timing=True) for k in ("o", "a0", "a1", "b0", "b1", "end")}
        ev["o"].record()
        s1.wait_stream(cur); s2.wait_stream(cur)
        def run_fma():
            with torch.cuda.stream(s1):
                ev["a0"].record()
                ops.search(za, E, 1)
                ev["a1"].record()
        def run_tensor():
            with torch.cuda.stream(s2):
                ev["b0"].record()
                ops.search(zb, E, 5)
                ev["b1"].record()
        order = os.environ.get("ORDER", "fma_first")
        if mode == "fma":
            run_fma()
        elif mode == "tensor":
            run_tensor()
        elif order == "fma_first":
            run_fma(); run_tensor()
        else:
            run_tensor(); run_fma()
        cur.wait_stream(s1); cur.wait_stream(s2)
        ev["end"].record()
        torch.cuda.synchronize()
        o = ev["o"]
        msg = f"cluster={cl} mode={mode}: total {o.elapsed_time(ev['end']):.3f} ms"
        if mode in ("fma", "both"):
            msg += f" | fma [{o.elapsed_time(ev['a0']):.3f}, {o.elapsed_time(ev['a1']):.3f}]"
        if mode in ("tensor", "both"):
            msg += f" | tensor [{o.elapsed_time(ev['b0']):.3f}, {o.elapsed_time(ev['b1']):.3f}]"
        print(msg, flush=True)
