"""Pretty-prints a bench.py JSON line (default workload: c2 + c3 / c4 / c5 blocks)."""
import json
import sys


def show(b, name):
    print(f"== {name}: value {b['value'] / 1e6:.1f} M tok/s, {b['ms_per_step']:.3f} ms/step; e2e {b['e2e']['value'] / 1e6:.1f} M "
          f"({b['e2e']['ms_per_step']:.3f} ms)")
    r = b["roofline"]
    print(f"   roofline {r['kernel']} {r['achieved']:.1f}/{r['peak']:.1f} {r['unit']} frac {r['frac']:.3f} kernel_ms {r['kernel_ms']:.3f} "
          f"share {r['step_share']}")
    for k, v in b.get("hbm_kernels", {}).items():
        print(f"    {k} {v['kernel_ms']:.3f} ms frac {v['frac']:.3f}")
    print("   clocks", b["clocks"], "launches", b["gpu_launches"])
    print("   parity", json.dumps(b["parity_check"])[:700])
    if "cpu_baseline" in b:
        print("   cpu", b["cpu_baseline"])
    print("   details", b.get("details"))


d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
show(d, "top")
for k in ("c3", "c5"):
    if k in d:
        show(d[k], k)
if "c3" in d and d["c3"].get("codebooks"):
    for k, v in d["c3"]["codebooks"].items():
        print("   c3 codebook", k, v)
if d.get("codebooks"):
    for k, v in d["codebooks"].items():
        print("   c2 codebook", k, v)
if "c4" in d:
    print("== c4:", json.dumps(d["c4"])[:900])
if "next_rows" in d:
    print("== next rows (N1 / N2):", json.dumps(d["next_rows"])[:2000])
