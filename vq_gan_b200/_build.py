"""Build recipe for libvqb200.so and libvqb200_bench.so (nvcc, sm_100a only, in-tree output).

    python -m vq_gan_b200._build [--force]

Two shared libraries are written to vq_gan_b200/lib/ so that they travel with the repo
snapshot to the GPU box (git-ignored):
  libvqb200.so        the product: the C ABI of include/vqb200.h, no mutable global state
  libvqb200_bench.so  the same sources + csrc/vqb_ubench.cu compiled with -DVQB_EXPERIMENTAL:
                      adds vqb_tune / microbenchmarks (include/vqb200_bench.h); only bench.py
                      and scripts/ load it
"""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIBDIR = os.path.join(PKG, "lib")
OBJDIR = os.path.join(PKG, "build")
LIB = os.path.join(LIBDIR, "libvqb200.so")
BENCH_LIB = os.path.join(LIBDIR, "libvqb200_bench.so")
SOURCES = ["vqb_api.cu", "vqb_prepare.cu", "vqb_search_lowd.cu", "vqb_search_fp32.cu",
           "vqb_search_tc.cu", "vqb_search_tc16.cu", "vqb_search_pruned.cu", "vqb_search_tclow.cu", "vqb_tail.cu", "vqb_indexio.cu", "vqb_stats.cu",
           "vqb_conv1x1.cu", "vqb_conv1x1_dw.cu", "vqb_norm.cu"]
BENCH_ONLY_SOURCES = ["vqb_ubench.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr"]


def _digest():
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(os.path.dirname(PKG), "include")):
        for name in sorted(os.listdir(root)):
            with open(os.path.join(root, name), "rb") as f:
                h.update(name.encode())
                h.update(f.read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    stamp = os.path.join(LIBDIR, "libvqb200.sha256")
    digest = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(BENCH_LIB) and os.path.exists(stamp) and open(stamp).read() == digest:
        return LIB
    if not os.path.exists(NVCC):
        raise RuntimeError(f"nvcc not found at {NVCC}; libvqb200.so must be prebuilt")

    def compile_one(job):
        src, experimental = job
        obj = os.path.join(OBJDIR, src.replace(".cu", ".x.o" if experimental else ".o"))
        cmd = [NVCC, *FLAGS, *(["-DVQB_EXPERIMENTAL"] if experimental else []), "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(r.stderr)
        return obj

    jobs = [(src, False) for src in SOURCES] + [(src, True) for src in SOURCES + BENCH_ONLY_SOURCES]
    with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
        objs = list(ex.map(compile_one, jobs))
    for lib_path, members in ((LIB, objs[:len(SOURCES)]), (BENCH_LIB, objs[len(SOURCES):])):
        cmd = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib_path, *members]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as f:
        f.write(digest)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
