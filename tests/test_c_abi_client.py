"""A plain-C program drives libvqb200 through include/vqb200.h (no Python, no torch on that path)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "c_abi", "c_abi_smoke.c")
LIBDIR = os.path.join(ROOT, "vq_gan_b200", "lib")
CUDA_LIB = "/usr/local/cuda/lib64"


def _build(tmp_path):
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    if not os.path.exists(os.path.join(LIBDIR, "libvqb200.so")):
        from vq_gan_b200 import _build as b
        b.build()
    exe = str(tmp_path / "c_abi_smoke")
    cmd = ["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), SRC, "-o", exe,
           "-L", LIBDIR, "-lvqb200", "-L", CUDA_LIB, "-lcudart", "-lm",
           f"-Wl,-rpath,{LIBDIR}", f"-Wl,-rpath,{CUDA_LIB}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_c_client_compiles_links_and_passes_host_checks(tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert "host checks OK" in r.stdout


@pytest.mark.gpu
def test_c_client_runs_the_forward_path_on_the_gpu(tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([exe, "gpu"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert "mismatches 0" in r.stdout
