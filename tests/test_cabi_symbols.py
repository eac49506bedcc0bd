"""The C-ABI library loads on a CPU-only host and exports every symbol that
include/vqb200.h declares (no compute calls here)."""
import ctypes
import os
import re

import pytest

import vq_gan_b200
from vq_gan_b200 import _cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols(header="vqb200.h"):
    text = open(os.path.join(ROOT, "include", header)).read()
    return sorted(set(re.findall(r"VQB_API\s+[\w\s\*]+?\b(vqb_\w+)\s*\(", text)))


def test_header_declares_expected_surface():
    names = declared_symbols()
    for must in ("vqb_search_f32", "vqb_gather_loss_st_f32", "vqb_backward_f32", "vqb_hist_i64",
                 "vqb_gather_f32", "vqb_codebook_prepare_f32", "vqb_last_error", "vqb_version"):
        assert must in names
    assert len(names) >= 17


def test_library_exports_every_declared_symbol():
    handle = ctypes.CDLL(_cabi.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(handle, name), f"{name} declared in vqb200.h but not exported"


def test_python_prototypes_cover_header():
    assert sorted(_cabi.PROTOTYPES) == declared_symbols()


def test_product_library_has_no_experiment_knobs_or_microbenchmarks():
    """vqb_tune / vqb_ubench_* / vqb_fma_peak_launch live in the measurement build only
    (include/vqb200_bench.h, libvqb200_bench.so); the product ABI has no process-global knobs."""
    handle = ctypes.CDLL(_cabi.LIB_PATH)
    bench_only = declared_symbols("vqb200_bench.h")
    assert sorted(bench_only) == sorted(_cabi.BENCH_PROTOTYPES)
    assert "vqb_tune" in bench_only and "vqb_fma_peak_launch" in bench_only
    for name in bench_only:
        assert not hasattr(handle, name), f"{name} must not be exported by libvqb200.so"
    bench = ctypes.CDLL(_cabi.BENCH_LIB_PATH)
    for name in bench_only + declared_symbols():
        assert hasattr(bench, name), f"{name} missing from libvqb200_bench.so"


def test_version_and_size_queries_without_gpu():
    lib = vq_gan_b200.lib()
    assert lib.vqb_version() >= 100
    # pure host arithmetic, no CUDA call
    small = lib.vqb_codebook_pack_bytes(128, 256)
    big = lib.vqb_codebook_pack_bytes(16384, 256)
    assert 0 < small < big
    assert lib.vqb_codebook_pack_bytes(0, 4) == 0
    assert lib.vqb_tail_partials_bytes(1 << 20) >= 8 * (1 << 15)
    assert lib.vqb_search_workspace_bytes(4, 4, 64, 128, 0) == 0
    assert lib.vqb_search_workspace_bytes(4, 32, 64, 128, 0) >= 8 * 256


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(_cabi, "_lib", None)
    monkeypatch.setattr(_cabi, "LIB_PATH", "/nonexistent/libvqb200.so")
    with pytest.raises(_cabi.VqbError, match="no CPU or eager fallback"):
        _cabi.lib()


def test_round2_host_side_queries_without_gpu():
    """Pure host arithmetic of the entry points added in round 2: which shapes the 1x1 convolution's parameter-gradient
    kernel takes, and that the search workspace / codebook pack grow by the pruned exact tier / the duplicate table
    exactly where those apply."""
    lib = vq_gan_b200.lib()
    assert lib.vqb_conv1x1_dw_supported(256, 256, 1024) == 1
    assert lib.vqb_conv1x1_dw_supported(16, 1, 32) == 1
    assert lib.vqb_conv1x1_dw_supported(24, 64, 1024) == 0      # Cin % 16
    assert lib.vqb_conv1x1_dw_supported(64, 257, 1024) == 0     # Cout > 256
    assert lib.vqb_conv1x1_dw_supported(64, 64, 1022) == 0      # HW % 4
    assert lib.vqb_conv1x1_dw_supported(64, 64, 16) == 0        # HW < 32
    f16 = _cabi.ALGO_TCGEN05_F16
    # pruned tier: 2048 <= K <= 16384 only
    small = lib.vqb_search_workspace_bytes(64, 256, 1024, 1024, f16)
    mid = lib.vqb_search_workspace_bytes(64, 256, 1024, 4096, f16)
    big = lib.vqb_search_workspace_bytes(64, 256, 1024, 32768, f16)
    assert mid > small and mid > big and small == big
    # duplicate table only with the fp16 image (D > 16)
    low = lib.vqb_codebook_pack_bytes(4096, 4)
    high = lib.vqb_codebook_pack_bytes(4096, 64)
    assert high - low > 8 * 2 * 4096          # at least the 2 * Kpad 64-bit slots
