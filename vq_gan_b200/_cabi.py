"""ctypes binding of libvqb200.so (the C ABI declared in include/vqb200.h).

There is no CPU fallback: importing this module never fails, but every call goes
through `lib()`, which raises if the shared library is missing.
"""
import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int64, c_size_t, c_void_p

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "lib", "libvqb200.so")

VQB_OK = 0
ALGO_AUTO, ALGO_LOWD_FMA, ALGO_FP32_TILE, ALGO_TCGEN05, ALGO_TCGEN05_F16, ALGO_TCGEN05_TF32X3 = 0, 1, 2, 3, 4, 5
ALGO_DUAL_LOWD = 6
SEARCH_PRESPLIT = 0x100  # flag: the token split in the workspace was written by conv1x1_split
ALGO_NAMES = {ALGO_AUTO: "auto", ALGO_LOWD_FMA: "lowd_fma", ALGO_FP32_TILE: "fp32_tile",
              ALGO_TCGEN05: "tcgen05", ALGO_TCGEN05_F16: "tcgen05_f16",
              ALGO_TCGEN05_TF32X3: "tcgen05_tf32x3", ALGO_DUAL_LOWD: "dual_lowd_fma+tf32x3"}

# name -> (restype, argtypes); must list every symbol include/vqb200.h declares
PROTOTYPES = {
    "vqb_version": (c_int, []),
    "vqb_last_error": (c_char_p, []),
    "vqb_device_query": (c_int, [c_int, POINTER(c_int), POINTER(c_int), POINTER(c_int),
                                 POINTER(c_size_t)]),
    "vqb_codebook_pack_bytes": (c_size_t, [c_int, c_int]),
    "vqb_codebook_prepare_f32": (c_int, [c_void_p, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "vqb_search_workspace_bytes": (c_size_t, [c_int64, c_int, c_int64, c_int, c_int]),
    "vqb_search_f32": (c_int, [c_void_p, c_int64, c_int, c_int64, c_void_p, c_int, c_void_p,
                               c_void_p, c_void_p, c_void_p, c_size_t, c_int, c_void_p, c_void_p]),
    "vqb_tail_partials_bytes": (c_size_t, [c_int64]),
    "vqb_gather_loss_st_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int64, c_int,
                                       c_float, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p,
                                       c_void_p]),
    "vqb_backward_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_int64,
                                 c_int, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "vqb_gather_f32": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int64, c_int, c_void_p,
                               c_void_p, c_void_p]),
    "vqb_hist_i64": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "vqb_code_sums_f32": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int64, c_int, c_void_p,
                                  c_void_p, c_void_p]),
    "vqb_ema_update_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                   c_float, c_float, c_void_p, c_void_p]),
    "vqb_pack_argmin_keys": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_void_p]),
    "vqb_unpack_argmin_keys": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    "vqb_stats_pack": (c_int, [c_void_p, c_int64, c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p]),
    "vqb_stats_unpack": (c_int, [c_void_p, c_int64, c_int, c_int, c_float, c_void_p, c_void_p, c_void_p, c_void_p]),
    "vqb_index_bytes": (c_int, [c_int]),
    "vqb_indices_narrow": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_int, c_void_p, c_void_p]),
    "vqb_indices_widen": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_void_p]),
    "vqb_conv1x1_workspace_bytes": (c_size_t, [c_int, c_int]),
    "vqb_conv1x1_f32": (c_int, [c_void_p, c_int64, c_int, c_int64, c_void_p, c_void_p, c_int, c_void_p,
                                c_void_p, c_size_t, c_int, c_void_p]),
    "vqb_conv1x1_split_f32": (c_int, [c_void_p, c_int64, c_int, c_int64, c_void_p, c_void_p, c_int, c_void_p, c_void_p,
                                      c_size_t, c_void_p, c_int, c_void_p, c_size_t, c_void_p]),
    "vqb_conv1x1_dw_supported": (c_int, [c_int, c_int, c_int64]),
    "vqb_conv1x1_dw_f32": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_int64, c_void_p, c_void_p, c_void_p]),
    "vqb_groupnorm_silu_f32": (c_int, [c_void_p, c_int64, c_int, c_int64, c_void_p, c_void_p, c_int, c_float, c_void_p,
                                       c_void_p, c_void_p, c_void_p]),
    "vqb_groupnorm_silu_backward_f32": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int64, c_void_p, c_void_p, c_int,
                                                c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
}

# measurement build (include/vqb200_bench.h): everything above plus the experiment knobs / microbenchmarks
BENCH_LIB_PATH = os.path.join(_PKG, "lib", "libvqb200_bench.so")
BENCH_PROTOTYPES = {
    "vqb_tune": (c_int, [c_char_p, c_int]),
    "vqb_ubench_launch": (c_int, [c_int, c_int, c_void_p, c_void_p, POINTER(c_double), c_void_p]),
    "vqb_ubench_copy": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int64, c_int, c_void_p]),
    "vqb_fma_peak_launch": (c_int, [c_int, c_int, c_void_p, POINTER(c_double), c_void_p]),
    "vqb_ubench_red": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p]),
}

_lib = None
_bench_lib = None


class VqbError(RuntimeError):
    pass


def lib() -> ctypes.CDLL:
    """Loads libvqb200.so once.  Raises (never falls back) if it is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise VqbError(
                f"{LIB_PATH} not found: build it with `python -m vq_gan_b200._build` "
                "(nvcc, sm_100a).  vq_gan_b200 has no CPU or eager fallback.")
        if os.environ.get("VQB200_EXPERIMENTAL") == "1":
            # A/B scripts only: route the ops through the measurement build so vqb_tune() reaches them
            _lib = bench_lib()
            return _lib
        handle = ctypes.CDLL(LIB_PATH)
        _bind(handle, PROTOTYPES)
        _lib = handle
    return _lib


def _bind(handle, prototypes) -> None:
    for name, (restype, argtypes) in prototypes.items():
        fn = getattr(handle, name)  # AttributeError if the symbol is not exported
        fn.restype = restype
        fn.argtypes = argtypes


def bench_lib() -> ctypes.CDLL:
    """libvqb200_bench.so (measurement build: vqb_tune, microbenchmarks).  bench.py / scripts only."""
    global _bench_lib
    if _bench_lib is None:
        if not os.path.exists(BENCH_LIB_PATH):
            raise VqbError(f"{BENCH_LIB_PATH} not found: build it with `python -m vq_gan_b200._build`")
        handle = ctypes.CDLL(BENCH_LIB_PATH)
        _bind(handle, PROTOTYPES)
        _bind(handle, BENCH_PROTOTYPES)
        _bench_lib = handle
    return _bench_lib


def check(rc: int, what: str) -> None:
    if rc != VQB_OK:
        msg = lib().vqb_last_error()
        raise VqbError(f"{what} failed (code {rc}): {msg.decode() if msg else '?'}")
