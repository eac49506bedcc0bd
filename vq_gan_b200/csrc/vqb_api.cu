// C ABI: library plumbing, codebook pre-pass and search dispatch.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "vqb_common.cuh"

namespace vqb {

static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
    set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
    return VQB_ERR_CUDA;
}

// immutable per-device capability cache (written once per device; racing writers store the same value)
static int g_sm_cached[64];

int device_ready() {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice");
    if (dev < 0 || dev >= 64) {
        set_error("device ordinal %d outside the capability cache (0..63)", dev);
        return VQB_ERR_UNSUPPORTED;
    }
    if (g_sm_cached[dev] == 0) {
        int n = 0, major = 0;
        e = cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess) return cuda_fail(e, "cudaDeviceGetAttribute(MultiProcessorCount)");
        e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
        if (e != cudaSuccess) return cuda_fail(e, "cudaDeviceGetAttribute(ComputeCapabilityMajor)");
        if (n <= 0 || major != 10) {
            set_error("libvqb200 needs an sm_100a device (B200); device %d reports %d SMs, compute capability %d.x", dev, n,
                      major);
            return VQB_ERR_UNSUPPORTED;
        }
        g_sm_cached[dev] = n;
    }
    return VQB_OK;
}

int sm_count() {
    // every launching entry point runs VQB_DEVICE_TRY() first, so the cache is filled here
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 0;
    return g_sm_cached[dev];
}

}  // namespace vqb

using namespace vqb;

extern "C" int vqb_version(void) { return 100; }  // 0.1.0

extern "C" const char* vqb_last_error(void) { return g_error; }

extern "C" int vqb_device_query(int device, int* sm, int* cc_major, int* cc_minor, size_t* smem_optin) {
    int v = 0;
    if (sm) {
        VQB_CUDA_TRY(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device));
        *sm = v;
    }
    if (cc_major) {
        VQB_CUDA_TRY(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMajor, device));
        *cc_major = v;
    }
    if (cc_minor) {
        VQB_CUDA_TRY(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMinor, device));
        *cc_minor = v;
    }
    if (smem_optin) {
        VQB_CUDA_TRY(cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
        *smem_optin = (size_t)v;
    }
    return VQB_OK;
}

#ifdef VQB_EXPERIMENTAL
// measurement build only (libvqb200_bench.so, include/vqb200_bench.h): process-global, not thread-safe
extern "C" int vqb_tune(const char* key, int value) {
    if (!key) {
        set_error("vqb_tune: null key");
        return VQB_ERR_INVALID_ARG;
    }
    if (strcmp(key, "lowd_variant") == 0 && value >= 0 && value <= 4) {
        set_lowd_variant(value);
        return VQB_OK;
    }
    if (strcmp(key, "lowd_ctas_per_sm") == 0 && value >= 0 && value <= 3) {
        set_lowd_variant(16 + value);
        return VQB_OK;
    }
    if (strcmp(key, "dual_permille") == 0 && value >= 0 && value <= 999) {  // 0 = the built-in model
        set_dual_permille(value);
        return VQB_OK;
    }
    if (strcmp(key, "tc16_branchy") == 0 && (value == 0 || value == 1)) {
        set_tc16_branchy(value);
        return VQB_OK;
    }
    if (strcmp(key, "dw_hw_trunc") == 0 && (value == 0 || value == 1)) {
        set_dw_hw_trunc(value);
        return VQB_OK;
    }
    if (strcmp(key, "norm_fwd_reg") == 0 && (value == 0 || value == 1)) {
        set_norm_fwd_reg(value);
        return VQB_OK;
    }
    if (strcmp(key, "norm_bwd2") == 0 && (value == 0 || value == 1)) {
        set_norm_bwd2(value);
        return VQB_OK;
    }
    if (strcmp(key, "norm_cluster") == 0 && value >= 0 && value <= 2) {
        set_norm_cluster(value);
        return VQB_OK;
    }
    if (strcmp(key, "tc16_pruned") == 0 && (value == 0 || value == 1)) {
        set_tc16_pruned(value);
        return VQB_OK;
    }
    if (strcmp(key, "tc16_group") == 0 && (value == 0 || value == 4 || value == 8)) {
        set_tc16_group(value);
        return VQB_OK;
    }
    if (strcmp(key, "tc16_cluster") == 0 && (value == 1 || value == 2 || value == 4)) {
        set_tc16_cluster(value);
        return VQB_OK;
    }
    if (strcmp(key, "tclow_cluster") == 0 && (value == 1 || value == 2 || value == 4)) {
        set_tclow_cluster(value);
        return VQB_OK;
    }
    if (strcmp(key, "tclow_skip_stages") == 0 && value >= 0 && value < 8) {
        set_tclow_cluster(16 + value);
        return VQB_OK;
    }
    if ((strcmp(key, "bwd_pass_channels") == 0 || strcmp(key, "fwd_pass_channels") == 0) &&
        (value == 64 || value == 128 || value == 192 || value == 256)) {
        set_tail_knob(key, value);
        return VQB_OK;
    }
    if ((strcmp(key, "bwd_warp") == 0 || strcmp(key, "tail_warp") == 0) && value >= 0 && value <= 2) {
        set_tail_knob(key, value);
        return VQB_OK;
    }
    if ((strcmp(key, "tail_pipe") == 0 || strcmp(key, "bwd_pipe") == 0) && value >= 0 && value <= 2) {
        set_tail_knob(key, value);
        return VQB_OK;
    }
    if (strcmp(key, "tail_tok128") == 0 && (value == 0 || value == 1)) {
        set_tail_knob(key, value);
        return VQB_OK;
    }
    if (strcmp(key, "conv_debug") == 0 && value >= 0 && value < 16) {
        set_conv_debug(value);
        return VQB_OK;
    }
    set_error("vqb_tune: unknown key or value (%s = %d)", key, value);
    return VQB_ERR_INVALID_ARG;
}
#endif

extern "C" size_t vqb_codebook_pack_bytes(int K, int D) {
    if (K <= 0 || D <= 0) return 0;
    return pack_layout(K, D).total;
}

extern "C" int vqb_codebook_prepare_f32(const float* E, int K, int D, void* pack, size_t pack_bytes,
                                        vqb_stream_t stream) {
    VQB_DEVICE_TRY();
    if (!E || !pack || K <= 0 || D <= 0) {
        set_error("vqb_codebook_prepare_f32: invalid argument (K=%d D=%d)", K, D);
        return VQB_ERR_INVALID_ARG;
    }
    if ((reinterpret_cast<uintptr_t>(pack) & 255u) != 0) {
        set_error("vqb_codebook_prepare_f32: pack must be 256-byte aligned");
        return VQB_ERR_INVALID_ARG;
    }
    const size_t need = pack_layout(K, D).total;
    if (pack_bytes < need) {
        set_error("codebook pack too small: %zu < %zu", pack_bytes, need);
        return VQB_ERR_WORKSPACE;
    }
    return launch_codebook_prepare(E, K, D, pack, static_cast<cudaStream_t>(stream));
}

// Measured on B200 (profiles/r01_tclow_vs_fma.txt): the CUDA-core kernel costs ~0.72 ms per 1M tokens x
// 16384 codes x dimension; the tf32x3 tensor kernel is bound by TMEM traffic at ~2.5-5 ms per 1M x 16384
// for any D <= 16, with a higher fixed cost.  So: FMA for D <= 4 and for small problems, tensor above.
static int resolve_algo(int algo, int D, int64_t N = 0, int K = 0, int64_t B = 0) {
    if (algo != VQB_ALGO_AUTO) return algo;
    if (D <= kLowDMax) {
        const bool large = (double)N * (double)K >= 268435456.0;  // 2^28 scores
        // enough images to split (D = 4 is config C2): both engines in one CTA, 2.32 ms against 2.94 (CUDA cores) and
        // 2.82 (tensor cores) per 1M tokens x 16384 codes at D = 4, identical results
        // (D <= 8 only: above that the 8-warp FMA role is too slow to be worth its share of the SM -- D = 12: 4.53 ms
        // against 4.41 for the tensor kernel alone, profiles/r02_dual_dims.txt)
        if (dual_eligible(B, D) && D <= 8 && B >= 16 && (double)N * (double)K >= 1073741824.0) return VQB_ALGO_DUAL_LOWD;
        return (D >= 5 && large) ? VQB_ALGO_TCGEN05_TF32X3 : VQB_ALGO_LOWD_FMA;
    }
    if (tc16_eligible_dim(D)) {
        // below ~1 GFLOP the five-launch tensor pipeline is latency-bound (80 us floor): one fp32 tile kernel wins
        // (reference default shape: 4096 tokens x 128 codes x 256 dims = 0.27 GFLOP)
        if (N > 0 && (double)N * (double)K * (double)D < 536870912.0) return VQB_ALGO_FP32_TILE;
        return VQB_ALGO_TCGEN05_F16;  // single fp16 pass + exact re-score: 2.4x the bf16x3 kernel
    }
    return VQB_ALGO_FP32_TILE;
}

extern "C" size_t vqb_search_workspace_bytes(int64_t B, int D, int64_t HW, int K, int algo) {
    if (B < 0 || HW < 0 || D <= 0 || K <= 0) return 0;
    algo &= ~VQB_SEARCH_PRESPLIT;
    const int a = resolve_algo(algo, D, B * HW, K, B);
    if (a == VQB_ALGO_TCGEN05) return search_tc_workspace_bytes(B * HW, D, K);
    if (a == VQB_ALGO_TCGEN05_F16) return search_tc16_workspace_bytes(B * HW, D, K);
    if (a == VQB_ALGO_TCGEN05_TF32X3) return search_tclow_workspace_bytes(B * HW, D, K);
    if (a == VQB_ALGO_DUAL_LOWD) return search_dual_workspace_bytes(B, D, HW, K);
    if (a == VQB_ALGO_FP32_TILE) return search_fp32_workspace_bytes(B * HW);
    return 0;
}

__global__ void write_stats_kernel(int64_t* stats, int64_t rescored, int64_t algo) {
    stats[0] = rescored;
    stats[1] = algo;
    stats[2] = 0;
    stats[3] = 0;
}

extern "C" int vqb_search_f32(const float* z, int64_t B, int D, int64_t HW, const float* E, int K,
                              const void* pack, int64_t* idx_out, float* dmin_out, void* workspace,
                              size_t workspace_bytes, int algo, int64_t* stats_out, vqb_stream_t stream) {
    VQB_DEVICE_TRY();
    if (B < 0 || HW < 0 || D <= 0 || K <= 0) {
        set_error("vqb_search_f32: invalid shape B=%lld D=%d HW=%lld K=%d", (long long)B, D, (long long)HW, K);
        return VQB_ERR_INVALID_ARG;
    }
    const int64_t N = B * HW;
    if (N == 0) return VQB_OK;
    if (!z || !E || !pack || !idx_out) {
        set_error("vqb_search_f32: null pointer");
        return VQB_ERR_INVALID_ARG;
    }
    if (N >= (1LL << 31)) {
        set_error("vqb_search_f32: at most 2^31-1 tokens per call, got %lld", (long long)N);
        return VQB_ERR_INVALID_ARG;
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const bool presplit = (algo & VQB_SEARCH_PRESPLIT) != 0;
    algo &= ~VQB_SEARCH_PRESPLIT;
    if (presplit && algo != VQB_ALGO_TCGEN05_F16) {
        set_error("VQB_SEARCH_PRESPLIT goes with VQB_ALGO_TCGEN05_F16 only (the split lives in that kernel's workspace)");
        return VQB_ERR_INVALID_ARG;
    }
    const int a = resolve_algo(algo, D, N, K, B);
    int rc;
    switch (a) {
        case VQB_ALGO_LOWD_FMA:
            if (D > kLowDMax) {
                set_error("VQB_ALGO_LOWD_FMA needs D <= %d, got %d", kLowDMax, D);
                return VQB_ERR_UNSUPPORTED;
            }
            rc = launch_search_lowd(z, B, D, HW, K, pack, idx_out, dmin_out, s);
            break;
        case VQB_ALGO_FP32_TILE:
            rc = launch_search_fp32(z, B, D, HW, E, K, pack, nullptr, nullptr, 0, workspace, workspace_bytes,
                                    idx_out, dmin_out, s);
            break;
        case VQB_ALGO_TCGEN05:
            if (!tc_eligible_dim(D)) {
                set_error("VQB_ALGO_TCGEN05 needs D %% 64 == 0 and %d <= D <= %d, got %d", kTcMinD, kTcMaxD, D);
                return VQB_ERR_UNSUPPORTED;
            }
            // writes its own stats (re-scored token count is only known on the device)
            return launch_search_tc(z, B, D, HW, E, K, pack, idx_out, dmin_out, workspace, workspace_bytes,
                                    stats_out, s);
        case VQB_ALGO_TCGEN05_F16:
            if (!tc16_eligible_dim(D)) {
                set_error("VQB_ALGO_TCGEN05_F16 needs %d < D <= %d, got %d", kLowDMax, kTc16MaxD, D);
                return VQB_ERR_UNSUPPORTED;
            }
            return launch_search_tc16(z, B, D, HW, E, K, pack, idx_out, dmin_out, workspace, workspace_bytes,
                                      stats_out, s, presplit);
        case VQB_ALGO_TCGEN05_TF32X3:
            if (D > kLowDMax) {
                set_error("VQB_ALGO_TCGEN05_TF32X3 needs D <= %d, got %d", kLowDMax, D);
                return VQB_ERR_UNSUPPORTED;
            }
            return launch_search_tclow(z, B, D, HW, E, K, pack, idx_out, dmin_out, workspace, workspace_bytes,
                                       stats_out, s);
        case VQB_ALGO_DUAL_LOWD:
            if (!dual_eligible(B, D)) {
                set_error("VQB_ALGO_DUAL_LOWD needs %d <= D <= %d and at least 2 images, got D=%d B=%lld", kDualMinD, kLowDMax, D,
                          (long long)B);
                return VQB_ERR_UNSUPPORTED;
            }
            return launch_search_dual(z, B, D, HW, E, K, pack, idx_out, dmin_out, workspace, workspace_bytes, stats_out, s);
        default:
            set_error("vqb_search_f32: unknown algo %d", algo);
            return VQB_ERR_INVALID_ARG;
    }
    if (rc != VQB_OK) return rc;
    if (stats_out) {
        write_stats_kernel<<<1, 1, 0, s>>>(stats_out, 0, a);
        VQB_LAUNCH_CHECK("write_stats_kernel");
    }
    return VQB_OK;
}
