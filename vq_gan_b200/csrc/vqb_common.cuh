// Shared declarations for libvqb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>

#include "../../include/vqb200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libvqb200 is written for sm_100a (B200) only"
#endif

namespace vqb {

// ---- error plumbing (thread-local message, never throws) -----------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
#define VQB_CUDA_TRY(expr)                                      \
    do {                                                        \
        cudaError_t _e = (expr);                                \
        if (_e != cudaSuccess) return ::vqb::cuda_fail(_e, #expr); \
    } while (0)
#define VQB_LAUNCH_CHECK(what)                                  \
    do {                                                        \
        cudaError_t _e = cudaGetLastError();                    \
        if (_e != cudaSuccess) return ::vqb::cuda_fail(_e, what); \
    } while (0)

int sm_count();  // cached per current device

// ---- packed codebook layout ----------------------------------------------
// header (int32[64]): [0] first NaN code (K if none) [1] K [2] D [4] float bits of max 0.5|e|^2
//                      [5] fp16 scale exponent se (e16 = fp16(E * 2^se)) [6] float bits of max|E| [7] bits of max |e - fp16(e)|
// half_norm: float[Kpad]  0.5|e_k|^2, +inf for k >= K.  Read as consecutive
//            (h[2p], h[2p+1]) pairs by the low-D kernel.
// pairs    : float[Kpad/2][2D]  (D <= 16)  e_d(2p), e_d(2p+1) interleaved per d
// ehi, elo : bf16[Kpad][D]      (tensor path) E ~= ehi + elo, zero rows for k >= K
// e16      : fp16[Kpad][D]      (single-pass tensor path) fp16(E * 2^se), zero rows for k >= K
constexpr int kPadCodes = 256;
constexpr int kHeaderBytes = 256;
constexpr int kLowDMax = 16;
constexpr int kTcMinD = 64;
constexpr int kTcMaxD = 256;

__host__ __device__ inline int round_up_i(int x, int m) { return (x + m - 1) / m * m; }
__host__ __device__ inline size_t round_up_z(size_t x, size_t m) { return (x + m - 1) / m * m; }

__host__ __device__ inline bool tc_eligible_dim(int D) {
    return D >= kTcMinD && D <= kTcMaxD && (D % 64) == 0;
}

struct PackLayout {
    int K, D, Kpad;
    size_t off_half_norm, off_pairs, off_ehi, off_elo, off_e16, off_half_norm_fin, total;
    bool has_pairs, has_bf16;
};

__host__ __device__ inline PackLayout pack_layout(int K, int D) {
    PackLayout L;
    L.K = K;
    L.D = D;
    L.Kpad = round_up_i(K, kPadCodes);
    L.has_pairs = D <= kLowDMax;
    L.has_bf16 = tc_eligible_dim(D);
    size_t off = kHeaderBytes;
    L.off_half_norm = off;
    off = round_up_z(off + sizeof(float) * L.Kpad, 1024);
    L.off_pairs = off;
    if (L.has_pairs) off = round_up_z(off + sizeof(float) * (size_t)L.Kpad * D, 1024);
    L.off_ehi = off;
    if (L.has_bf16) off = round_up_z(off + 2 * (size_t)L.Kpad * D, 1024);
    L.off_elo = off;
    if (L.has_bf16) off = round_up_z(off + 2 * (size_t)L.Kpad * D, 1024);
    L.off_e16 = off;
    if (L.has_bf16) off = round_up_z(off + 2 * (size_t)L.Kpad * D, 1024);
    L.off_half_norm_fin = off;  // half norms with a large FINITE pad (1e38) for the key-packing epilogue
    if (L.has_bf16) off = round_up_z(off + sizeof(float) * L.Kpad, 1024);
    L.total = off;
    return L;
}

// ---- small device helpers --------------------------------------------------
__device__ __forceinline__ unsigned long long pack_f32x2(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack_f32x2(unsigned long long v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
// d = a * b + c on two packed fp32 lanes (SASS FFMA2, sm_100+)
__device__ __forceinline__ unsigned long long fma_f32x2(unsigned long long a, unsigned long long b,
                                                        unsigned long long c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
// three-input minimum (SASS FMNMX3, sm_100+); NaN operands are ignored
__device__ __forceinline__ float min3_f32(float a, float b, float c) {
    float d;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}

__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// launch entry points implemented in the per-kernel translation units
int launch_codebook_prepare(const float* E, int K, int D, void* pack, cudaStream_t s);
int launch_search_lowd(const float* z, int64_t B, int D, int64_t HW, int K, const void* pack,
                       int64_t* idx_out, float* dmin_out, cudaStream_t s);
size_t search_fp32_workspace_bytes(int64_t n_rows);
int launch_search_fp32(const float* z, int64_t B, int D, int64_t HW, const float* E, int K,
                       const void* pack, const int32_t* token_list, const int32_t* list_count,
                       int64_t max_list, void* keys_ws, size_t keys_bytes, int64_t* idx_out,
                       float* dmin_out, cudaStream_t s);
size_t search_tc_workspace_bytes(int64_t n_tokens, int D, int K);
int launch_search_tc(const float* z, int64_t B, int D, int64_t HW, const float* E, int K,
                     const void* pack, int64_t* idx_out, float* dmin_out, void* ws, size_t ws_bytes,
                     int64_t* stats_out, cudaStream_t s);
size_t search_tc16_workspace_bytes(int64_t n_tokens, int D, int K);
int launch_search_tc16(const float* z, int64_t B, int D, int64_t HW, const float* E, int K,
                       const void* pack, int64_t* idx_out, float* dmin_out, void* ws, size_t ws_bytes,
                       int64_t* stats_out, cudaStream_t s);
void set_lowd_variant(int v);
void set_tc16_cluster(int c);

}  // namespace vqb
