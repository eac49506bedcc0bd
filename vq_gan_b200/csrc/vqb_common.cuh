// Shared declarations for libvqb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>

#include "../../include/vqb200.h"
#ifdef VQB_EXPERIMENTAL
#include "../../include/vqb200_bench.h"
#endif

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

int device_ready();  // VQB_OK once the current device is a queried sm_100 part, an error code otherwise
int sm_count();      // SM count of the current device (valid after device_ready())
#define VQB_DEVICE_TRY()                                  \
    do {                                                  \
        if (int _rc = ::vqb::device_ready()) return _rc;  \
    } while (0)

// Launch-shape knobs.  The product library (libvqb200.so) compiles them as constants: it has no
// process-global mutable state.  Only the measurement build (libvqb200_bench.so, -DVQB_EXPERIMENTAL,
// include/vqb200_bench.h) makes them variables behind vqb_tune() for A/B experiments.
#ifdef VQB_EXPERIMENTAL
#define VQB_KNOB static int
#else
#define VQB_KNOB static constexpr int
#endif

// ---- packed codebook layout ----------------------------------------------
// header (int32[64]): [0] first NaN code (K if none) [1] K [2] D [4] float bits of max 0.5|e|^2
//                      [5] fp16 scale exponent se (e16 = fp16(E * 2^se)) [6] float bits of max|E| [7] bits of max |e - fp16(e)|
//                      [8] bits of max_k |e_k - eh_k - el_k| and [9] bits of max_k |el_k| (tf32x3 image, D <= 16)
//                      residual bounds as a function of the code norm n = |e_k| (valid for EVERY code by construction:
//                      rho = max relative residual over codes with n >= theta, a = max absolute residual below theta,
//                      theta = 2^-12 max|E|):  [10],[11] fp16 image: |e - e16| <= rho16 n + a16;
//                      [12],[13] tf32x3 residual |r_e| <= rho_r n + a_r;  [14],[15] low part |el| <= rho_l n + a_l
// gmax     : float[Kpad/4]   max code norm per 4-code group  (fp16 tensor path, 16 < D <= 256)
// cmax     : float[Kpad/32]  max code norm per 32-code chunk (low-D tensor path, D <= 16)
// half_norm: float[Kpad]  0.5|e_k|^2, +inf for k >= K.  Read as consecutive
//            (h[2p], h[2p+1]) pairs by the low-D kernel.
// pairs    : float[Kpad/2][2D]  (D <= 16)  e_d(2p), e_d(2p+1) interleaved per d
// ehi, elo : bf16[Kpad][D]      (tensor path) E ~= ehi + elo, zero rows for k >= K
// e16      : fp16[Kpad][Dpad]   (single-pass tensor path, 16 < D <= 256) fp16(E * 2^se), zero rows for k >= K,
//            zero columns for d >= D (Dpad = D rounded up to 64)
// half_norm_fin: like half_norm with a large FINITE pad (1e38), read by the tensor kernel's key-packing epilogue; a code
//            whose row is bit-identical to that of a LOWER-indexed code is padded too (it can never be the answer:
//            ties go to the lowest index), so that a codebook full of copies does not fill every candidate slot with
//            the same score.  rowhash / dup: the hash table that finds them (codebook_shadow_kernel)
constexpr int kPadCodes = 256;
constexpr int kHeaderBytes = 256;
constexpr int kLowDMax = 16;
constexpr int kTcMinD = 64;
constexpr int kTcMaxD = 256;     // bf16x3 kernel (token tile hi + lo resident)
constexpr int kTc16MaxD = 512;   // fp16 kernel: a 128 x 512 fp16 token tile (128 KB) still leaves two 32 KB codebook stages

__host__ __device__ inline int round_up_i(int x, int m) { return (x + m - 1) / m * m; }
__host__ __device__ inline size_t round_up_z(size_t x, size_t m) { return (x + m - 1) / m * m; }

__host__ __device__ inline bool tc_eligible_dim(int D) {  // bf16x3 kernel: whole 64-channel blocks only
    return D >= kTcMinD && D <= kTcMaxD && (D % 64) == 0;
}
// single-pass fp16 kernel: any 16 < D <= 256, the fp16 operand images are zero-padded to 64-channel blocks
constexpr int kTc16WideGroupDpad = 64;   // fp16 kernel: candidate groups of 8 codes up to this padded D, of 4 above
__host__ __device__ inline bool tc16_eligible_dim(int D) { return D > kLowDMax && D <= kTc16MaxD; }
__host__ __device__ inline int tc16_dpad(int D) { return round_up_i(D, 64); }

struct PackLayout {
    int K, D, Kpad;
    size_t off_half_norm, off_pairs, off_ehi, off_elo, off_e16, off_half_norm_fin, off_img, off_gmax, off_cmax, off_rowhash,
        off_dup, total;
    bool has_pairs, has_bf16, has_e16;
    int Dpad;       // row length of the fp16 image (D rounded up to 64)
    int dup_slots;  // power of two >= 2 Kpad: open-addressing table of (row hash << 32 | lowest code index) words
};

// ---- tf32x3 operand images of the low-D tensor path (vqb_search_tclow.cu) -----------------
// A row of the contraction has 3D+3 slots, padded to `steps` k-steps of 8 tf32 values:
//   tokens: [ zh(D) | zl(D) | zh(D) | 1 1 1 | 0.. ]      codes: [ -eh(D) | -eh(D) | -el(D) | h1 h2 h3 | 0.. ]
// so that one accumulator element is h - z.e (hi*hi + lo*hi + hi*lo terms, half norm in three tf32
// pieces).  Rows are grouped in tiles of 128; a tile is stored as the exact shared-memory image
// of a K-major SWIZZLE_NONE UMMA operand: [16-byte k-chunk][8-row group][row][4 floats], i.e.
// leading byte offset 2048, stride byte offset 128 -- one contiguous bulk copy per tile.
constexpr int kLowRows = 128;
__host__ __device__ inline int tclow_steps(int D) { return (3 * D + 3 + 7) / 8; }
__host__ __device__ inline size_t tclow_tile_floats(int D) { return (size_t)kLowRows * 8 * tclow_steps(D); }
__host__ __device__ inline size_t tclow_slot_offset(int row_in_tile, int slot) {
    return (size_t)(slot >> 2) * (kLowRows * 4) + (size_t)(row_in_tile >> 3) * 32 + (row_in_tile & 7) * 4 + (slot & 3);
}

__host__ __device__ inline PackLayout pack_layout(int K, int D) {
    PackLayout L;
    L.K = K;
    L.D = D;
    L.Kpad = round_up_i(K, kPadCodes);
    L.has_pairs = D <= kLowDMax;
    L.has_bf16 = tc_eligible_dim(D);
    L.has_e16 = tc16_eligible_dim(D);
    L.Dpad = tc16_dpad(D);
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
    if (L.has_e16) off = round_up_z(off + 2 * (size_t)L.Kpad * L.Dpad, 1024);
    L.off_half_norm_fin = off;  // half norms with a large FINITE pad (1e38) for the key-packing epilogue
    if (L.has_e16) off = round_up_z(off + sizeof(float) * L.Kpad, 1024);
    L.off_img = off;  // tf32x3 codebook image (D <= 16)
    if (L.has_pairs) off = round_up_z(off + sizeof(float) * tclow_tile_floats(D) * (L.Kpad / kLowRows), 1024);
    L.off_gmax = off;
    if (L.has_e16) off = round_up_z(off + sizeof(float) * (L.Kpad / 4), 1024);
    L.off_cmax = off;
    if (L.has_pairs) off = round_up_z(off + sizeof(float) * (L.Kpad / 32), 1024);
    // exact-duplicate detection (fp16 tensor path): later copies of a code are hidden from the approximate pass.  (Tried
    // for the low-D tf32x3 image as well: it works -- a D = 4 codebook of 4x copies searches in 2.3 instead of 7.5 ms -- but the
    // hash insertions and look-ups add 8 us to a 27 us pre-pass, 0.4 % of every C2 step; a 3x slow-down on such a codebook
    // is not a cliff, so the headline path does not pay for it.)
    L.dup_slots = 1;
    while (L.dup_slots < 2 * L.Kpad) L.dup_slots <<= 1;
    L.off_rowhash = off;
    if (L.has_e16) off = round_up_z(off + 4 * (size_t)L.Kpad, 1024);
    L.off_dup = off;
    if (L.has_e16) off = round_up_z(off + 8 * (size_t)L.dup_slots, 1024);
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

// round-to-nearest tf32 image of an fp32 value (low 13 mantissa bits zero)
__device__ __forceinline__ float to_tf32(float v) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return __uint_as_float(r);
}

__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// launch entry points implemented in the per-kernel translation units
int launch_codebook_prepare(const float* E, int K, int D, void* pack, cudaStream_t s);
// ctas_per_sm: 0 = the kernel's own residency (2 CTAs per SM); 1 leaves half of every SM to a co-running kernel
int launch_search_lowd(const float* z, int64_t B, int D, int64_t HW, int K, const void* pack,
                       int64_t* idx_out, float* dmin_out, cudaStream_t s, int ctas_per_sm = 0);
size_t search_fp32_workspace_bytes(int64_t n_rows, int D = 0);
int launch_search_fp32(const float* z, int64_t B, int D, int64_t HW, const float* E, int K,
                       const void* pack, const int32_t* token_list, const int32_t* list_count,
                       int64_t max_list, void* keys_ws, size_t keys_bytes, int64_t* idx_out,
                       float* dmin_out, cudaStream_t s);
size_t search_tc_workspace_bytes(int64_t n_tokens, int D, int K);
int launch_search_tc(const float* z, int64_t B, int D, int64_t HW, const float* E, int K,
                     const void* pack, int64_t* idx_out, float* dmin_out, void* ws, size_t ws_bytes,
                     int64_t* stats_out, cudaStream_t s);
size_t search_tclow_workspace_bytes(int64_t n_tokens, int D, int K);
int launch_search_tclow(const float* z, int64_t B, int D, int64_t HW, const float* E, int K,
                        const void* pack, int64_t* idx_out, float* dmin_out, void* ws, size_t ws_bytes,
                        int64_t* stats_out, cudaStream_t s, bool share_sm = false);
// two engines in one CTA (D = 4): CUDA-core role + tensor role on disjoint images (vqb_search_tclow.cu)
// D >= 3: below that the tensor role (whose cost does not drop with D) cannot keep up with the CUDA cores
constexpr int kDualMinD = 3;
__host__ __device__ inline bool dual_eligible(int64_t B, int D) { return D >= kDualMinD && D <= kLowDMax && B >= 2; }
size_t search_dual_workspace_bytes(int64_t B, int D, int64_t HW, int K);
int launch_search_dual(const float* z, int64_t B, int D, int64_t HW, const float* E, int K, const void* pack,
                       int64_t* idx_out, float* dmin_out, void* ws, size_t ws_bytes, int64_t* stats_out, cudaStream_t s);
int launch_search_lowd_list(const float* z, int64_t B, int D, int64_t HW, int K, const void* pack,
                            const int32_t* list, const int32_t* list_count, int64_t* idx_out,
                            float* dmin_out, cudaStream_t s);
#ifdef VQB_EXPERIMENTAL
void set_tclow_cluster(int c);
void set_dual_permille(int v);
void set_tail_knob(const char* key, int value);
void set_conv_debug(int v);
#endif
// pruned exact tier of the fp16 tensor search (vqb_search_pruned.cu)
bool pruned_eligible(int K, int D);
size_t search_pruned_workspace_bytes(int64_t N, int K, int D);
float* search_pruned_ubound(void* ws, int64_t N, int K, int D);
int launch_search_pruned(const float* z, int64_t B, int D, int64_t HW, const float* E, int K, const void* pack,
                         const int32_t* token_list, const int32_t* list_count, const float* ubound, int32_t* state,
                         void* ws, size_t ws_bytes, int64_t* idx_out, float* dmin_out, const int32_t** remaining,
                         const int32_t** handled, cudaStream_t s);
constexpr int64_t kPrunedMinBatch = 16384;  // tokens per search call below which the tier is not even launched
constexpr int kPrunedStateInts = 16 + 2 * 128;  // zeroed by the caller: decision / counters, per-tile counts, cursors
size_t search_tc16_workspace_bytes(int64_t n_tokens, int D, int K);
int launch_search_tc16(const float* z, int64_t B, int D, int64_t HW, const float* E, int K,
                       const void* pack, int64_t* idx_out, float* dmin_out, void* ws, size_t ws_bytes,
                       int64_t* stats_out, cudaStream_t s, bool presplit = false);
// where the token split lives inside a search workspace of search_tc16_workspace_bytes(N, D, K) bytes
void tc16_split_pointers(void* ws, int64_t N, int D, __half** z16, float** inv_scale, float** znorm, float** zres);
void set_lowd_variant(int v);
void set_tc16_cluster(int c);
void set_tc16_branchy(int v);
void set_tc16_group(int v);
void set_tc16_pruned(int v);
void set_norm_cluster(int v);
void set_norm_bwd2(int v);
void set_norm_fwd_reg(int v);
void set_dw_hw_trunc(int v);

}  // namespace vqb
