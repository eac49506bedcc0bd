/*
 * vqb200 -- C ABI of the B200-native VQ bottleneck (libvqb200.so).
 *
 * The reference (heimaoqqq/vq-gan) has no FFI layer: its boundary for this path
 * is the Python class `VectorQuantizer` in vqgan_ldm_baseline/models/quantizer.py.
 * The entry points below are what a binding for that class needs; each cites the
 * reference lines it replaces (paths relative to the reference root).  The
 * Python host (`vq_gan_b200/`) binds them with ctypes and wraps them in
 * `torch.library` custom ops; INTEGRATION.md shows the reference-side stub.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless the name ends in `_host`;
 *  - latents use the reference layout z[B, D, HW] (NCHW contiguous, HW = H*W);
 *    token i = b*HW + hw, element (i, d) lives at z[(b*D + d)*HW + hw].  A
 *    token-major [N, D] matrix is the same layout with B = N, HW = 1;
 *  - the codebook is E[K, D] row-major fp32 (`embedding.weight`, quantizer.py:47);
 *  - indices are int64 like torch.argmin's (quantizer.py:76,101);
 *  - all work is enqueued on `stream`; nothing synchronises the device;
 *  - every function returns VQB_OK or a negative code, never throws or aborts;
 *    `vqb_last_error()` gives the thread-local message of the last failure;
 *  - the library allocates no user-visible memory: scratch is caller-provided
 *    and sized by the `*_bytes` queries;
 *  - the library holds no mutable process-global state (an immutable per-device
 *    SM-count cache only) and is re-entrant; every launching entry point fails with
 *    VQB_ERR_UNSUPPORTED on a device that is not sm_100.  Experiment knobs and
 *    microbenchmarks live in a separate measurement build (vqb200_bench.h).
 */
#ifndef VQB200_H
#define VQB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VQB_OK 0
#define VQB_ERR_INVALID_ARG (-1)
#define VQB_ERR_CUDA (-2)
#define VQB_ERR_UNSUPPORTED (-3)
#define VQB_ERR_WORKSPACE (-4)

/* search algorithms (vqb_search_f32 `algo`) */
#define VQB_ALGO_AUTO 0
#define VQB_ALGO_LOWD_FMA 1   /* D <= 16: packed-FFMA2 CUDA-core kernel         */
#define VQB_ALGO_FP32_TILE 2  /* any D: fp32 register-tiled CUDA-core kernel    */
#define VQB_ALGO_TCGEN05 3    /* D in {64,128,192,256}: bf16x3 tcgen05/TMEM     */
#define VQB_ALGO_TCGEN05_F16 4 /* any 16 < D <= 512 (zero-padded to 64-channel blocks): one fp16 tcgen05
                                 pass, certified candidates re-scored exactly in fp32 */
#define VQB_ALGO_TCGEN05_TF32X3 5 /* D <= 16: tf32x3 tcgen05 pass, certified 32-code chunk re-scored
                                     with the FMA chain of algo 1 (bit-identical indices) */

#define VQB_ALGO_DUAL_LOWD 6 /* 3 <= D <= 16, >= 2 images: CUDA-core role and tf32x3 tensor role in ONE CTA on disjoint images
                               (different pipes of the SM); bit-identical indices to VQB_ALGO_LOWD_FMA */

/* flag OR-ed into `algo` = VQB_ALGO_TCGEN05_F16: the token split in `workspace` was already written by the producer of z
 * (vqb_conv1x1_split_f32), skip the split pass */
#define VQB_SEARCH_PRESPLIT 0x100

typedef void* vqb_stream_t; /* a cudaStream_t */

#if defined(__GNUC__)
#define VQB_API __attribute__((visibility("default")))
#else
#define VQB_API
#endif

/* ---- library ---------------------------------------------------------- */
VQB_API int vqb_version(void);
VQB_API const char* vqb_last_error(void);
/* sm count, compute capability and opt-in shared memory of `device` */
VQB_API int vqb_device_query(int device, int* sm_count, int* cc_major, int* cc_minor,
                     size_t* smem_optin_bytes);

/* ---- codebook pre-pass -------------------------------------------------
 * Replaces `torch.sum(self.embedding.weight ** 2, dim=1)` (quantizer.py:70) and
 * stages the codebook in the layouts the search kernels read (half norms,
 * pair-interleaved fp32 rows for D <= 16, bf16 hi/lo split for the tensor
 * path).  Must be re-run whenever E changes (once per forward). */
VQB_API size_t vqb_codebook_pack_bytes(int K, int D);
VQB_API int vqb_codebook_prepare_f32(const float* E, int K, int D, void* pack, size_t pack_bytes,
                             vqb_stream_t stream);

/* ---- nearest-code search -----------------------------------------------
 * Replaces the distance matrix + argmin (quantizer.py:68-76) without
 * materialising [N, K].  Minimises 0.5|e|^2 - z.e (|z|^2 is constant per row),
 * lowest index on ties, NaN rows -> the first NaN code like ATen's argmin.
 * idx_out[N] int64; dmin_out[N] (nullable) receives min_k(0.5|e_k|^2 - z.e_k).
 * stats_out (nullable, int64[4]): [0] tokens fully re-scored in fp32 by a tensor
 * path (no certified winner), [1] algorithm actually used, [2] tokens with two or
 * three certified candidate groups (algo 4), [3] algo 4: how many of the [0] tokens
 * the pruned exact tier took (collapsed codebooks; the rest met the full fp32
 * re-search); algo 6: tokens handled by the tensor role. */
VQB_API size_t vqb_search_workspace_bytes(int64_t B, int D, int64_t HW, int K, int algo);
VQB_API int vqb_search_f32(const float* z, int64_t B, int D, int64_t HW, const float* E, int K,
                   const void* pack, int64_t* idx_out, float* dmin_out, void* workspace,
                   size_t workspace_bytes, int algo, int64_t* stats_out, vqb_stream_t stream);

/* ---- forward tail -------------------------------------------------------
 * Replaces gather + layout restore + both mse_loss calls + vq_loss +
 * straight-through value (quantizer.py:80-98) in one pass:
 *   zq_out[B,D,HW] = z + (E[idx] - z)        (fp32, two roundings, as :98)
 *   loss_out[0] = mean((E[idx]-z)^2)          (codebook_loss == commitment_loss)
 *   loss_out[1] = loss_out[0] + beta*loss_out[0]   (vq_loss, :95)
 * partials: scratch of vqb_tail_partials_bytes(N) bytes.
 * err_flag (nullable int32): set to 1 if an index is outside [0, K). */
VQB_API size_t vqb_tail_partials_bytes(int64_t n_tokens);
VQB_API int vqb_gather_loss_st_f32(const float* z, const float* E, const int64_t* idx, int64_t B, int D,
                           int64_t HW, int K, float beta, float* zq_out, float* loss_out,
                           void* partials, size_t partials_bytes, int* err_flag,
                           vqb_stream_t stream);

/* ---- backward -----------------------------------------------------------
 * The gradients autograd derives from quantizer.py:89-98:
 *   dz[B,D,HW]  = g_zq + g_vq*(2/n)*(z - E[idx])                 (g_zq nullable = 0)
 *   dE[K,D]    += g_vq*beta*(2/n) * sum_{i: idx_i = k} (E[k] - z_i)   (dense, ACCUMULATES)
 *   hist[K]    += bincount(idx)                                  (nullable, int64)
 * g_vq is a device scalar (nullable = 0); n = B*D*HW.  dE and hist must be
 * zeroed by the caller when a fresh gradient is wanted. */
VQB_API int vqb_backward_f32(const float* z, const float* E, const int64_t* idx, const float* g_zq,
                     const float* g_vq, float beta, int64_t B, int D, int64_t HW, int K,
                     float* dz_out, float* dE_accum, int64_t* hist_accum, vqb_stream_t stream);

/* ---- helper methods -----------------------------------------------------
 * get_codebook_entry (quantizer.py:112-132): out[B,D,HW] = E[idx[b,hw]][d]. */
VQB_API int vqb_gather_f32(const float* E, const int64_t* idx, int64_t B, int D, int64_t HW, int K,
                   float* out, int* err_flag, vqb_stream_t stream);
/* get_codebook_usage (quantizer.py:134-149): hist_out[K] = bincount (overwritten),
 * used_out[0] = number of codes with a non-zero count. */
VQB_API int vqb_hist_i64(const int64_t* idx, int64_t n_tokens, int K, int64_t* hist_out,
                 int64_t* used_out, int* err_flag, vqb_stream_t stream);

/* ---- extensions (no reference code; semantics in DESIGN.md) ------------- */
/* per-code token sums for an EMA codebook: counts[K] += bincount, sums[K,D] += sum z_i */
VQB_API int vqb_code_sums_f32(const float* z, const int64_t* idx, int64_t B, int D, int64_t HW, int K,
                      float* counts_accum, float* sums_accum, vqb_stream_t stream);
/* N <- g N + (1-g) counts ; m <- g m + (1-g) sums ; E = m / laplace(N) */
VQB_API int vqb_ema_update_f32(float* E, float* cluster_size, float* embed_sum, const float* counts,
                       const float* sums, int K, int D, float decay, float eps,
                       float* total_scratch, vqb_stream_t stream);
/* codebook-sharded search: key[i] = (monotone(dmin[i]) << 32) | (idx[i] + index_offset), a
 * signed int64 whose minimum over shards picks the nearest code, lowest index on ties */
VQB_API int vqb_pack_argmin_keys(const float* dmin, const int64_t* idx, int64_t n, int64_t index_offset,
                         int64_t* keys_out, vqb_stream_t stream);
VQB_API int vqb_unpack_argmin_keys(const int64_t* keys, int64_t n, int64_t* idx_out, float* dmin_out,
                           vqb_stream_t stream);

/* data-parallel statistics message (one fp32 all-reduce per step):
 * flat = [ dE (n_dE) | scalars | hist mod 4096 | hist div 4096 ]; unpack scales dE by dE_scale
 * (1/world reproduces DDP's gradient averaging) and rebuilds the int64 histogram exactly. */
VQB_API int vqb_stats_pack(const float* dE, int64_t n_dE, const float* scalars, int n_scalars,
                   const int64_t* hist, int n_hist, float* flat_out, vqb_stream_t stream);
VQB_API int vqb_stats_unpack(const float* flat, int64_t n_dE, int n_scalars, int n_hist, float dE_scale,
                     float* dE_out, float* scalars_out, int64_t* hist_out, vqb_stream_t stream);

/* ---- compact index maps (next row N3: VQVAE.encode_to_indices / decode_from_indices,
 * vq_vae.py:162-190, cached like preprocess_latents.py:236-238 but 1-4 bytes per token) ----
 * vqb_index_bytes(K): 1 (K <= 256), 2 (K <= 65536) or 4.
 * narrow: codes_out[i] = (uintN) idx[i]; err_flag set to 1 when an index is outside [0, K).
 * widen : idx_out[i] = (int64) codes[i]. */
VQB_API int vqb_index_bytes(int K);
VQB_API int vqb_indices_narrow(const int64_t* idx, int64_t n, int K, void* codes_out, int elem_bytes,
                       int* err_flag, vqb_stream_t stream);
VQB_API int vqb_indices_widen(const void* codes, int64_t n, int elem_bytes, int64_t* idx_out,
                      vqb_stream_t stream);

/* ---- 1x1 convolution either side of the quantizer (next row N1) -----------------------
 * `pre_quant_conv` / `post_quant_conv` = nn.Conv2d(Cin, Cout, kernel_size=1) (vq_vae.py:74-79,115,121):
 *   y[B, Cout, HW] = W[Cout, Cin] . x[B, Cin, HW] + bias[Cout]     (bias nullable)
 * algo 0 = auto, 1 = 3xTF32 tcgen05 path (needs Cin % 32 == 0, Cout % 16 == 0, Cout <= 256;
 * fp32-level accuracy: relative error ~2^-21), 2 = CUDA-core path (any shape).
 * workspace: vqb_conv1x1_workspace_bytes(Cin, Cout) bytes, 256-byte aligned (split weights). */
VQB_API size_t vqb_conv1x1_workspace_bytes(int Cin, int Cout);
VQB_API int vqb_conv1x1_f32(const float* x, int64_t B, int Cin, int64_t HW, const float* W,
                    const float* bias, int Cout, float* y, void* workspace, size_t workspace_bytes,
                    int algo, vqb_stream_t stream);

/* pre_quant_conv fused with the quantizer's token split (vq_vae.py:115 feeding :118): y as vqb_conv1x1_f32 (tensor path
 * only), plus -- while the accumulator is still in TMEM -- the fp16 token rows / scales / norms / rounding residuals the
 * fp16 tensor search would otherwise compute from y (one read of y and one launch less).  `codebook_pack` is the prepared
 * pack of the quantizer's codebook [K, Cout]; `search_workspace` (vqb_search_workspace_bytes(B, Cout, HW, K,
 * VQB_ALGO_TCGEN05_F16) bytes) is then passed to vqb_search_f32 with algo = VQB_ALGO_TCGEN05_F16 | VQB_SEARCH_PRESPLIT. */
VQB_API int vqb_conv1x1_split_f32(const float* x, int64_t B, int Cin, int64_t HW, const float* W,
                          const float* bias, int Cout, float* y, void* workspace, size_t workspace_bytes,
                          const void* codebook_pack, int K, void* search_workspace,
                          size_t search_workspace_bytes, vqb_stream_t stream);

/* parameter gradients of the same layer (autograd of vq_vae.py:115,121): a reduction over the tokens,
 *   dW_accum[Cout, Cin] += sum_{b,hw} dy[b, o, hw] * x[b, c, hw]      dbias_accum[Cout] += sum_{b,hw} dy[b, o, hw]
 * 3xTF32 tcgen05 (fp32-level accuracy), both operands straight from NCHW by TMA.  Both outputs are ACCUMULATED into
 * (the caller zeroes them; dbias_accum nullable).  Shapes: vqb_conv1x1_dw_supported(Cin, Cout, HW) != 0, i.e.
 * Cin % 16 == 0, Cout <= 256, HW >= 32, HW % 4 == 0; anything else returns VQB_ERR_UNSUPPORTED (no slow path: the
 * caller keeps its library GEMM for such shapes). */
VQB_API int vqb_conv1x1_dw_supported(int Cin, int Cout, int64_t HW);
VQB_API int vqb_conv1x1_dw_f32(const float* dy, const float* x, int64_t B, int Cin, int Cout, int64_t HW,
                       float* dW_accum, float* dbias_accum, vqb_stream_t stream);

/* ---- encoder tail / decoder tail normalisation + activation (next row N2) ---------------
 * `h = norm_out(h); h = F.silu(h)` (encoder_decoder.py:166-167 and 249-250; nn.GroupNorm(groups, C, eps, affine)):
 *   y[b,c,hw] = silu((x - mean[b,g]) * rstd[b,g] * gamma[c] + beta[c]),  g = c / (C / groups)
 * in one pass over HBM (the group is staged in shared memory when (C/groups)*HW <= 48 Ki floats).  mean_out / rstd_out
 * [B*groups] are saved for the backward, which returns dx and ACCUMULATES dgamma[C] / dbeta[C] (nullable). */
VQB_API int vqb_groupnorm_silu_f32(const float* x, int64_t B, int C, int64_t HW, const float* gamma,
                           const float* beta, int groups, float eps, float* y, float* mean_out,
                           float* rstd_out, vqb_stream_t stream);
VQB_API int vqb_groupnorm_silu_backward_f32(const float* dy, const float* x, int64_t B, int C, int64_t HW,
                                    const float* gamma, const float* beta, int groups, const float* mean,
                                    const float* rstd, float* dx, float* dgamma_accum, float* dbeta_accum,
                                    vqb_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* VQB200_H */
