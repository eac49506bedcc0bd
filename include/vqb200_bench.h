/*
 * vqb200_bench -- measurement-only additions of libvqb200_bench.so.
 *
 * libvqb200_bench.so is the SAME source tree compiled with -DVQB_EXPERIMENTAL: every symbol of
 * vqb200.h plus the ones below.  It exists for bench.py (FP32 FMA peak = the roofline
 * denominator of the low-D search) and the A/B scripts under scripts/.  The product library
 * (libvqb200.so) has none of these: its launch shapes are compile-time constants.
 *
 * vqb_tune() is PROCESS-GLOBAL and NOT thread-safe: a knob changes the launch shape (never the
 * results, except "tclow_skip_stages" and "conv_debug", which disable pipeline stages for timing
 * bisection and make results wrong by design) for every stream and module in the process.
 */
#ifndef VQB200_BENCH_H
#define VQB200_BENCH_H

#include "vqb200.h"

#ifdef __cplusplus
extern "C" {
#endif

/*   "lowd_variant"       0..4   launch shape of the CUDA-core low-D search (0 = 256 threads x 2 CTA/SM)
 *   "lowd_ctas_per_sm"   0..3   0 = the variant's own residency; 1 leaves room for a co-running kernel
 *   "dual_permille"      0..999 share of the images handed to the tensor role of the two-engine search (algo 6; 0 = model)
 *   "tc16_cluster"       1|2|4  cluster size (codebook-stage multicast) of the fp16 tensor search
 *   "tc16_pruned"        0|1    pruned exact tier of the fp16 tensor search for long lists of uncertified tokens (default 1)
 *   "tc16_group"         0|4|8  codes per candidate group of the fp16 tensor search (0 = 8 up to padded D 128, 4 above)
 *   "dw_hw_trunc"        0|1    1x1 conv parameter gradients: 1 = the landed fp32 tile is the hi image as it is (relies on the
 *                               tensor core ignoring the low 13 mantissa bits), only the residual image is written
 *   "norm_fwd_reg"       0|1    GroupNorm+SiLU forward (<= 16 items per group): staged in shared memory / register-resident (default 1)
 *   "norm_bwd2"          0|1    GroupNorm+SiLU backward (<= 16 items per group): one 512-thread CTA per SM with xhat and dxhat in
 *                               registers / two per SM with xhat in shared memory (default 1)
 *   "norm_cluster"       0|1|2  GroupNorm+SiLU: round-1 staged kernels only / product (default 1) / also clusters of
 *                               2-8 CTAs per (image, group) (measured slower)
 *   "tclow_cluster"      1|2|4  same for the low-D tensor search
 *   "tclow_skip_stages"  0..7   bit mask: 1 tensor kernel, 2 chunk re-score, 4 exact list search (WRONG RESULTS)
 *   "fwd_pass_channels" / "bwd_pass_channels"  64|128|192|256  channels per pass of the tiled tail kernels
 *   "tail_pipe" / "bwd_pipe"  0|1  pipelined (cp.async double-buffered) 128-token forward tail / backward for
 *                               D >= 128, D % 64 == 0 (default 1)
 *   "tail_tok128"        0|1    128-token float4 forward-tail kernel for D <= 64 (default 1)
 *   "bwd_warp"           0|1|2  warp-private backward kernel (default 0 = off; measured slower)
 *   "tail_warp"          0|1|2  warp-private forward-tail kernel: off / D >= 128 (default) / any D % 32 == 0
 *   "conv_debug"         0..15  bit mask for the 1x1 convolution: 1 no activation loads, 2 no stores,
 *                               4 one MMA in three (WRONG RESULTS) */
VQB_API int vqb_tune(const char* key, int value);

/* FP32 FMA peak microbenchmark (the low-D roofline denominator): launches a register-resident FFMA
 * (packed=0) or FFMA2 (packed=1) loop on every SM and returns the flop count issued; the caller
 * times it with CUDA events. */
VQB_API int vqb_fma_peak_launch(int packed, int iters, float* sink, double* flops_host,
                        vqb_stream_t stream);
/* access-pattern ceiling of the tiled tail kernels: mode 0 strided copy out = a,
 * mode 1 out = a + b, mode 2 linear float4 copy; tensors [B, D, HW] fp32, D % 64 == 0 */
VQB_API int vqb_ubench_copy(const float* a, const float* b, float* out, int64_t B, int D, int64_t HW, int mode,
                    vqb_stream_t stream);
/* scatter-reduction pattern of the codebook gradient: every token adds a D-float row of ones into
 * table[idx[token]] with red.global.add.v4.f32, lanes_per_row (1, 4 or 16) consecutive lanes per row piece */
VQB_API int vqb_ubench_red(float* table, const int64_t* idx, int64_t N, int D, int lanes_per_row, vqb_stream_t stream);
/* instruction-mix microbenchmarks of the low-D inner loop (modes in csrc/vqb_ubench.cu);
 * src: >= 10240 floats of finite data */
VQB_API int vqb_ubench_launch(int mode, int sweeps, const float* src, float* sink, double* flops_host,
                      vqb_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* VQB200_BENCH_H */
