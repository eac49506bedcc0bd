// HBM-bound kernels either side of the search: forward tail (gather + loss +
// straight-through), backward (dz + dE scatter + histogram), get_codebook_entry,
// get_codebook_usage, and the EMA / sharded-argmin extensions.
// Reference lines: quantizer.py:80-98 (tail), autograd of :89-98 (backward),
// :112-132 (entry), :134-149 (usage).
//
// Common mapping: the latent is z[B, D, HW]; a CTA is TX tokens (x, fastest) by
// S channel slices (y), TX*S = 256, so every global access to z / z_q / dz / g is
// a coalesced run along HW and every codebook access is a float4 walk along one
// row (rows are L2-resident: K*D*4 <= 64 MiB).
#include "vqb_common.cuh"

namespace vqb {

constexpr int kTailThreads = 256;

struct TailShape {
    int tx, slices, dims_per_slice;
};

static TailShape tail_shape(int D) {
    TailShape s;
    int slices = 1;
    // keep every slice a whole number of float4 groups so the vector path stays aligned
    if (D % 4 == 0)
        while (slices < 8 && D % (slices * 8) == 0) slices *= 2;
    s.slices = slices;
    s.tx = kTailThreads / slices;
    s.dims_per_slice = D / slices;
    return s;
}

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c),
                 "f"(d)
                 : "memory");
}

// ---------------------------------------------------------------------------
// forward tail
// ---------------------------------------------------------------------------
template <bool kVec4>
__global__ void __launch_bounds__(kTailThreads)
    gather_loss_st_kernel(const float* __restrict__ z, const float* __restrict__ E,
                          const int64_t* __restrict__ idx, int64_t N, int D, int64_t HW, int K,
                          int dims_per_slice, float* __restrict__ zq_out,
                          double* __restrict__ partials, int* __restrict__ err_flag) {
    const int64_t tok = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int d0 = threadIdx.y * dims_per_slice;
    float sq = 0.f;
    if (tok < N) {
        int64_t k = idx[tok];
        if (k < 0 || k >= K) {
            if (err_flag) *err_flag = 1;
            k = 0;
        }
        const int64_t b = tok / HW;
        const int64_t off = (b * D) * HW + (tok - b * HW);
        const float* erow = E + (size_t)k * D;
        if constexpr (kVec4) {
            for (int d = d0; d < d0 + dims_per_slice; d += 4) {
                const float4 ev = __ldg(reinterpret_cast<const float4*>(erow + d));
                const float e[4] = {ev.x, ev.y, ev.z, ev.w};
                float zv[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) zv[j] = __ldg(z + off + (int64_t)(d + j) * HW);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float diff = __fsub_rn(e[j], zv[j]);
                    zq_out[off + (int64_t)(d + j) * HW] = __fadd_rn(zv[j], diff);
                    sq = fmaf(diff, diff, sq);
                }
            }
        } else {
            for (int d = d0; d < d0 + dims_per_slice; ++d) {
                const float ev = __ldg(erow + d);
                const float zv = __ldg(z + off + (int64_t)d * HW);
                const float diff = __fsub_rn(ev, zv);
                zq_out[off + (int64_t)d * HW] = __fadd_rn(zv, diff);
                sq = fmaf(diff, diff, sq);
            }
        }
    }
    // block reduction in double, one partial per CTA (fixed order -> deterministic)
    __shared__ double warp_part[kTailThreads / 32];
    const int lin = threadIdx.y * blockDim.x + threadIdx.x;
    double v = warp_sum_f64((double)sq);
    if ((lin & 31) == 0) warp_part[lin >> 5] = v;
    __syncthreads();
    if (lin == 0) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < kTailThreads / 32; ++w) s += warp_part[w];
        partials[blockIdx.x] = s;
    }
}

__global__ void __launch_bounds__(256)
    loss_finalize_kernel(const double* __restrict__ partials, int64_t n_partials, double inv_n,
                         float beta, float* __restrict__ loss_out) {
    __shared__ double sh[256];
    double s = 0.0;
    for (int64_t i = threadIdx.x; i < n_partials; i += 256) s += partials[i];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const float mse = (float)(sh[0] * inv_n);
        loss_out[0] = mse;
        loss_out[1] = __fadd_rn(mse, __fmul_rn(beta, mse));  // fl(cb + fl(beta*commit)), quantizer.py:95
    }
}

// ---------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------
template <bool kVec4>
__global__ void __launch_bounds__(kTailThreads)
    backward_kernel(const float* __restrict__ z, const float* __restrict__ E,
                    const int64_t* __restrict__ idx, const float* __restrict__ g_zq,
                    const float* __restrict__ g_vq, float beta, float norm, int64_t N, int D,
                    int64_t HW, int K, int dims_per_slice, float* __restrict__ dz_out,
                    float* __restrict__ dE, unsigned long long* __restrict__ hist) {
    const int64_t tok = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int d0 = threadIdx.y * dims_per_slice;
    const float gv = g_vq ? __ldg(g_vq) : 0.f;
    const float gbeta = __fmul_rn(gv, beta);
    const bool live = tok < N;
    int64_t k = 0;
    if (live) {
        k = idx[tok];
        if (k < 0 || k >= K) k = 0;
    }
    if (hist != nullptr && threadIdx.y == 0) {
        // warp-aggregated histogram: one atomic per distinct code per warp
        const unsigned active = __ballot_sync(0xffffffffu, live);
        if (live) {
            const unsigned peers = __match_any_sync(active, (int)k);
            if ((threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(hist + k, (unsigned long long)__popc(peers));
        }
    }
    if (!live) return;
    const int64_t b = tok / HW;
    const int64_t off = (b * D) * HW + (tok - b * HW);
    const float* erow = E + (size_t)k * D;
    float* drow = dE + (size_t)k * D;
    if constexpr (kVec4) {
        for (int d = d0; d < d0 + dims_per_slice; d += 4) {
            const float4 ev = __ldg(reinterpret_cast<const float4*>(erow + d));
            const float e[4] = {ev.x, ev.y, ev.z, ev.w};
            float zv[4], gz[4], c[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                zv[j] = __ldg(z + off + (int64_t)(d + j) * HW);
                gz[j] = g_zq ? __ldg(g_zq + off + (int64_t)(d + j) * HW) : 0.f;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float t = __fmul_rn(__fmul_rn(norm, __fsub_rn(zv[j], e[j])), gv);
                dz_out[off + (int64_t)(d + j) * HW] = __fadd_rn(gz[j], t);
                c[j] = __fmul_rn(__fmul_rn(norm, __fsub_rn(e[j], zv[j])), gbeta);
            }
            if (dE) red_add_v4(drow + d, c[0], c[1], c[2], c[3]);
        }
    } else {
        for (int d = d0; d < d0 + dims_per_slice; ++d) {
            const float ev = __ldg(erow + d);
            const float zv = __ldg(z + off + (int64_t)d * HW);
            const float gz = g_zq ? __ldg(g_zq + off + (int64_t)d * HW) : 0.f;
            const float t = __fmul_rn(__fmul_rn(norm, __fsub_rn(zv, ev)), gv);
            dz_out[off + (int64_t)d * HW] = __fadd_rn(gz, t);
            if (dE) atomicAdd(drow + d, __fmul_rn(__fmul_rn(norm, __fsub_rn(ev, zv)), gbeta));
        }
    }
}

// ---------------------------------------------------------------------------
// D % 32 == 0: shared-memory transposed variants.  A CTA owns 32 consecutive tokens and up to
// 256 channels at a time: codebook rows are read as full 128-byte lines (lane = channel),
// transposed through a conflict-free [channels][33] tile, and combined with z / g in the
// token-major (lane = token) orientation in which z, z_q, dz are coalesced.  The dE scatter goes
// back through the same tile so each token row leaves as float4 reductions on consecutive
// addresses.  Two barriers per 256 channels; every thread keeps 16-32 loads in flight.
// ---------------------------------------------------------------------------
constexpr int kTileTok = 32;
constexpr int kTileDimMax = 256;

struct TileTokens {
    int64_t off[kTileTok];  // element offset of (token, channel 0); -1 for tokens past N
    int code[kTileTok];
};

__device__ __forceinline__ void tile_load_tokens(TileTokens& tt, const int64_t* __restrict__ idx, int64_t N,
                                                 int D, int64_t HW, int K, int* err_flag) {
    if (threadIdx.x < kTileTok) {
        const int64_t tok = (int64_t)blockIdx.x * kTileTok + threadIdx.x;
        int64_t off = -1;
        int k = 0;
        if (tok < N) {
            int64_t kk = idx[tok];
            if (kk < 0 || kk >= K) {
                if (err_flag) *err_flag = 1;
                kk = 0;
            }
            k = (int)kk;
            const int64_t b = tok / HW;
            off = (b * D) * HW + (tok - b * HW);
        }
        tt.off[threadIdx.x] = off;
        tt.code[threadIdx.x] = k;
    }
}

// codebook rows of the CTA's 32 tokens, channels [d0, d0+DC) -> tile[channel][token]
// (warp w stages tokens 4w..4w+3; lane = channel: full 128-byte lines, immediate offsets)
template <int DC>
__device__ __forceinline__ void tile_fill_codes(float (*tile)[kTileTok + 1], const TileTokens& tt,
                                                const float* __restrict__ E, int D, int d0) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int t = warp * 4 + i;
        const float* erow = E + (size_t)tt.code[t] * D + d0 + lane;
#pragma unroll
        for (int c = 0; c < DC / 32; ++c) tile[lane + 32 * c][t] = __ldg(erow + 32 * c);
    }
}

// DC = channels per pass (64, 128, 192 or 256, a divisor of D): every loop has a compile-time trip
// count, z (DRAM) is requested BEFORE the wait on the codebook rows (L2) so both latencies overlap.
template <int DC>
__global__ void __launch_bounds__(256, DC <= 128 ? 4 : 2)
    gather_loss_st_tiled_kernel(const float* __restrict__ z, const float* __restrict__ E,
                                const int64_t* __restrict__ idx, int64_t N, int D, int64_t HW, int K,
                                float* __restrict__ zq_out, double* __restrict__ partials,
                                int* __restrict__ err_flag) {
    constexpr int R = DC / 8;  // channels per thread per pass
    __shared__ float tile[DC][kTileTok + 1];
    __shared__ TileTokens tt;
    __shared__ double warp_part[8];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    tile_load_tokens(tt, idx, N, D, HW, K, err_flag);
    __syncthreads();
    const int64_t off = tt.off[tx];
    const int64_t step = 8 * HW;
    float sq = 0.f;
    for (int d0 = 0; d0 < D; d0 += DC) {
        float zv[R];
        if (off >= 0) {
            const float* zp = z + off + (int64_t)(d0 + ty) * HW;
#pragma unroll
            for (int i = 0; i < R; ++i) {
                zv[i] = __ldg(zp);
                zp += step;
            }
        }
        if (d0 > 0) __syncthreads();  // the previous pass is done with the tile
        tile_fill_codes<DC>(tile, tt, E, D, d0);
        __syncthreads();
        if (off >= 0) {
            float* qp = zq_out + off + (int64_t)(d0 + ty) * HW;
#pragma unroll
            for (int i = 0; i < R; ++i) {
                const float diff = __fsub_rn(tile[ty + 8 * i][tx], zv[i]);
                *qp = __fadd_rn(zv[i], diff);
                qp += step;
                sq = fmaf(diff, diff, sq);
            }
        }
    }
    double v = warp_sum_f64((double)sq);
    if (tx == 0) warp_part[ty] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += warp_part[w];
        partials[blockIdx.x] = s;
    }
}

template <int DC>
__global__ void __launch_bounds__(256, DC <= 64 ? 4 : (DC <= 128 ? 3 : 2))
    backward_tiled_kernel(const float* __restrict__ z, const float* __restrict__ E,
                          const int64_t* __restrict__ idx, const float* __restrict__ g_zq,
                          const float* __restrict__ g_vq, float beta, float norm, int64_t N, int D,
                          int64_t HW, int K, float* __restrict__ dz_out, float* __restrict__ dE,
                          unsigned long long* __restrict__ hist) {
    constexpr int R = DC / 8;
    __shared__ float tile[DC][kTileTok + 1];
    __shared__ TileTokens tt;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int lane = tx, warp = ty;
    tile_load_tokens(tt, idx, N, D, HW, K, nullptr);
    __syncthreads();
    const int64_t off = tt.off[tx];
    const int64_t step = 8 * HW;
    const float gv = g_vq ? __ldg(g_vq) : 0.f;
    const float gbeta = __fmul_rn(gv, beta);
    if (hist != nullptr && ty == 0) {
        const bool live = off >= 0;
        const unsigned active = __ballot_sync(0xffffffffu, live);
        if (live) {
            const int k = tt.code[tx];
            const unsigned peers = __match_any_sync(active, k);
            if (lane == __ffs(peers) - 1) atomicAdd(hist + k, (unsigned long long)__popc(peers));
        }
    }
    for (int d0 = 0; d0 < D; d0 += DC) {
        float zv[R], gz[R];
        if (off >= 0) {
            const int64_t o = off + (int64_t)(d0 + ty) * HW;
            const float* zp = z + o;
#pragma unroll
            for (int i = 0; i < R; ++i) {
                zv[i] = __ldg(zp);
                zp += step;
            }
            if (g_zq) {
                const float* gp = g_zq + o;
#pragma unroll
                for (int i = 0; i < R; ++i) {
                    gz[i] = __ldg(gp);
                    gp += step;
                }
            } else {
#pragma unroll
                for (int i = 0; i < R; ++i) gz[i] = 0.f;
            }
        }
        if (d0 > 0) __syncthreads();  // the scatter of the previous pass is done with the tile
        tile_fill_codes<DC>(tile, tt, E, D, d0);
        __syncthreads();
        if (off >= 0) {
            float* dp = dz_out + off + (int64_t)(d0 + ty) * HW;
#pragma unroll
            for (int i = 0; i < R; ++i) {
                const float e = tile[ty + 8 * i][tx];
                const float t = __fmul_rn(__fmul_rn(norm, __fsub_rn(zv[i], e)), gv);
                *dp = __fadd_rn(gz[i], t);
                dp += step;
                tile[ty + 8 * i][tx] = __fmul_rn(__fmul_rn(norm, __fsub_rn(e, zv[i])), gbeta);
            }
        }
        __syncthreads();
        if (dE != nullptr) {
            // half-warp per token: 16 lanes x float4 = 64 channels of one codebook-gradient row
            const int l16 = lane & 15, half = lane >> 4;
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int t = warp * 4 + i * 2 + half;
                if (tt.off[t] >= 0) {
                    float* drow = dE + (size_t)tt.code[t] * D + d0 + 4 * l16;
#pragma unroll
                    for (int c = 0; c < DC; c += 64)
                        if (DC % 64 == 0 || c + 4 * l16 < DC)  // DC = 32: the upper eight lanes of the half-warp idle
                            red_add_v4(drow + c, tile[c + 4 * l16 + 0][t], tile[c + 4 * l16 + 1][t],
                                       tile[c + 4 * l16 + 2][t], tile[c + 4 * l16 + 3][t]);
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------
// 128 tokens per CTA, one float4 (4 consecutive tokens) per lane: a warp instruction moves 512 contiguous bytes
// of one channel row.  Measured ceiling of this access pattern (scripts/pattern_copy.py): 6.6-6.7 TB/s, the
// same as a linear copy, against 4.2-5.7 TB/s for 128-byte segments.  Needs HW % 4 == 0 and 16-byte aligned
// tensors (a quad of tokens never straddles two images then); anything else takes the 32-token kernels above.
// ---------------------------------------------------------------------------
constexpr int kTok128 = 128;
constexpr int kTok128Stride = kTok128 + 4;  // floats; keeps rows 16-byte aligned

struct Tile128Tokens {
    int64_t off[kTok128 / 4];  // element offset of (first token of the quad, channel 0); -1 past N
    int code[kTok128];
};

__device__ __forceinline__ void tile128_load_tokens(Tile128Tokens& tt, const int64_t* __restrict__ idx, int64_t N,
                                                    int D, int64_t HW, int K, int* err_flag) {
    if (threadIdx.x < kTok128) {
        const int64_t tok = (int64_t)blockIdx.x * kTok128 + threadIdx.x;
        int k = 0;
        if (tok < N) {
            int64_t kk = idx[tok];
            if (kk < 0 || kk >= K) {
                if (err_flag) *err_flag = 1;
                kk = 0;
            }
            k = (int)kk;
        }
        tt.code[threadIdx.x] = k;
        if ((threadIdx.x & 3) == 0) {
            int64_t off = -1;
            if (tok < N) {
                const int64_t b = tok / HW;
                off = (b * D) * HW + (tok - b * HW);
            }
            tt.off[threadIdx.x >> 2] = off;
        }
    }
}

// codebook rows of the CTA's 128 tokens, channels [d0, d0+DC) -> tile[channel][token]; warp w stages tokens
// 16w .. 16w+15, lane = channel (full 128-byte lines from L2)
template <int DC>
__device__ __forceinline__ void tile128_fill_codes(float (*tile)[kTok128Stride], const Tile128Tokens& tt,
                                                   const float* __restrict__ E, int D, int d0) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll 4
    for (int i = 0; i < 16; ++i) {
        const int t = warp * 16 + i;
        const float* erow = E + (size_t)tt.code[t] * D + d0 + lane;
#pragma unroll
        for (int c = 0; c < DC / 32; ++c) tile[lane + 32 * c][t] = __ldg(erow + 32 * c);
    }
}

template <int DC>
__global__ void __launch_bounds__(256, 3)
    gather_loss_st_tok128_kernel(const float* __restrict__ z, const float* __restrict__ E,
                                 const int64_t* __restrict__ idx, int64_t N, int D, int64_t HW, int K,
                                 float* __restrict__ zq_out, double* __restrict__ partials,
                                 int* __restrict__ err_flag) {
    constexpr int R = DC / 8;  // channels per thread per pass
    __shared__ __align__(16) float tile[DC][kTok128Stride];
    __shared__ Tile128Tokens tt;
    __shared__ double warp_part[8];
    const int tq = threadIdx.x & 31, cy = threadIdx.x >> 5;
    tile128_load_tokens(tt, idx, N, D, HW, K, err_flag);
    __syncthreads();
    const int64_t off = tt.off[tq];
    const int64_t step = 8 * HW;
    float sq = 0.f;
    for (int d0 = 0; d0 < D; d0 += DC) {
        float4 zv[R];
        if (off >= 0) {
            const float* zp = z + off + (int64_t)(d0 + cy) * HW;
#pragma unroll
            for (int i = 0; i < R; ++i) {
                zv[i] = __ldg(reinterpret_cast<const float4*>(zp));
                zp += step;
            }
        }
        if (d0 > 0) __syncthreads();  // the previous pass is done with the tile
        tile128_fill_codes<DC>(tile, tt, E, D, d0);
        __syncthreads();
        if (off >= 0) {
            float* qp = zq_out + off + (int64_t)(d0 + cy) * HW;
#pragma unroll
            for (int i = 0; i < R; ++i) {
                const float4 e = *reinterpret_cast<const float4*>(&tile[cy + 8 * i][4 * tq]);
                float4 q;
                const float dx = __fsub_rn(e.x, zv[i].x), dy = __fsub_rn(e.y, zv[i].y);
                const float dz = __fsub_rn(e.z, zv[i].z), dw = __fsub_rn(e.w, zv[i].w);
                q.x = __fadd_rn(zv[i].x, dx);
                q.y = __fadd_rn(zv[i].y, dy);
                q.z = __fadd_rn(zv[i].z, dz);
                q.w = __fadd_rn(zv[i].w, dw);
                *reinterpret_cast<float4*>(qp) = q;
                qp += step;
                sq = fmaf(dx, dx, sq);
                sq = fmaf(dy, dy, sq);
                sq = fmaf(dz, dz, sq);
                sq = fmaf(dw, dw, sq);
            }
        }
    }
    double v = warp_sum_f64((double)sq);
    if (tq == 0) warp_part[cy] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += warp_part[w];
        partials[blockIdx.x] = s;
    }
}

// ---------------------------------------------------------------------------
// Pipelined 128-token kernels for large D (D % 64 == 0): the mapping of the tok128 kernel above (512 contiguous
// bytes of z / g / z_q / dz per warp instruction -- the access pattern that reaches the linear-copy ceiling), with
// the two things that made it lose at D = 256 removed:
//   * the codebook rows of the NEXT 64-channel pass are fetched by cp.async (LDGSTS, 16 bytes each, no register
//     staging) into the other half of a double buffer while the current pass streams, so the L2 latency of the
//     gathered rows is never exposed behind a barrier;
//   * the rows are stored token-major [token][64 + 4] exactly as they lie in the codebook -- no transpose on the
//     way in.  A lane owns 4 consecutive tokens (one float4 along the tokens) and 4 consecutive channels per
//     step: its 4x4 block of z arrives as four float4 (one per channel), its 4x4 block of e as four LDS.128 (one
//     per token); the row order in shared memory is permuted (token 4l+j -> row 32j + l) so that the 8 lanes of
//     a quarter-warp hit 8 distinct 16-byte bank groups (row stride 68 floats = 17 groups).
// The backward leaves the codebook-gradient terms as one red.global.add.v4.f32 per (token, 4 channels) straight
// from registers: no second trip through shared memory.
// ---------------------------------------------------------------------------
constexpr int kPipeDC = 64;
constexpr int kPipeStride = kPipeDC + 4;
constexpr int kPipeBufFloats = kTok128 * kPipeStride;

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// codebook rows of the CTA's 128 tokens, channels [d0, d0+64) -> buf[perm(token)][channel]: 2048 16-byte chunks,
// 8 per thread; 16 consecutive threads copy one 256-byte row segment
__device__ __forceinline__ void pipe_fill_async(float* buf, const Tile128Tokens& tt, const float* __restrict__ E, int D,
                                                int d0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int id = threadIdx.x + 256 * i;
        const int t = id >> 4, q = id & 15;
        const int r = ((t & 3) << 5) | (t >> 2);
        cp_async16(buf + r * kPipeStride + 4 * q, E + (size_t)tt.code[t] * D + d0 + 4 * q);
    }
}

__device__ __forceinline__ float4 ld_f4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

template <int MINB>
__global__ void __launch_bounds__(256, MINB)
    gather_loss_st_pipe_kernel(const float* __restrict__ z, const float* __restrict__ E, const int64_t* __restrict__ idx,
                               int64_t N, int D, int64_t HW, int K, float* __restrict__ zq_out,
                               double* __restrict__ partials, int* __restrict__ err_flag) {
    extern __shared__ __align__(16) float pipe_smem[];
    __shared__ Tile128Tokens tt;
    __shared__ double warp_part[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    tile128_load_tokens(tt, idx, N, D, HW, K, err_flag);
    __syncthreads();
    const int64_t off = tt.off[lane];
    const int n_pass = D / kPipeDC;
    pipe_fill_async(pipe_smem, tt, E, D, 0);
    cp_async_commit();
    float sq = 0.f;
    for (int p = 0; p < n_pass; ++p) {
        const int d0 = p * kPipeDC;
        float* buf = pipe_smem + (p & 1) * kPipeBufFloats;
        if (p + 1 < n_pass) pipe_fill_async(pipe_smem + ((p + 1) & 1) * kPipeBufFloats, tt, E, D, d0 + kPipeDC);
        cp_async_commit();
        // this thread's z: 2 chunks of 4 channels x 4 tokens, requested before waiting for the rows
        float4 zv[2][4];
        if (off >= 0) {
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    zv[h][i] = ld_f4(z + off + (int64_t)(d0 + 4 * (warp + 8 * h) + i) * HW);
        }
        cp_async_wait<1>();
        __syncthreads();  // pass p's rows are visible to every thread
        if (off >= 0) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int q = warp + 8 * h;
                float e[4][4];  // [token j][channel i]
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float4 v = *reinterpret_cast<const float4*>(buf + (32 * j + lane) * kPipeStride + 4 * q);
                    e[j][0] = v.x, e[j][1] = v.y, e[j][2] = v.z, e[j][3] = v.w;
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float zz[4] = {zv[h][i].x, zv[h][i].y, zv[h][i].z, zv[h][i].w};
                    float o[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float diff = __fsub_rn(e[j][i], zz[j]);
                        o[j] = __fadd_rn(zz[j], diff);
                        sq = fmaf(diff, diff, sq);
                    }
                    *reinterpret_cast<float4*>(zq_out + off + (int64_t)(d0 + 4 * q + i) * HW) =
                        make_float4(o[0], o[1], o[2], o[3]);
                }
            }
        }
        __syncthreads();  // done with buf before pass p+2's rows overwrite it
    }
    double v = warp_sum_f64((double)sq);
    if (lane == 0) warp_part[warp] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double sum = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) sum += warp_part[w];
        partials[blockIdx.x] = sum;
    }
}

// kStage: the codebook-gradient terms go back into the row buffer and leave as half-warp-per-token reductions (256
// contiguous bytes of one dE row per half-warp) instead of one 16-byte piece of 32 different rows per instruction
template <int MINB, bool kStage>
__global__ void __launch_bounds__(256, MINB)
    backward_pipe_kernel(const float* __restrict__ z, const float* __restrict__ E, const int64_t* __restrict__ idx,
                         const float* __restrict__ g_zq, const float* __restrict__ g_vq, float beta, float norm, int64_t N,
                         int D, int64_t HW, int K, float* __restrict__ dz_out, float* __restrict__ dE,
                         unsigned long long* __restrict__ hist) {
    extern __shared__ __align__(16) float pipe_smem[];
    __shared__ Tile128Tokens tt;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    tile128_load_tokens(tt, idx, N, D, HW, K, nullptr);
    __syncthreads();
    const int64_t off = tt.off[lane];
    const float gv = g_vq ? __ldg(g_vq) : 0.f;
    const float gbeta = __fmul_rn(gv, beta);
    if (hist != nullptr && warp < 4) {
        const int t = threadIdx.x;  // 0..127
        const bool live = (int64_t)blockIdx.x * kTok128 + t < N;
        const unsigned active = __ballot_sync(0xffffffffu, live);
        if (live) {
            const int k = tt.code[t];
            const unsigned peers = __match_any_sync(active, k);
            if (lane == __ffs(peers) - 1) atomicAdd(hist + k, (unsigned long long)__popc(peers));
        }
    }
    int code[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) code[j] = tt.code[4 * lane + j];
    const int n_pass = D / kPipeDC;
    pipe_fill_async(pipe_smem, tt, E, D, 0);
    cp_async_commit();
    for (int p = 0; p < n_pass; ++p) {
        const int d0 = p * kPipeDC;
        float* buf = pipe_smem + (p & 1) * kPipeBufFloats;
        if (p + 1 < n_pass) pipe_fill_async(pipe_smem + ((p + 1) & 1) * kPipeBufFloats, tt, E, D, d0 + kPipeDC);
        cp_async_commit();
        float4 zv[2][4], gz[2][4];
        if (off >= 0) {
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int64_t o = off + (int64_t)(d0 + 4 * (warp + 8 * h) + i) * HW;
                    zv[h][i] = ld_f4(z + o);
                    gz[h][i] = g_zq ? ld_f4(g_zq + o) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
        }
        cp_async_wait<1>();
        __syncthreads();
        if (off >= 0) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int q = warp + 8 * h;
                float e[4][4];  // [token j][channel i]
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float4 v = *reinterpret_cast<const float4*>(buf + (32 * j + lane) * kPipeStride + 4 * q);
                    e[j][0] = v.x, e[j][1] = v.y, e[j][2] = v.z, e[j][3] = v.w;
                }
                float ge[4][4];  // codebook-gradient terms [token j][channel i]
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float zz[4] = {zv[h][i].x, zv[h][i].y, zv[h][i].z, zv[h][i].w};
                    const float gg[4] = {gz[h][i].x, gz[h][i].y, gz[h][i].z, gz[h][i].w};
                    float o[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float t = __fmul_rn(__fmul_rn(norm, __fsub_rn(zz[j], e[j][i])), gv);
                        o[j] = __fadd_rn(gg[j], t);
                        ge[j][i] = __fmul_rn(__fmul_rn(norm, __fsub_rn(e[j][i], zz[j])), gbeta);
                    }
                    *reinterpret_cast<float4*>(dz_out + off + (int64_t)(d0 + 4 * q + i) * HW) =
                        make_float4(o[0], o[1], o[2], o[3]);
                }
                if (dE != nullptr) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        if constexpr (kStage)
                            *reinterpret_cast<float4*>(buf + (32 * j + lane) * kPipeStride + 4 * q) =
                                make_float4(ge[j][0], ge[j][1], ge[j][2], ge[j][3]);
                        else
                            red_add_v4(dE + (size_t)code[j] * D + d0 + 4 * q, ge[j][0], ge[j][1], ge[j][2], ge[j][3]);
                    }
                }
            }
        }
        __syncthreads();
        if constexpr (kStage) {
            if (dE != nullptr) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int id = threadIdx.x + 256 * i;
                    const int t = id >> 4, q = id & 15;
                    if (tt.off[t >> 2] >= 0) {
                        const int r = ((t & 3) << 5) | (t >> 2);
                        const float4 v = *reinterpret_cast<const float4*>(buf + r * kPipeStride + 4 * q);
                        red_add_v4(dE + (size_t)tt.code[t] * D + d0 + 4 * q, v.x, v.y, v.z, v.w);
                    }
                }
            }
            __syncthreads();
        }
    }
}

static bool pipe_ok(int D, int64_t HW, const void* a, const void* b, const void* c, const void* d) {
    const uintptr_t bits = reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(c) |
                           reinterpret_cast<uintptr_t>(d);
    return D % kPipeDC == 0 && HW % 4 == 0 && (bits & 15u) == 0;
}
constexpr size_t kPipeSmemBytes = 2 * sizeof(float) * kPipeBufFloats;  // 69 632 B: three CTAs per SM

static bool tok128_ok(int D, int64_t HW, const void* a, const void* b) {
    return D % 32 == 0 && HW % 4 == 0 && ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15u) == 0;
}

// ---------------------------------------------------------------------------
// D = 4 or 8 (the low-D latents of config C2): one thread owns FOUR consecutive tokens, every access to
// z / z_q / g / dz is a float4 along the tokens (512 contiguous bytes per warp instruction), the four codebook
// rows are whole float4 loads from L2.  Needs HW % 4 == 0 and 16-byte aligned tensors.
// ---------------------------------------------------------------------------
__device__ __forceinline__ float f4_get(const float4& v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : (i == 2 ? v.z : v.w)); }

template <int DQ>
__global__ void __launch_bounds__(256)
    gather_loss_st_quad_kernel(const float* __restrict__ z, const float* __restrict__ E, const int64_t* __restrict__ idx,
                               int64_t N, int64_t HW, int K, float* __restrict__ zq_out, double* __restrict__ partials,
                               int* __restrict__ err_flag) {
    constexpr int D = 4 * DQ;
    const int64_t tok = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 4;
    float sq = 0.f;
    if (tok < N) {  // N is a multiple of 4 here (HW % 4 == 0): the quad is complete
        const longlong2 i01 = __ldg(reinterpret_cast<const longlong2*>(idx + tok));
        const longlong2 i23 = __ldg(reinterpret_cast<const longlong2*>(idx + tok) + 1);
        long long k[4] = {i01.x, i01.y, i23.x, i23.y};
        float4 e[4][DQ];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            if (k[t] < 0 || k[t] >= K) {
                if (err_flag) *err_flag = 1;
                k[t] = 0;
            }
#pragma unroll
            for (int c = 0; c < DQ; ++c) e[t][c] = __ldg(reinterpret_cast<const float4*>(E + (size_t)k[t] * D) + c);
        }
        const int64_t b = tok / HW;
        const int64_t off = (b * D) * HW + (tok - b * HW);
        float4 zv[D];
#pragma unroll
        for (int d = 0; d < D; ++d) zv[d] = __ldg(reinterpret_cast<const float4*>(z + off + (int64_t)d * HW));
#pragma unroll
        for (int d = 0; d < D; ++d) {
            const float dx = __fsub_rn(f4_get(e[0][d >> 2], d & 3), zv[d].x);
            const float dy = __fsub_rn(f4_get(e[1][d >> 2], d & 3), zv[d].y);
            const float dz = __fsub_rn(f4_get(e[2][d >> 2], d & 3), zv[d].z);
            const float dw = __fsub_rn(f4_get(e[3][d >> 2], d & 3), zv[d].w);
            float4 q;
            q.x = __fadd_rn(zv[d].x, dx);
            q.y = __fadd_rn(zv[d].y, dy);
            q.z = __fadd_rn(zv[d].z, dz);
            q.w = __fadd_rn(zv[d].w, dw);
            *reinterpret_cast<float4*>(zq_out + off + (int64_t)d * HW) = q;
            sq = fmaf(dx, dx, sq);
            sq = fmaf(dy, dy, sq);
            sq = fmaf(dz, dz, sq);
            sq = fmaf(dw, dw, sq);
        }
    }
    __shared__ double warp_part[8];
    double v = warp_sum_f64((double)sq);
    if ((threadIdx.x & 31) == 0) warp_part[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += warp_part[w];
        partials[blockIdx.x] = s;
    }
}

template <int DQ>
__global__ void __launch_bounds__(256)
    backward_quad_kernel(const float* __restrict__ z, const float* __restrict__ E, const int64_t* __restrict__ idx,
                         const float* __restrict__ g_zq, const float* __restrict__ g_vq, float beta, float norm, int64_t N,
                         int64_t HW, int K, float* __restrict__ dz_out, float* __restrict__ dE,
                         unsigned long long* __restrict__ hist) {
    constexpr int D = 4 * DQ;
    const int64_t tok = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 4;
    const bool live = tok < N;
    const float gv = g_vq ? __ldg(g_vq) : 0.f;
    const float gbeta = __fmul_rn(gv, beta);
    long long k[4] = {0, 0, 0, 0};
    if (live) {
        const longlong2 i01 = __ldg(reinterpret_cast<const longlong2*>(idx + tok));
        const longlong2 i23 = __ldg(reinterpret_cast<const longlong2*>(idx + tok) + 1);
        k[0] = i01.x; k[1] = i01.y; k[2] = i23.x; k[3] = i23.y;
#pragma unroll
        for (int t = 0; t < 4; ++t)
            if (k[t] < 0 || k[t] >= K) k[t] = 0;
    }
    if (hist != nullptr) {  // warp-aggregated histogram, one token slot of every lane at a time
        const unsigned active = __ballot_sync(0xffffffffu, live);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            if (live) {
                const unsigned peers = __match_any_sync(active, (int)k[t]);
                if ((threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(hist + k[t], (unsigned long long)__popc(peers));
            }
        }
    }
    if (!live) return;
    float4 e[4][DQ];
#pragma unroll
    for (int t = 0; t < 4; ++t)
#pragma unroll
        for (int c = 0; c < DQ; ++c) e[t][c] = __ldg(reinterpret_cast<const float4*>(E + (size_t)k[t] * D) + c);
    const int64_t b = tok / HW;
    const int64_t off = (b * D) * HW + (tok - b * HW);
    float4 zv[D], gz[D];
#pragma unroll
    for (int d = 0; d < D; ++d) zv[d] = __ldg(reinterpret_cast<const float4*>(z + off + (int64_t)d * HW));
#pragma unroll
    for (int d = 0; d < D; ++d)
        gz[d] = g_zq ? __ldg(reinterpret_cast<const float4*>(g_zq + off + (int64_t)d * HW)) : make_float4(0.f, 0.f, 0.f, 0.f);
    float cg[4][D];  // codebook-gradient terms [token][channel]
#pragma unroll
    for (int d = 0; d < D; ++d) {
        const float ev[4] = {f4_get(e[0][d >> 2], d & 3), f4_get(e[1][d >> 2], d & 3), f4_get(e[2][d >> 2], d & 3),
                             f4_get(e[3][d >> 2], d & 3)};
        const float zz[4] = {zv[d].x, zv[d].y, zv[d].z, zv[d].w};
        const float gg[4] = {gz[d].x, gz[d].y, gz[d].z, gz[d].w};
        float o[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const float tt = __fmul_rn(__fmul_rn(norm, __fsub_rn(zz[t], ev[t])), gv);
            o[t] = __fadd_rn(gg[t], tt);
            cg[t][d] = __fmul_rn(__fmul_rn(norm, __fsub_rn(ev[t], zz[t])), gbeta);
        }
        *reinterpret_cast<float4*>(dz_out + off + (int64_t)d * HW) = make_float4(o[0], o[1], o[2], o[3]);
    }
    if (dE != nullptr) {
#pragma unroll
        for (int t = 0; t < 4; ++t)
#pragma unroll
            for (int c = 0; c < DQ; ++c)
                red_add_v4(dE + (size_t)k[t] * D + 4 * c, cg[t][4 * c], cg[t][4 * c + 1], cg[t][4 * c + 2], cg[t][4 * c + 3]);
    }
}

static bool quad_ok(int D, int64_t HW, const void* a, const void* b, const void* c, const void* d) {
    const uintptr_t bits = reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(c) |
                           reinterpret_cast<uintptr_t>(d);
    return (D == 4 || D == 8) && HW % 4 == 0 && (bits & 15u) == 0;
}

// ---------------------------------------------------------------------------
// Warp-private variant of the tiled forward tail (default for D >= 128; vqb_tune "tail_warp").  A warp
// owns 32 consecutive tokens for ALL channels and transposes its codebook rows through a private 32x33 tile, so
// the only synchronisation is __syncwarp (the CTA-wide kernel above spends most of its stall time in barriers
// and on the L2 latency of the gathered rows, profiles/r01_ncu_full_helpers_c3.txt).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 2)
    gather_loss_st_warp_kernel(const float* __restrict__ z, const float* __restrict__ E, const int64_t* __restrict__ idx,
                               int64_t N, int D, int64_t HW, int K, float* __restrict__ zq_out,
                               double* __restrict__ partials, int* __restrict__ err_flag) {
    __shared__ float tiles[8][32][33];
    __shared__ double warp_part[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float (*tile)[33] = tiles[warp];
    const int64_t tok = ((int64_t)blockIdx.x * 8 + warp) * 32 + lane;
    const bool live = tok < N;
    int code = 0;
    int64_t off = 0;
    if (live) {
        int64_t kk = idx[tok];
        if (kk < 0 || kk >= K) {
            if (err_flag) *err_flag = 1;
            kk = 0;
        }
        code = (int)kk;
        const int64_t b = tok / HW;
        off = (b * D) * HW + (tok - b * HW);
    }
    float sq = 0.f;
    for (int d0 = 0; d0 < D; d0 += 32) {
        float zv[32];
        const float* zp = z + off + (int64_t)d0 * HW;
#pragma unroll
        for (int i = 0; i < 32; ++i) zv[i] = live ? __ldg(zp + (int64_t)i * HW) : 0.f;
        float rv[32];
#pragma unroll
        for (int t = 0; t < 32; ++t) {
            const int ct = __shfl_sync(0xffffffffu, code, t);
            rv[t] = __ldg(E + (size_t)ct * D + d0 + lane);  // row of token t, channel d0 + lane: one 128-byte line
        }
#pragma unroll
        for (int t = 0; t < 32; ++t) tile[lane][t] = rv[t];
        __syncwarp();
        if (live) {
            float* qp = zq_out + off + (int64_t)d0 * HW;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const float diff = __fsub_rn(tile[i][lane], zv[i]);
                qp[(int64_t)i * HW] = __fadd_rn(zv[i], diff);
                sq = fmaf(diff, diff, sq);
            }
        }
        __syncwarp();
    }
    double v = warp_sum_f64((double)sq);
    if (lane == 0) warp_part[warp] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double sum = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) sum += warp_part[w];
        partials[blockIdx.x] = sum;
    }
}

// Warp-private backward (vqb_tune "bwd_warp": 0 off (default), 1 = D >= 128, 2 = any D % 32 == 0): same idea as
// gather_loss_st_warp_kernel.  A warp owns 32 consecutive tokens; per 32-channel block it transposes its codebook
// rows through a private tile, streams z / g / dz with lane = token, leaves the codebook-gradient terms in the
// tile and scatters them with lane = channel: one 128-byte coalesced red.global.add.f32 per token.
__global__ void __launch_bounds__(256, 2)
    backward_warp_kernel(const float* __restrict__ z, const float* __restrict__ E, const int64_t* __restrict__ idx,
                         const float* __restrict__ g_zq, const float* __restrict__ g_vq, float beta, float norm, int64_t N,
                         int D, int64_t HW, int K, float* __restrict__ dz_out, float* __restrict__ dE,
                         unsigned long long* __restrict__ hist) {
    __shared__ float tiles[8][32][33];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float (*tile)[33] = tiles[warp];
    const int64_t tok = ((int64_t)blockIdx.x * 8 + warp) * 32 + lane;
    const bool live = tok < N;
    const float gv = g_vq ? __ldg(g_vq) : 0.f;
    const float gbeta = __fmul_rn(gv, beta);
    int code = 0;
    int64_t off = 0;
    if (live) {
        int64_t kk = idx[tok];
        if (kk < 0 || kk >= K) kk = 0;
        code = (int)kk;
        const int64_t b = tok / HW;
        off = (b * D) * HW + (tok - b * HW);
    }
    const unsigned active = __ballot_sync(0xffffffffu, live);
    if (hist != nullptr && live) {
        const unsigned peers = __match_any_sync(active, code);
        if (lane == __ffs(peers) - 1) atomicAdd(hist + code, (unsigned long long)__popc(peers));
    }
    for (int d0 = 0; d0 < D; d0 += 32) {
        float rv[32];
#pragma unroll
        for (int t = 0; t < 32; ++t) {
            const int ct = __shfl_sync(0xffffffffu, code, t);
            rv[t] = __ldg(E + (size_t)ct * D + d0 + lane);
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            float zv[16], gz[16];
            const int64_t base = off + (int64_t)(d0 + 16 * h) * HW;
#pragma unroll
            for (int i = 0; i < 16; ++i) zv[i] = live ? __ldg(z + base + (int64_t)i * HW) : 0.f;
#pragma unroll
            for (int i = 0; i < 16; ++i) gz[i] = (live && g_zq) ? __ldg(g_zq + base + (int64_t)i * HW) : 0.f;
            if (h == 0) {
#pragma unroll
                for (int t = 0; t < 32; ++t) tile[lane][t] = rv[t];
                __syncwarp();
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const float e = tile[16 * h + i][lane];
                const float tt = __fmul_rn(__fmul_rn(norm, __fsub_rn(zv[i], e)), gv);
                if (live) dz_out[base + (int64_t)i * HW] = __fadd_rn(gz[i], tt);
                tile[16 * h + i][lane] = __fmul_rn(__fmul_rn(norm, __fsub_rn(e, zv[i])), gbeta);
            }
        }
        __syncwarp();
        if (dE != nullptr) {
#pragma unroll 8
            for (int t = 0; t < 32; ++t) {
                const int ct = __shfl_sync(0xffffffffu, code, t);
                if ((active >> t) & 1u) atomicAdd(dE + (size_t)ct * D + d0 + lane, tile[lane][t]);
            }
        }
        __syncwarp();
    }
}

// largest pass width (channels) that divides D; the backward keeps z and g in registers -> half of it
static int tiled_pass_width(int D, int cap) {
    for (int dc = cap; dc >= 64; dc -= 64)
        if (D % dc == 0) return dc;
    return 32;  // callers guarantee D % 32 == 0
}

// ---------------------------------------------------------------------------
// get_codebook_entry
// ---------------------------------------------------------------------------
template <bool kVec4>
__global__ void __launch_bounds__(kTailThreads)
    gather_kernel(const float* __restrict__ E, const int64_t* __restrict__ idx, int64_t N, int D,
                  int64_t HW, int K, int dims_per_slice, float* __restrict__ out,
                  int* __restrict__ err_flag) {
    const int64_t tok = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tok >= N) return;
    const int d0 = threadIdx.y * dims_per_slice;
    int64_t k = idx[tok];
    if (k < 0 || k >= K) {
        if (err_flag) *err_flag = 1;
        k = 0;
    }
    const int64_t b = tok / HW;
    const int64_t off = (b * D) * HW + (tok - b * HW);
    const float* erow = E + (size_t)k * D;
    if constexpr (kVec4) {
        for (int d = d0; d < d0 + dims_per_slice; d += 4) {
            const float4 ev = __ldg(reinterpret_cast<const float4*>(erow + d));
            out[off + (int64_t)(d + 0) * HW] = ev.x;
            out[off + (int64_t)(d + 1) * HW] = ev.y;
            out[off + (int64_t)(d + 2) * HW] = ev.z;
            out[off + (int64_t)(d + 3) * HW] = ev.w;
        }
    } else {
        for (int d = d0; d < d0 + dims_per_slice; ++d) out[off + (int64_t)d * HW] = __ldg(erow + d);
    }
}

// ---------------------------------------------------------------------------
// get_codebook_usage
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    hist_kernel(const int64_t* __restrict__ idx, int64_t N, int K, unsigned long long* __restrict__ hist,
                int* __restrict__ err_flag) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t base = (int64_t)blockIdx.x * blockDim.x; base < N; base += stride) {
        const int64_t i = base + threadIdx.x;
        const bool live = i < N;
        int64_t k = live ? idx[i] : 0;
        bool ok = live;
        if (live && (k < 0 || k >= K)) {
            if (err_flag) *err_flag = 1;
            ok = false;
        }
        const unsigned active = __ballot_sync(0xffffffffu, ok);
        if (ok) {
            const unsigned peers = __match_any_sync(active, (int)k);
            if ((threadIdx.x & 31) == __ffs(peers) - 1)
                atomicAdd(hist + k, (unsigned long long)__popc(peers));
        }
    }
}

__global__ void __launch_bounds__(256)
    count_used_kernel(const unsigned long long* __restrict__ hist, int K,
                      unsigned long long* __restrict__ used) {
    int c = 0;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < K; k += gridDim.x * blockDim.x)
        c += hist[k] != 0ull;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(used, (unsigned long long)c);
}

// ---------------------------------------------------------------------------
// extensions
// ---------------------------------------------------------------------------
template <bool kVec4>
__global__ void __launch_bounds__(kTailThreads)
    code_sums_kernel(const float* __restrict__ z, const int64_t* __restrict__ idx, int64_t N, int D,
                     int64_t HW, int K, int dims_per_slice, float* __restrict__ counts,
                     float* __restrict__ sums) {
    const int64_t tok = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int d0 = threadIdx.y * dims_per_slice;
    const bool live = tok < N;
    int64_t k = 0;
    if (live) {
        k = idx[tok];
        if (k < 0 || k >= K) k = 0;
    }
    if (threadIdx.y == 0) {
        const unsigned active = __ballot_sync(0xffffffffu, live);
        if (live) {
            const unsigned peers = __match_any_sync(active, (int)k);
            if ((threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(counts + k, (float)__popc(peers));
        }
    }
    if (!live) return;
    const int64_t b = tok / HW;
    const int64_t off = (b * D) * HW + (tok - b * HW);
    float* srow = sums + (size_t)k * D;
    if constexpr (kVec4) {
        for (int d = d0; d < d0 + dims_per_slice; d += 4) {
            float zv[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) zv[j] = __ldg(z + off + (int64_t)(d + j) * HW);
            red_add_v4(srow + d, zv[0], zv[1], zv[2], zv[3]);
        }
    } else {
        for (int d = d0; d < d0 + dims_per_slice; ++d)
            atomicAdd(srow + d, __ldg(z + off + (int64_t)d * HW));
    }
}

// ONE block, fixed summation order (fp64 per thread over a fixed stride, then a fixed tree): the total -- and with
// it the Laplace-smoothed update -- is bit-identical from run to run and from rank to rank
__global__ void __launch_bounds__(1024) ema_sizes_kernel(float* __restrict__ cluster_size,
                                                         const float* __restrict__ counts, int K, float decay,
                                                         float* __restrict__ total) {
    __shared__ double part_s[32];
    double part = 0.0;
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        const float v = cluster_size[k] * decay + counts[k] * (1.f - decay);
        cluster_size[k] = v;
        part += (double)v;
    }
    part = warp_sum_f64(part);
    if ((threadIdx.x & 31) == 0) part_s[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x < 32) {
        double v = threadIdx.x < (blockDim.x >> 5) ? part_s[threadIdx.x] : 0.0;
        v = warp_sum_f64(v);
        if (threadIdx.x == 0) *total = (float)v;
    }
}

__global__ void ema_embed_kernel(float* __restrict__ E, const float* __restrict__ cluster_size,
                                 float* __restrict__ embed_sum, const float* __restrict__ sums, int K,
                                 int D, float decay, float eps, const float* __restrict__ total) {
    const size_t n = (size_t)K * D;
    const float tot = *total;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x) {
        const int k = (int)(i / D);
        const float m = embed_sum[i] * decay + sums[i] * (1.f - decay);
        embed_sum[i] = m;
        const float smoothed = (cluster_size[k] + eps) / (tot + K * eps) * tot;
        E[i] = m / smoothed;
    }
}

__global__ void pack_keys_kernel(const float* __restrict__ dmin, const int64_t* __restrict__ idx,
                                 int64_t n, int64_t index_offset, int64_t* __restrict__ keys) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // ATen's argmin treats NaN as minimal (quantizer.py:76): a NaN score packs as the smallest key, so the MIN over
    // shards picks the NaN with the lowest global index; -0.0 is canonicalised to +0.0 (equal scores must tie)
    const float dv = dmin[i];
    int bits = (dv != dv) ? (int)0x80000000 : __float_as_int(dv + 0.f);
    if (bits < 0 && dv == dv) bits ^= 0x7fffffff;  // monotone signed order of IEEE floats
    keys[i] = ((int64_t)bits << 32) | (int64_t)(uint32_t)(idx[i] + index_offset);
}

__global__ void unpack_keys_kernel(const int64_t* __restrict__ keys, int64_t n,
                                   int64_t* __restrict__ idx_out, float* __restrict__ dmin_out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t key = keys[i];
    idx_out[i] = (int64_t)(uint32_t)(key & 0xffffffffll);
    if (dmin_out) {
        int bits = (int)(key >> 32);
        if (bits == (int)0x80000000) bits = 0x7fc00000;  // the NaN sentinel of pack_keys_kernel
        else if (bits < 0) bits ^= 0x7fffffff;
        dmin_out[i] = __int_as_float(bits);
    }
}

}  // namespace vqb

// ===========================================================================
// C ABI
// ===========================================================================
using namespace vqb;

static bool vec4_ok(int D, const void* E) {
    return (D % 4 == 0) && ((reinterpret_cast<uintptr_t>(E) & 15u) == 0);
}

static int check_shape(int64_t B, int D, int64_t HW, int K) {
    if (B < 0 || HW < 0 || D <= 0 || K <= 0) {
        set_error("invalid shape B=%lld D=%d HW=%lld K=%d", (long long)B, D, (long long)HW, K);
        return VQB_ERR_INVALID_ARG;
    }
    return VQB_OK;
}

VQB_KNOB g_bwd_pass_cap = 64;   // measured best on B200 (profiles/r01_tail_pass_width_sweep.txt)
VQB_KNOB g_fwd_pass_cap = 64;
VQB_KNOB g_bwd_warp = 0;    // warp-private backward (vqb_tune "bwd_warp")
VQB_KNOB g_tail_warp = 1;   // warp-private forward tail (vqb_tune "tail_warp": 0 off, 1 auto = D >= 128, 2 force)
VQB_KNOB g_tail_pipe = 1;   // pipelined 128-token forward tail for D >= 128, D % 64 == 0 (vqb_tune "tail_pipe", 0 = off)
VQB_KNOB g_bwd_pipe = 1;    // pipelined 128-token backward, same shapes (vqb_tune "bwd_pipe", 0 = off)
VQB_KNOB g_tail_tok128 = 1;  // 128-token float4 kernels when the layout allows (vqb_tune "tail_tok128", 0 = off)
#ifdef VQB_EXPERIMENTAL
namespace vqb {
void set_tail_knob(const char* key, int value) {
    if (key[0] == 't' && key[5] == 'p') g_tail_pipe = value;          // tail_pipe
    else if (key[0] == 'b' && key[4] == 'p' && key[5] == 'i') g_bwd_pipe = value;  // bwd_pipe
    else if (key[0] == 'b' && key[4] == 'p') g_bwd_pass_cap = value;  // bwd_pass_channels
    else if (key[0] == 'f') g_fwd_pass_cap = value;                   // fwd_pass_channels
    else if (key[0] == 'b') g_bwd_warp = value;                       // bwd_warp
    else if (key[5] == 'w') g_tail_warp = value;                      // tail_warp
    else g_tail_tok128 = value;                                       // tail_tok128
}
}  // namespace vqb
#endif

extern "C" size_t vqb_tail_partials_bytes(int64_t n_tokens) {
    return sizeof(double) * (size_t)((n_tokens + 31) / 32 + 1);
}

extern "C" int vqb_gather_loss_st_f32(const float* z, const float* E, const int64_t* idx, int64_t B,
                                      int D, int64_t HW, int K, float beta, float* zq_out,
                                      float* loss_out, void* partials, size_t partials_bytes,
                                      int* err_flag, vqb_stream_t stream) {
    VQB_DEVICE_TRY();
    if (int rc = check_shape(B, D, HW, K)) return rc;
    if (!z || !E || !idx || !zq_out || !loss_out || !partials) {
        set_error("vqb_gather_loss_st_f32: null pointer");
        return VQB_ERR_INVALID_ARG;
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int64_t N = B * HW;
    const TailShape sh = tail_shape(D);
    const int64_t blocks = (N + sh.tx - 1) / sh.tx;
    if (partials_bytes < sizeof(double) * (size_t)(blocks + 1)) {
        set_error("partials scratch too small: %zu < %zu", partials_bytes,
                  sizeof(double) * (size_t)(blocks + 1));
        return VQB_ERR_WORKSPACE;
    }
    if (N == 0) {
        VQB_CUDA_TRY(cudaMemsetAsync(loss_out, 0xff, 2 * sizeof(float), s));  // NaN like mean of empty
        return VQB_OK;
    }
    const dim3 block(sh.tx, sh.slices);
    double* parts = static_cast<double*>(partials);
    // measured (scripts/tail_ab.py): 0.083 -> 0.046 ms at D=32, 0.146 -> 0.138 at D=64 (512 K / 1 M tokens), but
    // 0.46 -> 0.59 ms at D=256 where the four fill/barrier rounds per CTA dominate: small D only
    if (g_tail_tok128 && D <= 64 && tok128_ok(D, HW, z, zq_out)) {
        const int64_t tb = (N + kTok128 - 1) / kTok128;
        if (D % 64 == 0)
            gather_loss_st_tok128_kernel<64><<<(unsigned)tb, 256, 0, s>>>(z, E, idx, N, D, HW, K, zq_out, parts, err_flag);
        else
            gather_loss_st_tok128_kernel<32><<<(unsigned)tb, 256, 0, s>>>(z, E, idx, N, D, HW, K, zq_out, parts, err_flag);
        VQB_LAUNCH_CHECK("gather_loss_st_tok128_kernel");
        loss_finalize_kernel<<<1, 256, 0, s>>>(parts, tb, 1.0 / ((double)N * D), beta, loss_out);
        VQB_LAUNCH_CHECK("loss_finalize_kernel");
        return VQB_OK;
    }
    if (g_tail_pipe && D >= 128 && pipe_ok(D, HW, z, zq_out, E, E)) {
        const int64_t pb = (N + kTok128 - 1) / kTok128;
        if (g_tail_pipe == 2) {  // 3 CTAs per SM (80 registers, small spill): A/B only, measured slower (0.46 vs 0.38 ms)
            VQB_CUDA_TRY(cudaFuncSetAttribute(gather_loss_st_pipe_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              (int)kPipeSmemBytes));
            gather_loss_st_pipe_kernel<3><<<(unsigned)pb, 256, kPipeSmemBytes, s>>>(z, E, idx, N, D, HW, K, zq_out, parts,
                                                                                   err_flag);
        } else {
            VQB_CUDA_TRY(cudaFuncSetAttribute(gather_loss_st_pipe_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              (int)kPipeSmemBytes));
            gather_loss_st_pipe_kernel<2><<<(unsigned)pb, 256, kPipeSmemBytes, s>>>(z, E, idx, N, D, HW, K, zq_out, parts,
                                                                                   err_flag);
        }
        VQB_LAUNCH_CHECK("gather_loss_st_pipe_kernel");
        loss_finalize_kernel<<<1, 256, 0, s>>>(parts, pb, 1.0 / ((double)N * D), beta, loss_out);
        VQB_LAUNCH_CHECK("loss_finalize_kernel");
        return VQB_OK;
    }
    // measured (scripts/tail_warp_ab.py): 0.44 ms / 4.9 TB/s at D=256 (CTA-wide kernel 0.46-0.50), 0.227 vs 0.247 ms at
    // D=128, no gain at D=64: large D only (g_tail_warp: 1 = auto, 0 = off, 2 = force for any D % 32 == 0)
    if (D % 32 == 0 && ((g_tail_warp == 1 && D >= 128) || g_tail_warp == 2)) {
        const int64_t wb = (N + 255) / 256;  // 8 warps x 32 tokens per CTA
        gather_loss_st_warp_kernel<<<(unsigned)wb, 256, 0, s>>>(z, E, idx, N, D, HW, K, zq_out, parts, err_flag);
        VQB_LAUNCH_CHECK("gather_loss_st_warp_kernel");
        loss_finalize_kernel<<<1, 256, 0, s>>>(parts, wb, 1.0 / ((double)N * D), beta, loss_out);
        VQB_LAUNCH_CHECK("loss_finalize_kernel");
        return VQB_OK;
    }
    if (D % 32 == 0) {
        const int64_t tb = (N + kTileTok - 1) / kTileTok;
#define VQB_GATHER(dc) \
    gather_loss_st_tiled_kernel<dc><<<(unsigned)tb, 256, 0, s>>>(z, E, idx, N, D, HW, K, zq_out, parts, err_flag)
        switch (tiled_pass_width(D, g_fwd_pass_cap)) {
            case 256: VQB_GATHER(256); break;
            case 192: VQB_GATHER(192); break;
            case 128: VQB_GATHER(128); break;
            case 64: VQB_GATHER(64); break;
            default: VQB_GATHER(32); break;
        }
#undef VQB_GATHER
        VQB_LAUNCH_CHECK("gather_loss_st_tiled_kernel");
        loss_finalize_kernel<<<1, 256, 0, s>>>(parts, tb, 1.0 / ((double)N * D), beta, loss_out);
        VQB_LAUNCH_CHECK("loss_finalize_kernel");
        return VQB_OK;
    }
    if (g_tail_tok128 && quad_ok(D, HW, z, zq_out, E, idx)) {
        const int64_t qb = (N / 4 + 255) / 256;  // <= blocks: the partials buffer is large enough
        if (D == 4)
            gather_loss_st_quad_kernel<1><<<(unsigned)qb, 256, 0, s>>>(z, E, idx, N, HW, K, zq_out, parts, err_flag);
        else
            gather_loss_st_quad_kernel<2><<<(unsigned)qb, 256, 0, s>>>(z, E, idx, N, HW, K, zq_out, parts, err_flag);
        VQB_LAUNCH_CHECK("gather_loss_st_quad_kernel");
        loss_finalize_kernel<<<1, 256, 0, s>>>(parts, qb, 1.0 / ((double)N * D), beta, loss_out);
        VQB_LAUNCH_CHECK("loss_finalize_kernel");
        return VQB_OK;
    }
    if (vec4_ok(D, E))
        gather_loss_st_kernel<true><<<(unsigned)blocks, block, 0, s>>>(
            z, E, idx, N, D, HW, K, sh.dims_per_slice, zq_out, parts, err_flag);
    else
        gather_loss_st_kernel<false><<<(unsigned)blocks, block, 0, s>>>(
            z, E, idx, N, D, HW, K, sh.dims_per_slice, zq_out, parts, err_flag);
    VQB_LAUNCH_CHECK("gather_loss_st_kernel");
    loss_finalize_kernel<<<1, 256, 0, s>>>(parts, blocks, 1.0 / ((double)N * D), beta, loss_out);
    VQB_LAUNCH_CHECK("loss_finalize_kernel");
    return VQB_OK;
}

extern "C" int vqb_backward_f32(const float* z, const float* E, const int64_t* idx, const float* g_zq,
                                const float* g_vq, float beta, int64_t B, int D, int64_t HW, int K,
                                float* dz_out, float* dE_accum, int64_t* hist_accum,
                                vqb_stream_t stream) {
    VQB_DEVICE_TRY();
    if (int rc = check_shape(B, D, HW, K)) return rc;
    if (!z || !E || !idx || !dz_out) {
        set_error("vqb_backward_f32: null pointer");
        return VQB_ERR_INVALID_ARG;
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int64_t N = B * HW;
    if (N == 0) return VQB_OK;
    const TailShape sh = tail_shape(D);
    const int64_t blocks = (N + sh.tx - 1) / sh.tx;
    const dim3 block(sh.tx, sh.slices);
    const float norm = (float)(2.0 / ((double)N * D));
    unsigned long long* hist = reinterpret_cast<unsigned long long*>(hist_accum);
    const bool v4 = vec4_ok(D, E) && (!dE_accum || (reinterpret_cast<uintptr_t>(dE_accum) & 15u) == 0);
    if (g_bwd_pipe && D >= 128 && pipe_ok(D, HW, z, dz_out, g_zq, dE_accum)) {
        const int64_t pb = (N + kTok128 - 1) / kTok128;
        if (g_bwd_pipe == 2) {  // direct 16-byte reductions from registers (A/B: slower, hot rows serialise)
            VQB_CUDA_TRY(cudaFuncSetAttribute(backward_pipe_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              (int)kPipeSmemBytes));
            backward_pipe_kernel<2, false><<<(unsigned)pb, 256, kPipeSmemBytes, s>>>(z, E, idx, g_zq, g_vq, beta, norm, N, D, HW,
                                                                                    K, dz_out, dE_accum, hist);
        } else {
            VQB_CUDA_TRY(cudaFuncSetAttribute(backward_pipe_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              (int)kPipeSmemBytes));
            backward_pipe_kernel<2, true><<<(unsigned)pb, 256, kPipeSmemBytes, s>>>(z, E, idx, g_zq, g_vq, beta, norm, N, D, HW,
                                                                                   K, dz_out, dE_accum, hist);
        }
        VQB_LAUNCH_CHECK("backward_pipe_kernel");
        return VQB_OK;
    }
    if (D % 32 == 0 && ((g_bwd_warp == 1 && D >= 128) || g_bwd_warp == 2)) {
        const int64_t wb = (N + 255) / 256;
        backward_warp_kernel<<<(unsigned)wb, 256, 0, s>>>(z, E, idx, g_zq, g_vq, beta, norm, N, D, HW, K, dz_out, dE_accum,
                                                         hist);
        VQB_LAUNCH_CHECK("backward_warp_kernel");
        return VQB_OK;
    }
    if (D % 32 == 0 && (!dE_accum || (reinterpret_cast<uintptr_t>(dE_accum) & 15u) == 0)) {
        const int64_t tb = (N + kTileTok - 1) / kTileTok;
#define VQB_BWD(dc)                                                                                            \
    backward_tiled_kernel<dc><<<(unsigned)tb, 256, 0, s>>>(z, E, idx, g_zq, g_vq, beta, norm, N, D, HW, K, dz_out, \
                                                           dE_accum, hist)
        switch (tiled_pass_width(D, g_bwd_pass_cap)) {
            case 256: VQB_BWD(256); break;
            case 192: VQB_BWD(192); break;
            case 128: VQB_BWD(128); break;
            case 64: VQB_BWD(64); break;
            default: VQB_BWD(32); break;
        }
#undef VQB_BWD
        VQB_LAUNCH_CHECK("backward_tiled_kernel");
        return VQB_OK;
    }
    // (measured: 18 -> 17 us at D=4, but 39 -> 43 us at D=8: the 32+32 float4 registers of z and g cost occupancy)
    if (g_tail_tok128 && D == 4 && quad_ok(D, HW, z, dz_out, E, idx) && (!g_zq || (reinterpret_cast<uintptr_t>(g_zq) & 15u) == 0) &&
        (!dE_accum || (reinterpret_cast<uintptr_t>(dE_accum) & 15u) == 0)) {
        const int64_t qb = (N / 4 + 255) / 256;
        if (D == 4)
            backward_quad_kernel<1><<<(unsigned)qb, 256, 0, s>>>(z, E, idx, g_zq, g_vq, beta, norm, N, HW, K, dz_out, dE_accum,
                                                                hist);
        else
            backward_quad_kernel<2><<<(unsigned)qb, 256, 0, s>>>(z, E, idx, g_zq, g_vq, beta, norm, N, HW, K, dz_out, dE_accum,
                                                                hist);
        VQB_LAUNCH_CHECK("backward_quad_kernel");
        return VQB_OK;
    }
    if (v4)
        backward_kernel<true><<<(unsigned)blocks, block, 0, s>>>(
            z, E, idx, g_zq, g_vq, beta, norm, N, D, HW, K, sh.dims_per_slice, dz_out, dE_accum, hist);
    else
        backward_kernel<false><<<(unsigned)blocks, block, 0, s>>>(
            z, E, idx, g_zq, g_vq, beta, norm, N, D, HW, K, sh.dims_per_slice, dz_out, dE_accum, hist);
    VQB_LAUNCH_CHECK("backward_kernel");
    return VQB_OK;
}

extern "C" int vqb_gather_f32(const float* E, const int64_t* idx, int64_t B, int D, int64_t HW, int K,
                              float* out, int* err_flag, vqb_stream_t stream) {
    VQB_DEVICE_TRY();
    if (int rc = check_shape(B, D, HW, K)) return rc;
    if (!E || !idx || !out) {
        set_error("vqb_gather_f32: null pointer");
        return VQB_ERR_INVALID_ARG;
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int64_t N = B * HW;
    if (N == 0) return VQB_OK;
    const TailShape sh = tail_shape(D);
    const int64_t blocks = (N + sh.tx - 1) / sh.tx;
    const dim3 block(sh.tx, sh.slices);
    if (vec4_ok(D, E))
        gather_kernel<true><<<(unsigned)blocks, block, 0, s>>>(E, idx, N, D, HW, K, sh.dims_per_slice,
                                                              out, err_flag);
    else
        gather_kernel<false><<<(unsigned)blocks, block, 0, s>>>(E, idx, N, D, HW, K, sh.dims_per_slice,
                                                               out, err_flag);
    VQB_LAUNCH_CHECK("gather_kernel");
    return VQB_OK;
}

extern "C" int vqb_hist_i64(const int64_t* idx, int64_t n_tokens, int K, int64_t* hist_out,
                            int64_t* used_out, int* err_flag, vqb_stream_t stream) {
    VQB_DEVICE_TRY();
    if (n_tokens < 0 || K <= 0 || !hist_out || (!idx && n_tokens > 0)) {
        set_error("vqb_hist_i64: invalid argument");
        return VQB_ERR_INVALID_ARG;
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    VQB_CUDA_TRY(cudaMemsetAsync(hist_out, 0, sizeof(int64_t) * (size_t)K, s));
    if (used_out) VQB_CUDA_TRY(cudaMemsetAsync(used_out, 0, sizeof(int64_t), s));
    unsigned long long* hist = reinterpret_cast<unsigned long long*>(hist_out);
    if (n_tokens > 0) {
        int64_t blocks = (n_tokens + 255) / 256;
        const int64_t cap = (int64_t)sm_count() * 16;
        if (blocks > cap) blocks = cap;
        hist_kernel<<<(unsigned)blocks, 256, 0, s>>>(idx, n_tokens, K, hist, err_flag);
        VQB_LAUNCH_CHECK("hist_kernel");
    }
    if (used_out) {
        int blocks = (K + 255) / 256;
        if (blocks > 1024) blocks = 1024;
        count_used_kernel<<<blocks, 256, 0, s>>>(hist, K, reinterpret_cast<unsigned long long*>(used_out));
        VQB_LAUNCH_CHECK("count_used_kernel");
    }
    return VQB_OK;
}

extern "C" int vqb_code_sums_f32(const float* z, const int64_t* idx, int64_t B, int D, int64_t HW, int K,
                                 float* counts_accum, float* sums_accum, vqb_stream_t stream) {
    VQB_DEVICE_TRY();
    if (int rc = check_shape(B, D, HW, K)) return rc;
    if (!z || !idx || !counts_accum || !sums_accum) {
        set_error("vqb_code_sums_f32: null pointer");
        return VQB_ERR_INVALID_ARG;
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int64_t N = B * HW;
    if (N == 0) return VQB_OK;
    const TailShape sh = tail_shape(D);
    const int64_t blocks = (N + sh.tx - 1) / sh.tx;
    const dim3 block(sh.tx, sh.slices);
    if (vec4_ok(D, sums_accum))
        code_sums_kernel<true><<<(unsigned)blocks, block, 0, s>>>(z, idx, N, D, HW, K, sh.dims_per_slice,
                                                                 counts_accum, sums_accum);
    else
        code_sums_kernel<false><<<(unsigned)blocks, block, 0, s>>>(z, idx, N, D, HW, K, sh.dims_per_slice,
                                                                  counts_accum, sums_accum);
    VQB_LAUNCH_CHECK("code_sums_kernel");
    return VQB_OK;
}

extern "C" int vqb_ema_update_f32(float* E, float* cluster_size, float* embed_sum, const float* counts,
                                  const float* sums, int K, int D, float decay, float eps,
                                  float* total_scratch, vqb_stream_t stream) {
    VQB_DEVICE_TRY();
    if (K <= 0 || D <= 0 || !E || !cluster_size || !embed_sum || !counts || !sums || !total_scratch) {
        set_error("vqb_ema_update_f32: invalid argument");
        return VQB_ERR_INVALID_ARG;
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    ema_sizes_kernel<<<1, 1024, 0, s>>>(cluster_size, counts, K, decay, total_scratch);
    VQB_LAUNCH_CHECK("ema_sizes_kernel");
    size_t n = (size_t)K * D;
    size_t b2 = (n + 255) / 256;
    if (b2 > (size_t)sm_count() * 16) b2 = (size_t)sm_count() * 16;
    ema_embed_kernel<<<(unsigned)b2, 256, 0, s>>>(E, cluster_size, embed_sum, sums, K, D, decay, eps,
                                                  total_scratch);
    VQB_LAUNCH_CHECK("ema_embed_kernel");
    return VQB_OK;
}

extern "C" int vqb_pack_argmin_keys(const float* dmin, const int64_t* idx, int64_t n,
                                    int64_t index_offset, int64_t* keys_out, vqb_stream_t stream) {
    VQB_DEVICE_TRY();
    if (n < 0 || (n > 0 && (!dmin || !idx || !keys_out))) {
        set_error("vqb_pack_argmin_keys: invalid argument");
        return VQB_ERR_INVALID_ARG;
    }
    if (n == 0) return VQB_OK;
    pack_keys_kernel<<<(unsigned)((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        dmin, idx, n, index_offset, keys_out);
    VQB_LAUNCH_CHECK("pack_keys_kernel");
    return VQB_OK;
}

extern "C" int vqb_unpack_argmin_keys(const int64_t* keys, int64_t n, int64_t* idx_out, float* dmin_out,
                                      vqb_stream_t stream) {
    VQB_DEVICE_TRY();
    if (n < 0 || (n > 0 && (!keys || !idx_out))) {
        set_error("vqb_unpack_argmin_keys: invalid argument");
        return VQB_ERR_INVALID_ARG;
    }
    if (n == 0) return VQB_OK;
    unpack_keys_kernel<<<(unsigned)((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        keys, n, idx_out, dmin_out);
    VQB_LAUNCH_CHECK("unpack_keys_kernel");
    return VQB_OK;
}
