// Instruction-mix microbenchmarks behind the low-D search design (DESIGN.md "FMA roofline").
// Each mode replays the inner loop of search_lowd_kernel<4> on one 2048-code tile held in shared
// memory, with parts of the instruction mix removed, so the cost of each ingredient can be read
// off the B200 directly:
//   mode 0  FFMA2 chains only (accumulators never reset; no minimum)
//   mode 1  FFMA2 + FMNMX3 (the shipped mix)
//   mode 2  scalar FFMA chains only
//   mode 3  scalar FFMA + FMNMX3
//   mode 4  FFMA2 + two 2-input FMNMX
//   mode 5  FFMA2 + FMNMX3 applied one code pair late (software-pipelined minimum)
// flops reported = 2 * 4 dims * tokens * codes (the algorithmic count of the search).
#include "vqb_common.cuh"

namespace vqb {

constexpr int kUbCodes = 2048;

template <int MODE, int T>
__global__ void __launch_bounds__(256, 2) ubench_kernel(int sweeps, const float* __restrict__ src, float* sink) {
    __shared__ __align__(16) float tile[kUbCodes * 5];
    for (int i = threadIdx.x; i < kUbCodes * 5; i += blockDim.x) tile[i] = src[i];
    __syncthreads();
    float m[T];
    float nzs[T][4];
    unsigned long long nz[T][4];
    unsigned long long acc2[T];
    float accs[T][2];
#pragma unroll
    for (int t = 0; t < T; ++t) {
        m[t] = INFINITY;
        acc2[t] = 0ull;
        accs[t][0] = accs[t][1] = 0.f;
#pragma unroll
        for (int d = 0; d < 4; ++d) {
            nzs[t][d] = src[(threadIdx.x * T + t) * 4 + d];
            nz[t][d] = pack_f32x2(nzs[t][d], nzs[t][d]);
        }
    }
    const float* hbuf = tile + kUbCodes * 4;
    for (int s = 0; s < sweeps; ++s) {
#pragma unroll 4
        for (int p = 0; p < kUbCodes / 2; ++p) {
            const ulonglong2 v0 = *reinterpret_cast<const ulonglong2*>(tile + p * 8);
            const ulonglong2 v1 = *reinterpret_cast<const ulonglong2*>(tile + p * 8 + 4);
            const unsigned long long h2 = *reinterpret_cast<const unsigned long long*>(hbuf + 2 * p);
            if constexpr (MODE == 4) {
#pragma unroll
                for (int t = 0; t < T; ++t) {
                    unsigned long long a = fma_f32x2(nz[t][0], v0.x, h2);
                    a = fma_f32x2(nz[t][1], v0.y, a);
                    a = fma_f32x2(nz[t][2], v1.x, a);
                    a = fma_f32x2(nz[t][3], v1.y, a);
                    float x, y;
                    unpack_f32x2(a, x, y);
                    m[t] = fminf(m[t], fminf(x, y));
                }
            } else if constexpr (MODE == 5) {
#pragma unroll
                for (int t = 0; t < T; ++t) {
                    float x, y;
                    unpack_f32x2(acc2[t], x, y);  // previous pair's scores
                    unsigned long long a = fma_f32x2(nz[t][0], v0.x, h2);
                    a = fma_f32x2(nz[t][1], v0.y, a);
                    m[t] = min3_f32(m[t], x, y);
                    a = fma_f32x2(nz[t][2], v1.x, a);
                    a = fma_f32x2(nz[t][3], v1.y, a);
                    acc2[t] = a;
                }
            } else if constexpr (MODE == 0 || MODE == 1) {
#pragma unroll
                for (int t = 0; t < T; ++t) {
                    unsigned long long a = fma_f32x2(nz[t][0], v0.x, MODE == 1 ? h2 : acc2[t]);
                    a = fma_f32x2(nz[t][1], v0.y, a);
                    a = fma_f32x2(nz[t][2], v1.x, a);
                    a = fma_f32x2(nz[t][3], v1.y, a);
                    if constexpr (MODE == 1) {
                        float x, y;
                        unpack_f32x2(a, x, y);
                        m[t] = min3_f32(m[t], x, y);
                    } else {
                        acc2[t] = a;
                    }
                }
            } else {
                float e[8], h[2];
                unpack_f32x2(v0.x, e[0], e[1]);
                unpack_f32x2(v0.y, e[2], e[3]);
                unpack_f32x2(v1.x, e[4], e[5]);
                unpack_f32x2(v1.y, e[6], e[7]);
                unpack_f32x2(h2, h[0], h[1]);
#pragma unroll
                for (int t = 0; t < T; ++t) {
                    float a0 = fmaf(nzs[t][0], e[0], MODE == 3 ? h[0] : accs[t][0]);
                    float a1 = fmaf(nzs[t][0], e[1], MODE == 3 ? h[1] : accs[t][1]);
#pragma unroll
                    for (int d = 1; d < 4; ++d) {
                        a0 = fmaf(nzs[t][d], e[2 * d], a0);
                        a1 = fmaf(nzs[t][d], e[2 * d + 1], a1);
                    }
                    if constexpr (MODE == 3) {
                        m[t] = min3_f32(m[t], a0, a1);
                    } else {
                        accs[t][0] = a0;
                        accs[t][1] = a1;
                    }
                }
            }
        }
    }
    float r = 0.f;
#pragma unroll
    for (int t = 0; t < T; ++t) {
        float x, y;
        unpack_f32x2(acc2[t], x, y);
        r += m[t] + x + y + accs[t][0] + accs[t][1];
    }
    if (r == 123.456f) sink[0] = r;
}

// Access-pattern ceiling of the tiled tail kernels: the same CTA mapping (32 consecutive tokens x 8 warps, each
// thread walks channels ty, ty+8, ... of its token, 128-byte requests 4*HW bytes apart) as a pure copy, i.e.
// without the codebook gather, the transposing tile and the barriers.  mode 0: out = in; mode 1: out = a + b
// (two input streams like the backward); mode 2: linear float4 copy of the same bytes for comparison.
__global__ void __launch_bounds__(256)
    pattern_copy_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, int64_t N,
                        int D, int64_t HW, int mode) {
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    if (mode == 2) {
        const int64_t n4 = N * D / 4;
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x)
            reinterpret_cast<float4*>(out)[i] = __ldg(reinterpret_cast<const float4*>(a) + i);
        return;
    }
    if (mode >= 3) {  // 3: out = a, 4: out = a + b with 128 tokens per CTA, one float4 (4 tokens) per lane
        const int64_t tok4 = (int64_t)blockIdx.x * 128 + 4 * tx;
        if (tok4 >= N) return;
        const int64_t b4 = tok4 / HW;
        const int64_t off4 = (b4 * D) * HW + (tok4 - b4 * HW) + (int64_t)ty * HW;
        const int64_t step4 = 8 * HW;
        for (int d0 = 0; d0 < D; d0 += 64) {
            float4 v[8], w[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = __ldg(reinterpret_cast<const float4*>(a + off4 + (int64_t)d0 * HW + i * step4));
            if (mode == 4) {
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    w[i] = __ldg(reinterpret_cast<const float4*>(b + off4 + (int64_t)d0 * HW + i * step4));
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float4 r = v[i];
                if (mode == 4) {
                    r.x += w[i].x;
                    r.y += w[i].y;
                    r.z += w[i].z;
                    r.w += w[i].w;
                }
                *reinterpret_cast<float4*>(out + off4 + (int64_t)d0 * HW + i * step4) = r;
            }
        }
        return;
    }
    const int64_t tok = (int64_t)blockIdx.x * 32 + tx;
    if (tok >= N) return;
    const int64_t bi = tok / HW;
    const int64_t off = (bi * D) * HW + (tok - bi * HW) + (int64_t)ty * HW;
    const int64_t step = 8 * HW;
    for (int d0 = 0; d0 < D; d0 += 64) {
        float v[8], w[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = __ldg(a + off + (int64_t)d0 * HW + i * step);
        if (mode == 1) {
#pragma unroll
            for (int i = 0; i < 8; ++i) w[i] = __ldg(b + off + (int64_t)d0 * HW + i * step);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) out[off + (int64_t)d0 * HW + i * step] = mode == 1 ? v[i] + w[i] : v[i];
    }
}

}  // namespace vqb

using namespace vqb;

extern "C" int vqb_ubench_copy(const float* a, const float* b, float* out, int64_t B, int D, int64_t HW, int mode,
                               vqb_stream_t stream) {
    VQB_DEVICE_TRY();
    if (!a || !out || ((mode == 1 || mode == 4) && !b) || B <= 0 || D <= 0 || D % 64 != 0 || HW <= 0 || HW % 4 != 0 || mode < 0 ||
        mode > 4) {
        set_error("vqb_ubench_copy: invalid argument");
        return VQB_ERR_INVALID_ARG;
    }
    const int64_t N = B * HW;
    const unsigned blocks = mode == 2 ? (unsigned)(sm_count() * 16)
                                      : (mode >= 3 ? (unsigned)((N + 127) / 128) : (unsigned)((N + 31) / 32));
    pattern_copy_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(a, b, out, N, D, HW, mode);
    VQB_LAUNCH_CHECK("pattern_copy_kernel");
    return VQB_OK;
}

extern "C" int vqb_ubench_launch(int mode, int sweeps, const float* src, float* sink, double* flops_host,
                                 vqb_stream_t stream) {
    VQB_DEVICE_TRY();
    if (mode < 0 || mode > 5 || sweeps <= 0 || !src || !sink) {
        set_error("vqb_ubench_launch: invalid argument");
        return VQB_ERR_INVALID_ARG;
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int blocks = sm_count() * 2;
    constexpr int T = 8;
    switch (mode) {
        case 0: ubench_kernel<0, T><<<blocks, 256, 0, s>>>(sweeps, src, sink); break;
        case 1: ubench_kernel<1, T><<<blocks, 256, 0, s>>>(sweeps, src, sink); break;
        case 2: ubench_kernel<2, T><<<blocks, 256, 0, s>>>(sweeps, src, sink); break;
        case 3: ubench_kernel<3, T><<<blocks, 256, 0, s>>>(sweeps, src, sink); break;
        case 4: ubench_kernel<4, T><<<blocks, 256, 0, s>>>(sweeps, src, sink); break;
        default: ubench_kernel<5, T><<<blocks, 256, 0, s>>>(sweeps, src, sink); break;
    }
    VQB_LAUNCH_CHECK("ubench_kernel");
    if (flops_host) *flops_host = 2.0 * 4 * (double)blocks * 256 * T * kUbCodes * sweeps;
    return VQB_OK;
}

// Scatter-reduction patterns of the codebook gradient: N tokens each add a D-float row into
// table[idx[token]][0..D).  lanes_per_row consecutive lanes cover lanes_per_row*4 consecutive floats of ONE row per
// instruction (red.global.add.v4.f32): 1 = lane-per-token (16-byte pieces of 32 different rows per warp instruction),
// 4 = 64 contiguous bytes per row, 16 = 256 contiguous bytes per row (half-warp per token).  No other memory traffic.
template <int LPR>
__global__ void __launch_bounds__(256) red_pattern_kernel(float* __restrict__ table, const int64_t* __restrict__ idx, int64_t N,
                                                          int D) {
    const int lane = threadIdx.x & 31;
    const int64_t warp_id = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    constexpr int TPW = 32 / LPR;                 // tokens per warp instruction
    const int sub = lane / LPR, l = lane % LPR;   // token slot / position within the row piece
    // a warp owns 32 consecutive tokens and walks them TPW at a time, each over all D channels
    for (int t0 = 0; t0 < 32; t0 += TPW) {
        const int64_t tok = warp_id * 32 + t0 + sub;
        if (tok >= N) continue;
        float* row = table + (size_t)idx[tok] * D;
        for (int c = 4 * l; c < D; c += 4 * LPR)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %1, %1, %1};" ::"l"(row + c), "f"(1.0f) : "memory");
    }
}

extern "C" int vqb_ubench_red(float* table, const int64_t* idx, int64_t N, int D, int lanes_per_row, vqb_stream_t stream) {
    VQB_DEVICE_TRY();
    if (!table || !idx || N <= 0 || D <= 0 || D % 64 != 0) {
        set_error("vqb_ubench_red: invalid argument");
        return VQB_ERR_INVALID_ARG;
    }
    const unsigned blocks = (unsigned)((N + 255) / 256);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    switch (lanes_per_row) {
        case 1: red_pattern_kernel<1><<<blocks, 256, 0, s>>>(table, idx, N, D); break;
        case 4: red_pattern_kernel<4><<<blocks, 256, 0, s>>>(table, idx, N, D); break;
        case 16: red_pattern_kernel<16><<<blocks, 256, 0, s>>>(table, idx, N, D); break;
        default: set_error("vqb_ubench_red: lanes_per_row must be 1, 4 or 16"); return VQB_ERR_INVALID_ARG;
    }
    VQB_LAUNCH_CHECK("red_pattern_kernel");
    return VQB_OK;
}

// FP32 FMA peak: 16 independent chains per thread, operands from registers
template <bool kPacked>
__global__ void __launch_bounds__(256) fma_peak_kernel(int iters, float a, float b, float* sink) {
    if constexpr (kPacked) {
        unsigned long long x[8];
        const unsigned long long a2 = pack_f32x2(a, a + 1e-3f), b2 = pack_f32x2(b, b - 1e-3f);
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = pack_f32x2(threadIdx.x * 1e-3f + i, i * 0.5f);
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int i = 0; i < 8; ++i) x[i] = fma_f32x2(x[i], a2, b2);
        }
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float lo, hi;
            unpack_f32x2(x[i], lo, hi);
            s += lo + hi;
        }
        if (s == 123.456f) sink[0] = s;
    } else {
        float x[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = threadIdx.x * 1e-3f + i;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int i = 0; i < 16; ++i) x[i] = fmaf(x[i], a, b);
        }
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) s += x[i];
        if (s == 123.456f) sink[0] = s;
    }
}

extern "C" int vqb_fma_peak_launch(int packed, int iters, float* sink, double* flops_host,
                                   vqb_stream_t stream) {
    VQB_DEVICE_TRY();
    if (iters <= 0 || !sink) {
        set_error("vqb_fma_peak_launch: invalid argument");
        return VQB_ERR_INVALID_ARG;
    }
    const int blocks = sm_count() * 8;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (packed)
        fma_peak_kernel<true><<<blocks, 256, 0, s>>>(iters, 0.999f, 0.001f, sink);
    else
        fma_peak_kernel<false><<<blocks, 256, 0, s>>>(iters, 0.999f, 0.001f, sink);
    VQB_LAUNCH_CHECK("fma_peak_kernel");
    // per thread per iteration: 8 rounds x 16 lanes of FMA = 128 FMA = 256 flop
    if (flops_host) *flops_host = 256.0 * iters * 256.0 * blocks;
    return VQB_OK;
}
