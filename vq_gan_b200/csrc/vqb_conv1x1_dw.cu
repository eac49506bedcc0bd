// Parameter gradients of the 1x1 convolution either side of the quantizer (SURVEY.md section 8f, row N1;
// vqgan_ldm_baseline/models/vq_vae.py:74-79: pre_quant_conv / post_quant_conv = nn.Conv2d(cin, cout, 1)):
//     dW[o, c]  = sum over tokens (b, hw) of dy[b, o, hw] * x[b, c, hw]        dbias[o] = sum over tokens of dy[b, o, hw]
// a GEMM whose reduction dimension is the TOKENS (10^6 long, 256 x 256 outputs): algorithmic bytes 4 (Cin + Cout) per
// token, 2 Cin Cout flop per token.  In NCHW both operands are already K-major for that product -- for a fixed (image,
// channel) the tokens are contiguous -- so TMA drops [channels x 32 tokens] boxes of dy and x straight into
// 128B-swizzled shared memory, no transposing converter as in the forward kernel.  3xTF32 for fp32-level accuracy
// (dyh.xh + dyl.xh + dyh.xl): the landed fp32 tile serves as the hi image as it is (the tensor core ignores the low 13
// mantissa bits), eight converter warps write the residual image tf32(x - trunc(x)) next to it (an elementwise pass: the
// swizzle does not matter) and sum dy for dbias on the way.  One persistent CTA per SM accumulates its share of the
// token blocks in TMEM: a 128-row tile of Cout as lanes, a chunk of <= 256 input channels as columns (blockIdx.y walks
// the (chunk, tile) pairs; 128 x 256 x 8 MMAs read 12 KB of operands per 128 tensor cycles, 128 x 128 x 8 ones the SM's
// whole 128 B/clk -- shared-memory bandwidth is what bounds this kernel), then adds its partial product to dW with
// 16-byte reductions (up to 148 partial sums per element: order-dependent in the last bits, like dE).
// Two experiments that did NOT help (1 M tokens, 256 -> 256, 0.79 ms as it stands): issuing the hi x hi products as soon as
// TMA lands, before the converters finish (0.80 ms: the conversion is not what the tensor pipe waits for), and a ring of
// three landed tiles with ONE residual slot (0.98 ms: conversion and residual products then serialise on that slot).
#include "vqb_tc_common.cuh"

namespace vqb {

constexpr int kDwThreads = 512;  // warp 0 TMA, 1 MMA, 2 TMEM alloc, 3 idle, 4-11 converters, 12-15 epilogue
constexpr int kDwTok = 32;       // tokens per k-block: one 128-byte swizzle row of fp32
constexpr int kDwTileBytes = 128 * kDwTok * 4;  // one 128-row operand tile: 16 KB
constexpr int kDwMaxStages = 4;

__device__ __forceinline__ void dw_umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ float dw_trunc_tf32(float v) { return __uint_as_float(__float_as_uint(v) & 0xffffe000u); }
__device__ __forceinline__ void dw_split(const float4& v, float4& h, float4& l, int hw_trunc) {
    if (hw_trunc) {
        h.x = dw_trunc_tf32(v.x); h.y = dw_trunc_tf32(v.y); h.z = dw_trunc_tf32(v.z); h.w = dw_trunc_tf32(v.w);
    } else {
        h.x = to_tf32(v.x); h.y = to_tf32(v.y); h.z = to_tf32(v.z); h.w = to_tf32(v.w);
    }
    l.x = to_tf32(v.x - h.x);
    l.y = to_tf32(v.y - h.y);
    l.z = to_tf32(v.z - h.z);
    l.w = to_tf32(v.w - h.w);
}

struct DwParams {
    int64_t B, HW;
    int Cin, Cout;
    int chunk;    // input channels per CTA (columns of the accumulator, the MMA's N), multiple of 16, <= 256
    int m_tiles;  // 128-row tiles of Cout: 1 or 2; blockIdx.y = chunk index * m_tiles + tile
    int stages;
    float* dW;     // [Cout, Cin], accumulated into
    float* dbias;  // [Cout] or null, accumulated into
    int hw_trunc;  // 1 (default): the landed fp32 tile IS the hi image -- the tensor core ignores the low 13 mantissa
                   // bits (measured: same 7e-7 error as explicit rounding) -- and only lo = tf32(x - trunc(x)) is written;
                   // 0 (vqb_tune "dw_hw_trunc"): hi = rna tf32(x) rewritten in place as well
};

__global__ void __launch_bounds__(kDwThreads, 1)
    conv1x1_dw_kernel(const __grid_constant__ CUtensorMap map_dy, const __grid_constant__ CUtensorMap map_x, DwParams p) {
    extern __shared__ unsigned char smem_unaligned[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_unaligned) + 1023) &
                                                           ~(uintptr_t)1023);
    const uint32_t a_bytes = kDwTileBytes;                                   // dy tile: 128 rows (this CTA's slice of Cout)
    const uint32_t b_bytes = (p.chunk > 128 ? 2u : 1u) * kDwTileBytes;        // x tile: chunk rows
    const uint32_t stage_bytes = 2 * a_bytes + 2 * b_bytes;                  // A hi | A lo | B hi | B lo
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_bytes);
    uint64_t* raw_full = bars + 0;                   // [stages] TMA -> converters
    uint64_t* cv_full = bars + kDwMaxStages;         // [stages] converters -> MMA
    uint64_t* s_empty = bars + 2 * kDwMaxStages;     // [stages] MMA -> TMA
    uint64_t* acc_full = bars + 3 * kDwMaxStages;    // MMA -> epilogue
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * kDwMaxStages + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t kb_per_img = (p.HW + kDwTok - 1) / kDwTok;
    const int64_t total_kb = p.B * kb_per_img;
    const int64_t n_local = blockIdx.x < total_kb ? (total_kb - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const int mt = (int)blockIdx.y % p.m_tiles;
    const int c_base = ((int)blockIdx.y / p.m_tiles) * p.chunk;
    const int o_base = mt * 128;
    const int a_rows = p.Cout < 128 ? p.Cout : 128;  // rows per dy box
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.chunk >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

    if (threadIdx.x == 0) {
        for (int i = 0; i < p.stages; ++i) {
            tc_mbar_init(raw_full + i, 1);
            tc_mbar_init(cv_full + i, 256);
            tc_mbar_init(s_empty + i, 1);
        }
        tc_mbar_init(acc_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(tmem_slot)), "r"(256)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer: raw dy / x boxes =====================
        if (lane == 0) {
            const uint32_t tx_bytes = (uint32_t)a_rows * 128u + (uint32_t)p.chunk * 128u;
            for (int64_t it = 0; it < n_local; ++it) {
                const int64_t kb = blockIdx.x + it * gridDim.x;
                const uint32_t stage = (uint32_t)(it % p.stages), ph = (uint32_t)((it / p.stages) & 1);
                tc_mbar_wait(s_empty + stage, ph ^ 1);
                tc_mbar_expect_tx(raw_full + stage, tx_bytes);
                const int64_t b = kb / kb_per_img;
                const int h0 = (int)(kb - b * kb_per_img) * kDwTok;
                unsigned char* ahi = smem + (size_t)stage * stage_bytes;
                unsigned char* bhi = ahi + 2 * a_bytes;
                tma_load_2d(ahi, &map_dy, raw_full + stage, h0, (int)(b * p.Cout) + o_base);
                tma_load_2d(bhi, &map_x, raw_full + stage, h0, (int)(b * p.Cin) + c_base);
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            for (int64_t it = 0; it < n_local; ++it) {
                const uint32_t stage = (uint32_t)(it % p.stages), ph = (uint32_t)((it / p.stages) & 1);
                tc_mbar_wait(cv_full + stage, ph);
                tc_fence_after();
                const uint32_t ahi = s32(smem + (size_t)stage * stage_bytes);
                const uint32_t alo = ahi + a_bytes;
                const uint32_t bhi = ahi + 2 * a_bytes;
                const uint32_t blo = bhi + b_bytes;
#pragma unroll
                for (int k4 = 0; k4 < kDwTok / 8; ++k4) {  // 8 tf32 = 32 bytes per MMA
                    const uint32_t accum = (it != 0 || k4 != 0) ? 1u : 0u;
                    const uint64_t dbh = umma_desc_sw128(bhi + k4 * 32), dbl = umma_desc_sw128(blo + k4 * 32);
                    const uint64_t dah = umma_desc_sw128(ahi + k4 * 32), dal = umma_desc_sw128(alo + k4 * 32);
                    dw_umma_tf32(tmem_base, dah, dbh, idesc, accum);
                    dw_umma_tf32(tmem_base, dal, dbh, idesc, 1);
                    dw_umma_tf32(tmem_base, dah, dbl, idesc, 1);
                }
                umma_commit(s_empty + stage);
            }
            if (n_local > 0) umma_commit(acc_full);
        }
    } else if (warp >= 4 && warp < 12) {
        // ===================== converters: residual image (and, hw_trunc = 0, the rounded hi image); dbias sums ====
        const int t = threadIdx.x - 128;  // 0..255
        const uint32_t a_chunks = (uint32_t)a_rows * 8u;  // 16-byte pieces of the dy tile
        const uint32_t b_chunks = (uint32_t)p.chunk * 8u;
        float bsum[4];  // piece id = t + 256 j -> row (t >> 3) + 32 j: the same rows for every token block
#pragma unroll
        for (int j = 0; j < 4; ++j) bsum[j] = 0.f;
        for (int64_t it = 0; it < n_local; ++it) {
            const uint32_t stage = (uint32_t)(it % p.stages), ph = (uint32_t)((it / p.stages) & 1);
            tc_mbar_wait(raw_full + stage, ph);
            unsigned char* ahi = smem + (size_t)stage * stage_bytes;
            unsigned char* alo = ahi + a_bytes;
            unsigned char* bhi = ahi + 2 * a_bytes;
            unsigned char* blo = bhi + b_bytes;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t id = (uint32_t)t + 256u * j;
                if (id < a_chunks) {
                    const float4 v = *reinterpret_cast<const float4*>(ahi + 16u * id);
                    bsum[j] += (v.x + v.y) + (v.z + v.w);
                    float4 h, l;
                    dw_split(v, h, l, p.hw_trunc);
                    if (!p.hw_trunc) *reinterpret_cast<float4*>(ahi + 16u * id) = h;
                    *reinterpret_cast<float4*>(alo + 16u * id) = l;
                }
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const uint32_t id = (uint32_t)t + 256u * j;
                if (id < b_chunks) {
                    const float4 v = *reinterpret_cast<const float4*>(bhi + 16u * id);
                    float4 h, l;
                    dw_split(v, h, l, p.hw_trunc);
                    if (!p.hw_trunc) *reinterpret_cast<float4*>(bhi + 16u * id) = h;
                    *reinterpret_cast<float4*>(blo + 16u * id) = l;
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic writes -> tensor-core reads
            tc_mbar_arrive(cv_full + stage);
        }
        if (p.dbias != nullptr && c_base == 0) {  // the CTAs of the first channel chunk own dbias (their tile's rows)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float s = bsum[j];
                s += __shfl_xor_sync(0xffffffffu, s, 1);
                s += __shfl_xor_sync(0xffffffffu, s, 2);
                s += __shfl_xor_sync(0xffffffffu, s, 4);
                const int row = o_base + (t >> 3) + 32 * j;
                if ((t & 7) == 0 && row < p.Cout && n_local > 0) atomicAdd(p.dbias + row, s);
            }
        }
    } else if (warp >= 12) {
        // ===================== epilogue: TMEM partial product -> dW (16-byte reductions) =====================
        if (n_local > 0) {
            const int q = warp & 3;
            const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
            tc_mbar_wait(acc_full, 0);
            tc_fence_after();
            const int o = o_base + q * 32 + lane;
            const int row_ok = o < p.Cout;
            float* wp = p.dW + (size_t)(row_ok ? o : 0) * p.Cin + c_base;
            for (int c0 = 0; c0 < p.chunk; c0 += 32) {
                uint32_t r[32];
                tmem_ld32(tmem_base + lane_addr + (uint32_t)c0, r);
#pragma unroll
                for (int g = 0; g < 8; ++g)
                    asm volatile(
                        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %5, 0;\n\t"
                        "@p red.global.add.v4.f32 [%0], {%1, %2, %3, %4};\n\t}" ::"l"(wp + c0 + 4 * g),
                        "f"(__uint_as_float(r[4 * g])), "f"(__uint_as_float(r[4 * g + 1])),
                        "f"(__uint_as_float(r[4 * g + 2])), "f"(__uint_as_float(r[4 * g + 3])),
                        "r"((int)(row_ok && c0 + 4 * g < p.chunk))
                        : "memory");
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256) : "memory");
    }
}

// input channels per CTA: the largest multiple of 16 that divides Cin and is <= 256 (0: no tensor path)
static int dw_chunk(int Cin) {
    if (Cin < 16 || Cin % 16 != 0) return 0;
    for (int c = 256; c >= 16; c -= 16)
        if (Cin % c == 0) return c;
    return 0;
}

static bool dw_supported(int Cin, int Cout, int64_t HW) {
    return dw_chunk(Cin) > 0 && Cout >= 1 && Cout <= 256 && HW >= kDwTok && HW % 4 == 0;
}

VQB_KNOB g_dw_hw_trunc = 1;
#ifdef VQB_EXPERIMENTAL
void set_dw_hw_trunc(int v) { g_dw_hw_trunc = v; }
#endif

}  // namespace vqb

using namespace vqb;

extern "C" int vqb_conv1x1_dw_supported(int Cin, int Cout, int64_t HW) { return dw_supported(Cin, Cout, HW) ? 1 : 0; }

extern "C" int vqb_conv1x1_dw_f32(const float* dy, const float* x, int64_t B, int Cin, int Cout, int64_t HW,
                                  float* dW_accum, float* dbias_accum, vqb_stream_t stream) {
    VQB_DEVICE_TRY();
    if (B < 0 || HW < 0 || Cin <= 0 || Cout <= 0) {
        set_error("vqb_conv1x1_dw_f32: invalid shape B=%lld Cin=%d Cout=%d HW=%lld", (long long)B, Cin, Cout, (long long)HW);
        return VQB_ERR_INVALID_ARG;
    }
    if (B * HW == 0) return VQB_OK;
    if (!dy || !x || !dW_accum) {
        set_error("vqb_conv1x1_dw_f32: null pointer");
        return VQB_ERR_INVALID_ARG;
    }
    if (!dw_supported(Cin, Cout, HW)) {
        set_error("vqb_conv1x1_dw_f32 needs Cin %% 16 == 0, Cout <= 256, HW >= 32 and HW %% 4 == 0 (Cin=%d Cout=%d HW=%lld)", Cin,
                  Cout, (long long)HW);
        return VQB_ERR_UNSUPPORTED;
    }
    if (((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dW_accum)) & 15u) != 0 ||
        B * (int64_t)(Cin > Cout ? Cin : Cout) >= (1LL << 31)) {
        set_error("vqb_conv1x1_dw_f32: tensors must be 16-byte aligned and B * channels < 2^31");
        return VQB_ERR_INVALID_ARG;
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    DwParams p;
    p.B = B;
    p.HW = HW;
    p.Cin = Cin;
    p.Cout = Cout;
    p.chunk = dw_chunk(Cin);
    p.m_tiles = Cout > 128 ? 2 : 1;
    p.dW = dW_accum;
    p.dbias = dbias_accum;
    p.hw_trunc = g_dw_hw_trunc;
    const size_t stage_bytes = 2 * (size_t)kDwTileBytes + 2 * (size_t)(p.chunk > 128 ? 2 : 1) * kDwTileBytes;
    int stages = (int)((kTcSmemBudget - 1024 - 512) / stage_bytes);
    p.stages = stages > kDwMaxStages ? kDwMaxStages : stages;
    const size_t smem = 1024 + (size_t)p.stages * stage_bytes + 512;
    CUtensorMap mdy, mx;
    if (int rc = make_tc_map_f32(&mdy, dy, (uint64_t)(B * Cout), (int)HW, (uint32_t)(Cout < 128 ? Cout : 128))) return rc;
    if (int rc = make_tc_map_f32(&mx, x, (uint64_t)(B * Cin), (int)HW, (uint32_t)p.chunk)) return rc;
    const int n_chunks = (Cin / p.chunk) * p.m_tiles;  // CTA kinds: (channel chunk, 128-row tile of Cout)
    const int64_t total_kb = B * ((HW + kDwTok - 1) / kDwTok);
    int gx = sm_count() / n_chunks;
    if (gx < 1) gx = 1;
    if ((int64_t)gx > total_kb) gx = (int)total_kb;
    VQB_CUDA_TRY(cudaFuncSetAttribute(conv1x1_dw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    conv1x1_dw_kernel<<<dim3((unsigned)gx, (unsigned)n_chunks), kDwThreads, smem, s>>>(mdy, mx, p);
    VQB_LAUNCH_CHECK("conv1x1_dw_kernel");
    return VQB_OK;
}
