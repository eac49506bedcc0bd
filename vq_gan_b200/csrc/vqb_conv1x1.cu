// 1x1 convolution on NCHW latents: the `pre_quant_conv` /
// `post_quant_conv` layers that sit directly either side of the quantizer when
// z_channels != embedding_dim (vqgan_ldm_baseline/models/vq_vae.py:74-79,115,121):
//     y[b, o, hw] = sum_c W[o, c] * x[b, c, hw] + bias[o]
//
// Tensor path (Cin % 32 == 0, Cout % 16 == 0, 16 <= Cout <= 256): tcgen05 with a 3xTF32 split for
// fp32-level accuracy,   x.w ~= xh.wh + xl.wh + xh.wl   (xh = tf32(x), xl = tf32(x - xh); the dropped
// xl.wl term and the second roundings are 2^-22 relative), fp32 accumulation in TMEM.
//   * tokens are the M dimension: 128-token tiles, accumulator = 128 x Cout fp32 (two TMEM buffers)
//   * A operand (activations) is token-contiguous in HBM but the MMA wants it K-major: eight
//     "converter" warps load 32 channels x 128 tokens (coalesced along the tokens), split into
//     hi/lo tf32 and write the 128B-swizzled K-major image straight into shared memory
//     (conflict-free 16-byte stores, fence.proxy.async before the MMA reads it)
//   * B operand (weights, hi/lo split once per call) arrives by TMA, multicast across a 2-CTA
//     cluster; 12 MMAs (128 x Cout x 8) per 32-channel block
//   * epilogue warps read the accumulator with tcgen05.ld, add the bias and store y coalesced
//     along the tokens.
// Generic path (any shape): CUDA-core kernel, thread = token, weights staged in shared memory.
#include "vqb_tc_common.cuh"

namespace vqb {

constexpr int kCvThreads = 512;  // warp 0 TMA, 1 MMA, 2 TMEM alloc, 3 idle, 4-11 converters, 12-15 epilogue
constexpr int kCvTok = 128;
constexpr int kCvKB = 32;                         // channels per k-block (128 bytes of tf32)
constexpr int kCvABytes = kCvTok * kCvKB * 4;     // 16 KB per hi or lo image
constexpr int kCvStages = 2;

__device__ __forceinline__ void umma_tf32_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}

__global__ void __launch_bounds__(256) conv_split_weights_kernel(const float* __restrict__ W, int n, float* __restrict__ hi,
                                                                 float* __restrict__ lo) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float v = W[i];
    const float h = to_tf32(v);
    hi[i] = h;
    lo[i] = to_tf32(v - h);
}

struct ConvParams {
    const float* x;
    float* y;
    const float* bias;  // nullable
    int64_t N, HW;
    int Cin, Cout;
    int dbg;  // experiments (vqb_tune "conv_debug"): 1 no activation loads, 2 no output stores, 4 one MMA in three
    // fused token split for the fp16 tensor search that consumes y (vqb_conv1x1_split_f32; all null otherwise): what
    // split16_tokens_kernel (vqb_search_tc16.cu) would compute from y, written while the accumulator is still in TMEM
    __half* z16;          // [N][Dpad] token-major fp16(y * 2^e_i)
    float* inv_scale;     // [N] 2^-(e_i + se)
    float* znorm;         // [N] |y_i|
    float* zres;          // [N] |y_i - fp16 image|
    const int* header;    // codebook pack header (se = header[5])
    int Dpad;
};

template <int CL>
__global__ void __launch_bounds__(kCvThreads, 1)
    conv1x1_tc_kernel(const __grid_constant__ CUtensorMap map_whi, const __grid_constant__ CUtensorMap map_wlo,
                      ConvParams p) {
    extern __shared__ unsigned char smem_unaligned[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_unaligned) + 1023) &
                                                           ~(uintptr_t)1023);
    const uint32_t w_bytes = (uint32_t)p.Cout * kCvKB * 4;             // one hi or lo weight block
    const uint32_t stage_bytes = 2 * kCvABytes + 2 * w_bytes;          // A hi | A lo | W hi | W lo
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kCvStages * stage_bytes);
    uint64_t* a_full = bars + 0;    // [2] converters -> MMA
    uint64_t* w_full = bars + 2;    // [2] TMA -> MMA
    uint64_t* s_empty = bars + 4;   // [2] MMA -> converters and TMA (this CTA's A, the cluster's W)
    uint64_t* w_empty = bars + 6;   // [2] MMA of every CTA in the cluster -> TMA
    uint64_t* tm_full = bars + 8;   // [2]
    uint64_t* tm_empty = bars + 10; // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);
    float* bias_sm = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(bars) + 256);  // [256]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_tiles = (int)((p.N + kCvTok - 1) / kCvTok);
    const int n_rounds = (n_tiles + (int)gridDim.x - 1) / (int)gridDim.x;
    const int n_kb = p.Cin / kCvKB;
    const uint32_t cta_rank = CL > 1 ? cluster_ctarank() : 0;
    constexpr uint16_t kMask = (uint16_t)((1u << CL) - 1);
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.Cout >> 3) << 17) |
                           ((uint32_t)(kCvTok >> 4) << 24);

    if (threadIdx.x == 0) {
        for (int i = 0; i < kCvStages; ++i) {
            tc_mbar_init(a_full + i, 256);
            tc_mbar_init(w_full + i, 1);
            tc_mbar_init(s_empty + i, 1);
            tc_mbar_init(w_empty + i, CL);
            tc_mbar_init(tm_full + i, 1);
            tc_mbar_init(tm_empty + i, 128);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(tmem_slot)), "r"(512)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (threadIdx.x >= 256) {
        const int o = threadIdx.x - 256;
        bias_sm[o] = (p.bias && o < p.Cout) ? p.bias[o] : 0.f;
    }
    tc_fence_before();
    __syncthreads();
    if constexpr (CL > 1) cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer: weight blocks =====================
        if (lane == 0) {
            uint32_t it = 0;
            for (int round = 0; round < n_rounds; ++round) {
                for (int kb = 0; kb < n_kb; ++kb, ++it) {
                    const uint32_t stage = it & 1, ph = (it >> 1) & 1;
                    tc_mbar_wait(w_empty + stage, ph ^ 1);
                    tc_mbar_expect_tx(w_full + stage, 2 * w_bytes);
                    unsigned char* whi = smem + stage * stage_bytes + 2 * kCvABytes;
                    unsigned char* wlo = whi + w_bytes;
                    if constexpr (CL > 1) {
                        const int rows = p.Cout / CL;
                        const uint32_t part = w_bytes / CL;
                        tma_load_2d_mc(whi + cta_rank * part, &map_whi, w_full + stage, kb * kCvKB, (int)cta_rank * rows, kMask);
                        tma_load_2d_mc(wlo + cta_rank * part, &map_wlo, w_full + stage, kb * kCvKB, (int)cta_rank * rows, kMask);
                    } else {
                        tma_load_2d(whi, &map_whi, w_full + stage, kb * kCvKB, 0);
                        tma_load_2d(wlo, &map_wlo, w_full + stage, kb * kCvKB, 0);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            uint32_t it = 0;
            for (int round = 0; round < n_rounds; ++round) {
                const uint32_t acc = round & 1;
                tc_mbar_wait(tm_empty + acc, ((round >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * 256;
                for (int kb = 0; kb < n_kb; ++kb, ++it) {
                    const uint32_t stage = it & 1, ph = (it >> 1) & 1;
                    tc_mbar_wait(a_full + stage, ph);
                    tc_mbar_wait(w_full + stage, ph);
                    tc_fence_after();
                    const uint32_t ahi = s32(smem + stage * stage_bytes);
                    const uint32_t alo = ahi + kCvABytes;
                    const uint32_t whi = ahi + 2 * kCvABytes;
                    const uint32_t wlo = whi + w_bytes;
#pragma unroll
                    for (int k4 = 0; k4 < kCvKB / 8; ++k4) {  // 8 tf32 = 32 bytes per MMA
                        umma_tf32_ss(d_tmem, umma_desc_sw128(ahi + k4 * 32), umma_desc_sw128(whi + k4 * 32), idesc,
                                     (kb | k4) != 0);
                        if (p.dbg & 4) continue;
                        umma_tf32_ss(d_tmem, umma_desc_sw128(alo + k4 * 32), umma_desc_sw128(whi + k4 * 32), idesc, 1);
                        umma_tf32_ss(d_tmem, umma_desc_sw128(ahi + k4 * 32), umma_desc_sw128(wlo + k4 * 32), idesc, 1);
                    }
                    umma_commit(s_empty + stage);
                    if constexpr (CL > 1)
                        umma_commit_mc(w_empty + stage, kMask);
                    else
                        umma_commit(w_empty + stage);
                }
                umma_commit(tm_full + acc);
            }
        }
    } else if (warp >= 4 && warp < 12) {
        // ===================== converters: x -> hi/lo tf32, K-major 128B-swizzled =====================
        // The loads of block i+1 are issued before block i is converted (two blocks = 32 KB per SM in
        // flight; with one block the kernel was latency-bound at 1.9 TB/s), across tile boundaries too.
        const int cw = warp - 4;  // 16-byte chunk (4 channels) of the 128-byte row this warp fills
        const uint32_t n_items = (uint32_t)n_rounds * (uint32_t)n_kb;
        auto token_offsets = [&](int round, int64_t (&off)[4]) {
            const int tile = blockIdx.x + round * gridDim.x;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const int64_t tok = (int64_t)tile * kCvTok + 32 * g + lane;
                if (tok < p.N) {
                    const int64_t b = tok / p.HW;
                    off[g] = (b * p.Cin) * p.HW + (tok - b * p.HW);
                } else {
                    off[g] = -1;
                }
            }
        };
        auto load_block = [&](const int64_t (&off)[4], int kb, float (&v)[4][4]) {
            const int64_t coff = (int64_t)(kb * kCvKB + 4 * cw) * p.HW;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    v[g][c] = (off[g] >= 0 && !(p.dbg & 1)) ? __ldg(p.x + off[g] + coff + (int64_t)c * p.HW) : 0.f;
            }
        };
        auto convert_block = [&](uint32_t it, const float (&v)[4][4]) {
            const uint32_t stage = it & 1, ph = (it >> 1) & 1;
            tc_mbar_wait(s_empty + stage, ph ^ 1);  // the MMAs that read this stage have retired
            unsigned char* ahi = smem + stage * stage_bytes;
            unsigned char* alo = ahi + kCvABytes;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const int t = 32 * g + lane;
                const uint32_t o = (uint32_t)(t >> 3) * 1024 + (uint32_t)(t & 7) * 128 + (uint32_t)((cw ^ (t & 7)) * 16);
                float4 h, l;
                h.x = to_tf32(v[g][0]); l.x = to_tf32(v[g][0] - h.x);
                h.y = to_tf32(v[g][1]); l.y = to_tf32(v[g][1] - h.y);
                h.z = to_tf32(v[g][2]); l.z = to_tf32(v[g][2] - h.z);
                h.w = to_tf32(v[g][3]); l.w = to_tf32(v[g][3] - h.w);
                *reinterpret_cast<float4*>(ahi + o) = h;
                *reinterpret_cast<float4*>(alo + o) = l;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic writes -> tensor-core reads
            tc_mbar_arrive(a_full + stage);
        };
        int64_t off[4];
        float va[4][4], vb[4][4];  // ping-pong register buffers (no copies: a copy would wait for the loads)
        int round_l = 0, kb_l = 0;  // position of the NEXT block to load
        auto load_next = [&](float (&v)[4][4]) {
            load_block(off, kb_l, v);
            if (++kb_l == n_kb) {
                kb_l = 0;
                ++round_l;
                if (round_l < n_rounds) token_offsets(round_l, off);
            }
        };
        token_offsets(0, off);
        load_next(va);
        for (uint32_t it = 0; it < n_items; it += 2) {
            if (it + 1 < n_items) load_next(vb);
            convert_block(it, va);
            if (it + 1 < n_items) {
                if (it + 2 < n_items) load_next(va);
                convert_block(it + 1, vb);
            }
        }
    } else if (warp >= 12) {
        // ===================== epilogue: TMEM -> + bias -> y (coalesced along tokens) =====================
        const int q = warp & 3;
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        for (int round = 0; round < n_rounds; ++round) {
            const int tile = blockIdx.x + round * gridDim.x;
            const uint32_t acc = round & 1;
            const int64_t tok = (int64_t)tile * kCvTok + q * 32 + lane;
            const bool ok = tok < p.N;
            int64_t yoff = 0;
            if (ok) {
                const int64_t b = tok / p.HW;
                yoff = (b * p.Cout) * p.HW + (tok - b * p.HW);
            }
            tc_mbar_wait(tm_full + acc, (round >> 1) & 1);
            tc_fence_after();
            // Fused token split (p.z16 != null): ONE pass over the accumulator.  The per-token power-of-two scale is fixed
            // from the first 32 channels (their largest magnitude times 8 lands in [512, 1024)): fp16 is a floating format, so
            // the scale only has to keep the row inside the normal range; an outlier channel 512x larger than anything in
            // the first chunk overflows to inf, the measured residual becomes NaN and the token takes the exact search --
            // slower, never wrong.  (A first version scaled by the row maximum like split16_tokens_kernel, which needs a
            // second TMEM pass: 0.98 -> 1.42 ms for 256 -> 256 channels, more than the 0.27 ms split pass it replaces.)
            // No `if (ok)` inside the loop: it lets the compiler unswitch the loop on `ok`, which puts the .sync.aligned
            // TMEM loads into divergent copies of the loop (observed: a hang whenever the last token tile is partial).
            float sq = 0.f, res = 0.f, fwd = 1.f, back = 1.f;
            int e = 0;
            bool finite = true;
            const bool stores = ok && !(p.dbg & 2);
            __half* zp = p.z16 ? p.z16 + (size_t)(ok ? tok : 0) * p.Dpad : nullptr;
            for (int c0 = 0; c0 < p.Cout; c0 += 32) {
                uint32_t r[32];
                tmem_ld32(tmem_base + lane_addr + acc * 256 + c0, r);
                float v[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = (c0 + j < p.Cout) ? __uint_as_float(r[j]) + bias_sm[c0 + j] : 0.f;
                float* yp = p.y + yoff + (int64_t)c0 * p.HW;
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %2, 0;\n\t@p st.global.f32 [%0], %1;\n\t}" ::"l"(yp + (int64_t)j * p.HW),
                                 "f"(v[j]), "r"((int)(stores && c0 + j < p.Cout))
                                 : "memory");
                if (p.z16) {  // kernel-uniform
                    if (c0 == 0) {
                        float mx = 0.f;
#pragma unroll
                        for (int j = 0; j < 32; ++j) mx = fmaxf(mx, fabsf(v[j]));
                        if (mx > 0.f && mx < INFINITY) {
                            e = 6 - ilogbf(mx);  // first-chunk maximum -> [64, 128): 512x headroom to fp16's largest
                            e = e < -100 ? -100 : (e > 100 ? 100 : e);
                        }
                        fwd = __int_as_float((127 + e) << 23);
                        back = __int_as_float((127 - e) << 23);
                    }
                    uint32_t packed[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float v0 = v[2 * j], v1 = v[2 * j + 1];
                        sq = fmaf(v0, v0, sq);
                        sq = fmaf(v1, v1, sq);
                        finite &= (fabsf(v0) < INFINITY) && (fabsf(v1) < INFINITY);
                        const __half h0 = __float2half_rn(v0 * fwd), h1 = __float2half_rn(v1 * fwd);
                        const float d0 = v0 - __half2float(h0) * back, d1 = v1 - __half2float(h1) * back;
                        res = fmaf(d0, d0, res);
                        res = fmaf(d1, d1, res);
                        packed[j] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        asm volatile(
                            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %5, 0;\n\t"
                            "@p st.global.v4.b32 [%0], {%1, %2, %3, %4};\n\t}" ::"l"(zp + c0 + 8 * j),
                            "r"(packed[4 * j]), "r"(packed[4 * j + 1]), "r"(packed[4 * j + 2]), "r"(packed[4 * j + 3]),
                            "r"((int)ok)
                            : "memory");
                }
            }
            if (p.z16 && ok) {
                for (int c = (p.Cout + 31) / 32 * 32; c < p.Dpad; c += 8)  // zero columns up to the 64-channel block
                    *reinterpret_cast<uint4*>(zp + c) = make_uint4(0u, 0u, 0u, 0u);
                // a row that overflowed fp16 has res = inf/NaN: NaN scale -> every approximate score NaN -> exact search
                const bool good = finite && res == res && res < INFINITY;
                p.znorm[tok] = sqrtf(sq);
                p.zres[tok] = good ? sqrtf(res) : 0.f;
                p.inv_scale[tok] = good ? ldexpf(1.f, -(e + p.header[5])) : __int_as_float(0x7fc00000);
            }
            tc_fence_before();
            tc_mbar_arrive(tm_empty + acc);
        }
    }

    tc_fence_before();
    __syncthreads();
    if constexpr (CL > 1) cluster_sync_all();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

// generic CUDA-core path: thread = token, OT output channels per pass (weights broadcast from shared
// memory), eight independent activation loads in flight per thread
template <int OT>
__global__ void __launch_bounds__(128)
    conv1x1_generic_kernel(const float* __restrict__ x, const float* __restrict__ W, const float* __restrict__ bias,
                           int64_t N, int64_t HW, int Cin, int Cout, float* __restrict__ y) {
    extern __shared__ float wsm[];  // [OT][Cin]
    const int64_t tok = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool ok = tok < N;
    const int64_t b = ok ? tok / HW : 0;
    const int64_t hw = ok ? tok - b * HW : 0;
    const float* xp = x + (b * Cin) * HW + hw;
    float* yp = y + (b * Cout) * HW + hw;
    for (int o0 = 0; o0 < Cout; o0 += OT) {
        const int on = (Cout - o0) < OT ? (Cout - o0) : OT;
        __syncthreads();
        for (int i = threadIdx.x; i < OT * Cin; i += blockDim.x) wsm[i] = i < on * Cin ? W[(size_t)o0 * Cin + i] : 0.f;
        __syncthreads();
        float acc[OT];
#pragma unroll
        for (int j = 0; j < OT; ++j) acc[j] = 0.f;
        for (int c0 = 0; c0 < Cin; c0 += 8) {
            float xv[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) xv[u] = (ok && c0 + u < Cin) ? __ldg(xp + (int64_t)(c0 + u) * HW) : 0.f;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                if (c0 + u < Cin) {
#pragma unroll
                    for (int j = 0; j < OT; ++j) acc[j] = fmaf(xv[u], wsm[j * Cin + c0 + u], acc[j]);
                }
            }
        }
        if (ok) {
#pragma unroll
            for (int j = 0; j < OT; ++j)
                if (j < on) yp[(int64_t)(o0 + j) * HW] = acc[j] + (bias ? bias[o0 + j] : 0.f);
        }
    }
}

template <int OT>
static int launch_conv_generic(const float* x, const float* W, const float* bias, int64_t N, int64_t HW, int Cin,
                               int Cout, float* y, cudaStream_t s) {
    const size_t smem = sizeof(float) * OT * (size_t)Cin;
    if (smem > 200 * 1024) {
        set_error("vqb_conv1x1_f32: Cin=%d too large for the generic path", Cin);
        return VQB_ERR_UNSUPPORTED;
    }
    VQB_CUDA_TRY(cudaFuncSetAttribute(conv1x1_generic_kernel<OT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    conv1x1_generic_kernel<OT><<<(unsigned)((N + 127) / 128), 128, smem, s>>>(x, W, bias, N, HW, Cin, Cout, y);
    VQB_LAUNCH_CHECK("conv1x1_generic_kernel");
    return VQB_OK;
}

static bool conv_tc_eligible(int Cin, int Cout) {
    return Cin % kCvKB == 0 && Cin >= kCvKB && Cout % 16 == 0 && Cout >= 16 && Cout <= 256;
}

}  // namespace vqb

using namespace vqb;

VQB_KNOB g_conv_debug = 0;
#ifdef VQB_EXPERIMENTAL
namespace vqb {
void set_conv_debug(int v) { g_conv_debug = v; }
}
#endif

extern "C" size_t vqb_conv1x1_workspace_bytes(int Cin, int Cout) {
    if (Cin <= 0 || Cout <= 0) return 0;
    return conv_tc_eligible(Cin, Cout) ? 2 * round_up_z(sizeof(float) * (size_t)Cin * Cout, 1024) : 0;
}

struct ConvSplitOut {
    __half* z16 = nullptr;
    float* inv_scale = nullptr;
    float* znorm = nullptr;
    float* zres = nullptr;
    const int* header = nullptr;
    int Dpad = 0;
};

static int conv1x1_impl(const float* x, int64_t B, int Cin, int64_t HW, const float* W, const float* bias, int Cout,
                        float* y, void* workspace, size_t workspace_bytes, int algo, const ConvSplitOut& split,
                        vqb_stream_t stream) {
    VQB_DEVICE_TRY();
    if (B < 0 || HW < 0 || Cin <= 0 || Cout <= 0) {
        set_error("vqb_conv1x1_f32: invalid shape B=%lld Cin=%d HW=%lld Cout=%d", (long long)B, Cin, (long long)HW, Cout);
        return VQB_ERR_INVALID_ARG;
    }
    const int64_t N = B * HW;
    if (N == 0) return VQB_OK;
    if (!x || !W || !y) {
        set_error("vqb_conv1x1_f32: null pointer");
        return VQB_ERR_INVALID_ARG;
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const bool tc = conv_tc_eligible(Cin, Cout) && algo != 2;
    if (split.z16 && !tc) {
        set_error("vqb_conv1x1_split_f32 needs the tensor path (Cin %% 32 == 0, Cout %% 16 == 0, Cout <= 256; Cin=%d Cout=%d)", Cin,
                  Cout);
        return VQB_ERR_UNSUPPORTED;
    }
    if (algo == 1 && !tc) {
        set_error("vqb_conv1x1_f32: the tensor path needs Cin %% 32 == 0 and Cout %% 16 == 0, Cout <= 256 (Cin=%d Cout=%d)", Cin, Cout);
        return VQB_ERR_UNSUPPORTED;
    }
    if (!tc) {
        if (Cout <= 4) return launch_conv_generic<4>(x, W, bias, N, HW, Cin, Cout, y, s);
        if (Cout <= 8) return launch_conv_generic<8>(x, W, bias, N, HW, Cin, Cout, y, s);
        if (Cout <= 16) return launch_conv_generic<16>(x, W, bias, N, HW, Cin, Cout, y, s);
        return launch_conv_generic<32>(x, W, bias, N, HW, Cin, Cout, y, s);
    }
    const size_t half = round_up_z(sizeof(float) * (size_t)Cin * Cout, 1024);
    if (!workspace || workspace_bytes < 2 * half || (reinterpret_cast<uintptr_t>(workspace) & 255u) != 0) {
        set_error("vqb_conv1x1_f32: workspace missing, too small (%zu < %zu) or not 256-byte aligned", workspace_bytes, 2 * half);
        return VQB_ERR_WORKSPACE;
    }
    float* whi = static_cast<float*>(workspace);
    float* wlo = reinterpret_cast<float*>(static_cast<unsigned char*>(workspace) + half);
    const int nw = Cin * Cout;
    conv_split_weights_kernel<<<(nw + 255) / 256, 256, 0, s>>>(W, nw, whi, wlo);
    VQB_LAUNCH_CHECK("conv_split_weights_kernel");

    constexpr int CL = 2;
    const bool use_cluster = (Cout % (8 * CL) == 0);
    CUtensorMap mhi, mlo;
    const uint32_t box_rows = (uint32_t)(use_cluster ? Cout / CL : Cout);
    if (int rc = make_tc_map_f32(&mhi, whi, (uint64_t)Cout, Cin, box_rows)) return rc;
    if (int rc = make_tc_map_f32(&mlo, wlo, (uint64_t)Cout, Cin, box_rows)) return rc;

    ConvParams p;
    p.x = x;
    p.y = y;
    p.bias = bias;
    p.N = N;
    p.HW = HW;
    p.Cin = Cin;
    p.Cout = Cout;
    p.dbg = g_conv_debug;
    p.z16 = split.z16;
    p.inv_scale = split.inv_scale;
    p.znorm = split.znorm;
    p.zres = split.zres;
    p.header = split.header;
    p.Dpad = split.Dpad;
    const size_t stage_bytes = 2 * (size_t)kCvABytes + 2 * (size_t)Cout * kCvKB * 4;
    const size_t smem = 1024 + kCvStages * stage_bytes + 256 + 1024;
    const int n_tiles = (int)((N + kCvTok - 1) / kCvTok);
    int grid = n_tiles < sm_count() ? n_tiles : sm_count();
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(kCvThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (use_cluster) {
        VQB_CUDA_TRY(cudaFuncSetAttribute(conv1x1_tc_kernel<CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        grid = (grid + CL - 1) / CL * CL;
        if (grid > sm_count()) grid = sm_count() / CL * CL;
        attr[0].val.clusterDim.x = CL;
        cfg.gridDim = dim3((unsigned)grid);
        int max_clusters = 0;
        VQB_CUDA_TRY(cudaOccupancyMaxActiveClusters(&max_clusters, conv1x1_tc_kernel<CL>, &cfg));
        if (max_clusters > 0 && grid > max_clusters * CL) cfg.gridDim = dim3((unsigned)(max_clusters * CL));
        VQB_CUDA_TRY(cudaLaunchKernelEx(&cfg, conv1x1_tc_kernel<CL>, mhi, mlo, p));
    } else {
        VQB_CUDA_TRY(cudaFuncSetAttribute(conv1x1_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr[0].val.clusterDim.x = 1;
        cfg.gridDim = dim3((unsigned)grid);
        VQB_CUDA_TRY(cudaLaunchKernelEx(&cfg, conv1x1_tc_kernel<1>, mhi, mlo, p));
    }
    return VQB_OK;
}

extern "C" int vqb_conv1x1_f32(const float* x, int64_t B, int Cin, int64_t HW, const float* W, const float* bias,
                               int Cout, float* y, void* workspace, size_t workspace_bytes, int algo,
                               vqb_stream_t stream) {
    return conv1x1_impl(x, B, Cin, HW, W, bias, Cout, y, workspace, workspace_bytes, algo, ConvSplitOut(), stream);
}

// pre_quant_conv fused with the quantizer's token split (vq_vae.py:115 feeding :118): y as above, plus -- while the
// accumulator is still in TMEM -- the fp16 token rows, scales, norms and rounding residuals that the fp16 tensor search
// would otherwise compute from y with split16_tokens_kernel (one read of y and a launch less).  `search_workspace` is the
// workspace the following vqb_search_f32(..., algo = VQB_ALGO_TCGEN05_F16 | VQB_SEARCH_PRESPLIT) call will be given.
extern "C" int vqb_conv1x1_split_f32(const float* x, int64_t B, int Cin, int64_t HW, const float* W, const float* bias,
                                     int Cout, float* y, void* workspace, size_t workspace_bytes, const void* codebook_pack,
                                     int K, void* search_workspace, size_t search_workspace_bytes, vqb_stream_t stream) {
    if (!codebook_pack || !search_workspace || K <= 0 || !tc16_eligible_dim(Cout)) {
        set_error("vqb_conv1x1_split_f32: needs a codebook pack, a search workspace and 16 < Cout <= %d", kTc16MaxD);
        return VQB_ERR_INVALID_ARG;
    }
    const int64_t N = B * HW;
    if (search_workspace_bytes < search_tc16_workspace_bytes(N, Cout, K) ||
        (reinterpret_cast<uintptr_t>(search_workspace) & 255u) != 0) {
        set_error("vqb_conv1x1_split_f32: search workspace too small or misaligned");
        return VQB_ERR_WORKSPACE;
    }
    ConvSplitOut so;
    tc16_split_pointers(search_workspace, N, Cout, &so.z16, &so.inv_scale, &so.znorm, &so.zres);
    so.header = reinterpret_cast<const int*>(codebook_pack);
    so.Dpad = tc16_dpad(Cout);
    return conv1x1_impl(x, B, Cin, HW, W, bias, Cout, y, workspace, workspace_bytes, 1, so, stream);
}
