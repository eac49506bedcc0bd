// Nearest-code search for high-dimensional latents on the 5th-gen tensor cores
// (tcgen05.mma, TMEM accumulators, TMA-fed), D in {64, 128, 192, 256}.
// Replaces quantizer.py:68-76 of the reference without writing [N, K] to HBM.
//
//   score[i,k] = 0.5|e_k|^2 - z_i.e_k,   z.e ~= zh.eh + zl.eh + zh.el   (bf16x3 split,
//   fp32 accumulation in TMEM, ~2^-16 relative), row-wise top-2 in the epilogue.
//
// Pipeline per CTA (persistent, one CTA per SM, 128-token tiles):
//   warp 0   TMA producer : A = zh|zl token tile (resident for the whole codebook
//                           sweep), B = 256-code x 64-dim blocks of ehi / elo through
//                           a ring of 32 KB stages (128B-swizzled, K-major)
//   warp 1   MMA issuer   : per 64-dim block 12 x tcgen05.mma 128x256x16 into one of
//                           two 256-column TMEM accumulators
//   warp 2   TMEM alloc / dealloc
//   warps 4-7 epilogue    : tcgen05.ld the finished accumulator while the next one is
//                           being computed; running (min1, idx1, min2) per token row
// Tokens whose top-2 gap is within the proven bf16x3 error bound are appended to
// a list and re-scored exactly by the fp32 tile kernel (vqb_search_fp32.cu).
#include "vqb_tc_common.cuh"

namespace vqb {

// ---------------------------------------------------------------------------
// pre-pass: z[B, D, HW] fp32 -> zh, zl [N, D] bf16 (token-major) and |z_i|
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    split_tokens_kernel(const float* __restrict__ z, int64_t N, int D, int64_t HW,
                        __nv_bfloat16* __restrict__ zh, __nv_bfloat16* __restrict__ zl,
                        float* __restrict__ znorm) {
    __shared__ float tile[64][33];
    __shared__ float part[8][32];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int64_t t0 = (int64_t)blockIdx.x * 32;
    const int64_t tok = t0 + tx;
    const bool ok = tok < N;
    int64_t off = 0;
    if (ok) {
        const int64_t b = tok / HW;
        off = (b * D) * HW + (tok - b * HW);
    }
    float sq = 0.f;
    for (int d0 = 0; d0 < D; d0 += 64) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int d = ty + 8 * i;
            const float v = ok ? __ldg(z + off + (int64_t)(d0 + d) * HW) : 0.f;
            tile[d][tx] = v;
            sq = fmaf(v, v, sq);
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = ty + 8 * i;  // token within the block
            const int64_t t = t0 + r;
            if (t < N) {
                const float a = tile[2 * tx][r], b2 = tile[2 * tx + 1][r];
                const __nv_bfloat16 ah = __float2bfloat16_rn(a), bh = __float2bfloat16_rn(b2);
                const __nv_bfloat16 al = __float2bfloat16_rn(a - __bfloat162float(ah));
                const __nv_bfloat16 bl = __float2bfloat16_rn(b2 - __bfloat162float(bh));
                const size_t o = (size_t)t * D + d0 + 2 * tx;
                *reinterpret_cast<__nv_bfloat162*>(zh + o) = __nv_bfloat162(ah, bh);
                *reinterpret_cast<__nv_bfloat162*>(zl + o) = __nv_bfloat162(al, bl);
            }
        }
        __syncthreads();
    }
    part[ty][tx] = sq;
    __syncthreads();
    if (ty == 0 && ok) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) s += part[i][tx];
        znorm[tok] = sqrtf(s);
    }
}

// ---------------------------------------------------------------------------
// main kernel
// ---------------------------------------------------------------------------
struct TcParams {
    int64_t N;
    int K, Kpad;
    const float* half_norm;
    const int* header;  // [0] first NaN code, [4] bits of max half norm
    const float* znorm;
    int64_t* idx_out;
    float* dmin_out;
    int32_t* list;
    int32_t* list_count;
    float tau_scale;
};

template <int NKB>
__global__ void __launch_bounds__(kTcThreads, 1)
    search_tc_kernel(const __grid_constant__ CUtensorMap map_zh, const __grid_constant__ CUtensorMap map_zl,
                     const __grid_constant__ CUtensorMap map_eh, const __grid_constant__ CUtensorMap map_el,
                     TcParams p) {
    constexpr int kStages = tc_stages(NKB);
    static_assert(kStages >= 2, "not enough shared memory for the B ring");
    extern __shared__ unsigned char smem_unaligned[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_unaligned) + 1023) &
                                                           ~(uintptr_t)1023);
    unsigned char* a_hi = smem;
    unsigned char* a_lo = smem + NKB * kTcABlockBytes;
    unsigned char* b_ring = smem + 2 * NKB * kTcABlockBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(b_ring + kStages * kTcBStageBytes);
    uint64_t* a_full = bars + 0;
    uint64_t* a_empty = bars + 1;
    uint64_t* tm_full = bars + 2;   // [2]
    uint64_t* tm_empty = bars + 4;  // [2]
    uint64_t* b_full = bars + 6;    // [kStages]
    uint64_t* b_empty = bars + 6 + kStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6 + 2 * kStages);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_m_tiles = (int)((p.N + kTcBM - 1) / kTcBM);
    const int n_n_tiles = p.Kpad / kTcBN;

    if (threadIdx.x == 0) {
        tc_mbar_init(a_full, 1);
        tc_mbar_init(a_empty, 1);
        for (int i = 0; i < 2; ++i) {
            tc_mbar_init(tm_full + i, 1);
            tc_mbar_init(tm_empty + i, 128);
        }
        for (int i = 0; i < kStages; ++i) {
            tc_mbar_init(b_full + i, 1);
            tc_mbar_init(b_empty + i, 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(tmem_slot)), "r"(512)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            uint32_t stage = 0, bphase = 0, aphase = 0;
            for (int mt = blockIdx.x; mt < n_m_tiles; mt += gridDim.x) {
                tc_mbar_wait(a_empty, aphase ^ 1);
                tc_mbar_expect_tx(a_full, 2 * NKB * kTcABlockBytes);
#pragma unroll
                for (int kb = 0; kb < NKB; ++kb) {
                    tma_load_2d(a_hi + kb * kTcABlockBytes, &map_zh, a_full, kb * kTcBK, mt * kTcBM);
                    tma_load_2d(a_lo + kb * kTcABlockBytes, &map_zl, a_full, kb * kTcBK, mt * kTcBM);
                }
                aphase ^= 1;
                for (int nt = 0; nt < n_n_tiles; ++nt) {
                    for (int kb = 0; kb < NKB; ++kb) {
#pragma unroll
                        for (int half = 0; half < 2; ++half) {
                            tc_mbar_wait(b_empty + stage, bphase ^ 1);
                            tc_mbar_expect_tx(b_full + stage, kTcBStageBytes);
                            tma_load_2d(b_ring + stage * kTcBStageBytes, half == 0 ? &map_eh : &map_el,
                                        b_full + stage, kb * kTcBK, nt * kTcBN);
                            if (++stage == kStages) {
                                stage = 0;
                                bphase ^= 1;
                            }
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            uint32_t stage = 0, bphase = 0, aphase = 0, acc = 0, accphase = 0;
            for (int mt = blockIdx.x; mt < n_m_tiles; mt += gridDim.x) {
                tc_mbar_wait(a_full, aphase);
                aphase ^= 1;
                tc_fence_after();
                for (int nt = 0; nt < n_n_tiles; ++nt) {
                    tc_mbar_wait(tm_empty + acc, accphase ^ 1);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + acc * kTcBN;
                    for (int kb = 0; kb < NKB; ++kb) {
                        const uint32_t ah = s32(a_hi + kb * kTcABlockBytes);
                        const uint32_t al = s32(a_lo + kb * kTcABlockBytes);
                        // ---- B = ehi block: zh.eh and zl.eh
                        tc_mbar_wait(b_full + stage, bphase);
                        tc_fence_after();
                        {
                            const uint32_t bs = s32(b_ring + stage * kTcBStageBytes);
#pragma unroll
                            for (int k4 = 0; k4 < kTcBK / 16; ++k4)
                                umma_bf16(d_tmem, umma_desc_sw128(ah + k4 * 32), umma_desc_sw128(bs + k4 * 32),
                                          kTcIdesc, (kb | k4) != 0);
#pragma unroll
                            for (int k4 = 0; k4 < kTcBK / 16; ++k4)
                                umma_bf16(d_tmem, umma_desc_sw128(al + k4 * 32), umma_desc_sw128(bs + k4 * 32),
                                          kTcIdesc, 1);
                        }
                        umma_commit(b_empty + stage);
                        if (++stage == kStages) {
                            stage = 0;
                            bphase ^= 1;
                        }
                        // ---- B = elo block: zh.el
                        tc_mbar_wait(b_full + stage, bphase);
                        tc_fence_after();
                        {
                            const uint32_t bs = s32(b_ring + stage * kTcBStageBytes);
#pragma unroll
                            for (int k4 = 0; k4 < kTcBK / 16; ++k4)
                                umma_bf16(d_tmem, umma_desc_sw128(ah + k4 * 32), umma_desc_sw128(bs + k4 * 32),
                                          kTcIdesc, 1);
                        }
                        umma_commit(b_empty + stage);
                        if (++stage == kStages) {
                            stage = 0;
                            bphase ^= 1;
                        }
                    }
                    umma_commit(tm_full + acc);
                    if (++acc == 2) {
                        acc = 0;
                        accphase ^= 1;
                    }
                }
                umma_commit(a_empty);  // every MMA reading this token tile has retired
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue =====================
        const int q = warp - 4;  // TMEM lane quarter this warp may read (warp % 4)
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        const int first_nan = p.header[0];
        const float e_max = sqrtf(2.f * __int_as_float(p.header[4]));
        uint32_t acc = 0, accphase = 0;
        for (int mt = blockIdx.x; mt < n_m_tiles; mt += gridDim.x) {
            float m1 = INFINITY, m2 = INFINITY;
            int i1 = 0;
            for (int nt = 0; nt < n_n_tiles; ++nt) {
                tc_mbar_wait(tm_full + acc, accphase);
                tc_fence_after();
                const uint32_t t_acc = tmem_base + lane_addr + acc * kTcBN;
                const float* hrow = p.half_norm + nt * kTcBN;
#pragma unroll 1
                for (int c0 = 0; c0 < kTcBN; c0 += 32) {
                    uint32_t r[32];
                    tmem_ld32(t_acc + c0, r);
#pragma unroll
                    for (int c4 = 0; c4 < 8; ++c4) {
                        const float4 h4 = __ldg(reinterpret_cast<const float4*>(hrow + c0) + c4);
                        const float hh[4] = {h4.x, h4.y, h4.z, h4.w};
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const float s = hh[u] - __uint_as_float(r[c4 * 4 + u]);
                            const int k = nt * kTcBN + c0 + c4 * 4 + u;
                            m2 = fminf(m2, fmaxf(s, m1));
                            i1 = (s < m1) ? k : i1;
                            m1 = fminf(m1, s);
                        }
                    }
                }
                tc_fence_before();
                tc_mbar_arrive(tm_empty + acc);
                if (++acc == 2) {
                    acc = 0;
                    accphase ^= 1;
                }
            }
            const int64_t row = (int64_t)mt * kTcBM + q * 32 + lane;
            if (row < p.N) {
                const float tau = p.tau_scale * p.znorm[row] * e_max;
                const bool sure = (m2 - m1) > tau;  // false for NaN / inf rows too
                p.idx_out[row] = i1;
                if (p.dmin_out) p.dmin_out[row] = m1;
                if (!sure || first_nan < p.K) {
                    const int slot = atomicAdd(p.list_count, 1);
                    p.list[slot] = (int32_t)row;
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

// exact fp32 score of the chosen code (same fmaf order as search_fp32_kernel, so the value is
// bit-identical to what that kernel reports): makes dmin comparable across codebook shards
__global__ void __launch_bounds__(256)
    exact_score_kernel(const float* __restrict__ z, const float* __restrict__ E, const float* __restrict__ half_norm,
                       const int64_t* __restrict__ idx, int64_t N, int D, int64_t HW, int K,
                       float* __restrict__ dmin_out) {
    const int64_t tok = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tok >= N) return;
    int64_t k = idx[tok];
    if (k < 0 || k >= K) k = 0;
    const int64_t b = tok / HW;
    const float* zp = z + (b * D) * HW + (tok - b * HW);
    const float4* erow = reinterpret_cast<const float4*>(E + (size_t)k * D);
    float acc = 0.f;
    for (int d = 0; d < D; d += 4) {
        const float4 e = __ldg(erow + (d >> 2));
        acc = fmaf(__ldg(zp + (int64_t)(d + 0) * HW), e.x, acc);
        acc = fmaf(__ldg(zp + (int64_t)(d + 1) * HW), e.y, acc);
        acc = fmaf(__ldg(zp + (int64_t)(d + 2) * HW), e.z, acc);
        acc = fmaf(__ldg(zp + (int64_t)(d + 3) * HW), e.w, acc);
    }
    const float d = half_norm[k] - acc;
    dmin_out[tok] = (d != d) ? INFINITY : d;  // all-NaN rows report +inf like the fp32 kernel
}

int launch_exact_score(const float* z, const float* E, const float* half_norm, const int64_t* idx, int64_t N, int D,
                       int64_t HW, int K, float* dmin_out, cudaStream_t s) {
    exact_score_kernel<<<(unsigned)((N + 255) / 256), 256, 0, s>>>(z, E, half_norm, idx, N, D, HW, K, dmin_out);
    VQB_LAUNCH_CHECK("exact_score_kernel");
    return VQB_OK;
}

__global__ void tc_stats_kernel(int64_t* stats, const int32_t* count) {
    stats[0] = *count;
    stats[1] = VQB_ALGO_TCGEN05;
    stats[2] = 0;
    stats[3] = 0;
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
struct TcWorkspace {
    size_t off_zh, off_zl, off_znorm, off_list, off_count, off_keys, keys_bytes, total;
};

static TcWorkspace tc_workspace(int64_t N, int D) {
    TcWorkspace w;
    size_t off = 0;
    w.off_zh = off;
    off = round_up_z(off + 2 * (size_t)N * D, 1024);
    w.off_zl = off;
    off = round_up_z(off + 2 * (size_t)N * D, 1024);
    w.off_znorm = off;
    off = round_up_z(off + 4 * (size_t)N, 1024);
    w.off_list = off;
    off = round_up_z(off + 4 * (size_t)N, 1024);
    w.off_count = off;
    off += 1024;
    w.off_keys = off;
    w.keys_bytes = search_fp32_workspace_bytes(N, D);
    off = round_up_z(off + w.keys_bytes, 1024);
    w.total = off;
    return w;
}

size_t search_tc_workspace_bytes(int64_t n_tokens, int D, int K) {
    (void)K;
    return tc_workspace(n_tokens, D).total;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

// rows x D 16-bit row-major, box = 64 columns x box_rows, 128B swizzle
int make_tc_map(CUtensorMap* map, const void* base, uint64_t rows, int D, uint32_t box_rows, bool fp16) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled is unavailable in this driver");
        return VQB_ERR_CUDA;
    }
    cuuint64_t dims[2] = {(cuuint64_t)D, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)D * 2};
    cuuint32_t box[2] = {(cuuint32_t)kTcBK, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with %d (rows=%llu D=%d)", (int)r, (unsigned long long)rows, D);
        return VQB_ERR_CUDA;
    }
    return VQB_OK;
}

// rows x cols fp32 row-major, box = 32 columns (128 B) x box_rows, 128B swizzle (tf32 operands)
int make_tc_map_f32(CUtensorMap* map, const void* base, uint64_t rows, int cols, uint32_t box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled is unavailable in this driver");
        return VQB_ERR_CUDA;
    }
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)cols * 4};
    cuuint32_t box[2] = {32u, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled (fp32) failed with %d (rows=%llu cols=%d)", (int)r, (unsigned long long)rows, cols);
        return VQB_ERR_CUDA;
    }
    return VQB_OK;
}

template <int NKB>
static int launch_tc_t(const CUtensorMap& mzh, const CUtensorMap& mzl, const CUtensorMap& meh,
                       const CUtensorMap& mel, const TcParams& p, cudaStream_t s) {
    constexpr int kStages = tc_stages(NKB);
    const size_t smem = 1024 + 2 * NKB * kTcABlockBytes + (size_t)kStages * kTcBStageBytes + kTcBarrierBytes;
    VQB_CUDA_TRY(cudaFuncSetAttribute(search_tc_kernel<NKB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int n_m_tiles = (int)((p.N + kTcBM - 1) / kTcBM);
    const int grid = n_m_tiles < sm_count() ? n_m_tiles : sm_count();
    search_tc_kernel<NKB><<<grid, kTcThreads, smem, s>>>(mzh, mzl, meh, mel, p);
    VQB_LAUNCH_CHECK("search_tc_kernel");
    return VQB_OK;
}

int launch_search_tc(const float* z, int64_t B, int D, int64_t HW, const float* E, int K, const void* pack,
                     int64_t* idx_out, float* dmin_out, void* ws, size_t ws_bytes, int64_t* stats_out,
                     cudaStream_t s) {
    const int64_t N = B * HW;
    const TcWorkspace w = tc_workspace(N, D);
    if (!ws || ws_bytes < w.total) {
        set_error("tcgen05 search workspace too small: %zu < %zu", ws_bytes, w.total);
        return VQB_ERR_WORKSPACE;
    }
    if ((reinterpret_cast<uintptr_t>(ws) & 255u) != 0) {
        set_error("tcgen05 search workspace must be 256-byte aligned");
        return VQB_ERR_INVALID_ARG;
    }
    const PackLayout L = pack_layout(K, D);
    unsigned char* wsb = static_cast<unsigned char*>(ws);
    const unsigned char* pk = static_cast<const unsigned char*>(pack);
    __nv_bfloat16* zh = reinterpret_cast<__nv_bfloat16*>(wsb + w.off_zh);
    __nv_bfloat16* zl = reinterpret_cast<__nv_bfloat16*>(wsb + w.off_zl);
    float* znorm = reinterpret_cast<float*>(wsb + w.off_znorm);
    int32_t* list = reinterpret_cast<int32_t*>(wsb + w.off_list);
    int32_t* count = reinterpret_cast<int32_t*>(wsb + w.off_count);

    VQB_CUDA_TRY(cudaMemsetAsync(count, 0, sizeof(int32_t), s));
    split_tokens_kernel<<<(unsigned)((N + 31) / 32), 256, 0, s>>>(z, N, D, HW, zh, zl, znorm);
    VQB_LAUNCH_CHECK("split_tokens_kernel");

    CUtensorMap mzh, mzl, meh, mel;
    if (int rc = make_tc_map(&mzh, zh, (uint64_t)N, D, kTcBM, false)) return rc;
    if (int rc = make_tc_map(&mzl, zl, (uint64_t)N, D, kTcBM, false)) return rc;
    if (int rc = make_tc_map(&meh, pk + L.off_ehi, (uint64_t)L.Kpad, D, kTcBN, false)) return rc;
    if (int rc = make_tc_map(&mel, pk + L.off_elo, (uint64_t)L.Kpad, D, kTcBN, false)) return rc;

    TcParams p;
    p.N = N;
    p.K = K;
    p.Kpad = L.Kpad;
    p.half_norm = reinterpret_cast<const float*>(pk + L.off_half_norm);
    p.header = reinterpret_cast<const int*>(pk);
    p.znorm = znorm;
    p.idx_out = idx_out;
    p.dmin_out = dmin_out;
    p.list = list;
    p.list_count = count;
    // |z.e - bf16x3(z.e)| <= 3 * 2^-16 |z||e| (two bf16 roundings per operand, dropped lo*lo
    // term); a top-2 gap above twice that bound cannot flip.  1.25 covers fp32 accumulation.
    p.tau_scale = 1.25f * 6.0f / 65536.0f;
    int rc;
    switch (D / kTcBK) {
        case 1: rc = launch_tc_t<1>(mzh, mzl, meh, mel, p, s); break;
        case 2: rc = launch_tc_t<2>(mzh, mzl, meh, mel, p, s); break;
        case 3: rc = launch_tc_t<3>(mzh, mzl, meh, mel, p, s); break;
        case 4: rc = launch_tc_t<4>(mzh, mzl, meh, mel, p, s); break;
        default:
            set_error("tcgen05 search supports D in {64,128,192,256}, got %d", D);
            return VQB_ERR_UNSUPPORTED;
    }
    if (rc != VQB_OK) return rc;
    // exact fp32 re-score of the flagged tokens: persistent CTAs over (token tile, code split)
    // work items, the token count stays on the device
    rc = launch_search_fp32(z, B, D, HW, E, K, pack, list, count, N, wsb + w.off_keys, w.keys_bytes, idx_out,
                            dmin_out, s);
    if (rc != VQB_OK) return rc;
    if (dmin_out) {
        exact_score_kernel<<<(unsigned)((N + 255) / 256), 256, 0, s>>>(z, E, p.half_norm, idx_out, N, D, HW, K,
                                                                      dmin_out);
        VQB_LAUNCH_CHECK("exact_score_kernel");
    }
    if (stats_out) {
        tc_stats_kernel<<<1, 1, 0, s>>>(stats_out, count);
        VQB_LAUNCH_CHECK("tc_stats_kernel");
    }
    return VQB_OK;
}

}  // namespace vqb
