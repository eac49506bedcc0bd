// Low-D nearest-code search on the tensor cores (D <= 16): tf32x3 split, certified winning chunk,
// exact fp32 re-score.  Replaces quantizer.py:68-76 of the reference with the same contract (and,
// by construction, the same indices) as the CUDA-core kernel of vqb_search_lowd.cu.
//
// Why: at D = 4 the CUDA-core kernel spends 4 FFMA2 + 1 FMNMX3 per (token, code pair) and is bound
// by the FP32 pipe.  Here the multiply-adds move to tcgen05 and the CUDA cores only keep the
// running minimum: ~0.66 ALU instructions per score instead of ~5 issue slots.
//
//   approx[i,k] = sum over 3D+3 tf32 slots of A[i,s] * B[k,s]  (fp32 accumulation in TMEM)
//     A[i] = [ zh | zl | zh | 1 1 1 ]   B[k] = [ -eh | -eh | -el | h1 h2 h3 ],  z = zh + zl + rz, e = eh + el + re
//   |approx - exact| <= eps_i = |rz_i| max|e| + |z_i| max|re| + |zl_i| max|el| + accumulation term
//   (Cauchy-Schwarz on the MEASURED split residuals, pre-passes below).
//
// Epilogue (8 warps; the two warps of a TMEM lane quarter alternate code tiles): per 32-code chunk
// the minimum of the 32 scores (16 FMNMX3), then a top-2 update over chunk minima with the chunk id
// of the best (5 ALU ops).  A token is SURE iff second-best chunk minimum - best > 2 eps_i: then
// every code that could win lies in the best chunk, whose 32 codes are re-scored exactly (same FMA
// chain as the CUDA-core kernel, lowest index on ties).  Unsure tokens (~1 %), NaN/inf inputs and
// NaN codebooks go through the CUDA-core kernel itself (token-list mode), so the result is
// bit-identical to VQB_ALGO_LOWD_FMA.
//
// Pipeline: persistent CTA per SM, 128-token tiles (A double buffered), 128-code B stages streamed
// with 1-D bulk copies of pre-built shared-memory images (SWIZZLE_NONE K-major; cluster multicast:
// every CTA of a cluster fetches 1/CL of each stage), four 128-column TMEM accumulators.
#include "vqb_lowd_body.cuh"
#include "vqb_tc_common.cuh"

namespace vqb {

constexpr int kLowThreads = 384;     // warp 0 producer, 1 MMA issuer, 2 TMEM allocator, 4-11 epilogue
constexpr int kLowEpiThreads = 256;
constexpr int kLowBN = 128;          // codes per B stage / accumulator
constexpr int kLowAcc = 4;           // TMEM accumulators (4 x 128 columns)
constexpr int kLowMaxStages = 8;
constexpr int kLowChunk = 32;        // codes per certified chunk (one tcgen05.ld.x32)

// A, B = tf32 (format 2), accumulator f32, both K-major, N = 128, M = 128
constexpr uint32_t kLowIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kLowBN >> 3) << 17) |
                               ((uint32_t)(kLowRows >> 4) << 24);

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// K-major, no swizzle: 8-row x 16-byte core matrices, k-chunks 2048 B apart (LBO), row groups 128 B apart (SBO)
__device__ __forceinline__ uint64_t umma_desc_interleave(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3fff);
    d |= (uint64_t)((kLowRows * 16) >> 4) << 16;  // leading byte offset: next 16-byte chunk along K
    d |= (uint64_t)(128 >> 4) << 32;              // stride byte offset: next 8-row group
    d |= (uint64_t)1 << 46;                       // descriptor version (sm_100)
    return d;                                     // layout type 0 = SWIZZLE_NONE
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(dst)),
        "l"(src), "r"(bytes), "r"(s32(bar))
        : "memory");
}
__device__ __forceinline__ void bulk_load_mc(void* dst, const void* src, uint32_t bytes, uint64_t* bar,
                                             uint16_t mask) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::
            "r"(s32(dst)),
        "l"(src), "r"(bytes), "r"(s32(bar)), "h"(mask)
        : "memory");
}

// ---------------------------------------------------------------------------
// pre-pass: z[B, D, HW] fp32 -> tf32x3 token image (tiles of 128 rows) and tau_i = 2 eps_i
// ---------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(128)
    split_tf32_tokens_kernel(const float* __restrict__ z, int64_t N, int64_t HW, const int* __restrict__ header,
                             float* __restrict__ img, float4* __restrict__ tok_norms) {
    constexpr int kSlots = 8 * ((3 * D + 3 + 7) / 8);
    const int64_t tok = (int64_t)blockIdx.x * 128 + threadIdx.x;  // block = one 128-row tile
    const bool ok = tok < N;
    float* tile = img + (size_t)blockIdx.x * (kLowRows * kSlots);
    const int r = threadIdx.x;
    float v[kSlots];
    float zz = 0.f, res = 0.f, lo2 = 0.f;
    const int64_t b = ok ? tok / HW : 0;
    const float* zp = z + (b * D) * HW + (ok ? tok - b * HW : 0);
#pragma unroll
    for (int d = 0; d < D; ++d) {
        const float x = ok ? __ldg(zp + (int64_t)d * HW) : 0.f;
        const float zh = to_tf32(x);
        const float rem = x - zh;
        const float zl = to_tf32(rem);
        const float r2 = rem - zl;
        zz = fmaf(x, x, zz);
        res = fmaf(r2, r2, res);
        lo2 = fmaf(zl, zl, lo2);
        v[d] = zh;
        v[D + d] = zl;
        v[2 * D + d] = zh;
    }
    v[3 * D] = v[3 * D + 1] = v[3 * D + 2] = 1.f;
#pragma unroll
    for (int s = 3 * D + 3; s < kSlots; ++s) v[s] = 0.f;
#pragma unroll
    for (int c = 0; c < kSlots / 4; ++c)  // one 16-byte chunk per store, consecutive rows are adjacent
        *reinterpret_cast<float4*>(tile + c * (kLowRows * 4) + (r >> 3) * 32 + (r & 7) * 4) =
            make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
    (void)header;
    // |z|, |z - zh - zl|, |zl| (NaN / inf coordinates propagate: the token is then re-searched exactly)
    if (ok) tok_norms[tok] = make_float4(sqrtf(zz), sqrtf(res), sqrtf(lo2), 0.f);
}

// ---------------------------------------------------------------------------
// main kernel
// ---------------------------------------------------------------------------
struct LowParams {
    int64_t N;
    int K, Kpad, steps;        // steps = k-steps of 8 tf32 slots
    int n_stages;
    const float* a_img;        // token image, n_m_tiles tiles
    const float* b_img;        // codebook image, Kpad / 128 tiles
    const int* header;
    const float4* tok_norms;   // per token: |z|, |z - zh - zl|, |zl|
    const float* cmax;         // [Kpad/32] max code norm per 32-code chunk
    int32_t* chunk;            // [N] winning 32-code chunk
    int32_t* list;             // tokens that need the exact full search
    int32_t* list_count;
    int share_sm;              // host only: launched next to the CUDA-core kernel (two-engine search)
};

// barrier over the 384 threads of the tensor role: the whole CTA (BAR = 0) or a named barrier (search_dual_kernel)
template <int BAR>
__device__ __forceinline__ void tclow_sync() {
    if constexpr (BAR == 0)
        __syncthreads();
    else
        asm volatile("bar.sync %0, %1;" ::"n"(BAR), "n"(kLowThreads) : "memory");
}

// The CTA program of the tensor search (384 threads: TID0 .. TID0 + 383): the body of search_tclow_kernel and the
// tensor role of search_dual_kernel.  The cluster barriers inside are executed by EVERY thread of the cluster's CTAs
// (search_dual_kernel makes its FMA role arrive at the same two points).
template <int CL, int TID0, int BAR>
__device__ __forceinline__ void tclow_cta_body(const LowParams& p, unsigned char* smem_unaligned, int cta_index, int n_ctas) {  // <= 85 registers: can share an SM with the CUDA-core kernel
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_unaligned) + 1023) &
                                                           ~(uintptr_t)1023);
    const uint32_t tile_bytes = (uint32_t)p.steps * (kLowRows * 32);  // A tile and B stage have the same shape
    unsigned char* a_buf = smem;                        // 2 tiles
    unsigned char* b_ring = smem + 2 * tile_bytes;      // n_stages tiles
    uint64_t* bars = reinterpret_cast<uint64_t*>(b_ring + (size_t)p.n_stages * tile_bytes);
    uint64_t* a_full = bars + 0;                        // [2]
    uint64_t* a_empty = bars + 2;                       // [2]
    uint64_t* tm_full = bars + 4;                       // [4]
    uint64_t* tm_empty = bars + 8;                      // [4]
    uint64_t* b_full = bars + 12;                       // [kLowMaxStages]
    uint64_t* b_empty = bars + 12 + kLowMaxStages;      // [kLowMaxStages]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12 + 2 * kLowMaxStages);
    float* xchg = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(bars) + 512);  // [2][128][4]

    const int tid = (int)threadIdx.x - TID0;  // thread index within the tensor role
    const int warp = tid >> 5, lane = tid & 31;
    const int n_m_tiles = (int)((p.N + kLowRows - 1) / kLowRows);
    const int n_n_tiles = p.Kpad / kLowBN;  // even: Kpad is a multiple of 256
    const int n_rounds = (n_m_tiles + n_ctas - 1) / n_ctas;
    const uint32_t cta_rank = CL > 1 ? cluster_ctarank() : 0;
    constexpr uint16_t kMask = (uint16_t)((1u << CL) - 1);
    const int n_stages = p.n_stages;

    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            tc_mbar_init(a_full + i, 1);
            tc_mbar_init(a_empty + i, 1);
        }
        for (int i = 0; i < kLowAcc; ++i) {
            tc_mbar_init(tm_full + i, 1);
            tc_mbar_init(tm_empty + i, kLowEpiThreads / 2);  // the four warps that own this accumulator's parity
        }
        for (int i = 0; i < n_stages; ++i) {
            tc_mbar_init(b_full + i, 1);
            tc_mbar_init(b_empty + i, CL);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(tmem_slot)), "r"(512)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    tclow_sync<BAR>();
    if constexpr (CL > 1) cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== producer =====================
        if (lane == 0) {
            uint32_t stage = 0, bphase = 0;
            for (int round = 0; round < n_rounds; ++round) {
                int mt = cta_index + round * n_ctas;
                if (mt >= n_m_tiles) mt = n_m_tiles - 1;  // padding round of a cluster: rows are dropped later
                const int ab = round & 1;
                tc_mbar_wait(a_empty + ab, ((round >> 1) & 1) ^ 1);
                tc_mbar_expect_tx(a_full + ab, tile_bytes);
                bulk_load(a_buf + ab * tile_bytes, reinterpret_cast<const unsigned char*>(p.a_img) + (size_t)mt * tile_bytes,
                          tile_bytes, a_full + ab);
                for (int nt = 0; nt < n_n_tiles; ++nt) {
                    tc_mbar_wait(b_empty + stage, bphase ^ 1);
                    tc_mbar_expect_tx(b_full + stage, tile_bytes);
                    const unsigned char* src = reinterpret_cast<const unsigned char*>(p.b_img) + (size_t)nt * tile_bytes;
                    unsigned char* dst = b_ring + (size_t)stage * tile_bytes;
                    if constexpr (CL > 1) {
                        const uint32_t part = tile_bytes / CL;
                        bulk_load_mc(dst + cta_rank * part, src + cta_rank * part, part, b_full + stage, kMask);
                    } else {
                        bulk_load(dst, src, tile_bytes, b_full + stage);
                    }
                    if (++stage == (uint32_t)n_stages) {
                        stage = 0;
                        bphase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            uint32_t stage = 0, bphase = 0, g = 0;  // g = accumulator use counter
            for (int round = 0; round < n_rounds; ++round) {
                const int ab = round & 1;
                tc_mbar_wait(a_full + ab, (round >> 1) & 1);
                tc_fence_after();
                const uint32_t as = s32(a_buf + ab * tile_bytes);
                for (int nt = 0; nt < n_n_tiles; ++nt, ++g) {
                    const uint32_t acc = g & (kLowAcc - 1);
                    tc_mbar_wait(tm_empty + acc, ((g >> 2) & 1) ^ 1);
                    tc_mbar_wait(b_full + stage, bphase);
                    tc_fence_after();
                    const uint32_t bs = s32(b_ring + (size_t)stage * tile_bytes);
                    const uint32_t d_tmem = tmem_base + acc * kLowBN;
                    for (int ks = 0; ks < p.steps; ++ks)  // one k-step = 8 tf32 = two 16-byte chunks
                        umma_tf32(d_tmem, umma_desc_interleave(as + ks * (2 * kLowRows * 16)),
                                  umma_desc_interleave(bs + ks * (2 * kLowRows * 16)), kLowIdesc, ks != 0);
                    if constexpr (CL > 1)
                        umma_commit_mc(b_empty + stage, kMask);
                    else
                        umma_commit(b_empty + stage);
                    umma_commit(tm_full + acc);
                    if (++stage == (uint32_t)n_stages) {
                        stage = 0;
                        bphase ^= 1;
                    }
                }
                umma_commit(a_empty + ab);
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue (8 warps) =====================
        const int q = warp & 3;            // TMEM lane quarter this warp may read
        const int par = (warp - 4) >> 2;   // code tiles with (nt & 1) == par
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        const int first_nan = p.header[0];
        const float e_max = sqrtf(2.f * __int_as_float(p.header[4]));
        const float rho_r = __int_as_float(p.header[12]), rho_l = __int_as_float(p.header[14]);
        const int row_in_tile = q * 32 + lane;
        for (int round = 0; round < n_rounds; ++round) {
            const int mt = cta_index + round * n_ctas;
            const int64_t row = (int64_t)mt * kLowRows + row_in_tile;
            float m1 = INFINITY, m2 = INFINITY;
            int id = 0;
            const uint32_t g0 = (uint32_t)round * (uint32_t)n_n_tiles;
            for (int nt = par; nt < n_n_tiles; nt += 2) {
                const uint32_t g = g0 + nt;
                const uint32_t acc = g & (kLowAcc - 1);
                tc_mbar_wait(tm_full + acc, (g >> 2) & 1);
                tc_fence_after();
                const uint32_t t_acc = tmem_base + lane_addr + acc * kLowBN;
                // one TMEM load in flight per warp: the partner warp of this scheduler (other tile
                // parity) covers the latency; double buffering measured no faster -- the kernel is
                // bound by TMEM traffic (64 KB written by the MMAs + 64 KB read back per 128x128 tile)
                uint32_t r[32];
#pragma unroll
                for (int j = 0; j < kLowBN / kLowChunk; ++j) {
                    tmem_ld32_issue(t_acc + j * kLowChunk, r);
                    tmem_ld_wait();
                    // minimum of the 32 scores: five independent FMNMX3 chains, then combined
                    float c0 = min3_f32(__uint_as_float(r[0]), __uint_as_float(r[1]), __uint_as_float(r[2]));
                    float c1 = min3_f32(__uint_as_float(r[3]), __uint_as_float(r[4]), __uint_as_float(r[5]));
                    float c2 = min3_f32(__uint_as_float(r[6]), __uint_as_float(r[7]), __uint_as_float(r[8]));
                    float c3 = min3_f32(__uint_as_float(r[9]), __uint_as_float(r[10]), __uint_as_float(r[11]));
                    float c4 = min3_f32(__uint_as_float(r[12]), __uint_as_float(r[13]), __uint_as_float(r[14]));
                    c0 = min3_f32(c0, __uint_as_float(r[15]), __uint_as_float(r[16]));
                    c1 = min3_f32(c1, __uint_as_float(r[17]), __uint_as_float(r[18]));
                    c2 = min3_f32(c2, __uint_as_float(r[19]), __uint_as_float(r[20]));
                    c3 = min3_f32(c3, __uint_as_float(r[21]), __uint_as_float(r[22]));
                    c4 = min3_f32(c4, __uint_as_float(r[23]), __uint_as_float(r[24]));
                    c0 = min3_f32(c0, __uint_as_float(r[25]), __uint_as_float(r[26]));
                    c1 = min3_f32(c1, __uint_as_float(r[27]), __uint_as_float(r[28]));
                    c2 = min3_f32(c2, __uint_as_float(r[29]), __uint_as_float(r[30]));
                    c3 = min3_f32(c3, c4, __uint_as_float(r[31]));
                    const float c = fminf(min3_f32(c0, c1, c2), c3);
                    // top-2 over chunk minima, chunk id of the best (earlier chunk wins ties)
                    m2 = fminf(m2, fmaxf(m1, c));
                    id = (c < m1) ? (nt * (kLowBN / kLowChunk) + j) : id;
                    m1 = fminf(m1, c);
                }
                tc_fence_before();
                tc_mbar_arrive(tm_empty + acc);
            }
            // merge the two parities of each row
            float* x = xchg + ((round & 1) * kLowRows + row_in_tile) * 4;
            if (par == 1) {
                x[0] = m1;
                x[1] = m2;
                x[2] = __int_as_float(id);
            }
            asm volatile("bar.sync 1, %0;" ::"n"(kLowEpiThreads) : "memory");
            if (par == 0 && row < p.N) {
                const float o1 = x[0], o2 = x[1];
                const int oid = __float_as_int(x[2]);
                const float b1 = fminf(m1, o1);
                const float b2 = fminf(fminf(m2, o2), fmaxf(m1, o1));
                const int bid = (o1 < m1) ? oid : id;
                p.chunk[row] = bid;
                // Error of the approximate score of a code of norm n (Cauchy-Schwarz on the measured split
                // residuals: |r_e| <= rho_r n, |el| <= rho_l n for every code; fp32 accumulation of <= 8*steps
                // exact products inside the tensor core 2^-18 of the term magnitudes, 2^-20 for the
                // three-piece half norm and this estimate itself):
                //   eps(n) = (|r_z| + |z| rho_r + |zl| rho_l) n + (2^-18 + 2^-20) (|z| n + n^2/2)
                // Winner-norm argument as in vqb_search_tc16.cu: the approximate winner sits in chunk bid
                // (norm <= n1), exact_best <= U = b1 + eps(n1); codes with norm > n0 = |z| + sqrt(|z|^2 + 2U)
                // cannot win; all others have error <= eps(n0)  =>  tau = eps(n0) + eps(n1).
                const float4 tn = p.tok_norms[row];
                const float zn = tn.x;
                const float alpha = tn.y + zn * rho_r + tn.z * rho_l;
                const float kacc = 1.f / 262144.f + 1.f / 1048576.f;
                const float n1 = fminf(__ldg(p.cmax + bid), e_max);
                const float eps1 = alpha * n1 + kacc * (zn * n1 + 0.5f * n1 * n1);
                const float U = b1 + eps1;
                const float n0 = fminf(zn + sqrtf(fmaxf(zn * zn + 2.f * U, 0.f)), e_max);
                const float eps0 = alpha * n0 + kacc * (zn * n0 + 0.5f * n0 * n0);
                const float tau = 1.0001f * (eps0 + eps1);
                // sure iff no code outside the winning chunk can beat it (false for NaN / inf anywhere)
                const bool sure = (b2 - b1 > tau) && (b1 < 1e37f) && (first_nan >= p.K);
                if (!sure) {
                    const int slot = atomicAdd(p.list_count, 1);
                    p.list[slot] = (int32_t)row;
                }
            }
        }
    }

    tc_fence_before();
    tclow_sync<BAR>();
    if constexpr (CL > 1) cluster_sync_all();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

template <int CL>
__global__ void __launch_bounds__(kLowThreads, 2) search_tclow_kernel(LowParams p) {  // <= 85 registers
    extern __shared__ unsigned char smem_unaligned[];
    tclow_cta_body<CL, 0, 0>(p, smem_unaligned, (int)blockIdx.x, (int)gridDim.x);
}

// ---------------------------------------------------------------------------
// Two engines in ONE CTA (D = 4, config C2): warps 0-7 run the CUDA-core search (FP32 FMA pipe; bound by the one
// ALU-pipe minimum per score), warps 8-19 the tensor search (tcgen05 tf32x3; bound by TMEM read-back).  They bind
// different resources of the SM, work on disjoint token ranges, and return bit-identical results, so the split is
// invisible in the output.  (Launching the two kernels side by side on two streams does NOT work on B200: the second
// kernel only starts when the first drains -- DESIGN.md section 4.2 -- hence one kernel, where co-residency is by
// construction.)  Registers: the CTA starts at 96 per thread; the tensor warpgroups drop to 80 and the FMA warpgroups
// take 120 with setmaxnreg (the pool is what the CTA itself released: 384 x 16 = 256 x 24).
// ---------------------------------------------------------------------------
constexpr int kDualFmaThreads = 256;
constexpr int kDualThreads = kDualFmaThreads + kLowThreads;  // 640

struct DualFmaParams {
    const float* z;        // first image of the FMA range
    int64_t N, HW;
    int K;
    const unsigned char* pack;
    PackLayout L;
    int64_t tokens_per_cta;
    int64_t* idx_out;      // of the FMA range
    float* dmin_out;       // nullable
    uint32_t tensor_smem_offset;  // dynamic shared memory: [FMA tiles | tensor role]
};

template <int D, int CL>
__global__ void __launch_bounds__(kDualThreads, 1) search_dual_kernel(LowParams pt, DualFmaParams pf) {
    extern __shared__ __align__(1024) unsigned char dual_smem[];
    if (threadIdx.x < kDualFmaThreads) {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 120;");  // 256 x (120 - 96) = 6144 = what the tensor warps release
        if constexpr (CL > 1) cluster_sync_all();  // matches the tensor role's barrier after its mbarrier init
        lowd_cta_body<D, 0, false, 2>(pf.z, pf.N, pf.HW, pf.K, pf.pack, pf.L, pf.tokens_per_cta, nullptr, nullptr, pf.idx_out,
                                      pf.dmin_out, dual_smem, (int)blockIdx.x, (int)gridDim.x);
        if constexpr (CL > 1) cluster_sync_all();  // ... and the one before its CTA may leave
    } else {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 80;");
        tclow_cta_body<CL, kDualFmaThreads, 3>(pt, dual_smem + pf.tensor_smem_offset, (int)blockIdx.x, (int)gridDim.x);
    }
}

// ---------------------------------------------------------------------------
// exact re-score of the winning 32-code chunk of every token: lane = code, the FMA chain of the
// CUDA-core kernel (h - z0 e0 - z1 e1 ...), first lane that equals the minimum wins
// ---------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(256)
    rescore_chunk_kernel(const float* __restrict__ z, const float* __restrict__ E, const float* __restrict__ half_norm,
                         const int32_t* __restrict__ chunk, int64_t N, int64_t HW, int K, int64_t* __restrict__ idx_out,
                         float* __restrict__ dmin_out) {
    const int lane = threadIdx.x & 31;
    const int64_t warp_base = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 32;
    if (warp_base >= N) return;
    // lane t owns token warp_base + t for loading (coalesced along HW)
    const int64_t tok = warp_base + lane;
    const bool ok = tok < N;
    float nz[D];
    int my_chunk = 0;
    {
        const int64_t b = ok ? tok / HW : 0;
        const float* zp = z + (b * D) * HW + (ok ? tok - b * HW : 0);
#pragma unroll
        for (int d = 0; d < D; ++d) nz[d] = ok ? -__ldg(zp + (int64_t)d * HW) : 0.f;
        if (ok) my_chunk = __ldg(chunk + tok);
    }
    const int n_here = (N - warp_base) < 32 ? (int)(N - warp_base) : 32;
    int my_idx = 0;
    float my_min = INFINITY;
#pragma unroll 4
    for (int t = 0; t < n_here; ++t) {
        const int chunk_t = __shfl_sync(0xffffffffu, my_chunk, t);
        const int k = chunk_t * kLowChunk + lane;
        float zt[D];
#pragma unroll
        for (int d = 0; d < D; ++d) zt[d] = __shfl_sync(0xffffffffu, nz[d], t);  // all lanes take part
        float s = INFINITY;
        if (k < K) {
            s = __ldg(half_norm + k);
            const float* e = E + (size_t)k * D;
            if constexpr (D % 4 == 0) {
#pragma unroll
                for (int d = 0; d < D; d += 4) {
                    const float4 ev = __ldg(reinterpret_cast<const float4*>(e + d));
                    s = fmaf(zt[d], ev.x, s);
                    s = fmaf(zt[d + 1], ev.y, s);
                    s = fmaf(zt[d + 2], ev.z, s);
                    s = fmaf(zt[d + 3], ev.w, s);
                }
            } else {
#pragma unroll
                for (int d = 0; d < D; ++d) s = fmaf(zt[d], __ldg(e + d), s);
            }
        }
        float m = (s == s) ? s : INFINITY;  // NaN scores never win (such tokens are re-searched anyway)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fminf(m, __shfl_xor_sync(0xffffffffu, m, o));
        const unsigned hit = __ballot_sync(0xffffffffu, s == m && k < K);
        const int first = hit ? (__ffs(hit) - 1) : 0;
        if (lane == t) {
            my_idx = chunk_t * kLowChunk + first;
            my_min = m;
        }
    }
    if (ok) {
        idx_out[tok] = my_idx < K ? my_idx : 0;
        if (dmin_out) dmin_out[tok] = my_min;
    }
}

__global__ void tclow_stats_kernel(int64_t* stats, const int32_t* list_count) {
    stats[0] = *list_count;
    stats[1] = VQB_ALGO_TCGEN05_TF32X3;
    stats[2] = 0;
    stats[3] = 0;
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
struct LowWorkspace {
    size_t off_img, off_tau, off_chunk, off_list, off_count, total;
};

static LowWorkspace low_workspace(int64_t N, int D) {
    LowWorkspace w;
    const size_t n_m_tiles = (size_t)((N + kLowRows - 1) / kLowRows);
    size_t off = 0;
    w.off_img = off;
    off = round_up_z(off + sizeof(float) * tclow_tile_floats(D) * n_m_tiles, 1024);
    w.off_tau = off;
    off = round_up_z(off + 16 * (size_t)N, 1024);
    w.off_chunk = off;
    off = round_up_z(off + 4 * (size_t)N, 1024);
    w.off_list = off;
    off = round_up_z(off + 4 * (size_t)N, 1024);
    w.off_count = off;
    off += 1024;
    w.total = off;
    return w;
}

size_t search_tclow_workspace_bytes(int64_t n_tokens, int D, int K) {
    (void)K;
    return low_workspace(n_tokens, D).total;
}

VQB_KNOB g_tclow_cluster = 2;
VQB_KNOB g_tclow_debug = 0;  // bit 0: skip the tensor kernel, bit 1: skip the chunk re-score, bit 2: skip the list search
#ifdef VQB_EXPERIMENTAL
void set_tclow_cluster(int c) {
    if (c >= 16) g_tclow_debug = c - 16; else g_tclow_cluster = c;
}
#endif

template <int CL>
static int launch_tclow_cl(const LowParams& p0, cudaStream_t s) {
    LowParams p = p0;
    const size_t tile_bytes = (size_t)p.steps * (kLowRows * 32);
    int stages = (int)((200 * 1024 - 2 * tile_bytes) / tile_bytes);
    stages = stages > kLowMaxStages ? kLowMaxStages : stages;
    const int n_n_tiles = p.Kpad / kLowBN;
    if (stages > n_n_tiles) stages = n_n_tiles;
    if (stages < 2) stages = 2;
    p.n_stages = stages;
    size_t smem = 1024 + (2 + (size_t)stages) * tile_bytes + 512 + 2 * kLowRows * 4 * sizeof(float);
    // every CTA allocates all 512 TMEM columns: never let two of them share an SM
    // (111 KB when it shares the SM with the CUDA-core kernel of the two-engine search, which takes 113 KB)
    const size_t floor_bytes = (size_t)(p.share_sm ? 111 : 116) * 1024;
    if (smem < floor_bytes) smem = floor_bytes;
    VQB_CUDA_TRY(cudaFuncSetAttribute(search_tclow_kernel<CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // same L1 / shared split as search_lowd_kernel so that both can be resident on one SM (vqb_search_dual_f32)
    VQB_CUDA_TRY(cudaFuncSetAttribute(search_tclow_kernel<CL>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                      cudaSharedmemCarveoutMaxShared));
    VQB_CUDA_TRY(cudaFuncSetAttribute(search_tclow_kernel<CL>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                      (int)cudaSharedmemCarveoutMaxShared));
    const int n_m_tiles = (int)((p.N + kLowRows - 1) / kLowRows);
    int grid = n_m_tiles < sm_count() ? n_m_tiles : sm_count();
    grid = (grid + CL - 1) / CL * CL;
    if (grid > sm_count()) grid = sm_count() / CL * CL;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(kLowThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (CL > 1) {
        int max_clusters = 0;
        VQB_CUDA_TRY(cudaOccupancyMaxActiveClusters(&max_clusters, search_tclow_kernel<CL>, &cfg));
        if (max_clusters > 0 && grid > max_clusters * CL) {
            grid = max_clusters * CL;
            cfg.gridDim = dim3((unsigned)grid);
        }
    }
    VQB_CUDA_TRY(cudaLaunchKernelEx(&cfg, search_tclow_kernel<CL>, p));
    return VQB_OK;
}

template <int D>
static int launch_tclow_d(const float* z, int64_t N, int64_t HW, const float* E, int K, const unsigned char* pk,
                          const PackLayout& L, unsigned char* wsb, const LowWorkspace& w, int64_t* idx_out,
                          float* dmin_out, cudaStream_t s, bool share_sm) {
    float* img = reinterpret_cast<float*>(wsb + w.off_img);
    float4* tok_norms = reinterpret_cast<float4*>(wsb + w.off_tau);
    int32_t* chunk = reinterpret_cast<int32_t*>(wsb + w.off_chunk);
    int32_t* list = reinterpret_cast<int32_t*>(wsb + w.off_list);
    int32_t* count = reinterpret_cast<int32_t*>(wsb + w.off_count);
    const unsigned n_m_tiles = (unsigned)((N + kLowRows - 1) / kLowRows);
    VQB_CUDA_TRY(cudaMemsetAsync(count, 0, sizeof(int32_t), s));
    split_tf32_tokens_kernel<D><<<n_m_tiles, 128, 0, s>>>(z, N, HW, reinterpret_cast<const int*>(pk), img, tok_norms);
    VQB_LAUNCH_CHECK("split_tf32_tokens_kernel");
    LowParams p;
    p.N = N;
    p.K = K;
    p.Kpad = L.Kpad;
    p.steps = tclow_steps(D);
    p.n_stages = 0;
    p.a_img = img;
    p.b_img = reinterpret_cast<const float*>(pk + L.off_img);
    p.header = reinterpret_cast<const int*>(pk);
    p.tok_norms = tok_norms;
    p.cmax = reinterpret_cast<const float*>(pk + L.off_cmax);
    p.chunk = chunk;
    p.list = list;
    p.list_count = count;
    p.share_sm = share_sm ? 1 : 0;
    int rc = VQB_OK;
    if (g_tclow_debug & 1) {
        VQB_CUDA_TRY(cudaMemsetAsync(chunk, 0, sizeof(int32_t) * (size_t)N, s));
    } else {
        switch (g_tclow_cluster) {
            case 1: rc = launch_tclow_cl<1>(p, s); break;
            case 4: rc = launch_tclow_cl<4>(p, s); break;
            default: rc = launch_tclow_cl<2>(p, s); break;
        }
    }
    if (rc != VQB_OK) return rc;
    if (g_tclow_debug & 2) return VQB_OK;
    const unsigned warps = (unsigned)((N + 31) / 32);
    rescore_chunk_kernel<D><<<(warps + 7) / 8, 256, 0, s>>>(z, E, reinterpret_cast<const float*>(pk + L.off_half_norm),
                                                            chunk, N, HW, K, idx_out, dmin_out);
    VQB_LAUNCH_CHECK("rescore_chunk_kernel");
    return VQB_OK;
}

int launch_search_tclow(const float* z, int64_t B, int D, int64_t HW, const float* E, int K, const void* pack,
                        int64_t* idx_out, float* dmin_out, void* ws, size_t ws_bytes, int64_t* stats_out,
                        cudaStream_t s, bool share_sm) {
    const int64_t N = B * HW;
    const LowWorkspace w = low_workspace(N, D);
    if (!ws || ws_bytes < w.total) {
        set_error("low-D tensor search workspace too small: %zu < %zu", ws_bytes, w.total);
        return VQB_ERR_WORKSPACE;
    }
    if ((reinterpret_cast<uintptr_t>(ws) & 255u) != 0) {
        set_error("low-D tensor search workspace must be 256-byte aligned");
        return VQB_ERR_INVALID_ARG;
    }
    const PackLayout L = pack_layout(K, D);
    unsigned char* wsb = static_cast<unsigned char*>(ws);
    const unsigned char* pk = static_cast<const unsigned char*>(pack);
    int rc;
    switch (D) {
#define VQB_CASE(d)                                                                            \
    case d:                                                                                    \
        rc = launch_tclow_d<d>(z, N, HW, E, K, pk, L, wsb, w, idx_out, dmin_out, s, share_sm); \
        break;
        VQB_CASE(1) VQB_CASE(2) VQB_CASE(3) VQB_CASE(4) VQB_CASE(5) VQB_CASE(6) VQB_CASE(7) VQB_CASE(8)
        VQB_CASE(9) VQB_CASE(10) VQB_CASE(11) VQB_CASE(12) VQB_CASE(13) VQB_CASE(14) VQB_CASE(15)
        VQB_CASE(16)
#undef VQB_CASE
        default:
            set_error("low-D tensor search supports 1 <= D <= 16, got %d", D);
            return VQB_ERR_UNSUPPORTED;
    }
    if (rc != VQB_OK) return rc;
    // unsure tokens: exact search by the CUDA-core kernel itself (bit-identical to VQB_ALGO_LOWD_FMA)
    const int32_t* list = reinterpret_cast<const int32_t*>(wsb + w.off_list);
    const int32_t* count = reinterpret_cast<const int32_t*>(wsb + w.off_count);
    if (!(g_tclow_debug & 4)) {
        rc = launch_search_lowd_list(z, B, D, HW, K, pack, list, count, idx_out, dmin_out, s);
        if (rc != VQB_OK) return rc;
    }
    if (stats_out) {
        tclow_stats_kernel<<<1, 1, 0, s>>>(stats_out, count);
        VQB_LAUNCH_CHECK("tclow_stats_kernel");
    }
    return VQB_OK;
}

// ---------------------------------------------------------------------------
// two engines in one CTA: host side
// ---------------------------------------------------------------------------
// Share of the images handed to the tensor role: both roles should finish together.  Single-engine times per 1M tokens x
// 16384 codes measured on B200 (profiles/r02_dim_sweep.txt): CUDA cores 0.735 ms x D; tensor cores 1.95 ms + 0.44 ms per
// k-step of 8 tf32 slots ((3D+3)/8 of them).  D = 4: 0.51, measured optimum 0.50 (profiles/r02_dual_search_split_sweep.txt).
VQB_KNOB g_dual_permille = 0;  // 0 = the model above; vqb_tune "dual_permille" overrides it
#ifdef VQB_EXPERIMENTAL
void set_dual_permille(int v) { g_dual_permille = v; }
#endif

static int64_t dual_tensor_images(int64_t B, int D) {
    double share = g_dual_permille / 1000.0;
    if (g_dual_permille <= 0) {
        // (the FMA role has 8 warps here instead of the stand-alone kernel's 16: measured optimum 0.50 at D = 4 and 0.65
        // at D = 8, profiles/r02_dual_dims.txt -> a correction linear in D)
        const double t_fma = 0.735 * D * (0.96 + 0.0525 * (D - 4)), t_tc = 1.95 + 0.44 * tclow_steps(D);
        share = t_fma / (t_fma + t_tc);
    }
    int64_t bt = (int64_t)((double)B * share + 0.5);
    return bt < 1 ? 1 : (bt > B - 1 ? B - 1 : bt);
}

size_t search_dual_workspace_bytes(int64_t B, int D, int64_t HW, int K) {
    if (!dual_eligible(B, D)) return 0;
    return search_tclow_workspace_bytes(dual_tensor_images(B, D) * HW, D, K);
}

__global__ void dual_stats_kernel(int64_t* stats, const int32_t* list_count, int64_t tensor_tokens) {
    stats[0] = *list_count;
    stats[1] = VQB_ALGO_DUAL_LOWD;
    stats[2] = 0;
    stats[3] = tensor_tokens;
}

template <int DD>
static int launch_search_dual_d(const float* z, int64_t B, int64_t HW, const float* E, int K, const void* pack,
                                int64_t* idx_out, float* dmin_out, void* ws, size_t ws_bytes, int64_t* stats_out,
                                cudaStream_t s) {
    constexpr int D = DD;
    constexpr int CL = 2;
    const int64_t Bt = dual_tensor_images(B, D);
    const int64_t Nt = Bt * HW, Nf = (B - Bt) * HW;
    const LowWorkspace w = low_workspace(Nt, D);
    if (!ws || ws_bytes < w.total || (reinterpret_cast<uintptr_t>(ws) & 255u) != 0) {
        set_error("two-engine search workspace too small or misaligned: %zu < %zu", ws_bytes, w.total);
        return VQB_ERR_WORKSPACE;
    }
    const PackLayout L = pack_layout(K, D);
    unsigned char* wsb = static_cast<unsigned char*>(ws);
    const unsigned char* pk = static_cast<const unsigned char*>(pack);
    float* img = reinterpret_cast<float*>(wsb + w.off_img);
    float4* tok_norms = reinterpret_cast<float4*>(wsb + w.off_tau);
    int32_t* chunk = reinterpret_cast<int32_t*>(wsb + w.off_chunk);
    int32_t* list = reinterpret_cast<int32_t*>(wsb + w.off_list);
    int32_t* count = reinterpret_cast<int32_t*>(wsb + w.off_count);
    const unsigned n_m_tiles = (unsigned)((Nt + kLowRows - 1) / kLowRows);
    VQB_CUDA_TRY(cudaMemsetAsync(count, 0, sizeof(int32_t), s));
    split_tf32_tokens_kernel<DD><<<n_m_tiles, 128, 0, s>>>(z, Nt, HW, reinterpret_cast<const int*>(pk), img, tok_norms);
    VQB_LAUNCH_CHECK("split_tf32_tokens_kernel");

    LowParams p;
    p.N = Nt;
    p.K = K;
    p.Kpad = L.Kpad;
    p.steps = tclow_steps(D);
    p.a_img = img;
    p.b_img = reinterpret_cast<const float*>(pk + L.off_img);
    p.header = reinterpret_cast<const int*>(pk);
    p.tok_norms = tok_norms;
    p.cmax = reinterpret_cast<const float*>(pk + L.off_cmax);
    p.chunk = chunk;
    p.list = list;
    p.list_count = count;
    p.share_sm = 1;
    const size_t tile_bytes = (size_t)p.steps * (kLowRows * 32);
    const size_t fma_smem = round_up_z(LowDCfg<DD, 0>::kSmemBytes, 1024);
    const size_t tc_fixed = 1024 + 512 + 2 * kLowRows * 4 * sizeof(float);
    // codebook stages of the tensor role: what the FMA role's tiles leave of the 227 KB (8 at D = 4, 2 at D = 16)
    int stages = (int)(((size_t)225 * 1024 - fma_smem - tc_fixed - 2 * tile_bytes) / tile_bytes);
    stages = stages > kLowMaxStages ? kLowMaxStages : stages;
    const int n_n_tiles = p.Kpad / kLowBN;
    if (stages > n_n_tiles) stages = n_n_tiles;
    if (stages < 2) stages = 2;
    p.n_stages = stages;
    const size_t tc_smem = tc_fixed + (2 + (size_t)stages) * tile_bytes;
    const size_t smem = fma_smem + tc_smem;

    DualFmaParams pf;
    pf.z = z + Nt * D;  // images [Bt, B): the latent is [B, D, HW]
    pf.N = Nf;
    pf.HW = HW;
    pf.K = K;
    pf.pack = pk;
    pf.L = L;
    pf.idx_out = idx_out + Nt;
    pf.dmin_out = dmin_out ? dmin_out + Nt : nullptr;
    pf.tensor_smem_offset = (uint32_t)fma_smem;

    auto kernel = search_dual_kernel<DD, CL>;
    VQB_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int grid = sm_count() / CL * CL;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(kDualThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int max_clusters = 0;
    VQB_CUDA_TRY(cudaOccupancyMaxActiveClusters(&max_clusters, kernel, &cfg));
    if (max_clusters > 0 && grid > max_clusters * CL) {
        grid = max_clusters * CL;
        cfg.gridDim = dim3((unsigned)grid);
    }
    int64_t per_cta = (Nf + grid - 1) / grid;
    per_cta = (per_cta + kDualFmaThreads - 1) / kDualFmaThreads * kDualFmaThreads;
    pf.tokens_per_cta = per_cta;
    VQB_CUDA_TRY(cudaLaunchKernelEx(&cfg, kernel, p, pf));

    const unsigned warps = (unsigned)((Nt + 31) / 32);
    rescore_chunk_kernel<DD><<<(warps + 7) / 8, 256, 0, s>>>(z, E, reinterpret_cast<const float*>(pk + L.off_half_norm),
                                                             chunk, Nt, HW, K, idx_out, dmin_out);
    VQB_LAUNCH_CHECK("rescore_chunk_kernel");
    // unsure tokens of the tensor role: exact search by the CUDA-core kernel (token-list mode)
    if (int rc = launch_search_lowd_list(z, Bt, D, HW, K, pack, list, count, idx_out, dmin_out, s)) return rc;
    if (stats_out) {
        dual_stats_kernel<<<1, 1, 0, s>>>(stats_out, count, Nt);
        VQB_LAUNCH_CHECK("dual_stats_kernel");
    }
    return VQB_OK;
}

int launch_search_dual(const float* z, int64_t B, int D, int64_t HW, const float* E, int K, const void* pack,
                       int64_t* idx_out, float* dmin_out, void* ws, size_t ws_bytes, int64_t* stats_out, cudaStream_t s) {
    switch (D) {
#define VQB_CASE(d) \
    case d:         \
        return launch_search_dual_d<d>(z, B, HW, E, K, pack, idx_out, dmin_out, ws, ws_bytes, stats_out, s);
        VQB_CASE(3) VQB_CASE(4) VQB_CASE(5) VQB_CASE(6) VQB_CASE(7) VQB_CASE(8) VQB_CASE(9) VQB_CASE(10)
        VQB_CASE(11) VQB_CASE(12) VQB_CASE(13) VQB_CASE(14) VQB_CASE(15) VQB_CASE(16)
#undef VQB_CASE
        default:
            set_error("two-engine search supports %d <= D <= %d, got %d", kDualMinD, kLowDMax, D);
            return VQB_ERR_UNSUPPORTED;
    }
}

}  // namespace vqb
