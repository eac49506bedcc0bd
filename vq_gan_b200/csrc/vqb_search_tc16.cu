// Single-pass fp16 tensor-core search with exact candidate re-scoring, D in {64,128,192,256}.
// Replaces quantizer.py:68-76 of the reference; same contract as the fp32 kernel, one third of
// the tensor work of the bf16x3 kernel (vqb_search_tc.cu).
//
//   approx[i,k] = 0.5|e_k|^2 - (zs_i . es_k) / (s_i * se),  zs = fp16(z_i * s_i), es = fp16(e_k * se)
//   s_i, se powers of two (exact scaling).  With dz_i = z_i - zs_i/s_i, de_k = e_k - es_k/se:
//   |approx - exact| <= eps_i = |dz_i| max|e| + (|z_i| + |dz_i|) max|de|  (Cauchy-Schwarz on the
//   ACTUAL rounding residuals, measured by the pre-passes) + fp32 accumulation and key-mask terms.
//
// The epilogue works on groups of 4 consecutive codes: two FMNMX give the group minimum, the
// group id rides in the low 8 mantissa bits, and a 7-FMNMX sorted insert keeps the four smallest
// group minima per token -- 2.5 ALU ops per score, so the epilogue hides behind the MMAs.  Then,
// with tau = 2 eps (any code whose exact score beats the approximate winner lies within tau):
//   m2 - m1 > tau   -> every candidate is in group g1: its 4 codes are re-scored in fp32
//   m3 - m1 > tau   -> candidates are in g1, g2: 8 codes;   m4 - m1 > tau -> g1..g3: 12 codes
//   otherwise (or NaN) -> full fp32 re-score of the token (vqb_search_fp32.cu)
//
// Pipeline: persistent CTA per SM, 128-token tiles; warp 0 TMA producer (A = fp16 token tile,
// resident; B = 256-code x 64-dim blocks through a ring of 32 KB stages), warp 1 MMA issuer
// (tcgen05.mma kind::f16, 128x256x16, two 256-column TMEM accumulators), warp 2 TMEM allocator,
// warps 4-11 epilogue (two warpgroups, 128 columns each).
#include "vqb_tc_common.cuh"

namespace vqb {

constexpr int kT16Threads = 384;
constexpr int kT16EpiThreads = 256;

__host__ __device__ constexpr int tc16_stages(int nkb) {
    int s = (kTcSmemBudget - 1024 - kTcBarrierBytes - 12288 - nkb * kTcABlockBytes) / kTcBStageBytes;
    return s > 6 ? 6 : s;
}

constexpr uint32_t kT16Idesc = (1u << 4)  // accumulator f32; A, B = f16 (format 0), K-major
                               | ((uint32_t)(kTcBN >> 3) << 17) | ((uint32_t)(kTcBM >> 4) << 24);

// ---------------------------------------------------------------------------
// pre-pass: z[B, D, HW] fp32 -> z16[N, D] fp16 scaled per token, 1/(s_i*se), |z_i|, |dz_i|
// (one global read: the 32-token x D tile waits in shared memory while the scale is found)
// ---------------------------------------------------------------------------
// NCH = padded D / 32 (2, 4, 6, 8): every loop below has a compile-time trip count, so the addresses are
// immediates (the run-time-D version spent 3 of 4 issue slots on address arithmetic)
template <int NCH, bool kExact>
__global__ void __launch_bounds__(256)
    split16_tokens_kernel(const float* __restrict__ z, int64_t N, int Dreal, int64_t HW, const int* __restrict__ header,
                          __half* __restrict__ z16, float* __restrict__ inv_scale, float* __restrict__ znorm,
                          float* __restrict__ zres) {
    constexpr int D = 32 * NCH;  // padded row length (multiple of 64); channels >= Dreal are zeros
    extern __shared__ float split_smem[];  // [D][33]: 33.8 KB at D = 256, 67.6 KB at D = 512
    float (*tile)[33] = reinterpret_cast<float (*)[33]>(split_smem);
    __shared__ float part_a[8][32];
    __shared__ float part_b[8][32];
    __shared__ int tok_exp[32];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int64_t t0 = (int64_t)blockIdx.x * 32;
    const int64_t tok = t0 + tx;
    const bool ok = tok < N;
    // pass 1: stage the tile, per-token max |z| and sum of squares
    float sq = 0.f, mx = 0.f;
    bool finite = true;
    {
        const int64_t b = ok ? tok / HW : 0;
        const float* zp = z + (b * Dreal) * HW + (ok ? tok - b * HW : 0) + (int64_t)ty * HW;
        const int64_t step = 8 * HW;
#pragma unroll
        for (int i = 0; i < D / 8; ++i) {
            const float v = (ok && (kExact || ty + 8 * i < Dreal)) ? __ldg(zp) : 0.f;
            zp += step;
            tile[ty + 8 * i][tx] = v;
            sq = fmaf(v, v, sq);
            const float a = fabsf(v);
            finite &= (a < INFINITY);  // false for NaN as well
            mx = fmaxf(mx, a);
        }
    }
    part_a[ty][tx] = sq;
    part_b[ty][tx] = finite ? mx : INFINITY;
    __syncthreads();
    if (ty == 0) {
        float s = 0.f, m = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            s += part_a[i][tx];
            m = fmaxf(m, part_b[i][tx]);
        }
        int e = 0;
        if (m > 0.f && m < INFINITY) {
            e = 9 - ilogbf(m);  // max |z_i| * 2^e in [512, 1024)
            e = e < -100 ? -100 : (e > 100 ? 100 : e);
        }
        tok_exp[tx] = e;
        if (ok) {
            znorm[tok] = sqrtf(s);
            // a token with a NaN / inf coordinate gets a NaN scale: every approximate score becomes
            // NaN and the token is re-scored by the fp32 kernel, which implements ATen's NaN rules
            inv_scale[tok] = (m < INFINITY) ? ldexpf(1.f, -(e + header[5])) : __int_as_float(0x7fc00000);
        }
    }
    __syncthreads();
    // pass 2: scale, convert, write token-major rows; accumulate the rounding residual.  Lane L reads
    // channels L and L+32 of a 64-channel block (conflict-free), one shuffle pairs neighbours up so
    // that even lanes store channels (L, L+1) and odd lanes (L+31, L+32) as one half2 each.
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = ty + 8 * i;  // token within the block
        const int64_t t = t0 + r;
        const int ex = tok_exp[r];
        // |ex| <= 100: both factors are normal floats, so the products round exactly like ldexpf
        const float fwd = __int_as_float((127 + ex) << 23), back = __int_as_float((127 - ex) << 23);
        float res = 0.f;
#pragma unroll
        for (int j = 0; j < NCH / 2; ++j) {
            const float a = tile[64 * j + tx][r], b2 = tile[64 * j + 32 + tx][r];
            const __half ha = __float2half_rn(a * fwd), hb = __float2half_rn(b2 * fwd);
            const float da = a - __half2float(ha) * back, db = b2 - __half2float(hb) * back;
            res = fmaf(da, da, res);
            res = fmaf(db, db, res);
            const bool odd = tx & 1;
            const unsigned short mine = __half_as_ushort(odd ? ha : hb);  // what the partner needs
            const unsigned short got = (unsigned short)__shfl_xor_sync(0xffffffffu, (int)mine, 1);
            const __half2 h = odd ? __halves2half2(__ushort_as_half(got), hb) : __halves2half2(ha, __ushort_as_half(got));
            const int ch = 64 * j + (odd ? 32 + tx - 1 : tx);
            if (t < N) *reinterpret_cast<__half2*>(z16 + (size_t)t * D + ch) = h;
        }
        // the 32 lanes of this warp hold the residual of token r: reduce and store
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) res += __shfl_xor_sync(0xffffffffu, res, o);
        if (tx == 0 && t < N) zres[t] = sqrtf(res);
    }
}

// ---------------------------------------------------------------------------
// main kernel
// ---------------------------------------------------------------------------
struct T16Params {
    int64_t N;
    int K, Kpad;
    const float* half_norm;      // exact, +inf padded (re-score kernels)
    const float* half_norm_fin;  // finite padding (epilogue keys must never be NaN)
    const int* header;
    const float* gmax;       // [Kpad/4] max code norm per 4-code group
    const float* znorm;
    const float* zres;       // |z_i - fp16 image of z_i|
    const float* inv_scale;
    int32_t* group1;         // [N] best group (code / 4)
    int32_t* group2;         // [N] second group, or -1 when group1 alone holds every candidate
    int32_t* group3;         // [N] third group, or -1
    int32_t* pair_count;     // statistics only
    int32_t* full_list;      // tokens needing a full fp32 re-score
    int32_t* full_count;
    float acc_eps;           // relative fp32-accumulation allowance: 2^-15 up to D = 256, 2^-14 above (sums twice as long)
};

// four smallest group minima of a token and the groups of the best three
struct Top4 {
    float v1, v2, v3, v4;
    int i1, i2, i3;
};

// (value, index) lexicographic insert: equal keys from different tiles keep the lower index first
__device__ __forceinline__ void top4_insert(Top4& g, float v, int i) {
    if (v < g.v1 || (v == g.v1 && i < g.i1)) {
        g.v4 = g.v3;
        g.v3 = g.v2;
        g.i3 = g.i2;
        g.v2 = g.v1;
        g.i2 = g.i1;
        g.v1 = v;
        g.i1 = i;
    } else if (v < g.v2 || (v == g.v2 && i < g.i2)) {
        g.v4 = g.v3;
        g.v3 = g.v2;
        g.i3 = g.i2;
        g.v2 = v;
        g.i2 = i;
    } else if (v < g.v3 || (v == g.v3 && i < g.i3)) {
        g.v4 = g.v3;
        g.v3 = v;
        g.i3 = i;
    } else if (v < g.v4) {
        g.v4 = v;
    }
}

// CL = thread-block cluster size (1, 2 or 4).  With CL > 1 the CTAs of a cluster work on different
// token tiles but sweep the codebook in lockstep: each CTA fetches 1/CL of every B stage and TMA
// multicasts it to all of them, so the L2 -> SM operand traffic (the limiter of the single-pass
// kernel: 32 KB of B per 512 tensor cycles per SM) drops by CL.
// BR: the sorted insert runs only when the group key beats the current fourth-best (a warp-divergent branch taken by
// ~14 % of the warp-steps: the probability that group n of a token enters its top-4 is 4/n) instead of unconditionally
// GS: codes per candidate group (4 or 8).  The epilogue costs 4 FFMA + (GS/2 + 8) ALU-pipe operations per group, i.e.
// 2.5 ALU operations per score with groups of 4 and 1.5 with groups of 8; the price of wider groups is a re-score
// over 8 / 16 / 24 exact candidates instead of 4 / 8 / 12.  Measured (scripts/tc16_group_ab.py, 1 M tokens x 16 384
// codes, whole search): D = 32 4.19 -> 3.56 ms, D = 64 4.23 -> 3.66 ms, D = 128 4.8 -> 4.6-5.1 ms, D = 256 7.2-7.7 -> 7.9-8.1 ms:
// 40 % fewer ALU operations buy 16 % -- below D = 128 the kernel is bound by the TMEM read-back of the scores (128 KB
// per tile at ~80-96 B/clk/SM, about 22 scores per clock per SM whatever D is), not by the ALU pipe.  GS = 8 up to
// padded D = 64 (tc16_group_size), 4 above.
template <int NKB, int CL, bool BR = false, int GS = 4>
__global__ void __launch_bounds__(kT16Threads, 1)
    search_tc16_kernel(const __grid_constant__ CUtensorMap map_z, const __grid_constant__ CUtensorMap map_e,
                       T16Params p) {
    constexpr int kStages = tc16_stages(NKB);
    static_assert(kStages >= 2, "not enough shared memory for the B ring");
    extern __shared__ unsigned char smem_unaligned[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_unaligned) + 1023) &
                                                           ~(uintptr_t)1023);
    unsigned char* a_tile = smem;
    unsigned char* b_ring = smem + NKB * kTcABlockBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(b_ring + kStages * kTcBStageBytes);
    uint64_t* a_full = bars + 0;
    uint64_t* a_empty = bars + 1;
    uint64_t* tm_full = bars + 2;   // [2]
    uint64_t* tm_empty = bars + 4;  // [2]
    uint64_t* b_full = bars + 6;    // [kStages]
    uint64_t* b_empty = bars + 6 + kStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6 + 2 * kStages);
    // hand-over of the upper-column warpgroup's top-4 to the lower one: 128 rows x 8 words (4 KB),
    // then two 1024-float buffers of half norms for the current / next super-tile of 4 code tiles (8 KB)
    float* xchg = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(bars) + kTcBarrierBytes);
    float* hbuf = xchg + 128 * 8;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_m_tiles = (int)((p.N + kTcBM - 1) / kTcBM);
    const int n_n_tiles = p.Kpad / kTcBN;
    // every CTA of a cluster runs the same number of rounds (tiles past the end are all padding:
    // TMA zero-fills them and the epilogue drops their rows)
    const int n_rounds = (n_m_tiles + (int)gridDim.x - 1) / (int)gridDim.x;
    const uint32_t cta_rank = CL > 1 ? cluster_ctarank() : 0;
    constexpr uint16_t kMask = (uint16_t)((1u << CL) - 1);

    if (threadIdx.x == 0) {
        tc_mbar_init(a_full, 1);
        tc_mbar_init(a_empty, 1);
        for (int i = 0; i < 2; ++i) {
            tc_mbar_init(tm_full + i, 1);
            tc_mbar_init(tm_empty + i, kT16EpiThreads);
        }
        for (int i = 0; i < kStages; ++i) {
            tc_mbar_init(b_full + i, 1);
            tc_mbar_init(b_empty + i, CL);  // released by the MMA issuer of every CTA in the cluster
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
    if constexpr (CL > 1) cluster_sync_all();  // peers' barriers exist before anything remote touches them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            uint32_t stage = 0, bphase = 0, aphase = 0;
            for (int round = 0; round < n_rounds; ++round) {
                const int mt = blockIdx.x + round * gridDim.x;
                tc_mbar_wait(a_empty, aphase ^ 1);
                tc_mbar_expect_tx(a_full, NKB * kTcABlockBytes);
#pragma unroll
                for (int kb = 0; kb < NKB; ++kb)
                    tma_load_2d(a_tile + kb * kTcABlockBytes, &map_z, a_full, kb * kTcBK, mt * kTcBM);
                aphase ^= 1;
                for (int nt = 0; nt < n_n_tiles; ++nt) {
                    for (int kb = 0; kb < NKB; ++kb) {
                        tc_mbar_wait(b_empty + stage, bphase ^ 1);
                        tc_mbar_expect_tx(b_full + stage, kTcBStageBytes);
                        if constexpr (CL > 1) {
                            constexpr int kRows = kTcBN / CL;  // this CTA's slice of the stage
                            tma_load_2d_mc(b_ring + stage * kTcBStageBytes + cta_rank * (kTcBStageBytes / CL), &map_e,
                                           b_full + stage, kb * kTcBK, nt * kTcBN + (int)cta_rank * kRows, kMask);
                        } else {
                            tma_load_2d(b_ring + stage * kTcBStageBytes, &map_e, b_full + stage, kb * kTcBK,
                                        nt * kTcBN);
                        }
                        if (++stage == kStages) {
                            stage = 0;
                            bphase ^= 1;
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            uint32_t stage = 0, bphase = 0, aphase = 0, acc = 0, accphase = 0;
            for (int round = 0; round < n_rounds; ++round) {
                tc_mbar_wait(a_full, aphase);
                aphase ^= 1;
                tc_fence_after();
                for (int nt = 0; nt < n_n_tiles; ++nt) {
                    tc_mbar_wait(tm_empty + acc, accphase ^ 1);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + acc * kTcBN;
                    for (int kb = 0; kb < NKB; ++kb) {
                        const uint32_t as = s32(a_tile + kb * kTcABlockBytes);
                        tc_mbar_wait(b_full + stage, bphase);
                        tc_fence_after();
                        const uint32_t bs = s32(b_ring + stage * kTcBStageBytes);
#pragma unroll
                        for (int k4 = 0; k4 < kTcBK / 16; ++k4)
                            umma_bf16(d_tmem, umma_desc_sw128(as + k4 * 32), umma_desc_sw128(bs + k4 * 32), kT16Idesc,
                                      (kb | k4) != 0);
                        if constexpr (CL > 1)
                            umma_commit_mc(b_empty + stage, kMask);
                        else
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
                umma_commit(a_empty);
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue (8 warps) =====================
        const int q = warp & 3;           // TMEM lane quarter this warp may read
        const int hsel = (warp - 4) >> 2;  // column half: 0 -> [0,128), 1 -> [128,256)
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        const int first_nan = p.header[0];
        const float h_max = __int_as_float(p.header[4]);
        const float e_max = sqrtf(2.f * h_max);
        // |e_k - fp16 image of e_k| <= rho16 |e_k| + a16 for every code (measured by the prepare kernel)
        const float rho16 = __int_as_float(p.header[10]), a16 = __int_as_float(p.header[11]);
        const int row_in_tile = q * 32 + lane;
        const int epi_tid = threadIdx.x - 128;  // 0..255
        const int n_super = (n_n_tiles + 3) / 4;
        const uint32_t hbuf_addr = s32(hbuf);
        uint32_t hsel_buf = 0;  // which half-norm buffer holds the current super-tile
#pragma unroll
        for (int j = 0; j < 4; ++j)  // super-tile 0
            hbuf[j * kTcBN + epi_tid] = (j < n_n_tiles) ? __ldg(p.half_norm_fin + j * kTcBN + epi_tid) : 1e38f;
        asm volatile("bar.sync 1, %0;" ::"n"(kT16EpiThreads) : "memory");
        uint32_t acc = 0, accphase = 0;
        for (int round = 0; round < n_rounds; ++round) {
            const int mt = blockIdx.x + round * gridDim.x;
            const int64_t row = (int64_t)mt * kTcBM + row_in_tile;
            const float neg_inv = (row < p.N) ? -p.inv_scale[row] : 0.f;
            Top4 g;
            g.v1 = g.v2 = g.v3 = g.v4 = INFINITY;
            g.i1 = g.i2 = g.i3 = 0;
            // code tiles are processed in super-tiles of 4 (1024 codes = 256 groups): the group id
            // within the super-tile rides in the low 8 mantissa bits, and the running top-4 is merged
            // (and the half-norm buffer rotated, one barrier) once per super-tile
            for (int st = 0; st < n_super; ++st) {
                const int tiles_here = (n_n_tiles - st * 4) < 4 ? (n_n_tiles - st * 4) : 4;
                // prefetch this thread's share of the NEXT super-tile's half norms
                const int next_st = (st + 1 < n_super) ? st + 1 : 0;
                float h_next[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int t = next_st * 4 + j;
                    h_next[j] = (t < n_n_tiles) ? __ldg(p.half_norm_fin + t * kTcBN + epi_tid) : 1e38f;
                }
                float k1 = INFINITY, k2 = INFINITY, k3 = INFINITY, k4 = INFINITY;
                const uint32_t hs_base = hbuf_addr + (hsel_buf & 1) * (4 * kTcBN * 4) + hsel * 128 * 4;
#pragma unroll 1
                for (int tj = 0; tj < tiles_here; ++tj) {
                    tc_mbar_wait(tm_full + acc, accphase);
                    tc_fence_after();
                    const uint32_t t_acc = tmem_base + lane_addr + acc * kTcBN + hsel * 128;
                    const uint32_t hs = hs_base + tj * (kTcBN * 4);
                    const uint32_t id_base = (uint32_t)(tj * (kTcBN / GS) + hsel * (128 / GS));
                    uint32_t ra[32], rb[32];
                    tmem_ld32_issue(t_acc, ra);
#pragma unroll
                    for (int c0 = 0; c0 < 128; c0 += 32) {
                        uint32_t(&r)[32] = ((c0 >> 5) & 1) ? rb : ra;
                        uint32_t(&rn)[32] = ((c0 >> 5) & 1) ? ra : rb;
                        tmem_ld_wait();
                        if (c0 + 32 < 128) tmem_ld32_issue(t_acc + c0 + 32, rn);
#pragma unroll
                        for (int cg = 0; cg < 32 / GS; ++cg) {
                            float sc[GS];
#pragma unroll
                            for (int v = 0; v < GS / 4; ++v) {
                                float4 h4;
                                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                                             : "=f"(h4.x), "=f"(h4.y), "=f"(h4.z), "=f"(h4.w)
                                             : "r"(hs + (c0 + cg * GS + v * 4) * 4));
                                sc[4 * v + 0] = fmaf(__uint_as_float(r[cg * GS + v * 4 + 0]), neg_inv, h4.x);
                                sc[4 * v + 1] = fmaf(__uint_as_float(r[cg * GS + v * 4 + 1]), neg_inv, h4.y);
                                sc[4 * v + 2] = fmaf(__uint_as_float(r[cg * GS + v * 4 + 2]), neg_inv, h4.z);
                                sc[4 * v + 3] = fmaf(__uint_as_float(r[cg * GS + v * 4 + 3]), neg_inv, h4.w);
                            }
                            float gm;
                            if constexpr (GS == 4)
                                gm = fminf(min3_f32(sc[0], sc[1], sc[2]), sc[3]);
                            else
                                gm = fminf(min3_f32(min3_f32(sc[0], sc[1], sc[2]), min3_f32(sc[3], sc[4], sc[5]), sc[6]),
                                           sc[7]);
                            const uint32_t gid = id_base + (uint32_t)(c0 / GS + cg);
                            const float key = __uint_as_float((__float_as_uint(gm) & 0xffffff00u) | gid);
                            // sorted insert of key into (k1 <= k2 <= k3 <= k4)
                            if (!BR || key < k4) {
                                const float n2 = fminf(k2, fmaxf(k1, key));
                                const float n3 = fminf(k3, fmaxf(k2, key));
                                const float n4 = fminf(k4, fmaxf(k3, key));
                                k1 = fminf(k1, key);
                                k2 = n2;
                                k3 = n3;
                                k4 = n4;
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
                // publish the next super-tile's half norms; the barrier also fences the buffer reuse
                {
                    float* nb = hbuf + ((hsel_buf + 1) & 1) * (4 * kTcBN);
#pragma unroll
                    for (int j = 0; j < 4; ++j) nb[j * kTcBN + epi_tid] = h_next[j];
                }
                ++hsel_buf;
                asm volatile("bar.sync 1, %0;" ::"n"(kT16EpiThreads) : "memory");
                // merge the super-tile's top-4 into the running top-4 (earlier codes win ties)
                if (k1 < g.v4) {
                    const int base = st * (4 * kTcBN / GS);
                    top4_insert(g, k1, base + (int)(__float_as_uint(k1) & 0xffu));
                    top4_insert(g, k2, base + (int)(__float_as_uint(k2) & 0xffu));
                    top4_insert(g, k3, base + (int)(__float_as_uint(k3) & 0xffu));
                    top4_insert(g, k4, base + (int)(__float_as_uint(k4) & 0xffu));
                }
            }
            // ---- combine the two column halves of each row (named barrier over the epilogue warps)
            if (hsel == 1) {
                float* x = xchg + row_in_tile * 8;
                x[0] = g.v1;
                x[1] = g.v2;
                x[2] = g.v3;
                x[3] = g.v4;
                x[4] = __int_as_float(g.i1);
                x[5] = __int_as_float(g.i2);
                x[6] = __int_as_float(g.i3);
            }
            asm volatile("bar.sync 1, %0;" ::"n"(kT16EpiThreads) : "memory");
            if (hsel == 0 && row < p.N) {
                const float* x = xchg + row_in_tile * 8;
                Top4 m = g;
                top4_insert(m, x[0], __float_as_int(x[4]));
                top4_insert(m, x[1], __float_as_int(x[5]));
                top4_insert(m, x[2], __float_as_int(x[6]));
                if (x[3] < m.v4) m.v4 = x[3];
                // Error of the approximate score of a code of norm n (Cauchy-Schwarz on the MEASURED fp16
                // residuals; fp32 accumulation 2^-15 relative, covering this kernel's and the re-score's sums;
                // the 8 masked mantissa bits are 2^-15 of the score, |score| <= n^2/2 + |z| n):
                //   eps(n) = (|dz| + (|z| + |dz|) rho16 + 2^-15 |z|) n + (|z| + |dz|) a16 + 2^-14 (n^2/2 + |z| n)
                // The approximate winner k1 sits in group i1 (norm <= n1): exact_best <= U = m1 + eps(n1).
                // A code of norm n scores at least n^2/2 - |z| n exactly, so codes with n > n0 =
                // |z| + sqrt(|z|^2 + 2U) cannot win whatever their approximate score; every other code has
                // error <= eps(n0).  Hence a code outside the kept groups is excluded once its group minimum
                // exceeds m1 by tau = eps(min(n0, max|e|)) + eps(n1) -- the threshold follows the norm of
                // the WINNER, not the largest norm in the codebook (one outlier code no longer inflates it).
                const float zn = p.znorm[row], zr = p.zres[row];
                const float alpha = zr + (zn + zr) * rho16 + p.acc_eps * zn;
                const float beta0 = (zn + zr) * a16;
                float n1 = __ldg(p.gmax + m.i1 * (GS / 4));  // gmax: per 4 codes
                if constexpr (GS == 8) n1 = fmaxf(n1, __ldg(p.gmax + m.i1 * 2 + 1));
                n1 = fminf(n1, e_max);
                const float eps1 = alpha * n1 + beta0 + 2.f * p.acc_eps * (0.5f * n1 * n1 + zn * n1);
                const float U = m.v1 + eps1;
                const float n0 = fminf(zn + sqrtf(fmaxf(zn * zn + 2.f * U, 0.f)), e_max);
                const float eps0 = alpha * n0 + beta0 + 2.f * p.acc_eps * (0.5f * n0 * n0 + zn * n0);
                const float tau = 1.0001f * (eps0 + eps1);
                const bool only1 = (m.v2 - m.v1) > tau;  // every candidate is in group 1
                const bool only2 = (m.v3 - m.v1) > tau;  // ... in groups 1-2
                const bool only3 = (m.v4 - m.v1) > tau;  // ... in groups 1-3
                p.group1[row] = m.i1;
                p.group2[row] = only1 ? -1 : m.i2;
                p.group3[row] = only2 ? -1 : m.i3;
                if (first_nan < p.K || !only3) {
                    const int slot = atomicAdd(p.full_count, 1);
                    p.full_list[slot] = (int32_t)row;
                } else if (!only1) {
                    atomicAdd(p.pair_count, 1);
                }
            }
            // the hand-over buffer is reused by the next token tile
            asm volatile("bar.sync 1, %0;" ::"n"(kT16EpiThreads) : "memory");
        }
    }

    tc_fence_before();
    __syncthreads();
    if constexpr (CL > 1) cluster_sync_all();  // no CTA leaves while a peer may still multicast into it
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

// Exact fp32 re-score of the candidate groups of every token: the 4 codes of group1 and, when
// present, those of group2 and group3.  A CTA stages 32 tokens x D of z in shared memory (coalesced
// along the tokens), then each warp takes 4 tokens AT ONCE: lanes run along the channels, the 16
// (token, code) partial sums are reduced with a transposed butterfly (16 shuffles instead of 80) and
// the 16 codebook-row streams (full 128-byte lines from L2) are all in flight together.  Tokens with
// a second / third candidate group (a few %) take a warp-uniform slow path.  Lowest index wins ties.
__device__ __forceinline__ void lex_min(float& s, int& k, float s2, int k2) {
    if (s2 < s || (s2 == s && k2 < k)) {
        s = s2;
        k = k2;
    }
}

// kExact: D == 32 * NCH, no channel guards (the guards cost the batched immediate-offset loads: 0.52 -> 0.79 ms)
// GS: codes per candidate group (4 or 8).  A warp owns 4 tokens and keeps 16 (token, code) row streams in flight per
// step: 4 tokens x 4 codes in one step (GS = 4), or 2 tokens x 8 codes in each of two steps (GS = 8).
template <int NCH, bool kExact, int GS>
__global__ void __launch_bounds__(256)
    rescore_groups_kernel(const float* __restrict__ z, const float* __restrict__ E, const float* __restrict__ half_norm,
                          const int32_t* __restrict__ group1, const int32_t* __restrict__ group2,
                          const int32_t* __restrict__ group3, int64_t N, int Dreal, int64_t HW, int K,
                          int64_t* __restrict__ idx_out, float* __restrict__ dmin_out) {
    static_assert(GS == 4 || GS == 8, "group size");
    constexpr int TPS = 16 / GS;   // tokens per step
    constexpr int STEPS = 4 / TPS;
    constexpr int LPT = 2 * GS;    // lanes per token after the butterfly
    constexpr int DT = 32 * NCH;  // D rounded up to 32; tile rows >= D are zeros
    const int D = kExact ? DT : Dreal;  // a compile-time constant in the exact instantiations
    // row stride 36 floats: the 4 tokens of a warp are one aligned 16-byte read, stores stay conflict-free
    extern __shared__ __align__(16) float rescore_smem[];  // [DT][36]: 36.9 KB at D = 256, 73.7 KB at D = 512
    float (*tile)[36] = reinterpret_cast<float (*)[36]>(rescore_smem);
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int64_t t0 = (int64_t)blockIdx.x * 32;
    {
        const int64_t tok = t0 + tx;
        if (tok < N) {
            const int64_t b = tok / HW;
            const float* zp = z + (b * D) * HW + (tok - b * HW) + (int64_t)ty * HW;
            const int64_t step = 8 * HW;
#pragma unroll
            for (int i = 0; i < DT / 8; ++i) {
                tile[ty + 8 * i][tx] = (kExact || ty + 8 * i < D) ? __ldg(zp) : 0.f;
                zp += step;
            }
        }
    }
    const int lane = tx, r0 = ty * 4;
    int g1[4], g2[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const int64_t tok = t0 + r0 + t;
        const bool ok = tok < N;
        g1[t] = ok ? __ldg(group1 + tok) : 0;
        g2[t] = ok ? __ldg(group2 + tok) : -1;
    }
    __syncthreads();
#pragma unroll
    for (int st = 0; st < STEPS; ++st) {
        const float* row[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            int k = g1[st * TPS + j / GS] * GS + (j % GS);
            k = k < K ? k : K - 1;  // rows past the end are read in bounds and discarded below
            row[j] = E + (size_t)k * D + lane;
        }
        float acc[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j] = 0.f;
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            const float4 z4 = *reinterpret_cast<const float4*>(&tile[32 * c + lane][r0]);
            const float zv[4] = {z4.x, z4.y, z4.z, z4.w};
            if (kExact || 32 * c + lane < D) {  // only the last block of a D that is not a multiple of 32 is partial
#pragma unroll
                for (int j = 0; j < 16; ++j) acc[j] = fmaf(zv[st * TPS + j / GS], __ldg(row[j] + 32 * c), acc[j]);
            }
        }
        // transposed butterfly: afterwards lane L holds the full sum of accumulator (L >> 1) & 15
        {
            const bool up = lane & 16;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float keep = up ? acc[j + 8] : acc[j], send = up ? acc[j] : acc[j + 8];
                acc[j] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
            }
        }
        {
            const bool up = lane & 8;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float keep = up ? acc[j + 4] : acc[j], send = up ? acc[j] : acc[j + 4];
                acc[j] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
            }
        }
        {
            const bool up = lane & 4;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const float keep = up ? acc[j + 2] : acc[j], send = up ? acc[j] : acc[j + 2];
                acc[j] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
            }
        }
        {
            const bool up = lane & 2;
            const float keep = up ? acc[1] : acc[0], send = up ? acc[0] : acc[1];
            acc[0] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
        }
        const float dot = acc[0] + __shfl_xor_sync(0xffffffffu, acc[0], 1);
        const int my_t = lane / LPT, my_c = (lane >> 1) & (GS - 1);  // token / code within the group of this lane's sum
        int my_g1 = g1[st * TPS];
#pragma unroll
        for (int t = 1; t < TPS; ++t) my_g1 = my_t == t ? g1[st * TPS + t] : my_g1;
        int best_k = my_g1 * GS + my_c;
        float best = INFINITY;
        if (best_k < K) {
            best = __ldg(half_norm + best_k) - dot;
            if (best != best) {  // NaN scores never win
                best = INFINITY;
                best_k = 0x7fffffff;
            }
        } else {
            best_k = 0x7fffffff;
        }
        // (score, index) lexicographic minimum over the GS codes of the token (lane bits 1 .. log2(GS))
#pragma unroll
        for (int o = 2; o <= GS; o <<= 1) {
            const float s2 = __shfl_xor_sync(0xffffffffu, best, o);
            const int k2 = __shfl_xor_sync(0xffffffffu, best_k, o);
            lex_min(best, best_k, s2, k2);
        }
        // slow path: tokens with more candidate groups (warp-uniform branches)
#pragma unroll
        for (int t = 0; t < TPS; ++t) {
            if (g2[st * TPS + t] < 0) continue;
            const int64_t tok = t0 + r0 + st * TPS + t;
            float b = __shfl_sync(0xffffffffu, best, t * LPT);
            int bk = __shfl_sync(0xffffffffu, best_k, t * LPT);
#pragma unroll 1
            for (int pass = 0; pass < 2; ++pass) {
                const int grp = pass == 0 ? g2[st * TPS + t] : __ldg(group3 + tok);
                if (grp < 0) break;
                const int k0 = grp * GS;
                float a4[GS];
#pragma unroll
                for (int c = 0; c < GS; ++c) a4[c] = 0.f;
                for (int d = lane; d < D; d += 32) {
                    const float zv = tile[d][r0 + st * TPS + t];
#pragma unroll
                    for (int c = 0; c < GS; ++c) {
                        const int k = k0 + c;
                        a4[c] = fmaf(zv, (k < K) ? __ldg(E + (size_t)k * D + d) : 0.f, a4[c]);
                    }
                }
#pragma unroll
                for (int c = 0; c < GS; ++c) {
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) a4[c] += __shfl_xor_sync(0xffffffffu, a4[c], o);
                    const int k = k0 + c;
                    if (k < K) {
                        const float dsc = __ldg(half_norm + k) - a4[c];
                        if (dsc == dsc) lex_min(b, bk, dsc, k);
                    }
                }
            }
            if (lane / LPT == t) {
                best = b;
                best_k = bk;
            }
        }
        if ((lane & (LPT - 1)) == 0) {
            const int64_t tok = t0 + r0 + st * TPS + my_t;
            if (tok < N) {
                if (best_k == 0x7fffffff) best_k = my_g1 * GS < K ? my_g1 * GS : 0;
                idx_out[tok] = best_k;
                if (dmin_out) dmin_out[tok] = best;
            }
        }
    }
}

// stats[3]: how many of the stats[0] uncertified tokens the pruned exact tier took (the rest met the full re-search)
__global__ void tc16_stats_kernel(int64_t* stats, const int32_t* full_count, const int32_t* pair_count, const int32_t* handled) {
    stats[0] = *full_count;
    stats[1] = VQB_ALGO_TCGEN05_F16;
    stats[2] = *pair_count;
    stats[3] = handled ? *handled : 0;
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
struct T16Workspace {
    size_t off_z16, off_inv, off_znorm, off_zres, off_g1, off_g2, off_g3, off_full, off_counts, off_keys, keys_bytes,
        off_pruned, pruned_bytes, total;
};

static T16Workspace t16_workspace(int64_t N, int D, int K) {
    T16Workspace w;
    size_t off = 0;
    w.off_z16 = off;
    off = round_up_z(off + 2 * (size_t)N * tc16_dpad(D), 1024);
    w.off_inv = off;
    off = round_up_z(off + 4 * (size_t)N, 1024);
    w.off_znorm = off;
    off = round_up_z(off + 4 * (size_t)N, 1024);
    w.off_zres = off;
    off = round_up_z(off + 4 * (size_t)N, 1024);
    w.off_g1 = off;
    off = round_up_z(off + 4 * (size_t)N, 1024);
    w.off_g2 = off;
    off = round_up_z(off + 4 * (size_t)N, 1024);
    w.off_g3 = off;
    off = round_up_z(off + 4 * (size_t)N, 1024);
    w.off_full = off;
    off = round_up_z(off + 4 * (size_t)N, 1024);
    w.off_counts = off;  // int32[2] (full_count, pair_count) | int32[kPrunedStateInts] at +64 bytes: ONE memset per call
    off += 2048;
    w.off_keys = off;
    w.keys_bytes = search_fp32_workspace_bytes(N, D);
    off = round_up_z(off + w.keys_bytes, 1024);
    w.off_pruned = off;  // pruned exact tier (vqb_search_pruned.cu); 0 bytes where it does not apply
    w.pruned_bytes = search_pruned_workspace_bytes(N, K, D);
    off = round_up_z(off + w.pruned_bytes, 1024);
    w.total = off;
    return w;
}

void tc16_split_pointers(void* ws, int64_t N, int D, __half** z16, float** inv_scale, float** znorm, float** zres) {
    const T16Workspace w = t16_workspace(N, D, 0);  // (the split lives in front of everything that depends on K)
    unsigned char* b = static_cast<unsigned char*>(ws);
    *z16 = reinterpret_cast<__half*>(b + w.off_z16);
    *inv_scale = reinterpret_cast<float*>(b + w.off_inv);
    *znorm = reinterpret_cast<float*>(b + w.off_znorm);
    *zres = reinterpret_cast<float*>(b + w.off_zres);
}

size_t search_tc16_workspace_bytes(int64_t n_tokens, int D, int K) { return t16_workspace(n_tokens, D, K).total; }

VQB_KNOB g_tc16_cluster = 2;
VQB_KNOB g_tc16_branchy = 0;  // vqb_tune "tc16_branchy": conditional top-4 insert in the epilogue (cluster 2 only)
VQB_KNOB g_tc16_pruned = 1;   // vqb_tune "tc16_pruned": 0 = no pruned exact tier (A/B)
VQB_KNOB g_tc16_group = 0;    // vqb_tune "tc16_group": 0 = tc16_group_size(Dpad), 4 or 8 = forced (cluster 2, Dpad <= 256)
#ifdef VQB_EXPERIMENTAL
void set_tc16_cluster(int c) { g_tc16_cluster = c; }
void set_tc16_branchy(int v) { g_tc16_branchy = v; }
void set_tc16_group(int v) { g_tc16_group = v; }
void set_tc16_pruned(int v) { g_tc16_pruned = v; }
#endif

// codes per candidate group: 8 where the epilogue, not the tensor pipe, bounds the kernel (see search_tc16_kernel)
static int tc16_group_size(int Dpad) {
    if (g_tc16_group == 4 || g_tc16_group == 8) return Dpad <= 256 ? g_tc16_group : 4;
    return Dpad <= kTc16WideGroupDpad ? 8 : 4;
}

template <int NKB, int CL, bool BR = false, int GS = 4>
static int launch_tc16_cl(const CUtensorMap& mz, const CUtensorMap& me, const T16Params& p, cudaStream_t s) {
    constexpr int kStages = tc16_stages(NKB);
    const size_t smem = 1024 + NKB * kTcABlockBytes + (size_t)kStages * kTcBStageBytes + kTcBarrierBytes + 12288;
    VQB_CUDA_TRY(cudaFuncSetAttribute(search_tc16_kernel<NKB, CL, BR, GS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem));
    const int n_m_tiles = (int)((p.N + kTcBM - 1) / kTcBM);
    int grid = n_m_tiles < sm_count() ? n_m_tiles : sm_count();
    grid = (grid + CL - 1) / CL * CL;               // whole clusters
    if (grid > sm_count()) grid = sm_count() / CL * CL;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(kT16Threads);
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
        // persistent kernel: never launch more clusters than can be co-resident (GPC granularity
        // strands some SMs for larger clusters)
        int max_clusters = 0;
        VQB_CUDA_TRY(cudaOccupancyMaxActiveClusters(&max_clusters, search_tc16_kernel<NKB, CL, BR, GS>, &cfg));
        if (max_clusters > 0 && grid > max_clusters * CL) {
            grid = max_clusters * CL;
            cfg.gridDim = dim3((unsigned)grid);
        }
    }
    VQB_CUDA_TRY(cudaLaunchKernelEx(&cfg, search_tc16_kernel<NKB, CL, BR, GS>, mz, me, p));
    return VQB_OK;
}

template <int NKB>
static int launch_tc16_t(const CUtensorMap& mz, const CUtensorMap& me1, const CUtensorMap& me2, const CUtensorMap& me4,
                         const T16Params& p, cudaStream_t s, int gs) {
#ifdef VQB_EXPERIMENTAL
    constexpr bool kWide = true;  // measurement build: either group size at any Dpad <= 256 (A/B)
#else
    constexpr bool kWide = NKB * kTcBK <= kTc16WideGroupDpad;
#endif
    if constexpr (kWide) {
        if (gs == 8) return launch_tc16_cl<NKB, 2, false, 8>(mz, me2, p, s);
    }
    switch (g_tc16_cluster) {
        case 1: return launch_tc16_cl<NKB, 1>(mz, me1, p, s);
        case 4: return launch_tc16_cl<NKB, 4>(mz, me4, p, s);
        default:
#ifdef VQB_EXPERIMENTAL  // measured slower (4.82 vs 4.32 ms at D=64, 7.56 vs 7.26 at D=256): measurement build only
            if (g_tc16_branchy) return launch_tc16_cl<NKB, 2, true>(mz, me2, p, s);
#endif
            return launch_tc16_cl<NKB, 2, false>(mz, me2, p, s);
    }
}

// exact re-score launch for a D of NCH 32-channel blocks; the wide-group instantiations exist only where tc16_group_size
// can ask for them (every Dpad <= 256 in the measurement build)
template <int NCH, bool kExact, int GS>
static int launch_rescore_g(const float* z, const float* E, const T16Params& p, int64_t N, int D, int64_t HW, int K,
                            int64_t* idx_out, float* dmin, cudaStream_t s) {
    constexpr int kSm = NCH * 32 * 36 * (int)sizeof(float);
    if (kSm > 48 * 1024)
        VQB_CUDA_TRY(cudaFuncSetAttribute(rescore_groups_kernel<NCH, kExact, GS>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSm));
    rescore_groups_kernel<NCH, kExact, GS><<<(unsigned)((N + 31) / 32), 256, kSm, s>>>(z, E, p.half_norm, p.group1, p.group2,
                                                                                      p.group3, N, D, HW, K, idx_out, dmin);
    return VQB_OK;
}

template <int NCH>
static int launch_rescore(const float* z, const float* E, const T16Params& p, int64_t N, int D, int64_t HW, int K,
                          int64_t* idx_out, float* dmin, int gs, cudaStream_t s) {
#ifdef VQB_EXPERIMENTAL
    constexpr bool kWide = NCH <= 8;
#else
    constexpr bool kWide = NCH * 32 <= kTc16WideGroupDpad;
#endif
    if constexpr (kWide) {
        if (gs == 8)
            return D == 32 * NCH ? launch_rescore_g<NCH, true, 8>(z, E, p, N, D, HW, K, idx_out, dmin, s)
                                 : launch_rescore_g<NCH, false, 8>(z, E, p, N, D, HW, K, idx_out, dmin, s);
    }
    return D == 32 * NCH ? launch_rescore_g<NCH, true, 4>(z, E, p, N, D, HW, K, idx_out, dmin, s)
                         : launch_rescore_g<NCH, false, 4>(z, E, p, N, D, HW, K, idx_out, dmin, s);
}

int launch_search_tc16(const float* z, int64_t B, int D, int64_t HW, const float* E, int K, const void* pack,
                       int64_t* idx_out, float* dmin_out, void* ws, size_t ws_bytes, int64_t* stats_out,
                       cudaStream_t s, bool presplit) {
    const int64_t N = B * HW;
    const T16Workspace w = t16_workspace(N, D, K);
    if (!ws || ws_bytes < w.total) {
        set_error("fp16 tensor search workspace too small: %zu < %zu", ws_bytes, w.total);
        return VQB_ERR_WORKSPACE;
    }
    if ((reinterpret_cast<uintptr_t>(ws) & 255u) != 0) {
        set_error("fp16 tensor search workspace must be 256-byte aligned");
        return VQB_ERR_INVALID_ARG;
    }
    const PackLayout L = pack_layout(K, D);
    unsigned char* wsb = static_cast<unsigned char*>(ws);
    const unsigned char* pk = static_cast<const unsigned char*>(pack);
    __half* z16 = reinterpret_cast<__half*>(wsb + w.off_z16);
    float* inv = reinterpret_cast<float*>(wsb + w.off_inv);
    float* znorm = reinterpret_cast<float*>(wsb + w.off_znorm);
    float* zres = reinterpret_cast<float*>(wsb + w.off_zres);
    int32_t* counts = reinterpret_cast<int32_t*>(wsb + w.off_counts);

    VQB_CUDA_TRY(cudaMemsetAsync(counts, 0, 64 + sizeof(int32_t) * kPrunedStateInts, s));
    const int Dpad = L.Dpad;
    if (!presplit) {  // (presplit: the producer of z -- vqb_conv1x1_split_f32 -- already filled z16 / inv / znorm / zres)
        const unsigned blocks = (unsigned)((N + 31) / 32);
        const int* hdr = reinterpret_cast<const int*>(pk);
#define VQB_SPLIT(nch)                                                                                                  \
    do {                                                                                                                \
        constexpr int kSm = (nch) * 32 * 33 * (int)sizeof(float);                                                       \
        if (D == Dpad) {                                                                                                \
            if (kSm > 48 * 1024)                                                                                        \
                VQB_CUDA_TRY(cudaFuncSetAttribute(split16_tokens_kernel<nch, true>,                                     \
                                                  cudaFuncAttributeMaxDynamicSharedMemorySize, kSm));                   \
            split16_tokens_kernel<nch, true><<<blocks, 256, kSm, s>>>(z, N, D, HW, hdr, z16, inv, znorm, zres);          \
        } else {                                                                                                        \
            if (kSm > 48 * 1024)                                                                                        \
                VQB_CUDA_TRY(cudaFuncSetAttribute(split16_tokens_kernel<nch, false>,                                    \
                                                  cudaFuncAttributeMaxDynamicSharedMemorySize, kSm));                   \
            split16_tokens_kernel<nch, false><<<blocks, 256, kSm, s>>>(z, N, D, HW, hdr, z16, inv, znorm, zres);         \
        }                                                                                                               \
    } while (0)
        switch (Dpad / 32) {
            case 2: VQB_SPLIT(2); break;
            case 4: VQB_SPLIT(4); break;
            case 6: VQB_SPLIT(6); break;
            case 8: VQB_SPLIT(8); break;
            case 10: VQB_SPLIT(10); break;
            case 12: VQB_SPLIT(12); break;
            case 14: VQB_SPLIT(14); break;
            default: VQB_SPLIT(16); break;
        }
#undef VQB_SPLIT
    }
    VQB_LAUNCH_CHECK("split16_tokens_kernel");

    // codebook maps: the box is the slice one CTA of the cluster fetches (256, 128 or 64 rows)
    CUtensorMap mz, me1, me2, me4;
    if (int rc = make_tc_map(&mz, z16, (uint64_t)N, Dpad, kTcBM, true)) return rc;
    if (int rc = make_tc_map(&me1, pk + L.off_e16, (uint64_t)L.Kpad, Dpad, kTcBN, true)) return rc;
    if (int rc = make_tc_map(&me2, pk + L.off_e16, (uint64_t)L.Kpad, Dpad, kTcBN / 2, true)) return rc;
    if (int rc = make_tc_map(&me4, pk + L.off_e16, (uint64_t)L.Kpad, Dpad, kTcBN / 4, true)) return rc;

    T16Params p;
    p.N = N;
    p.K = K;
    p.Kpad = L.Kpad;
    p.half_norm = reinterpret_cast<const float*>(pk + L.off_half_norm);
    p.half_norm_fin = reinterpret_cast<const float*>(pk + L.off_half_norm_fin);
    p.header = reinterpret_cast<const int*>(pk);
    p.gmax = reinterpret_cast<const float*>(pk + L.off_gmax);
    p.znorm = znorm;
    p.zres = zres;
    p.inv_scale = inv;
    p.group1 = reinterpret_cast<int32_t*>(wsb + w.off_g1);
    p.group2 = reinterpret_cast<int32_t*>(wsb + w.off_g2);
    p.group3 = reinterpret_cast<int32_t*>(wsb + w.off_g3);
    p.pair_count = counts + 1;
    p.full_list = reinterpret_cast<int32_t*>(wsb + w.off_full);
    p.full_count = counts + 0;
    p.acc_eps = D > 256 ? 1.f / 16384.f : 1.f / 32768.f;
    int rc;
    const int gs = tc16_group_size(Dpad);
    switch (Dpad / kTcBK) {
        case 1: rc = launch_tc16_t<1>(mz, me1, me2, me4, p, s, gs); break;
        case 2: rc = launch_tc16_t<2>(mz, me1, me2, me4, p, s, gs); break;
        case 3: rc = launch_tc16_t<3>(mz, me1, me2, me4, p, s, gs); break;
        case 4: rc = launch_tc16_t<4>(mz, me1, me2, me4, p, s, gs); break;
        // 256 < D <= 512: the resident token tile takes 80-128 KB, 4-2 codebook stages remain; clusters of 2 only
        case 5: rc = launch_tc16_cl<5, 2>(mz, me2, p, s); break;
        case 6: rc = launch_tc16_cl<6, 2>(mz, me2, p, s); break;
        case 7: rc = launch_tc16_cl<7, 2>(mz, me2, p, s); break;
        case 8: rc = launch_tc16_cl<8, 2>(mz, me2, p, s); break;
        default:
            set_error("fp16 tensor search supports 16 < D <= %d, got %d", kTc16MaxD, D);
            return VQB_ERR_UNSUPPORTED;
    }
    if (rc != VQB_OK) return rc;
    // exact fp32 choice among the 4 (or 8) certified candidates of every token; its score is also the upper bound the
    // pruned exact tier starts from, so it is kept in the workspace when the caller does not ask for it
    // (idle cost: five launches that return at once, ~18 us per search -- 0.2 % at C3, 1 % of a 458 K-token C5 step;
    // below kPrunedMinBatch tokens even the worst case of the full re-search is a few milliseconds, so the tier is left out)
    const bool pruned = g_tc16_pruned && w.pruned_bytes > 0 && N >= kPrunedMinBatch;
    float* dmin_eff = dmin_out;
    if (pruned && !dmin_eff) dmin_eff = search_pruned_ubound(wsb + w.off_pruned, N, K, D);
    switch ((D + 31) / 32) {
#define VQB_RESCORE(nch) case nch: rc = launch_rescore<nch>(z, E, p, N, D, HW, K, idx_out, dmin_eff, gs, s); break
        VQB_RESCORE(1); VQB_RESCORE(2); VQB_RESCORE(3); VQB_RESCORE(4); VQB_RESCORE(5); VQB_RESCORE(6); VQB_RESCORE(7);
        VQB_RESCORE(8); VQB_RESCORE(9); VQB_RESCORE(10); VQB_RESCORE(11); VQB_RESCORE(12); VQB_RESCORE(13); VQB_RESCORE(14);
        VQB_RESCORE(15);
#undef VQB_RESCORE
        default: rc = launch_rescore<16>(z, E, p, N, D, HW, K, idx_out, dmin_eff, gs, s); break;
    }
    if (rc != VQB_OK) return rc;
    VQB_LAUNCH_CHECK("rescore_groups_kernel");
    // ambiguous tokens: a LONG list (collapsed codebooks) first meets the pruned exact tier, which either takes all of
    // it or declines; what is left goes to the full exact fp32 search
    const int32_t* remaining = p.full_count;
    const int32_t* handled = nullptr;
    if (pruned) {
        rc = launch_search_pruned(z, B, D, HW, E, K, pack, p.full_list, p.full_count, dmin_eff, counts + 16,
                                  wsb + w.off_pruned, w.pruned_bytes, idx_out, dmin_out, &remaining, &handled, s);
        if (rc != VQB_OK) return rc;
    }
    rc = launch_search_fp32(z, B, D, HW, E, K, pack, p.full_list, remaining, N, wsb + w.off_keys, w.keys_bytes,
                            idx_out, dmin_out, s);
    if (rc != VQB_OK) return rc;
    if (stats_out) {
        tc16_stats_kernel<<<1, 1, 0, s>>>(stats_out, p.full_count, p.pair_count, handled);
        VQB_LAUNCH_CHECK("tc16_stats_kernel");
    }
    return VQB_OK;
}

}  // namespace vqb
