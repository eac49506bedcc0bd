// Pruned exact tier of the fp16 tensor search (vqb_search_tc16.cu), between the certified candidate re-score and the
// full fp32 re-search.  Replaces nothing in the reference (quantizer.py:68-76 is one dense matmul + argmin); it exists
// for COLLAPSED codebooks -- many codes a rounding error apart -- where no low-precision first pass can certify a
// winner and every token used to fall through to the CUDA-core search over all K codes (300 ms instead of 7 ms per
// 1 M tokens at D = 256).  Exact arithmetic is still required, but not over the whole codebook:
//   * the codes are sorted along a fixed +-1 projection (near-duplicates land next to each other) and cut into tiles
//     of 128; every tile gets a bounding ball (centre c_t = midpoint of its extreme members, radius r_t = max |e - c_t|);
//   * a token already has an exactly scored candidate (the certified re-score's best, score U): a code that beats or
//     ties it lies within R = sqrt(2 U + |z|^2) of z, so tiles with |z - c_t| - r_t > R cannot hold the winner;
//   * tokens are bucketed by their first surviving tile (tokens of one cluster share it); 128 tokens of a bucket are
//     gathered ONCE into a token-major scratch and scored against the union of their surviving tiles by the same
//     register-tiled fp32 FMA chain as search_fp32_kernel (bit-identical scores; the sorted order is not the index
//     order, so ties are broken on the ORIGINAL code index explicitly).  Every token sits in exactly one bucket, so the
//     winners are written straight to the outputs.
// Everything is decided on the device: the kernels are always launched and return at once unless the list of
// uncertified tokens is long (>= kPrunedMinTokens); if the pruning does not pay (on average more than half of the
// tiles survive per token) the tier declines and the full re-search runs as before.
#include "vqb_common.cuh"

namespace vqb {

constexpr int kPrTile = 128;             // codes per tile (= the fp32 kernel's code tile)
constexpr int kPrMaxTiles = 128;         // K <= 16384
constexpr int kPrMinK = 2048;
constexpr int kPrunedMinTokens = 4096;   // shorter lists: the plain list search is cheaper than sorting the codebook
constexpr int kPrTokens = 128, kPrDk = 16, kPrThreads = 256, kPrPad = 4;

bool pruned_eligible(int K, int D) { return K >= kPrMinK && K <= kPrTile * kPrMaxTiles && D > kLowDMax && D <= kTc16MaxD; }

struct PrunedWorkspace {
    size_t off_perm, off_es, off_hs, off_cent, off_rad, off_cn2, off_ub, off_mask, off_lists, off_scratch, total;
    int p2, n_tiles, search_ctas;
};

static PrunedWorkspace pruned_workspace(int64_t N, int K, int D, int sms) {
    PrunedWorkspace w;
    w.n_tiles = (K + kPrTile - 1) / kPrTile;
    int p2 = 1;
    while (p2 < K) p2 <<= 1;
    w.p2 = p2;
    w.search_ctas = 2 * sms;
    const size_t kp = (size_t)w.n_tiles * kPrTile;
    size_t off = 0;
    w.off_perm = off;
    off = round_up_z(off + 4 * kp, 1024);
    w.off_es = off;
    off = round_up_z(off + 4 * kp * D, 1024);
    w.off_hs = off;
    off = round_up_z(off + 4 * kp, 1024);
    w.off_cent = off;  // [kPrMaxTiles][D] row-major (zero rows beyond n_tiles)
    off = round_up_z(off + 4 * (size_t)D * kPrMaxTiles, 1024);
    w.off_rad = off;
    off = round_up_z(off + 4 * kPrMaxTiles, 1024);
    w.off_cn2 = off;
    off = round_up_z(off + 4 * kPrMaxTiles, 1024);
    w.off_ub = off;
    off = round_up_z(off + 4 * (size_t)N, 1024);
    w.off_mask = off;  // uint4 per list row: surviving tiles
    off = round_up_z(off + 16 * (size_t)N, 1024);
    w.off_lists = off;  // list rows bucketed by first surviving tile
    off = round_up_z(off + 4 * (size_t)N, 1024);
    w.off_scratch = off;  // per search CTA: 128 tokens x D, token-major
    off = round_up_z(off + 4 * (size_t)w.search_ctas * kPrTokens * D, 1024);
    w.total = off;
    return w;
}

// (sized for the largest sm_100 part so that the byte count does not depend on the device that later runs the search)
constexpr int kPrMaxSms = 160;
size_t search_pruned_workspace_bytes(int64_t N, int K, int D) {
    return pruned_eligible(K, D) ? pruned_workspace(N, K, D, kPrMaxSms).total : 0;
}

float* search_pruned_ubound(void* ws, int64_t N, int K, int D) {
    return reinterpret_cast<float*>(static_cast<unsigned char*>(ws) + pruned_workspace(N, K, D, kPrMaxSms).off_ub);
}

__device__ __forceinline__ bool pruned_gate(const int32_t* list_count, const int* header, int K) {
    return *list_count >= kPrunedMinTokens && header[0] >= K;  // long list, no NaN code
}

__device__ __forceinline__ unsigned long long pr_score_key(float d, int idx) {
    int bits = __float_as_int(d);
    if (bits < 0) bits ^= 0x7fffffff;
    const unsigned int u = (unsigned int)bits ^ 0x80000000u;
    return ((unsigned long long)u << 32) | (unsigned int)idx;
}

// ---- 1 + 2. projection keys key_k = sum_d s_d E[k, d], s_d = +-1 (a fixed hash of d), and a bitonic sort of the
// (key, code) words in shared memory: ONE CTA (p2 <= 16384; ~0.3 ms at K = 16384, D = 256, only when the tier runs) ----
__global__ void __launch_bounds__(1024)
    pruned_sort_kernel(const float* __restrict__ E, int K, int D, int p2, const int32_t* __restrict__ list_count,
                       const int* __restrict__ header, int32_t* __restrict__ perm) {
    if (!pruned_gate(list_count, header, K)) return;
    extern __shared__ unsigned long long sk[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int k = warp; k < p2; k += 32) {
        if (k >= K) {
            if (lane == 0) sk[k] = ~0ull;  // padding sorts last
            continue;
        }
        float s = 0.f;
        for (int d = lane; d < D; d += 32) {
            const unsigned h = (unsigned)d * 2654435761u;
            const float v = __ldg(E + (size_t)k * D + d);
            s += ((h >> 15) & 1u) ? v : -v;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) sk[k] = pr_score_key(s == s ? s : INFINITY, k);
    }
    __syncthreads();
    for (int size = 2; size <= p2; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = threadIdx.x; i < (p2 >> 1); i += 1024) {
                const int lo = ((i / stride) * stride * 2) + (i % stride);
                const int hi = lo + stride;
                const bool up = ((lo & size) == 0);
                const unsigned long long a = sk[lo], b = sk[hi];
                if ((a > b) == up) {
                    sk[lo] = b;
                    sk[hi] = a;
                }
            }
            __syncthreads();
        }
    }
    for (int i = threadIdx.x; i < K; i += 1024) perm[i] = (int32_t)(sk[i] & 0xffffffffull);
}

// ---- 3. per tile of 128 sorted codes: gathered rows, half norms, bounding ball ------------------------------------
__global__ void __launch_bounds__(256)
    pruned_tiles_kernel(const float* __restrict__ E, const float* __restrict__ half_norm, int K, int D,
                        int n_tiles, const int32_t* __restrict__ list_count, const int* __restrict__ header,
                        const int32_t* __restrict__ perm, float* __restrict__ Es, float* __restrict__ hs,
                        float* __restrict__ cent, float* __restrict__ rad, float* __restrict__ cn2) {
    if (!pruned_gate(list_count, header, K)) return;
    if ((int)blockIdx.x >= n_tiles) {  // rows of the centre table beyond the last tile: zeros (the block product reads all 128)
        for (int d = threadIdx.x; d < D; d += 256) cent[(size_t)blockIdx.x * D + d] = 0.f;
        return;
    }
    __shared__ int src[kPrTile];
    __shared__ float c_s[kTc16MaxD];
    __shared__ float wmax[8], wsum[8];
    const int t = blockIdx.x;
    const int n_here = (K - t * kPrTile) < kPrTile ? (K - t * kPrTile) : kPrTile;
    if (threadIdx.x < kPrTile) {
        const int j = threadIdx.x;
        const int k = j < n_here ? perm[t * kPrTile + j] : -1;
        src[j] = k;
        hs[t * kPrTile + j] = k >= 0 ? half_norm[k] : INFINITY;
    }
    __syncthreads();
    // columns: gather the rows (coalesced along d).  Centre = midpoint of the tile's first and last member in projection
    // order (any point is a valid centre; the radius below is measured from it).  NOT the mean: a tile that straddles two
    // tight clusters 98 : 2 has its mean next to the big one and a radius reaching the small one -- a ball through which
    // EVERY token of every cluster passes; from the midpoint the radius is half the distance between the clusters whatever
    // the mixture (measured on the collapsed C3 codebook: 14.5 -> 9.5 surviving tiles per token).
    for (int d = threadIdx.x; d < D; d += 256) {
        float first = 0.f, last = 0.f;
        for (int j = 0; j < kPrTile; ++j) {
            const int k = src[j];
            const float v = k >= 0 ? __ldg(E + (size_t)k * D + d) : 0.f;
            Es[((size_t)t * kPrTile + j) * D + d] = v;
            if (j == 0) first = v;
            if (j == n_here - 1) last = v;
        }
        const float c = 0.5f * first + 0.5f * last;
        c_s[d] = c;
        cent[(size_t)t * D + d] = c;
    }
    __syncthreads();
    // rows: radius = max |e - c| over the members; |c|^2
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float mx = 0.f;
    for (int j = warp; j < n_here; j += 8) {
        float q = 0.f;
        for (int d = lane; d < D; d += 32) {
            const float df = Es[((size_t)t * kPrTile + j) * D + d] - c_s[d];
            q = fmaf(df, df, q);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
        mx = fmaxf(mx, q);
    }
    float cq = 0.f;
    for (int d = threadIdx.x; d < D; d += 256) cq = fmaf(c_s[d], c_s[d], cq);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cq += __shfl_xor_sync(0xffffffffu, cq, o);
    if (lane == 0) {
        wmax[warp] = mx;
        wsum[warp] = cq;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float m = 0.f, s = 0.f;
        for (int w = 0; w < 8; ++w) {
            m = fmaxf(m, wmax[w]);
            s += wsum[w];
        }
        // inflated: fp32 rounding of the differences / the sum (relative 2^-16 at D = 512) must never shrink the ball;
        // a non-finite member makes the ball infinite (nothing is pruned against it)
        rad[t] = (m == m && m < INFINITY) ? sqrtf(m) * 1.001f + 1e-30f : INFINITY;
        cn2[t] = s;
    }
}

// ---- the 128 x 128 x D block product of search_fp32_kernel: acc[i][j] = sum_d A[row_i][d] * B[col_j][d], d ascending,
// one fmaf per term (the chain that defines the exact score).  B is row-major [128][D]; A is either gathered from the
// latents (kGather: element (r, d) at z + tok_off[r] + d * HW, rows with tok_off < 0 read as zero) or row-major
// [128][D].  kNorm also accumulates |A row|^2 for the thread's 8 rows.  Thread (tx, ty) owns rows ty*4+i / 64+ty*4+i-4
// and columns tx*4+j / 64+tx*4+j-4.
struct PrSmem {
    float As[2][kPrDk][kPrTokens + kPrPad];
    float Bs[2][kPrDk][kPrTile + kPrPad];
};

__device__ __forceinline__ void pr_cp4(float* dst, const float* src, bool ok) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src),
                 "r"(ok ? 4 : 0)
                 : "memory");
}

template <bool kGather, bool kNorm>
__device__ __forceinline__ void pr_block_product(PrSmem& sm, const float* __restrict__ a_base, const int64_t* tok_off,
                                                 int64_t HW, const float* __restrict__ b_base, int D, float (&acc)[8][8],
                                                 float (&zz)[8]) {
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        zz[i] = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    }
    constexpr int kReg = (kPrDk * kPrTokens) / kPrThreads;  // 8 copies per thread and operand
    const int g_mm = tid % kPrTokens, g_k = tid / kPrTokens;  // gather mapping: consecutive threads -> consecutive rows
    const int r_nn = tid / kPrDk, r_k = tid % kPrDk;          // row-major mapping: consecutive threads -> consecutive dims
    int64_t a_off = 0;
    if (kGather) a_off = tok_off[g_mm];
    const float* pa = kGather ? a_base + (a_off >= 0 ? a_off : 0) + (int64_t)g_k * HW : a_base + (size_t)r_nn * D + r_k;
    const float* pb = b_base + (size_t)r_nn * D + r_k;
    auto issue = [&](int buf, int k0) {
#pragma unroll
        for (int i = 0; i < kReg; ++i) {
            if (kGather) {
                const int kk = g_k + i * (kPrThreads / kPrTokens);
                const bool ok = a_off >= 0 && k0 + kk < D;
                pr_cp4(&sm.As[buf][kk][g_mm], ok ? pa + (int64_t)(k0 + i * (kPrThreads / kPrTokens)) * HW : a_base, ok);
            } else {
                const int nn = r_nn + i * (kPrThreads / kPrDk);
                const bool ok = k0 + r_k < D;
                pr_cp4(&sm.As[buf][r_k][nn], ok ? pa + (size_t)i * (kPrThreads / kPrDk) * D + k0 : a_base, ok);
            }
        }
#pragma unroll
        for (int i = 0; i < kReg; ++i) {
            const int nn = r_nn + i * (kPrThreads / kPrDk);
            const bool ok = k0 + r_k < D;
            pr_cp4(&sm.Bs[buf][r_k][nn], ok ? pb + (size_t)i * (kPrThreads / kPrDk) * D + k0 : b_base, ok);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    int buf = 0;
    issue(0, 0);
    for (int k0 = 0; k0 < D; k0 += kPrDk) {
        if (k0 + kPrDk < D) {
            issue(buf ^ 1, k0 + kPrDk);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < kPrDk; ++kk) {
            const float4 a0 = *reinterpret_cast<const float4*>(&sm.As[buf][kk][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&sm.As[buf][kk][64 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&sm.Bs[buf][kk][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4*>(&sm.Bs[buf][kk][64 + tx * 4]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (kNorm) zz[i] = fmaf(a[i], a[i], zz[i]);
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
            }
        }
        __syncthreads();  // everyone is done with `buf` before the next iteration's copies overwrite it
        buf ^= 1;
    }
}

__device__ __forceinline__ int pr_row_of(int ty, int i) { return i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4); }
__device__ __forceinline__ int pr_col_of(int tx, int j) { return j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4); }

// ---- 4. which tiles can hold a code that beats or ties the token's exactly scored candidate ------------------------
// 128 listed tokens x the (up to) 128 tile centres as one block product; the tile set of a token leaves as 128 bits.
__global__ void __launch_bounds__(kPrThreads, 2)
    pruned_select_kernel(const float* __restrict__ z, int D, int64_t HW, int K, int n_tiles,
                         const int32_t* __restrict__ token_list, const int32_t* __restrict__ list_count,
                         const int* __restrict__ header, const float* __restrict__ ubound,
                         const float* __restrict__ cent, const float* __restrict__ rad, const float* __restrict__ cn2,
                         uint32_t* __restrict__ mask, int32_t* __restrict__ counts, unsigned long long* __restrict__ pairs) {
    if (!pruned_gate(list_count, header, K)) return;
    __shared__ __align__(16) PrSmem sm;
    __shared__ int64_t tok_off[kPrTokens];
    __shared__ int64_t tok_id[kPrTokens];
    __shared__ uint32_t mask_s[kPrTokens][4];
    __shared__ int cnt_s[kPrMaxTiles];
    __shared__ unsigned int pair_s;
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int64_t count = *list_count;
    const int64_t chunks = (count + kPrTokens - 1) / kPrTokens;
    const float e_max = sqrtf(2.f * __int_as_float(header[4]));
    if (tid < kPrMaxTiles) cnt_s[tid] = 0;
    if (tid == 0) pair_s = 0u;
    float rd[8], c2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int t = pr_col_of(tx, j);
        rd[j] = t < n_tiles ? rad[t] : INFINITY;
        c2[j] = t < n_tiles ? cn2[t] : 0.f;
    }
    for (int64_t ch = blockIdx.x; ch < chunks; ch += gridDim.x) {
        __syncthreads();  // the previous chunk is done with tok_off / mask_s
        if (tid < kPrTokens) {
            const int64_t r = ch * kPrTokens + tid;
            int64_t tok = -1, o = -1;
            if (r < count) {
                tok = token_list[r];
                const int64_t b = tok / HW;
                o = (b * D) * HW + (tok - b * HW);
            }
            tok_id[tid] = tok;
            tok_off[tid] = o;
            mask_s[tid][0] = mask_s[tid][1] = mask_s[tid][2] = mask_s[tid][3] = 0u;
        }
        __syncthreads();
        float acc[8][8], zz[8];
        pr_block_product<true, true>(sm, z, tok_off, HW, cent, D, acc, zz);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int row = pr_row_of(ty, i);
            const int64_t tok = tok_id[row];
            if (tok < 0) continue;
            const float U = __ldg(ubound + tok), zn2 = zz[i];
            // a code e with fp32 score s(e) <= s(candidate) has real score <= U + 2 eps (both within eps of their fp32
            // values; eps: an fp32 FMA chain of D <= 512 terms, with a 4x margin), i.e. |z - e|^2 <= 2 (U + 2 eps) + |z|^2
            const float eps = 1.220703125e-4f * (sqrtf(zn2) * e_max + 0.5f * e_max * e_max);
            float R = sqrtf(fmaxf(2.f * (U + 2.f * eps) + zn2, 0.f)) * 1.001f;
            if (!(fabsf(U) < INFINITY) || !(zn2 < INFINITY)) R = INFINITY;  // NaN / inf tokens: nothing is pruned
            uint32_t w0 = 0u, w1 = 0u, w2 = 0u, w3 = 0u;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int t = pr_col_of(tx, j);
                const float d2 = zn2 - 2.f * acc[i][j] + c2[j];
                const float delta = 3.814697265625e-6f * (zn2 + c2[j]);  // 2^-18: cancellation in d2 never inflates it
                const float lo = sqrtf(fmaxf(d2 - delta, 0.f));
                const bool prune = (lo - rd[j]) > R;  // NaN anywhere: keep the tile
                if (t < n_tiles && !prune) {
                    const uint32_t bit = 1u << (t & 31);
                    if (t < 32) w0 |= bit; else if (t < 64) w1 |= bit; else if (t < 96) w2 |= bit; else w3 |= bit;
                }
            }
            if (w0) atomicOr(&mask_s[row][0], w0);
            if (w1) atomicOr(&mask_s[row][1], w1);
            if (w2) atomicOr(&mask_s[row][2], w2);
            if (w3) atomicOr(&mask_s[row][3], w3);
        }
        __syncthreads();
        if (tid < kPrTokens && tok_id[tid] >= 0) {
            const uint4 m = make_uint4(mask_s[tid][0], mask_s[tid][1], mask_s[tid][2], mask_s[tid][3]);
            *reinterpret_cast<uint4*>(mask + 4 * (ch * kPrTokens + tid)) = m;
            // bucket = first surviving tile (the candidate's own tile always survives, so there is one)
            const int first = m.x ? __ffs(m.x) - 1 : (m.y ? 31 + __ffs(m.y) : (m.z ? 63 + __ffs(m.z) : (m.w ? 95 + __ffs(m.w) : 0)));
            atomicAdd(&cnt_s[first], 1);
            atomicAdd(&pair_s, (unsigned)(__popc(m.x) + __popc(m.y) + __popc(m.z) + __popc(m.w)));
        }
    }
    __syncthreads();
    if (tid < kPrMaxTiles && cnt_s[tid]) atomicAdd(counts + tid, cnt_s[tid]);
    if (tid == 0 && pair_s) atomicAdd(pairs, (unsigned long long)pair_s);
}

// prefix of the per-bucket token counts (every CTA recomputes it: 128 values)
__device__ __forceinline__ void pruned_offsets(const int32_t* __restrict__ counts, int n_tiles, int64_t* off, int64_t* item0) {
    if (threadIdx.x == 0) {
        int64_t o = 0, it = 0;
        for (int t = 0; t < n_tiles; ++t) {
            off[t] = o;
            item0[t] = it;
            o += counts[t];
            it += (counts[t] + kPrTokens - 1) / kPrTokens;
        }
        off[n_tiles] = o;
        item0[n_tiles] = it;
    }
    __syncthreads();
}

// ---- 5. decision + bucket lists ------------------------------------------------------------------------------------
// state: [0] decided, [1] tokens left for the full re-search, [2] tokens taken (statistics), [4..5] surviving pairs (u64)
__global__ void __launch_bounds__(256)
    pruned_scatter_kernel(int K, int n_tiles, const int32_t* __restrict__ list_count, const int* __restrict__ header,
                          const int32_t* __restrict__ counts, int32_t* __restrict__ cursors,
                          const uint32_t* __restrict__ mask, int32_t* __restrict__ lists, int32_t* __restrict__ state) {
    __shared__ int64_t off[kPrMaxTiles + 1], item0[kPrMaxTiles + 1];
    const bool gate = pruned_gate(list_count, header, K);
    bool decided = false;
    if (gate) {
        // worth it when, on average, at most half of the tiles survive per token (the tier's own passes cost about two
        // tiles' worth of work per token)
        const unsigned long long pairs = *reinterpret_cast<const unsigned long long*>(state + 4);
        const int per_tok = n_tiles / 2 > 2 ? n_tiles / 2 : 2;
        decided = pairs <= (unsigned long long)(*list_count) * (unsigned long long)per_tok;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        state[0] = decided ? 1 : 0;
        state[1] = decided ? 0 : *list_count;
        state[2] = decided ? *list_count : 0;
    }
    if (!decided) return;
    pruned_offsets(counts, n_tiles, off, item0);
    const int lane = threadIdx.x & 31;
    const int64_t count = *list_count;
    const int64_t batches = (count + 31) / 32;
    const unsigned lt = (1u << lane) - 1u;
    for (int64_t bt = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); bt < batches; bt += (int64_t)gridDim.x * 8) {
        const int64_t row = bt * 32 + lane;
        const bool live = row < count;
        int first = -1;
        if (live) {
            const uint4 m = *reinterpret_cast<const uint4*>(mask + 4 * row);
            first = m.x ? __ffs(m.x) - 1 : (m.y ? 31 + __ffs(m.y) : (m.z ? 63 + __ffs(m.z) : (m.w ? 95 + __ffs(m.w) : 0)));
        }
        const unsigned active = __ballot_sync(0xffffffffu, live);
        if (live) {
            const unsigned peers = __match_any_sync(active, first);
            const int leader = __ffs(peers) - 1;
            int base = 0;
            if (lane == leader) base = atomicAdd(cursors + first, __popc(peers));
            base = __shfl_sync(peers, base, leader);
            lists[off[first] + base + __popc(peers & lt)] = (int32_t)row;
        }
    }
}

// ---- 6. exact scores: 128 tokens of a bucket x the union of their surviving tiles ---------------------------------
__global__ void __launch_bounds__(kPrThreads, 2)
    pruned_search_kernel(const float* __restrict__ z, int D, int64_t HW, int K, int n_tiles,
                         const int32_t* __restrict__ token_list, const int32_t* __restrict__ list_count,
                         const int* __restrict__ header, const int32_t* __restrict__ state,
                         const int32_t* __restrict__ counts, const int32_t* __restrict__ lists,
                         const uint32_t* __restrict__ mask, const float* __restrict__ Es, const float* __restrict__ hs,
                         const int32_t* __restrict__ perm, float* __restrict__ scratch_all,
                         int64_t* __restrict__ idx_out, float* __restrict__ dmin_out) {
    if (!pruned_gate(list_count, header, K) || state[0] == 0) return;
    __shared__ __align__(16) PrSmem sm;
    __shared__ int64_t tok_off[kPrTokens];
    __shared__ int64_t tok_id[kPrTokens];
    __shared__ uint32_t uni[4];
    __shared__ int64_t off[kPrMaxTiles + 1], item0[kPrMaxTiles + 1];
    pruned_offsets(counts, n_tiles, off, item0);
    const int64_t n_items = item0[n_tiles];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    float* scratch = scratch_all + (size_t)blockIdx.x * kPrTokens * D;
    for (int64_t item = blockIdx.x; item < n_items; item += gridDim.x) {
        int lo_t = 0, hi_t = n_tiles;  // last bucket with item0[t] <= item
        while (hi_t - lo_t > 1) {
            const int mid = (lo_t + hi_t) >> 1;
            if (item0[mid] <= item) lo_t = mid; else hi_t = mid;
        }
        const int64_t start = off[lo_t] + (item - item0[lo_t]) * kPrTokens;
        const int64_t left = off[lo_t + 1] - start;
        const int n_tok = left < kPrTokens ? (int)left : kPrTokens;
        __syncthreads();  // the previous item is done with tok_off / tok_id / uni / scratch
        if (tid < 4) uni[tid] = 0u;
        __syncthreads();
        if (tid < kPrTokens) {
            int64_t tok = -1, o = -1;
            if (tid < n_tok) {
                const int32_t r = lists[start + tid];
                tok = token_list[r];
                const int64_t b = tok / HW;
                o = (b * D) * HW + (tok - b * HW);
                const uint4 m = *reinterpret_cast<const uint4*>(mask + 4 * (int64_t)r);
                if (m.x) atomicOr(&uni[0], m.x);
                if (m.y) atomicOr(&uni[1], m.y);
                if (m.z) atomicOr(&uni[2], m.z);
                if (m.w) atomicOr(&uni[3], m.w);
            }
            tok_id[tid] = tok;
            tok_off[tid] = o;
        }
        __syncthreads();
        // gather the 128 tokens once (4-byte reads scattered over the latents) into the token-major scratch: every
        // tile of the union then reads them back in whole 64-byte runs
        for (int k0 = 0; k0 < D; k0 += kPrDk) {
            {
                const int mm = tid % kPrTokens;
                const int64_t o = tok_off[mm];
#pragma unroll
                for (int i = 0; i < (kPrDk * kPrTokens) / kPrThreads; ++i) {
                    const int kk = tid / kPrTokens + i * (kPrThreads / kPrTokens);
                    sm.As[0][kk][mm] = (o >= 0 && k0 + kk < D) ? __ldg(z + o + (int64_t)(k0 + kk) * HW) : 0.f;
                }
            }
            __syncthreads();
            {
                // 128 rows x 16 dims: thread -> (row = tid / 2, 8 dims): two 16-byte stores
                const int row = tid >> 1, h = (tid & 1) * 8;
                float v[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) v[q] = sm.As[0][h + q][row];
                if ((D & 3) == 0 && k0 + h + 8 <= D) {
                    *reinterpret_cast<float4*>(scratch + (size_t)row * D + k0 + h) = make_float4(v[0], v[1], v[2], v[3]);
                    *reinterpret_cast<float4*>(scratch + (size_t)row * D + k0 + h + 4) = make_float4(v[4], v[5], v[6], v[7]);
                } else {
#pragma unroll
                    for (int q = 0; q < 8; ++q)
                        if (k0 + h + q < D) scratch[(size_t)row * D + k0 + h + q] = v[q];
                }
            }
            __syncthreads();
        }
        __threadfence_block();
        float best[8];
        int bidx[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            best[i] = INFINITY;
            bidx[i] = 0;
        }
        for (int w = 0; w < 4; ++w) {
            uint32_t bits = uni[w];
            while (bits) {
                const int t = 32 * w + __ffs(bits) - 1;
                bits &= bits - 1;
                float acc[8][8], zz[8];
                pr_block_product<false, false>(sm, scratch, nullptr, 0, Es + (size_t)t * kPrTile * D, D, acc, zz);
                // the same d = h - acc as the full search; ties are broken on the original code index
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int n = t * kPrTile + pr_col_of(tx, j);
                    const float h = hs[n];  // +inf for the padding of the last tile
                    const int orig = n < K ? perm[n] : 0x7fffffff;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float d = h - acc[i][j];
                        if (d < best[i] || (d == best[i] && d < INFINITY && orig < bidx[i])) {
                            best[i] = d;
                            bidx[i] = orig;
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, best[i], o);
                const int oi = __shfl_xor_sync(0xffffffffu, bidx[i], o);
                if (ov < best[i] || (ov == best[i] && oi < bidx[i])) {
                    best[i] = ov;
                    bidx[i] = oi;
                }
            }
        }
        if (tx == 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int64_t tok = tok_id[pr_row_of(ty, i)];
                if (tok < 0) continue;
                idx_out[tok] = bidx[i];
                if (dmin_out) dmin_out[tok] = best[i];
            }
        }
    }
}

// The tier.  `ubound[token]` = exact fp32 score of a code already chosen for the token (the certified re-score's result).
// On return *remaining points at the device-side count the full re-search must use instead of list_count (0 when the
// tier took the list, list_count when it declined) and *handled at the number of tokens the tier took (statistics).
// `state`: kPrunedStateInts int32 zeroed by the caller before the main kernel (decision, counters, per-tile counts / cursors).
int launch_search_pruned(const float* z, int64_t B, int D, int64_t HW, const float* E, int K, const void* pack,
                         const int32_t* token_list, const int32_t* list_count, const float* ubound, int32_t* state,
                         void* ws, size_t ws_bytes, int64_t* idx_out, float* dmin_out, const int32_t** remaining,
                         const int32_t** handled, cudaStream_t s) {
    const int64_t N = B * HW;
    const int sms = sm_count();
    const PrunedWorkspace w = pruned_workspace(N, K, D, kPrMaxSms);
    if (!pruned_eligible(K, D) || !ws || ws_bytes < w.total) {
        set_error("pruned exact tier: not eligible or workspace too small (%zu < %zu)", ws_bytes, w.total);
        return VQB_ERR_WORKSPACE;
    }
    unsigned char* b = static_cast<unsigned char*>(ws);
    const unsigned char* pk = static_cast<const unsigned char*>(pack);
    const PackLayout L = pack_layout(K, D);
    const int* header = reinterpret_cast<const int*>(pk);
    const float* half_norm = reinterpret_cast<const float*>(pk + L.off_half_norm);
    int32_t* counts = state + 16;
    int32_t* cursors = counts + kPrMaxTiles;
    int32_t* perm = reinterpret_cast<int32_t*>(b + w.off_perm);
    float* Es = reinterpret_cast<float*>(b + w.off_es);
    float* hs = reinterpret_cast<float*>(b + w.off_hs);
    float* cent = reinterpret_cast<float*>(b + w.off_cent);
    float* rad = reinterpret_cast<float*>(b + w.off_rad);
    float* cn2 = reinterpret_cast<float*>(b + w.off_cn2);
    uint32_t* mask = reinterpret_cast<uint32_t*>(b + w.off_mask);
    int32_t* lists = reinterpret_cast<int32_t*>(b + w.off_lists);
    float* scratch = reinterpret_cast<float*>(b + w.off_scratch);

    // (state / counts / cursors were zeroed by the caller together with its own counters: one memset per call)
    const size_t sort_smem = 8 * (size_t)w.p2;
    if (sort_smem > 48 * 1024)
        VQB_CUDA_TRY(cudaFuncSetAttribute(pruned_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sort_smem));
    pruned_sort_kernel<<<1, 1024, sort_smem, s>>>(E, K, D, w.p2, list_count, header, perm);
    VQB_LAUNCH_CHECK("pruned_sort_kernel");
    pruned_tiles_kernel<<<kPrMaxTiles, 256, 0, s>>>(E, half_norm, K, D, w.n_tiles, list_count, header, perm, Es, hs, cent, rad, cn2);
    VQB_LAUNCH_CHECK("pruned_tiles_kernel");
    pruned_select_kernel<<<sms * 2, kPrThreads, 0, s>>>(z, D, HW, K, w.n_tiles, token_list, list_count, header, ubound, cent, rad,
                                                       cn2, mask, counts, reinterpret_cast<unsigned long long*>(state + 4));
    VQB_LAUNCH_CHECK("pruned_select_kernel");
    pruned_scatter_kernel<<<sms * 2, 256, 0, s>>>(K, w.n_tiles, list_count, header, counts, cursors, mask, lists, state);
    VQB_LAUNCH_CHECK("pruned_scatter_kernel");
    const int search_ctas = 2 * (sms < kPrMaxSms ? sms : kPrMaxSms);  // one scratch slice per CTA (workspace sized for kPrMaxSms)
    pruned_search_kernel<<<search_ctas, kPrThreads, 0, s>>>(z, D, HW, K, w.n_tiles, token_list, list_count, header, state, counts,
                                                       lists, mask, Es, hs, perm, scratch, idx_out, dmin_out);
    VQB_LAUNCH_CHECK("pruned_search_kernel");
    *remaining = state + 1;
    *handled = state + 2;
    return VQB_OK;
}

}  // namespace vqb
