// Pruned exact tier of the fp16 tensor search (vqb_search_tc16.cu), between the certified candidate re-score and the
// full fp32 re-search.  Replaces nothing in the reference (quantizer.py:68-76 is one dense matmul + argmin); it exists
// for COLLAPSED codebooks -- many codes a rounding error apart -- where no low-precision first pass can certify a
// winner and every token used to fall through to the CUDA-core search over all K codes (300 ms instead of 7 ms per
// 1 M tokens at D = 256).  Exact arithmetic is still required, but not over the whole codebook:
//   * the codes are sorted along a fixed +-1 projection (near-duplicates land next to each other) and cut into tiles
//     of 128; every tile gets a bounding ball (centre c_t = mean of its members, radius r_t = max |e - c_t|);
//   * a token already has an exactly scored candidate (the certified re-score's best, score U): a code that beats or
//     ties it lies within R = sqrt(2 U + |z|^2) of z, so tiles with |z - c_t| - r_t > R cannot hold the winner;
//   * the surviving (token, tile) pairs are bucketed per tile and scored by the same register-tiled fp32 FMA chain as
//     search_fp32_kernel (bit-identical scores; ties broken on the ORIGINAL code index), winners meet in a 64-bit
//     atomicMin on (ordered score bits, index).
// Everything is decided on the device: the kernels are always launched and return at once unless the list of
// uncertified tokens is long (>= kPrunedMinTokens); if the pruning does not pay (more than kPrunedPairsPerToken
// surviving tiles per token on average) the tier declines and the full re-search runs as before.
#include "vqb_common.cuh"

namespace vqb {

constexpr int kPrTile = 128;             // codes per tile (= the fp32 kernel's code tile)
constexpr int kPrMaxTiles = 128;         // K <= 16384
constexpr int kPrMinK = 2048;
constexpr int kPrunedMinTokens = 4096;   // shorter lists: the plain list search is cheaper than sorting the codebook
constexpr int kPrunedPairsPerToken = 16; // capacity of the per-tile token lists, in tiles per listed token
constexpr int kPrTokens = 128, kPrDk = 16, kPrThreads = 256, kPrPad = 4;

bool pruned_eligible(int K, int D) { return K >= kPrMinK && K <= kPrTile * kPrMaxTiles && D > kLowDMax && D <= kTc16MaxD; }

struct PrunedWorkspace {
    size_t off_perm, off_es, off_hs, off_cent, off_rad, off_cn2, off_ub, off_mask, off_lists, total;
    int64_t cap;
    int p2, n_tiles;
};

static PrunedWorkspace pruned_workspace(int64_t N, int K, int D) {
    PrunedWorkspace w;
    w.n_tiles = (K + kPrTile - 1) / kPrTile;
    int p2 = 1;
    while (p2 < K) p2 <<= 1;
    w.p2 = p2;
    w.cap = (int64_t)kPrunedPairsPerToken * N;
    const size_t kp = (size_t)w.n_tiles * kPrTile;
    size_t off = 0;
    w.off_perm = off;
    off = round_up_z(off + 4 * kp, 1024);
    w.off_es = off;
    off = round_up_z(off + 4 * kp * D, 1024);
    w.off_hs = off;
    off = round_up_z(off + 4 * kp, 1024);
    w.off_cent = off;
    off = round_up_z(off + 4 * (size_t)D * kPrMaxTiles, 1024);
    w.off_rad = off;
    off = round_up_z(off + 4 * kPrMaxTiles, 1024);
    w.off_cn2 = off;
    off = round_up_z(off + 4 * kPrMaxTiles, 1024);
    w.off_ub = off;
    off = round_up_z(off + 4 * (size_t)N, 1024);
    w.off_mask = off;
    off = round_up_z(off + 16 * (size_t)N, 1024);
    w.off_lists = off;
    off = round_up_z(off + 4 * (size_t)w.cap, 1024);
    w.total = off;
    return w;
}

size_t search_pruned_workspace_bytes(int64_t N, int K, int D) { return pruned_eligible(K, D) ? pruned_workspace(N, K, D).total : 0; }

float* search_pruned_ubound(void* ws, int64_t N, int K, int D) {
    return reinterpret_cast<float*>(static_cast<unsigned char*>(ws) + pruned_workspace(N, K, D).off_ub);
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
__device__ __forceinline__ void pr_key_unpack(unsigned long long key, float& d, int& idx) {
    idx = (int)(unsigned int)(key & 0xffffffffull);
    int bits = (int)((unsigned int)(key >> 32) ^ 0x80000000u);
    if (bits < 0) bits ^= 0x7fffffff;
    d = __int_as_float(bits);
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
                        const int32_t* __restrict__ list_count, const int* __restrict__ header,
                        const int32_t* __restrict__ perm, float* __restrict__ Es, float* __restrict__ hs,
                        float* __restrict__ centT, float* __restrict__ rad, float* __restrict__ cn2) {
    if (!pruned_gate(list_count, header, K)) return;
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
    // columns: gather the rows (coalesced along d), centre = mean of the members
    for (int d = threadIdx.x; d < D; d += 256) {
        float s = 0.f;
        for (int j = 0; j < kPrTile; ++j) {
            const int k = src[j];
            const float v = k >= 0 ? __ldg(E + (size_t)k * D + d) : 0.f;
            Es[((size_t)t * kPrTile + j) * D + d] = v;
            s += v;
        }
        const float c = s / (float)n_here;
        c_s[d] = c;
        centT[(size_t)d * kPrMaxTiles + t] = c;
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

// ---- 4. which tiles can hold a code that beats or ties the token's exactly scored candidate ------------------------
// 32 listed tokens per step are staged dim-major; warp w owns rows 4w..4w+3, lane l the tiles l, l+32, l+64, l+96.
__global__ void __launch_bounds__(256)
    pruned_select_kernel(const float* __restrict__ z, int D, int64_t HW, int K, int n_tiles,
                         const int32_t* __restrict__ token_list, const int32_t* __restrict__ list_count,
                         const int* __restrict__ header, const float* __restrict__ ubound,
                         const float* __restrict__ centT, const float* __restrict__ rad, const float* __restrict__ cn2,
                         unsigned long long* __restrict__ keys, uint32_t* __restrict__ mask, int32_t* __restrict__ counts) {
    if (!pruned_gate(list_count, header, K)) return;
    extern __shared__ float zt_raw[];  // [D][36]
    float (*zt)[36] = reinterpret_cast<float (*)[36]>(zt_raw);
    __shared__ int64_t tok_s[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t count = *list_count;
    const int64_t groups = (count + 31) / 32;
    const float e_max = sqrtf(2.f * __int_as_float(header[4]));
    int cnt[4] = {0, 0, 0, 0};
    float rd[4], c2[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int t = lane + 32 * j;
        rd[j] = t < n_tiles ? rad[t] : INFINITY;
        c2[j] = t < n_tiles ? cn2[t] : 0.f;
    }
    for (int64_t g = blockIdx.x; g < groups; g += gridDim.x) {
        __syncthreads();
        {
            const int64_t r = g * 32 + lane;
            const int64_t tok = r < count ? (int64_t)token_list[r] : -1;
            if (warp == 0) tok_s[lane] = tok;
            int64_t off = 0;
            if (tok >= 0) {
                const int64_t b = tok / HW;
                off = (b * D) * HW + (tok - b * HW);
            }
            for (int d = warp; d < D; d += 8) zt[d][lane] = tok >= 0 ? __ldg(z + off + (int64_t)d * HW) : 0.f;
        }
        __syncthreads();
        float acc[4][4], zz[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
        for (int d = 0; d < D; ++d) {
            const float4 z4 = *reinterpret_cast<const float4*>(&zt[d][4 * warp]);
            const float zv[4] = {z4.x, z4.y, z4.z, z4.w};
            float c[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) c[j] = __ldg(centT + (size_t)d * kPrMaxTiles + lane + 32 * j);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                zz[i] = fmaf(zv[i], zv[i], zz[i]);
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(zv[i], c[j], acc[i][j]);
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int64_t row = g * 32 + 4 * warp + i;
            if (row >= count) continue;  // warp-uniform
            const int64_t tok = tok_s[4 * warp + i];
            const float U = __ldg(ubound + tok), zn2 = zz[i];
            // a code e with fp32 score s(e) <= U has real score <= U + eps, i.e. |z - e|^2 <= 2 (U + eps) + |z|^2; the
            // candidate's own real score is within eps of U as well, hence 2 eps.  eps: fp32 FMA chain of D <= 512 terms.
            const float eps = 1.220703125e-4f * (sqrtf(zn2) * e_max + 0.5f * e_max * e_max);
            float R = sqrtf(fmaxf(2.f * (U + 2.f * eps) + zn2, 0.f)) * 1.001f;
            if (!(fabsf(U) < INFINITY) || !(zn2 < INFINITY)) R = INFINITY;  // NaN / inf tokens: nothing is pruned
            uint32_t words[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float d2 = zn2 - 2.f * acc[i][j] + c2[j];
                const float delta = 3.814697265625e-6f * (zn2 + c2[j]);  // 2^-18: cancellation in d2 never inflates it
                const float lo = sqrtf(fmaxf(d2 - delta, 0.f));
                const bool prune = (lo - rd[j]) > R;  // NaN anywhere: keep the tile
                const bool surv = (lane + 32 * j < n_tiles) && !prune;
                words[j] = __ballot_sync(0xffffffffu, surv);
                cnt[j] += surv ? 1 : 0;
            }
            if (lane == 0) {
                *reinterpret_cast<uint4*>(mask + 4 * row) = make_uint4(words[0], words[1], words[2], words[3]);
                keys[row] = ~0ull;
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
        if (cnt[j]) atomicAdd(counts + lane + 32 * j, cnt[j]);
}

// prefix of the per-tile pair counts (every CTA recomputes it: 128 values)
__device__ __forceinline__ int64_t pruned_offsets(const int32_t* __restrict__ counts, int n_tiles, int64_t* off, int64_t* item0) {
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
    return off[n_tiles];
}

// ---- 5. decision + per-tile lists of list rows -----------------------------------------------------------------
__global__ void __launch_bounds__(256)
    pruned_scatter_kernel(int K, int n_tiles, int64_t cap, const int32_t* __restrict__ list_count,
                          const int* __restrict__ header, const int32_t* __restrict__ counts, int32_t* __restrict__ cursors,
                          const uint32_t* __restrict__ mask, int32_t* __restrict__ lists, int32_t* __restrict__ state) {
    __shared__ int64_t off[kPrMaxTiles + 1], item0[kPrMaxTiles + 1];
    const bool gate = pruned_gate(list_count, header, K);
    int64_t total = 0;
    if (gate) total = pruned_offsets(counts, n_tiles, off, item0);
    const bool decided = gate && total <= cap;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        state[0] = decided ? 1 : 0;
        state[1] = decided ? 0 : *list_count;  // what is left for the full re-search
        state[2] = decided ? *list_count : 0;  // statistics
    }
    if (!decided) return;
    const int lane = threadIdx.x & 31;
    const int64_t count = *list_count;
    const int64_t batches = (count + 31) / 32;
    const unsigned lt = (1u << lane) - 1u;
    for (int64_t bt = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); bt < batches; bt += (int64_t)gridDim.x * 8) {
        const int64_t row = bt * 32 + lane;
        uint4 m = make_uint4(0u, 0u, 0u, 0u);
        if (row < count) m = *reinterpret_cast<const uint4*>(mask + 4 * row);
        const uint32_t w[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (__ballot_sync(0xffffffffu, w[j] != 0u) == 0u) continue;
            for (int b = 0; b < 32; ++b) {
                const bool has = (w[j] >> b) & 1u;
                const unsigned bal = __ballot_sync(0xffffffffu, has);
                if (bal == 0u) continue;
                const int t = b + 32 * j;
                const int leader = __ffs(bal) - 1;
                int base = 0;
                if (lane == leader) base = atomicAdd(cursors + t, __popc(bal));
                base = __shfl_sync(0xffffffffu, base, leader);
                if (has) lists[off[t] + base + __popc(bal & lt)] = (int32_t)row;
            }
        }
    }
}

// ---- 6. exact scores of the surviving (token, tile) pairs: the FMA chain of search_fp32_kernel ----------------------
__global__ void __launch_bounds__(kPrThreads, 2)
    pruned_search_kernel(const float* __restrict__ z, int D, int64_t HW, int K, int n_tiles,
                         const int32_t* __restrict__ token_list, const int32_t* __restrict__ list_count,
                         const int* __restrict__ header, const int32_t* __restrict__ state,
                         const int32_t* __restrict__ counts, const int32_t* __restrict__ lists,
                         const float* __restrict__ Es, const float* __restrict__ hs, const int32_t* __restrict__ perm,
                         unsigned long long* __restrict__ keys, int32_t* __restrict__ done_ctas,
                         int64_t* __restrict__ idx_out, float* __restrict__ dmin_out) {
    if (!pruned_gate(list_count, header, K) || state[0] == 0) return;
    __shared__ __align__(16) float As[2][kPrDk][kPrTokens + kPrPad];
    __shared__ __align__(16) float Bs[2][kPrDk][kPrTile + kPrPad];
    __shared__ int64_t tok_off[kPrTokens];
    __shared__ int32_t row_id[kPrTokens];
    __shared__ int64_t off[kPrMaxTiles + 1], item0[kPrMaxTiles + 1];
    pruned_offsets(counts, n_tiles, off, item0);
    const int64_t n_items = item0[n_tiles];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    for (int64_t item = blockIdx.x; item < n_items; item += gridDim.x) {
        int lo_t = 0, hi_t = n_tiles;  // last tile with item0[t] <= item
        while (hi_t - lo_t > 1) {
            const int mid = (lo_t + hi_t) >> 1;
            if (item0[mid] <= item) lo_t = mid; else hi_t = mid;
        }
        const int t = lo_t;
        const int64_t start = off[t] + (item - item0[t]) * kPrTokens;
        const int64_t left = off[t + 1] - start;
        const int n_tok = left < kPrTokens ? (int)left : kPrTokens;
        __syncthreads();  // previous item is done with tok_off / row_id
        if (tid < kPrTokens) {
            int32_t r = -1;
            int64_t o = -1;
            if (tid < n_tok) {
                r = lists[start + tid];
                const int64_t tok = token_list[r];
                const int64_t b = tok / HW;
                o = (b * D) * HW + (tok - b * HW);
            }
            row_id[tid] = r;
            tok_off[tid] = o;
        }
        __syncthreads();
        float acc[8][8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
        constexpr int kRegA = (kPrDk * kPrTokens) / kPrThreads, kRegB = (kPrDk * kPrTile) / kPrThreads;
        const int a_mm = tid % kPrTokens, a_k = tid / kPrTokens;
        const int64_t a_off = tok_off[a_mm];
        const float* pa = z + (a_off >= 0 ? a_off : 0) + (int64_t)a_k * HW;
        const int b_nn = tid / kPrDk, b_k = tid % kPrDk;
        const float* pb = Es + ((size_t)t * kPrTile + b_nn) * D + b_k;
        auto issue = [&](int buf, int k0) {
#pragma unroll
            for (int i = 0; i < kRegA; ++i) {
                const int kk = a_k + i * (kPrThreads / kPrTokens);
                const bool ok = a_off >= 0 && k0 + kk < D;
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(&As[buf][kk][a_mm])),
                             "l"(ok ? pa + (int64_t)(k0 + i * (kPrThreads / kPrTokens)) * HW : z), "r"(ok ? 4 : 0)
                             : "memory");
            }
#pragma unroll
            for (int i = 0; i < kRegB; ++i) {
                const int nn = b_nn + i * (kPrThreads / kPrDk);
                const bool ok = k0 + b_k < D;
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(&Bs[buf][b_k][nn])),
                             "l"(ok ? pb + (size_t)i * (kPrThreads / kPrDk) * D + k0 : Es), "r"(ok ? 4 : 0)
                             : "memory");
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
                const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
                const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + ty * 4]);
                const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
                const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][64 + tx * 4]);
                const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
            }
            __syncthreads();
            buf ^= 1;
        }
        // epilogue: the same d = h - acc as the full search; the sorted order is not the index order, so ties are
        // broken on the original code index explicitly
        float best[8];
        int bidx[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            best[i] = INFINITY;
            bidx[i] = 0;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int n = t * kPrTile + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
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
                const int row = i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4);
                const int32_t r = row_id[row];
                if (r >= 0) atomicMin(keys + r, pr_score_key(best[i], bidx[i]));
            }
        }
    }
    // winners -> outputs: the CTA that finishes last unpacks every key (one CTA, a few hundred microseconds at 1 M rows;
    // a separate kernel would cost a launch on every call whether the tier runs or not)
    __shared__ int last;
    __threadfence();
    __syncthreads();
    if (tid == 0) last = (atomicAdd(done_ctas, 1) == (int)gridDim.x - 1);
    __syncthreads();
    if (!last) return;
    __threadfence();
    const int64_t count = *list_count;
    constexpr int kU = 16;  // independent rows in flight per thread (one row at a time ran at DRAM latency: 6 ms per 1 M rows)
    for (int64_t r0 = 0; r0 < count; r0 += (int64_t)kU * kPrThreads) {
        unsigned long long kv[kU];
        int64_t tk[kU];
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            const int64_t r = r0 + (int64_t)u * kPrThreads + tid;
            kv[u] = r < count ? __ldcg(keys + r) : ~0ull;
            tk[u] = r < count ? (int64_t)token_list[r] : -1;
        }
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            float d;
            int idx;
            pr_key_unpack(kv[u], d, idx);
            if (tk[u] < 0 || idx < 0 || idx >= K) continue;  // (idx: cannot happen, the candidate's own tile always survives)
            idx_out[tk[u]] = idx;
            if (dmin_out) dmin_out[tk[u]] = d;
        }
    }
}

// The tier.  `ubound[token]` = exact fp32 score of a code already chosen for the token (the certified re-score's result);
// `keys` = 8 bytes per list row of scratch (shared with the full re-search that follows).  On return *remaining points
// at the device-side count the full re-search must use instead of list_count (0 when the tier took the list, list_count
// when it declined) and *handled at the number of tokens the tier took (statistics).
// `state`: kPrunedStateInts int32 zeroed by the caller before the main kernel (decision, counters, per-tile counts / cursors).
int launch_search_pruned(const float* z, int64_t B, int D, int64_t HW, const float* E, int K, const void* pack,
                         const int32_t* token_list, const int32_t* list_count, const float* ubound,
                         unsigned long long* keys, int32_t* state, void* ws, size_t ws_bytes, int64_t* idx_out,
                         float* dmin_out, const int32_t** remaining, const int32_t** handled, cudaStream_t s) {
    const int64_t N = B * HW;
    const PrunedWorkspace w = pruned_workspace(N, K, D);
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
    float* centT = reinterpret_cast<float*>(b + w.off_cent);
    float* rad = reinterpret_cast<float*>(b + w.off_rad);
    float* cn2 = reinterpret_cast<float*>(b + w.off_cn2);
    uint32_t* mask = reinterpret_cast<uint32_t*>(b + w.off_mask);
    int32_t* lists = reinterpret_cast<int32_t*>(b + w.off_lists);
    const int sms = sm_count();

    // (state / counts / cursors were zeroed by the caller together with its own counters: one memset per call)
    const size_t sort_smem = 8 * (size_t)w.p2;
    if (sort_smem > 48 * 1024)
        VQB_CUDA_TRY(cudaFuncSetAttribute(pruned_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sort_smem));
    pruned_sort_kernel<<<1, 1024, sort_smem, s>>>(E, K, D, w.p2, list_count, header, perm);
    VQB_LAUNCH_CHECK("pruned_sort_kernel");
    pruned_tiles_kernel<<<w.n_tiles, 256, 0, s>>>(E, half_norm, K, D, list_count, header, perm, Es, hs, centT, rad, cn2);
    VQB_LAUNCH_CHECK("pruned_tiles_kernel");
    const size_t sel_smem = sizeof(float) * 36 * (size_t)D;
    if (sel_smem > 48 * 1024)
        VQB_CUDA_TRY(cudaFuncSetAttribute(pruned_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sel_smem));
    pruned_select_kernel<<<sms * 2, 256, sel_smem, s>>>(z, D, HW, K, w.n_tiles, token_list, list_count, header, ubound, centT,
                                                       rad, cn2, keys, mask, counts);
    VQB_LAUNCH_CHECK("pruned_select_kernel");
    pruned_scatter_kernel<<<sms * 2, 256, 0, s>>>(K, w.n_tiles, w.cap, list_count, header, counts, cursors, mask, lists, state);
    VQB_LAUNCH_CHECK("pruned_scatter_kernel");
    pruned_search_kernel<<<sms * 2, kPrThreads, 0, s>>>(z, D, HW, K, w.n_tiles, token_list, list_count, header, state, counts,
                                                       lists, Es, hs, perm, keys, state + 3, idx_out, dmin_out);
    VQB_LAUNCH_CHECK("pruned_search_kernel");
    *remaining = state + 1;
    *handled = state + 2;
    return VQB_OK;
}

}  // namespace vqb
