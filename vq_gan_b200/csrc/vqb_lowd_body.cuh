// Device code of the CUDA-core low-D search (see vqb_search_lowd.cu for the design notes): shared by
// search_lowd_kernel and by the FMA role of search_dual_kernel (vqb_search_tclow.cu).
#pragma once
#include "vqb_common.cuh"

namespace vqb {

constexpr int kChunkCodes = 64;  // codes per index-tracking chunk (re-scored as kChunkCodes/64 rounds of 32 pairs, one per lane;
                                 // 128 measured 3 % slower: 3.00 vs 2.91 ms at D=4)

// launch shape variants (selected at run time by vqb_tune("lowd_variant", v); 0 is the default)
//   V0: 256 threads x 2 CTA/SM, up to 8 tokens/thread (measured best: 47.5 TFLOP/s at D=4)
//   V1: 256 x 1 CTA/SM, 8 tokens/thread (45.6)            V2: 512 x 1 CTA/SM, 4 tokens/thread (47.1)
//   V3: 256 x 2 CTA/SM, 4 tokens/thread (44.5)            V4: 256 x 3 CTA/SM, 4 tokens/thread (42.9)
template <int V>
struct LowDVariant {
    static constexpr int kThreads = V == 2 ? 512 : 256;
    static constexpr int kMinBlocks = (V == 0 || V == 3) ? 2 : (V == 4 ? 3 : 1);
    static constexpr int kTcap = (V >= 2) ? 4 : 8;
};

template <int D, int V>
struct LowDCfg {
    static constexpr int kThreads = LowDVariant<V>::kThreads;
    static constexpr int kMinBlocks = LowDVariant<V>::kMinBlocks;
    static constexpr int kTbase = D <= 4 ? 8 : (D <= 8 ? 4 : 2);
    static constexpr int kTmax = kTbase < LowDVariant<V>::kTcap ? kTbase : LowDVariant<V>::kTcap;
    static constexpr int kTileCodes = D <= 4 ? 2048 : (D <= 8 ? 1024 : 512);
    static constexpr int kTileFloats = kTileCodes * (D + 1);
    static constexpr size_t kSmemBytes = 2 * sizeof(float) * kTileFloats + 64;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes,
                                         uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

template <int D, int V>
struct LowDCtx {  // (kList kernels fill `list`; the dense kernels leave it null and never read it)
    const float* z;
    int64_t N, HW;             // N = work items: all tokens, or the entries of `list`
    const int32_t* list;       // optional token list (exact re-search of flagged tokens)
    const float* g_pairs;      // [Kpad/2][2D]
    const float* g_half_norm;  // [Kpad]
    int codes_padded;          // K rounded up to kChunkCodes
    int n_tiles;
    int K, first_nan;
    int64_t* idx_out;
    float* dmin_out;
    float* smem_tiles;  // 2 x kTileFloats
    uint64_t* bars;     // 2
    uint32_t visit;     // tile visits consumed so far (all threads agree)
    uint32_t total_visits;
};

template <int D, int V>
__device__ __forceinline__ int tile_codes(const LowDCtx<D, V>& c, int tile) {
    const int left = c.codes_padded - tile * LowDCfg<D, V>::kTileCodes;
    return left < LowDCfg<D, V>::kTileCodes ? left : LowDCfg<D, V>::kTileCodes;
}

// one thread: enqueue the bulk copies of tile-visit `v` into buffer v&1
template <int D, int V>
__device__ __forceinline__ void issue_visit(const LowDCtx<D, V>& c, uint32_t v) {
    using Cfg = LowDCfg<D, V>;
    const int tile = v % c.n_tiles;
    const int codes = tile_codes<D, V>(c, tile);
    float* buf = c.smem_tiles + (v & 1) * Cfg::kTileFloats;
    uint64_t* bar = c.bars + (v & 1);
    const uint32_t bytes_e = codes * D * sizeof(float);
    const uint32_t bytes_h = codes * sizeof(float);
    mbar_expect_tx(bar, bytes_e + bytes_h);
    bulk_g2s(buf, c.g_pairs + (size_t)tile * Cfg::kTileCodes * D, bytes_e, bar);
    bulk_g2s(buf + Cfg::kTileCodes * D, c.g_half_norm + (size_t)tile * Cfg::kTileCodes, bytes_h, bar);
}

// score of one code pair for one token: D chained FFMA2 seeded with the half norms
template <int D>
__device__ __forceinline__ unsigned long long pair_score(const unsigned long long (&nz)[D],
                                                         const unsigned long long (&ev)[D],
                                                         unsigned long long h2) {
    unsigned long long acc = fma_f32x2(nz[0], ev[0], h2);
#pragma unroll
    for (int d = 1; d < D; ++d) acc = fma_f32x2(nz[d], ev[d], acc);
    return acc;
}

template <int D>
__device__ __forceinline__ void load_pair(const float* p, unsigned long long (&ev)[D]) {
    if constexpr (D % 2 == 0) {
#pragma unroll
        for (int d = 0; d < D; d += 2) {
            const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(p + 2 * d);
            ev[d] = v.x;
            ev[d + 1] = v.y;
        }
    } else {
#pragma unroll
        for (int d = 0; d < D; ++d) ev[d] = *reinterpret_cast<const unsigned long long*>(p + 2 * d);
    }
}

// barrier over the 256 threads that run the CUDA-core search: the whole CTA (BAR = 0), or a named barrier when the
// kernel shares its CTA with the tensor-core role (search_dual_kernel, vqb_search_tclow.cu)
template <int BAR, int THREADS>
__device__ __forceinline__ void lowd_sync() {
    if constexpr (BAR == 0)
        __syncthreads();
    else
        asm volatile("bar.sync %0, %1;" ::"n"(BAR), "n"(THREADS) : "memory");
}

// T tokens per thread: tokens seg_base + t*kThreads + tid
template <int D, int V, int T, bool kList, int BAR = 0>
__device__ __forceinline__ void lowd_segment(LowDCtx<D, V>& c, int64_t seg_base) {
    using Cfg = LowDCfg<D, V>;
    constexpr int kLowDThreads = Cfg::kThreads;
    const int tid = threadIdx.x;
    const int lane = tid & 31;

    unsigned long long nz[T][D];
    int64_t tok[T];
#pragma unroll
    for (int t = 0; t < T; ++t) {
        const int64_t item = seg_base + (int64_t)t * kLowDThreads + tid;
        const bool ok = item < c.N;
        tok[t] = ok ? (kList ? (int64_t)__ldg(c.list + item) : item) : -1;
        const int64_t b = ok ? tok[t] / c.HW : 0;
        const int64_t hw = ok ? tok[t] - b * c.HW : 0;
        const float* zp = c.z + (b * D) * c.HW + hw;
#pragma unroll
        for (int d = 0; d < D; ++d) {
            const float v = ok ? -__ldg(zp + (int64_t)d * c.HW) : 0.f;
            nz[t][d] = pack_f32x2(v, v);
        }
    }

    float m[T], mprev[T];
    int cid[T];
    bool any_tok = false;
#pragma unroll
    for (int t = 0; t < T; ++t) {
        m[t] = INFINITY;
        mprev[t] = INFINITY;
        cid[t] = 0;
        any_tok |= tok[t] >= 0;
    }
    // a warp without tokens (the tail of a range, most warps of a list-mode CTA) only keeps the barriers company
    const bool warp_live = __ballot_sync(0xffffffffu, any_tok) != 0;

    for (int tile = 0; tile < c.n_tiles; ++tile) {
        const uint32_t v = c.visit;
        const float* buf = c.smem_tiles + (v & 1) * Cfg::kTileFloats;
        mbar_wait(c.bars + (v & 1), (v >> 1) & 1);
        const int chunks = warp_live ? tile_codes<D, V>(c, tile) / kChunkCodes : 0;
        const float* hbuf = buf + Cfg::kTileCodes * D;
        for (int ch = 0; ch < chunks; ++ch) {
            const float* ep = buf + (size_t)ch * kChunkCodes * D;
            const float* hp = hbuf + ch * kChunkCodes;
#pragma unroll 4
            for (int p = 0; p < kChunkCodes / 2; ++p) {
                unsigned long long ev[D];
                load_pair<D>(ep + p * 2 * D, ev);
                const unsigned long long h2 = *reinterpret_cast<const unsigned long long*>(hp + 2 * p);
#pragma unroll
                for (int t = 0; t < T; ++t) {
                    float x, y;
                    unpack_f32x2(pair_score<D>(nz[t], ev, h2), x, y);
                    m[t] = min3_f32(m[t], x, y);
                }
            }
            const int chunk_id = tile * (Cfg::kTileCodes / kChunkCodes) + ch;
#pragma unroll
            for (int t = 0; t < T; ++t) {
                cid[t] = (m[t] < mprev[t]) ? chunk_id : cid[t];
                mprev[t] = m[t];
            }
        }
        lowd_sync<BAR, kLowDThreads>();  // every warp is done with this buffer
        c.visit = v + 1;
        if (tid == 0 && v + 2 < c.total_visits) issue_visit<D, V>(c, v + 2);
    }

    if (!warp_live) return;
    // ---- resolve the index: the warp re-scores chunk cid[t] of each token ----
    int best[T];
#pragma unroll
    for (int t = 0; t < T; ++t) {
        int mine = 0;
        for (int owner = 0; owner < 32; ++owner) {
            const float ms = __shfl_sync(0xffffffffu, m[t], owner);
            const int cs = __shfl_sync(0xffffffffu, cid[t], owner);
            unsigned long long nzo[D];
#pragma unroll
            for (int d = 0; d < D; ++d) {
                float lo, hi;
                unpack_f32x2(nz[t][d], lo, hi);
                const float s = __shfl_sync(0xffffffffu, lo, owner);
                nzo[d] = pack_f32x2(s, s);
            }
            int res = 0;  // all-NaN row: ATen's argmin returns 0
            bool found = false;
#pragma unroll
            for (int sub = 0; sub < kChunkCodes / 64; ++sub) {
                const int pair = cs * (kChunkCodes / 2) + sub * 32 + lane;
                unsigned long long ev[D];
                load_pair<D>(c.g_pairs + (size_t)pair * 2 * D, ev);
                const unsigned long long h2 =
                    *reinterpret_cast<const unsigned long long*>(c.g_half_norm + 2 * (size_t)pair);
                float x, y;
                unpack_f32x2(pair_score<D>(nzo, ev, h2), x, y);
                const bool hx = (x == ms), hy = (y == ms);
                const unsigned hit = __ballot_sync(0xffffffffu, hx || hy);
                const int cand = 2 * pair + (hx ? 0 : 1);
                const int first = __shfl_sync(0xffffffffu, cand, hit ? (__ffs(hit) - 1) : 0);
                if (!found && hit) {  // warp-uniform: the first round with a hit holds the lowest index
                    res = first;
                    found = true;
                }
            }
            if (lane == owner) mine = res;
        }
        best[t] = mine;
    }

#pragma unroll
    for (int t = 0; t < T; ++t) {
        if (tok[t] >= 0) {
            int r = best[t];
            float dm = m[t];
            if (c.first_nan < c.K && m[t] != INFINITY) {  // NaN code is minimal, and so is its score (sharded MIN keys)
                r = c.first_nan;
                dm = __int_as_float(0x7fc00000);
            } else if (c.first_nan < c.K) {
                r = 0;
            }
            c.idx_out[tok[t]] = r;
            if (c.dmin_out) c.dmin_out[tok[t]] = dm;
        }
    }
}

// The whole CTA program of the CUDA-core search (256 threads, threadIdx.x in [0, 256)): the body of
// search_lowd_kernel, and the "FMA role" of search_dual_kernel.  `tokens_per_cta` tokens starting at cta_index * that.
template <int D, int V, bool kList, int BAR>
__device__ __forceinline__ void lowd_cta_body(const float* __restrict__ z, int64_t N, int64_t HW, int K,
                                              const unsigned char* __restrict__ pack, const PackLayout& L,
                                              int64_t tokens_per_cta, const int32_t* __restrict__ list,
                                              const int32_t* __restrict__ list_count, int64_t* __restrict__ idx_out,
                                              float* __restrict__ dmin_out, unsigned char* smem_raw, int cta_index,
                                              int n_ctas) {

    using Cfg = LowDCfg<D, V>;
    constexpr int kLowDThreads = Cfg::kThreads;
    LowDCtx<D, V> c;
    c.z = z;
    c.HW = HW;
    c.list = list;
    if constexpr (kList) {  // the number of flagged tokens is only known on the device
        N = *list_count;
        // a list is short (a few thousand unsure tokens): spread it over ALL CTAs in whole warps instead of filling
        // a dozen CTAs with 256 tokens each (measured 195 us per call at C2 before, latency of 12 busy SMs)
        tokens_per_cta = (N + n_ctas - 1) / n_ctas;
        tokens_per_cta = (tokens_per_cta + 31) / 32 * 32;
    }
    c.N = N;
    c.g_pairs = reinterpret_cast<const float*>(pack + L.off_pairs);
    c.g_half_norm = reinterpret_cast<const float*>(pack + L.off_half_norm);
    c.codes_padded = round_up_i(K, kChunkCodes);
    c.n_tiles = (c.codes_padded + Cfg::kTileCodes - 1) / Cfg::kTileCodes;
    c.K = K;
    c.first_nan = reinterpret_cast<const int*>(pack)[0];
    c.idx_out = idx_out;
    c.dmin_out = dmin_out;
    c.smem_tiles = reinterpret_cast<float*>(smem_raw);
    c.bars = reinterpret_cast<uint64_t*>(smem_raw + 2 * sizeof(float) * Cfg::kTileFloats);
    c.visit = 0;

    const int64_t start = (int64_t)cta_index * tokens_per_cta;
    int64_t end = start + tokens_per_cta;
    if (end > N) end = N;
    if (start >= end) return;
    int q = (int)((end - start + kLowDThreads - 1) / kLowDThreads);  // tokens per thread
    constexpr int TM = Cfg::kTmax;
    int n_seg = q / TM + __popc(q % TM);
    c.total_visits = (uint32_t)n_seg * c.n_tiles;

    if (threadIdx.x == 0) {
        mbar_init(c.bars + 0, 1);
        mbar_init(c.bars + 1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    lowd_sync<BAR, kLowDThreads>();
    if (threadIdx.x == 0) {
        issue_visit<D, V>(c, 0);
        if (c.total_visits > 1) issue_visit<D, V>(c, 1);
    }

    int64_t base = start;
    while (q >= TM) {
        lowd_segment<D, V, TM, kList, BAR>(c, base);
        base += (int64_t)TM * kLowDThreads;
        q -= TM;
    }
    if constexpr (TM >= 8) {
        if (q & 4) {
            lowd_segment<D, V, 4, kList, BAR>(c, base);
            base += 4 * kLowDThreads;
        }
    }
    if constexpr (TM >= 4) {
        if (q & 2) {
            lowd_segment<D, V, 2, kList, BAR>(c, base);
            base += 2 * kLowDThreads;
        }
    }
    if (q & 1) lowd_segment<D, V, 1, kList, BAR>(c, base);
}

}  // namespace vqb
