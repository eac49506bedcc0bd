// Nearest-code search for low-dimensional latents (D <= 16), CUDA-core FMA bound.
// Replaces the [N,K] distance matrix + argmin of the reference
// (vqgan_ldm_baseline/models/quantizer.py:68-76).
//
// Design (DESIGN.md "search, low-D"):
//  * persistent CTAs (two per SM, 256 threads); every thread owns T tokens whose
//    negated coordinates sit in registers;
//  * the codebook streams through shared memory in tiles (1-D bulk-async copies
//    completing on an mbarrier, double buffered) in a pair-interleaved layout so
//    that one FFMA2 evaluates one coordinate of TWO codes for one token:
//        acc.xy = (-z_d, -z_d) * (e_d(c0), e_d(c1)) + acc.xy,  acc seeded with 0.5|e|^2
//    -> D FFMA2 per (token, code pair);
//  * the running minimum is a single FMNMX3 per (token, code pair); the index is
//    NOT tracked in the hot loop.  Per 64-code chunk one compare records the
//    chunk in which the minimum last improved; after the sweep the warp
//    re-evaluates that one chunk cooperatively (bit-identical FMA chain) and
//    takes the first code whose score equals the minimum -> lowest index on ties.
//  * token ranges are split evenly over the SMs and each thread's token count is
//    decomposed as 8+8+...+4+2+1 so no SM idles in a partial wave.
#include "vqb_lowd_body.cuh"

namespace vqb {

VQB_KNOB g_lowd_variant = 0;
VQB_KNOB g_lowd_ctas_per_sm = 0;  // 0 = the variant's own residency; 1 leaves room for a co-running kernel
#ifdef VQB_EXPERIMENTAL
void set_lowd_variant(int v) {
    if (v >= 16) g_lowd_ctas_per_sm = v - 16; else g_lowd_variant = v;
}
#endif

template <int D, int V, bool kList>
__global__ void __launch_bounds__(LowDCfg<D, V>::kThreads, LowDCfg<D, V>::kMinBlocks)
    search_lowd_kernel(const float* __restrict__ z, int64_t N, int64_t HW, int K,
                       const unsigned char* __restrict__ pack, PackLayout L, int64_t tokens_per_cta,
                       const int32_t* __restrict__ list, const int32_t* __restrict__ list_count,
                       int64_t* __restrict__ idx_out, float* __restrict__ dmin_out) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    lowd_cta_body<D, V, kList, 0>(z, N, HW, K, pack, L, tokens_per_cta, list, list_count, idx_out, dmin_out, smem_raw,
                                  (int)blockIdx.x, (int)gridDim.x);
}

template <int D, int V>
static int launch_lowd_t(const float* z, int64_t N, int64_t HW, int K, const void* pack,
                         int64_t* idx_out, float* dmin_out, cudaStream_t s, const int32_t* list = nullptr,
                         const int32_t* list_count = nullptr, int ctas_per_sm = 0) {
    using Cfg = LowDCfg<D, V>;
    constexpr int kLowDThreads = Cfg::kThreads;
    auto kernel = list ? search_lowd_kernel<D, V, true> : search_lowd_kernel<D, V, false>;
    // the largest shared-memory carve-out, whatever this kernel needs by itself: an SM only runs kernels that agree on
    // the L1 / shared split, and the two-engine search (vqb_search_dual_f32) co-schedules this kernel with the tensor
    // kernel (measured: with the default carve-outs the second kernel waits for the first to drain)
    VQB_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    const PackLayout L = pack_layout(K, D);
    // persistent: one wave of resident CTAs
    if (ctas_per_sm <= 0) ctas_per_sm = g_lowd_ctas_per_sm;
    // one CTA per SM on purpose (two-engine search): ask for more than half of the SM's shared memory so that the
    // hardware cannot stack two of these CTAs on one SM and leave others empty -- the co-running tensor kernel needs
    // a slot on EVERY SM (113 + 111 KB + bookkeeping = the whole 228 KB)
    const size_t smem_bytes = ctas_per_sm == 1 && Cfg::kSmemBytes < (size_t)113 * 1024 ? (size_t)113 * 1024 : Cfg::kSmemBytes;
    VQB_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
    const int slots = sm_count() * ((ctas_per_sm > 0 && ctas_per_sm < Cfg::kMinBlocks) ? ctas_per_sm : Cfg::kMinBlocks);
    int64_t per_cta = (N + slots - 1) / slots;
    per_cta = (per_cta + kLowDThreads - 1) / kLowDThreads * kLowDThreads;
    const int grid = list ? slots : (int)((N + per_cta - 1) / per_cta);
    kernel<<<grid, kLowDThreads, smem_bytes, s>>>(
        z, N, HW, K, static_cast<const unsigned char*>(pack), L, per_cta, list, list_count, idx_out, dmin_out);
    VQB_LAUNCH_CHECK("search_lowd_kernel");
    return VQB_OK;
}

int launch_search_lowd(const float* z, int64_t B, int D, int64_t HW, int K, const void* pack,
                       int64_t* idx_out, float* dmin_out, cudaStream_t s, int ctas_per_sm) {
    const int64_t N = B * HW;
    if (D == 4 && g_lowd_variant != 0) {  // tuning variants are only instantiated for the headline D
        switch (g_lowd_variant) {
            case 1: return launch_lowd_t<4, 1>(z, N, HW, K, pack, idx_out, dmin_out, s);
            case 2: return launch_lowd_t<4, 2>(z, N, HW, K, pack, idx_out, dmin_out, s);
            case 3: return launch_lowd_t<4, 3>(z, N, HW, K, pack, idx_out, dmin_out, s);
            case 4: return launch_lowd_t<4, 4>(z, N, HW, K, pack, idx_out, dmin_out, s);
            default: break;
        }
    }
    switch (D) {
#define VQB_CASE(d) \
    case d:         \
        return launch_lowd_t<d, 0>(z, N, HW, K, pack, idx_out, dmin_out, s, nullptr, nullptr, ctas_per_sm);
        VQB_CASE(1) VQB_CASE(2) VQB_CASE(3) VQB_CASE(4) VQB_CASE(5) VQB_CASE(6) VQB_CASE(7) VQB_CASE(8)
        VQB_CASE(9) VQB_CASE(10) VQB_CASE(11) VQB_CASE(12) VQB_CASE(13) VQB_CASE(14) VQB_CASE(15)
        VQB_CASE(16)
#undef VQB_CASE
        default:
            set_error("low-D search supports 1 <= D <= 16, got %d", D);
            return VQB_ERR_UNSUPPORTED;
    }
}

// ---------------------------------------------------------------------------
// exact search of the tokens in a device-side list (the unsure tokens of the tensor paths): a few thousand tokens at
// most, so the sweep-with-tokens-in-registers kernel above is the wrong shape (one warp needs ~75 us to walk 16384
// codes for its 32 tokens whatever the list length: 207 us per call at C2).  Here a CTA takes 32 listed tokens
// (lane = token) and its 8 warps split the CODEBOOK: warp w scores the 64-code chunks w, w+8, ... straight from
// L2 (every lane reads the same code pair: one broadcast transaction), with the same FFMA2 chain, the same
// "chunk where the minimum first appeared" bookkeeping and the same cooperative chunk re-score as the main kernel,
// so indices and minimum scores stay bit-identical to VQB_ALGO_LOWD_FMA.
// ---------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(256)
    search_lowd_list_kernel(const float* __restrict__ z, int64_t HW, int K, const unsigned char* __restrict__ pack,
                            PackLayout L, const int32_t* __restrict__ list, const int32_t* __restrict__ list_count,
                            int64_t* __restrict__ idx_out, float* __restrict__ dmin_out) {
    __shared__ float sm_m[8][32];
    __shared__ int sm_cid[8][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int count = *list_count;
    const int n_groups = (count + 31) / 32;
    const float* g_pairs = reinterpret_cast<const float*>(pack + L.off_pairs);
    const float* g_half_norm = reinterpret_cast<const float*>(pack + L.off_half_norm);
    const int n_chunks = round_up_i(K, kChunkCodes) / kChunkCodes;
    const int first_nan = reinterpret_cast<const int*>(pack)[0];
    for (int g = blockIdx.x; g < n_groups; g += gridDim.x) {
        const int item = g * 32 + lane;
        const bool ok = item < count;
        const int64_t tok = ok ? (int64_t)__ldg(list + item) : -1;
        unsigned long long nz[D];
        {
            const int64_t b = ok ? tok / HW : 0;
            const float* zp = z + (b * D) * HW + (ok ? tok - b * HW : 0);
#pragma unroll
            for (int d = 0; d < D; ++d) {
                const float v = ok ? -__ldg(zp + (int64_t)d * HW) : 0.f;
                nz[d] = pack_f32x2(v, v);
            }
        }
        float m = INFINITY, mprev = INFINITY;
        int cid = 0;
        for (int ch = warp; ch < n_chunks; ch += 8) {
            const float* ep = g_pairs + (size_t)ch * kChunkCodes * D;
            const float* hp = g_half_norm + (size_t)ch * kChunkCodes;
#pragma unroll 8
            for (int p = 0; p < kChunkCodes / 2; ++p) {
                unsigned long long ev[D];
                load_pair<D>(ep + p * 2 * D, ev);
                const unsigned long long h2 = *reinterpret_cast<const unsigned long long*>(hp + 2 * p);
                float x, y;
                unpack_f32x2(pair_score<D>(nz, ev, h2), x, y);
                m = min3_f32(m, x, y);
            }
            cid = (m < mprev) ? ch : cid;
            mprev = m;
        }
        sm_m[warp][lane] = m;
        sm_cid[warp][lane] = cid;
        __syncthreads();
        if (warp == 0) {
            // minimum over the warps; equal minima -> the lowest chunk (= first occurrence in code order)
            float bm = sm_m[0][lane];
            int bc = sm_cid[0][lane];
#pragma unroll
            for (int w = 1; w < 8; ++w) {
                const float om = sm_m[w][lane];
                const int oc = sm_cid[w][lane];
                if (om < bm || (om == bm && oc < bc)) {
                    bm = om;
                    bc = oc;
                }
            }
            // the warp re-scores chunk bc of each token and takes the first code that equals the minimum
            int mine = 0;
            for (int owner = 0; owner < 32; ++owner) {
                const float ms = __shfl_sync(0xffffffffu, bm, owner);
                const int cs = __shfl_sync(0xffffffffu, bc, owner);
                unsigned long long nzo[D];
#pragma unroll
                for (int d = 0; d < D; ++d) {
                    float lo, hi;
                    unpack_f32x2(nz[d], lo, hi);
                    const float sv = __shfl_sync(0xffffffffu, lo, owner);
                    nzo[d] = pack_f32x2(sv, sv);
                }
                int res = 0;  // all-NaN row: ATen's argmin returns 0
                bool found = false;
#pragma unroll
                for (int sub = 0; sub < kChunkCodes / 64; ++sub) {
                    const int pair = cs * (kChunkCodes / 2) + sub * 32 + lane;
                    unsigned long long ev[D];
                    load_pair<D>(g_pairs + (size_t)pair * 2 * D, ev);
                    const unsigned long long h2 = *reinterpret_cast<const unsigned long long*>(g_half_norm + 2 * (size_t)pair);
                    float x, y;
                    unpack_f32x2(pair_score<D>(nzo, ev, h2), x, y);
                    const bool hx = (x == ms), hy = (y == ms);
                    const unsigned hit = __ballot_sync(0xffffffffu, hx || hy);
                    const int cand = 2 * pair + (hx ? 0 : 1);
                    const int first = __shfl_sync(0xffffffffu, cand, hit ? (__ffs(hit) - 1) : 0);
                    if (!found && hit) {
                        res = first;
                        found = true;
                    }
                }
                if (lane == owner) mine = res;
            }
            if (ok) {
                int r = mine;
                float dm = bm;
                if (first_nan < K && bm != INFINITY) {  // NaN code is minimal, and so is its score
                    r = first_nan;
                    dm = __int_as_float(0x7fc00000);
                } else if (first_nan < K) {
                    r = 0;
                }
                idx_out[tok] = r;
                if (dmin_out) dmin_out[tok] = dm;
            }
        }
        __syncthreads();  // the exchange buffers are reused by the next group
    }
}

template <int D>
static int launch_lowd_list_t(const float* z, int64_t HW, int K, const void* pack, const int32_t* list,
                              const int32_t* list_count, int64_t* idx_out, float* dmin_out, cudaStream_t s) {
    const PackLayout L = pack_layout(K, D);
    search_lowd_list_kernel<D><<<sm_count() * 4, 256, 0, s>>>(z, HW, K, static_cast<const unsigned char*>(pack), L, list,
                                                             list_count, idx_out, dmin_out);
    VQB_LAUNCH_CHECK("search_lowd_list_kernel");
    return VQB_OK;
}

// exact search of the tokens in a device-side list (the unsure tokens of the tensor path)
int launch_search_lowd_list(const float* z, int64_t B, int D, int64_t HW, int K, const void* pack,
                            const int32_t* list, const int32_t* list_count, int64_t* idx_out, float* dmin_out,
                            cudaStream_t s) {
    (void)B;
    switch (D) {
#define VQB_CASE(d) \
    case d:         \
        return launch_lowd_list_t<d>(z, HW, K, pack, list, list_count, idx_out, dmin_out, s);
        VQB_CASE(1) VQB_CASE(2) VQB_CASE(3) VQB_CASE(4) VQB_CASE(5) VQB_CASE(6) VQB_CASE(7) VQB_CASE(8)
        VQB_CASE(9) VQB_CASE(10) VQB_CASE(11) VQB_CASE(12) VQB_CASE(13) VQB_CASE(14) VQB_CASE(15)
        VQB_CASE(16)
#undef VQB_CASE
        default:
            set_error("low-D search supports 1 <= D <= 16, got %d", D);
            return VQB_ERR_UNSUPPORTED;
    }
}

}  // namespace vqb
