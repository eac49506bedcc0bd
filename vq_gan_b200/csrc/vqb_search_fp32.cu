// Nearest-code search for any embedding dimension on the CUDA cores (fp32).
// Replaces quantizer.py:68-76 of the reference for 16 < D < 64 or D not a
// multiple of 64, and is the exact fp32 re-score of the near-ties flagged by the
// bf16x3 tensor kernel (then it runs on a compacted token list whose length is
// only known on the device).
//
// Work item = (128-token tile, code split).  Register-tiled 128 tokens x 128
// codes, 8x8 scores per thread, the score tile never leaves registers: epilogue
// d = 0.5|e|^2 - z.e (z.e accumulated with one fmaf per dimension, d ascending)
// feeds a per-thread running (min, index), reduced over the 16 threads that
// share a token row.  With more than one code split the partial winners meet in
// a 64-bit atomicMin on (ordered score bits, index) keys -- lowest index wins
// ties -- and a finalize pass unpacks them.  CTAs are persistent and stride over
// the work items so a device-side token count needs no host sync.
#include "vqb_common.cuh"

namespace vqb {

constexpr int kFtTokens = 128;
constexpr int kFtCodes = 128;
constexpr int kFtDk = 16;
constexpr int kFtThreads = 256;
constexpr int kFtPad = 4;

__device__ __forceinline__ unsigned long long score_key(float d, int idx) {
    int bits = __float_as_int(d);
    if (bits < 0) bits ^= 0x7fffffff;                 // monotone signed order
    const unsigned int u = (unsigned int)bits ^ 0x80000000u;  // -> monotone unsigned order
    return ((unsigned long long)u << 32) | (unsigned int)idx;
}
__device__ __forceinline__ void key_unpack(unsigned long long key, float& d, int& idx) {
    idx = (int)(unsigned int)(key & 0xffffffffull);
    int bits = (int)((unsigned int)(key >> 32) ^ 0x80000000u);
    if (bits < 0) bits ^= 0x7fffffff;
    d = __int_as_float(bits);
}

// 4-byte asynchronous global -> shared copy; `valid` false writes a zero without touching `src`
__device__ __forceinline__ void cp_async4(float* smem_dst, const float* src, bool valid) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(src),
                 "r"(valid ? 4 : 0)
                 : "memory");
}

constexpr int kFtCompactCap = 4096;  // flagged tokens staged dim-major ([D][cap]) so every code split reads them coalesced

// list mode pre-pass: zc[d][r] = z[token list[r]][d] for r < count (skipped when the list is longer than the buffer)
__global__ void __launch_bounds__(256)
    gather_list_rows_kernel(const float* __restrict__ z, int D, int64_t HW, const int32_t* __restrict__ token_list,
                            const int32_t* __restrict__ list_count, float* __restrict__ zc) {
    const int count = *list_count;
    if (count > kFtCompactCap) return;
    const int r = blockIdx.x * 32 + (threadIdx.x & 31);
    if (r >= count) return;
    const int64_t t = token_list[r];
    const int64_t b = t / HW;
    const float* zp = z + (b * D) * HW + (t - b * HW);
    for (int d = threadIdx.x >> 5; d < D; d += 8) zc[(size_t)d * kFtCompactCap + r] = __ldg(zp + (int64_t)d * HW);
}

__global__ void __launch_bounds__(kFtThreads, 2)
    search_fp32_kernel(const float* __restrict__ z, int64_t N, int D, int64_t HW, const float* __restrict__ zc,
                       const float* __restrict__ E, int K, const unsigned char* __restrict__ pack,
                       PackLayout L, const int32_t* __restrict__ token_list,
                       const int32_t* __restrict__ list_count, int splits, int codes_per_split,
                       unsigned long long* __restrict__ keys, int64_t* __restrict__ idx_out,
                       float* __restrict__ dmin_out) {
    // double buffered: cp.async fills buffer b^1 with k-step k+1 while k-step k is multiplied out of buffer b
    __shared__ __align__(16) float As[2][kFtDk][kFtTokens + kFtPad];
    __shared__ __align__(16) float Bs[2][kFtDk][kFtCodes + kFtPad];
    __shared__ int64_t tok_off[kFtTokens];
    __shared__ int64_t tok_id[kFtTokens];

    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int64_t count = token_list ? (int64_t)(*list_count) : N;
    const int64_t n_tiles = (count + kFtTokens - 1) / kFtTokens;
    const int64_t n_items = n_tiles * splits;
    const float* half_norm = reinterpret_cast<const float*>(pack + L.off_half_norm);
    const int first_nan = reinterpret_cast<const int*>(pack)[0];
    // a short list was gathered into zc (dim-major, row = list position): read that instead of the
    // scattered tokens (which cost a 32-byte sector per 4-byte value, once per code split)
    const bool compact = token_list != nullptr && zc != nullptr && count <= kFtCompactCap;
    const float* zsrc = compact ? zc : z;
    const int64_t zstride = compact ? (int64_t)kFtCompactCap : HW;

    for (int64_t item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int64_t tile = item / splits;
        const int split = (int)(item - tile * splits);
        const int64_t m0 = tile * kFtTokens;
        const int n_begin = split * codes_per_split;
        int n_end = n_begin + codes_per_split;
        if (n_end > K) n_end = K;

        __syncthreads();  // previous item is done with tok_off / tok_id
        if (tid < kFtTokens) {
            const int64_t r = m0 + tid;
            int64_t t = -1;
            if (r < count) t = token_list ? (int64_t)token_list[r] : r;
            tok_id[tid] = t;
            if (t >= 0) {
                const int64_t b = t / HW;
                tok_off[tid] = compact ? r : (b * D) * HW + (t - b * HW);
            } else {
                tok_off[tid] = -1;
            }
        }
        __syncthreads();

        float best[8];
        int bidx[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            best[i] = INFINITY;
            bidx[i] = 0;
        }

        for (int n0 = n_begin; n0 < n_end; n0 += kFtCodes) {
            float acc[8][8];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

            // software pipeline (cp.async, 4-byte copies with zero fill, no register staging): the unpipelined loop
            // exposed one L2 round trip per 16 dimensions, which is what made the short token-list runs of the tensor
            // paths latency-bound (175 us for ~400 tokens) and held the dense kernel at 34 TFLOP/s
            constexpr int kRegA = (kFtDk * kFtTokens) / kFtThreads, kRegB = (kFtDk * kFtCodes) / kFtThreads;
            const int a_mm = tid % kFtTokens, a_k = tid / kFtTokens;   // row of A is thread-constant (256 % 128 == 0)
            const int64_t a_off = tok_off[a_mm];
            const float* pa = zsrc + (a_off >= 0 ? a_off : 0) + (int64_t)a_k * zstride;
            const int b_nn = tid / kFtDk, b_k = tid % kFtDk;
            const float* pb = E + (size_t)(n0 + b_nn < K ? n0 + b_nn : 0) * D + b_k;
            auto issue = [&](int buf, int k0) {
#pragma unroll
                for (int i = 0; i < kRegA; ++i) {
                    const int kk = a_k + i * (kFtThreads / kFtTokens);
                    const bool ok = a_off >= 0 && k0 + kk < D;
                    cp_async4(&As[buf][kk][a_mm], ok ? pa + (int64_t)(k0 + i * (kFtThreads / kFtTokens)) * zstride : zsrc, ok);
                }
#pragma unroll
                for (int i = 0; i < kRegB; ++i) {
                    const int nn = b_nn + i * (kFtThreads / kFtDk);
                    const bool ok = n0 + nn < K && k0 + b_k < D;
                    cp_async4(&Bs[buf][b_k][nn], ok ? pb + (size_t)i * (kFtThreads / kFtDk) * D + k0 : E, ok);
                }
                asm volatile("cp.async.commit_group;" ::: "memory");
            };
            int buf = 0;
            issue(0, 0);
            for (int k0 = 0; k0 < D; k0 += kFtDk) {
                if (k0 + kFtDk < D) {
                    issue(buf ^ 1, k0 + kFtDk);
                    asm volatile("cp.async.wait_group 1;" ::: "memory");
                } else {
                    asm volatile("cp.async.wait_group 0;" ::: "memory");
                }
                __syncthreads();
#pragma unroll
                for (int kk = 0; kk < kFtDk; ++kk) {
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
                __syncthreads();  // everyone is done with `buf` before the next iteration's copies overwrite it
                buf ^= 1;
            }
            // epilogue: codes visited in increasing index order per thread
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
                const float h = half_norm[n];  // +inf beyond K (Kpad is a multiple of 256)
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float d = h - acc[i][j];
                    if (d < best[i]) {
                        best[i] = d;
                        bidx[i] = n;
                    }
                }
            }
        }

        // reduce over the 16 threads (tx) sharing each token row
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
                const int64_t t = tok_id[row];
                if (t < 0) continue;
                if (keys != nullptr) {
                    atomicMin(keys + (m0 + row), score_key(best[i], bidx[i]));
                } else {
                    int r = bidx[i];
                    float dm = best[i];
                    if (first_nan < K) {
                        r = (best[i] == INFINITY) ? 0 : first_nan;
                        if (best[i] != INFINITY) dm = __int_as_float(0x7fc00000);  // the NaN code's score
                    }
                    idx_out[t] = r;
                    if (dmin_out) dmin_out[t] = dm;
                }
            }
        }
    }
}

__global__ void __launch_bounds__(256)
    fp32_finalize_kernel(const unsigned long long* __restrict__ keys, const int32_t* __restrict__ token_list,
                         const int32_t* __restrict__ list_count, int64_t N, int K,
                         const unsigned char* __restrict__ pack, int64_t* __restrict__ idx_out,
                         float* __restrict__ dmin_out) {
    const int64_t count = token_list ? (int64_t)(*list_count) : N;
    const int first_nan = reinterpret_cast<const int*>(pack)[0];
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < count;
         r += (int64_t)gridDim.x * blockDim.x) {
        const int64_t t = token_list ? (int64_t)token_list[r] : r;
        float d;
        int idx;
        key_unpack(keys[r], d, idx);
        if (first_nan < K) {
            idx = (d == INFINITY) ? 0 : first_nan;
            if (d != INFINITY) d = __int_as_float(0x7fc00000);  // the NaN code's score
        }
        idx_out[t] = idx;
        if (dmin_out) dmin_out[t] = d;
    }
}

// 8 bytes of key per row, then (D > 0) the compact buffer of the list mode
size_t search_fp32_workspace_bytes(int64_t n_rows, int D) {
    return round_up_z(8 * (size_t)n_rows + 1024, 1024) + (D > 0 ? sizeof(float) * (size_t)kFtCompactCap * D : 0);
}

// keys_ws: scratch of search_fp32_workspace_bytes(rows) bytes, needed when the codebook is split
int launch_search_fp32(const float* z, int64_t B, int D, int64_t HW, const float* E, int K,
                       const void* pack, const int32_t* token_list, const int32_t* list_count,
                       int64_t max_list, void* keys_ws, size_t keys_bytes, int64_t* idx_out,
                       float* dmin_out, cudaStream_t s) {
    const int64_t N = B * HW;
    const int64_t rows = token_list ? max_list : N;
    if (rows <= 0) return VQB_OK;
    const PackLayout L = pack_layout(K, D);
    const int sms = sm_count();
    const int64_t tiles = (rows + kFtTokens - 1) / kFtTokens;
    const int code_tiles = (K + kFtCodes - 1) / kFtCodes;

    // splits: a token list is short (re-score) -> spread each tile over the SMs; a dense run
    // only splits when there are too few token tiles to fill the chip twice
    int splits = 1;
    if (token_list) {
        splits = code_tiles < 128 ? code_tiles : 128;  // a few hundred tokens: one item per SM slot is latency-bound
    } else if (tiles < 2 * sms) {
        int64_t want = (2 * sms + tiles - 1) / tiles;
        splits = (int)(want < code_tiles ? want : code_tiles);
    }
    if (splits < 1) splits = 1;
    if (splits > 1 && (keys_ws == nullptr || keys_bytes < search_fp32_workspace_bytes(rows, 0))) splits = 1;
    float* zc = nullptr;
    if (token_list && keys_ws && keys_bytes >= search_fp32_workspace_bytes(rows, D)) {
        zc = reinterpret_cast<float*>(static_cast<unsigned char*>(keys_ws) + round_up_z(8 * (size_t)rows + 1024, 1024));
        gather_list_rows_kernel<<<kFtCompactCap / 32, 256, 0, s>>>(z, D, HW, token_list, list_count, zc);
        VQB_LAUNCH_CHECK("gather_list_rows_kernel");
    }
    const int tiles_per_split = (code_tiles + splits - 1) / splits;
    const int codes_per_split = tiles_per_split * kFtCodes;
    splits = (code_tiles + tiles_per_split - 1) / tiles_per_split;

    unsigned long long* keys = nullptr;
    if (splits > 1) {
        keys = static_cast<unsigned long long*>(keys_ws);
        VQB_CUDA_TRY(cudaMemsetAsync(keys, 0xff, 8 * (size_t)rows, s));
    }
    int64_t items = tiles * splits;
    const int64_t cap = (int64_t)sms * 2 * (token_list ? 1 : 64);  // persistent for lists
    const unsigned grid = (unsigned)(items < cap ? items : cap);
    search_fp32_kernel<<<grid, kFtThreads, 0, s>>>(z, N, D, HW, zc, E, K, static_cast<const unsigned char*>(pack),
                                                   L, token_list, list_count, splits, codes_per_split, keys,
                                                   idx_out, dmin_out);
    VQB_LAUNCH_CHECK("search_fp32_kernel");
    if (splits > 1) {
        int64_t fb = (rows + 255) / 256;
        if (fb > (int64_t)sms * 8) fb = (int64_t)sms * 8;
        fp32_finalize_kernel<<<(unsigned)fb, 256, 0, s>>>(keys, token_list, list_count, N, K,
                                                         static_cast<const unsigned char*>(pack), idx_out,
                                                         dmin_out);
        VQB_LAUNCH_CHECK("fp32_finalize_kernel");
    }
    return VQB_OK;
}

}  // namespace vqb
