// Codebook pre-pass: half norms, NaN scan, pair-interleaved fp32 rows (D <= 16)
// and the bf16 hi/lo split for the tensor path.  Replaces
// `torch.sum(self.embedding.weight ** 2, dim=1)` (quantizer.py:70 of the
// reference).  HBM-bound and tiny: reads K*D*4 bytes once.
#include "vqb_common.cuh"

namespace vqb {

// header + everything the later kernels accumulate into with atomics (one launch instead of a kernel and three memsets):
// the per-group / per-chunk norm maxima (zeros) and the duplicate-detection table (all ones = empty)
__global__ void __launch_bounds__(256) prepare_header_kernel(int* header, int K, int D, unsigned char* pack, PackLayout L) {
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
    if (L.has_e16)
        for (size_t i = tid; i < (size_t)L.Kpad / 4; i += nth) reinterpret_cast<float*>(pack + L.off_gmax)[i] = 0.f;
    if (L.has_pairs)
        for (size_t i = tid; i < (size_t)L.Kpad / 32; i += nth) reinterpret_cast<float*>(pack + L.off_cmax)[i] = 0.f;
    if (L.has_e16)
        for (size_t i = tid; i < (size_t)L.dup_slots / 2; i += nth)
            reinterpret_cast<ulonglong2*>(pack + L.off_dup)[i] = make_ulonglong2(~0ull, ~0ull);
    if (tid == 0) {
        header[0] = K;  // first NaN code (atomicMin below)
        header[1] = K;
        header[2] = D;
        header[4] = 0;  // bits of max_k 0.5|e_k|^2 (atomicMax below; non-negative floats order as ints)
        header[5] = 0;  // fp16 scale exponent, written by codebook_prepare_kernel
        header[6] = 0;  // bits of max |E| (codebook_absmax_kernel)
        header[7] = 0;  // bits of max_k |e_k - fp16 image of e_k| (residual of the single-pass tensor path)
        header[8] = 0;  // bits of max_k |e_k - eh_k - el_k| (tf32x3 image)
        header[9] = 0;  // bits of max_k |el_k|
        for (int i = 10; i < 16; ++i) header[i] = 0;  // norm-dependent residual bounds (vqb_common.cuh)
    }
}

__global__ void __launch_bounds__(256) codebook_absmax_kernel(const float* __restrict__ E, size_t n, int* header) {
    float m = 0.f;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float v = fabsf(E[i]);
        if (v == v && v < INFINITY) m = fmaxf(m, v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(&header[6], __float_as_int(m));
}

// power-of-two scale that puts max|E| into [512, 1024): keeps small entries out of the fp16
// subnormal range and far from overflow; exact (no rounding) because it is a power of two
__device__ __forceinline__ int fp16_scale_exponent(float maxabs) {
    if (!(maxabs > 0.f) || maxabs == INFINITY) return 0;
    int e = 9 - ilogbf(maxabs);
    return e < -100 ? -100 : (e > 100 ? 100 : e);
}

// one warp per code row (padded rows included)
__global__ void __launch_bounds__(256) codebook_prepare_kernel(const float* __restrict__ E, int K,
                                                               int D, unsigned char* pack,
                                                               PackLayout L) {
    const int warps_per_block = blockDim.x >> 5;
    const int k = blockIdx.x * warps_per_block + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (k >= L.Kpad) return;
    int* header = reinterpret_cast<int*>(pack);
    float* half_norm = reinterpret_cast<float*>(pack + L.off_half_norm);
    float* pairs = reinterpret_cast<float*>(pack + L.off_pairs);
    __nv_bfloat16* ehi = reinterpret_cast<__nv_bfloat16*>(pack + L.off_ehi);
    __nv_bfloat16* elo = reinterpret_cast<__nv_bfloat16*>(pack + L.off_elo);
    __half* e16 = reinterpret_cast<__half*>(pack + L.off_e16);
    const bool live = k < K;
    const int se = fp16_scale_exponent(__int_as_float(header[6]));
    if (k == 0 && lane == 0) header[5] = se;

    float sq = 0.f, res = 0.f;
    bool bad = false;
    uint32_t hsum = 0u;  // order-independent row hash: sum over d of mix(bits(v_d), d)
    for (int d = lane; d < D; d += 32) {
        const float v = live ? E[(size_t)k * D + d] : 0.f;
        sq = fmaf(v, v, sq);
        if (L.has_e16) {
            uint32_t x = __float_as_uint(v) ^ ((uint32_t)d * 0x9E3779B9u);
            x *= 0x85EBCA6Bu;
            x ^= x >> 13;
            x *= 0xC2B2AE35u;
            x ^= x >> 16;
            hsum += x;
        }
        bad |= (v != v);
        if (L.has_pairs) {
            // pair p = k/2 holds e_d(2p), e_d(2p+1) adjacent for every d
            pairs[(size_t)(k >> 1) * (2 * D) + 2 * d + (k & 1)] = v;
        }
        if (L.has_bf16) {
            const __nv_bfloat16 hi = __float2bfloat16_rn(v);
            const float rem = v - __bfloat162float(hi);
            ehi[(size_t)k * D + d] = hi;
            elo[(size_t)k * D + d] = __float2bfloat16_rn(rem);
        }
        if (L.has_e16) {
            const __half hv = __float2half_rn(ldexpf(v, se));
            e16[(size_t)k * L.Dpad + d] = hv;
            const float dv = v - ldexpf(__half2float(hv), -se);
            res = fmaf(dv, dv, res);
        }
    }
    if (L.has_e16)  // zero padding up to whole 64-channel blocks
        for (int d = D + lane; d < L.Dpad; d += 32) e16[(size_t)k * L.Dpad + d] = __float2half_rn(0.f);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) res += __shfl_xor_sync(0xffffffffu, res, o);
    bad = __any_sync(0xffffffffu, bad);
    if (L.has_e16) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) hsum += __shfl_xor_sync(0xffffffffu, hsum, o);
    }
    if (lane == 0) {
        if (L.has_e16) {
            reinterpret_cast<uint32_t*>(pack + L.off_rowhash)[k] = hsum;
            if (live) {
                // first-fit open addressing, no deletions: every row with this hash ends up in the same slot, which keeps
                // the LOWEST code index among them (hash in the high word, so a 64-bit min orders equal hashes by index)
                unsigned long long* table = reinterpret_cast<unsigned long long*>(pack + L.off_dup);
                const unsigned long long key = ((unsigned long long)hsum << 32) | (unsigned)k;
                for (int pr = 0; pr < 64; ++pr) {
                    unsigned long long* slot = table + ((hsum + (uint32_t)pr) & (uint32_t)(L.dup_slots - 1));
                    const unsigned long long old = atomicCAS(slot, ~0ull, key);
                    if (old == ~0ull) break;
                    if ((uint32_t)(old >> 32) == hsum) {
                        atomicMin(slot, key);
                        break;
                    }
                }
            }
        }
        half_norm[k] = live ? 0.5f * sq : INFINITY;
        if (L.has_e16) {
            const float h = 0.5f * sq;
            reinterpret_cast<float*>(pack + L.off_half_norm_fin)[k] = (live && h < 1e38f) ? h : 1e38f;
        }
        if (live && (bad || sq != sq)) atomicMin(&header[0], k);
        if (live && sq == sq) atomicMax(&header[4], __float_as_int(0.5f * sq));
        if (live && res == res && res > 0.f) atomicMax(&header[7], __float_as_int(sqrtf(res)));
        if (L.has_e16 && live && sq == sq && sq < INFINITY) {
            const float n = sqrtf(sq);
            atomicMax(reinterpret_cast<int*>(pack + L.off_gmax) + (k >> 2), __float_as_int(n));
            if (res == res && res > 0.f) {
                const float theta = __int_as_float(header[6]) * (1.f / 4096.f);
                const float r = sqrtf(res);
                if (n >= theta && n > 0.f)
                    atomicMax(&header[10], __float_as_int(r / n * 1.0000002f));  // rounded up
                else
                    atomicMax(&header[11], __float_as_int(r));
            }
        }
    }
}

// Hides every code whose row is bit-identical to that of a lower-indexed code from the fp16 tensor pass (half_norm_fin
// -> the finite pad): such a code can never be the answer (lowest index wins ties), but FOUR or more copies of a popular
// code -- a codebook restarted by copying live codes onto dead ones -- would occupy all four candidate slots of the
// epilogue with the same score and send every token near it to the exact full search.  The exact tiers read half_norm,
// which is untouched.  One thread per code; the row comparison only runs for hash matches.
__device__ __forceinline__ bool code_is_shadowed(const float* __restrict__ E, int k, int D, const unsigned char* pack,
                                                 const PackLayout& L) {
    const uint32_t h = reinterpret_cast<const uint32_t*>(pack + L.off_rowhash)[k];
    const unsigned long long* table = reinterpret_cast<const unsigned long long*>(pack + L.off_dup);
    int k0 = k;
    for (int pr = 0; pr < 64; ++pr) {
        const unsigned long long e = table[(h + (uint32_t)pr) & (uint32_t)(L.dup_slots - 1)];
        if (e == ~0ull) break;
        if ((uint32_t)(e >> 32) == h) {
            k0 = (int)(uint32_t)(e & 0xffffffffull);
            break;
        }
    }
    if (k0 >= k) return false;
    const uint32_t* a = reinterpret_cast<const uint32_t*>(E + (size_t)k0 * D);
    const uint32_t* b = reinterpret_cast<const uint32_t*>(E + (size_t)k * D);
    for (int d = 0; d < D; ++d)
        if (a[d] != b[d]) return false;
    return true;
}

__global__ void __launch_bounds__(256) codebook_shadow_kernel(const float* __restrict__ E, int K, int D, unsigned char* pack,
                                                              PackLayout L) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    if (code_is_shadowed(E, k, D, pack, L)) reinterpret_cast<float*>(pack + L.off_half_norm_fin)[k] = 1e38f;
}

// tf32x3 image of the codebook for the low-D tensor path: one thread per (padded) code row.
// Runs after codebook_prepare_kernel (reads the half norms it wrote).
__global__ void __launch_bounds__(128) codebook_image_kernel(const float* __restrict__ E, int K, int D,
                                                             unsigned char* pack, PackLayout L) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= L.Kpad) return;
    int* header = reinterpret_cast<int*>(pack);
    const float* half_norm = reinterpret_cast<const float*>(pack + L.off_half_norm);
    float* img = reinterpret_cast<float*>(pack + L.off_img) + (size_t)(k / kLowRows) * tclow_tile_floats(D);
    const int r = k % kLowRows;
    const int slots = 8 * tclow_steps(D);
    const bool live = k < K;
    float res = 0.f, lo2 = 0.f;
    for (int d = 0; d < D; ++d) {
        const float v = live ? E[(size_t)k * D + d] : 0.f;
        const float eh = to_tf32(v);
        const float rem = v - eh;
        const float el = to_tf32(rem);
        const float r2 = rem - el;
        res = fmaf(r2, r2, res);
        lo2 = fmaf(el, el, lo2);
        img[tclow_slot_offset(r, d)] = -eh;
        img[tclow_slot_offset(r, D + d)] = -eh;
        img[tclow_slot_offset(r, 2 * D + d)] = -el;
    }
    // half norm in three tf32 pieces (33 significant bits >= fp32); padded rows score ~1e38
    const float h = live ? half_norm[k] : 1e38f;
    const float h1 = to_tf32(h);
    const float h2 = to_tf32(h - h1);
    const float h3 = to_tf32((h - h1) - h2);
    img[tclow_slot_offset(r, 3 * D + 0)] = h1;
    img[tclow_slot_offset(r, 3 * D + 1)] = h2;
    img[tclow_slot_offset(r, 3 * D + 2)] = h3;
    for (int sl = 3 * D + 3; sl < slots; ++sl) img[tclow_slot_offset(r, sl)] = 0.f;
    if (live && res == res && res > 0.f && res < INFINITY) atomicMax(&header[8], __float_as_int(sqrtf(res)));
    if (live && lo2 == lo2 && lo2 > 0.f && lo2 < INFINITY) atomicMax(&header[9], __float_as_int(sqrtf(lo2)));
    if (live && h == h && h < INFINITY) {
        const float n = sqrtf(2.f * h);
        atomicMax(reinterpret_cast<int*>(pack + L.off_cmax) + (k >> 5), __float_as_int(n));
        // theta from the half-norm maximum is not known yet (same kernel): max|E| is not computed for D <= 16,
        // so classify against the code's own scale: relative bound for every code with n > 0
        if (n > 0.f) {
            if (res == res && res < INFINITY) atomicMax(&header[12], __float_as_int(sqrtf(res) / n * 1.0000002f));
            if (lo2 == lo2 && lo2 < INFINITY) atomicMax(&header[14], __float_as_int(sqrtf(lo2) / n * 1.0000002f));
        }
    }
}

int launch_codebook_prepare(const float* E, int K, int D, void* pack, cudaStream_t s) {
    const PackLayout L = pack_layout(K, D);
    {
        const size_t words = L.has_e16 ? (size_t)L.dup_slots / 2 : 1;
        const unsigned hb = (unsigned)((words + 255) / 256 < 64 ? (words + 255) / 256 : 64);
        prepare_header_kernel<<<hb ? hb : 1, 256, 0, s>>>(reinterpret_cast<int*>(pack), K, D, static_cast<unsigned char*>(pack), L);
        VQB_LAUNCH_CHECK("prepare_header_kernel");
    }
    if (L.has_e16) {
        const size_t n = (size_t)K * D;
        size_t ab = (n + 255) / 256;
        if (ab > (size_t)sm_count() * 8) ab = (size_t)sm_count() * 8;
        codebook_absmax_kernel<<<(unsigned)ab, 256, 0, s>>>(E, n, reinterpret_cast<int*>(pack));
        VQB_LAUNCH_CHECK("codebook_absmax_kernel");
    }
    const int warps = 8;
    const int blocks = (L.Kpad + warps - 1) / warps;
    codebook_prepare_kernel<<<blocks, warps * 32, 0, s>>>(E, K, D, static_cast<unsigned char*>(pack), L);
    VQB_LAUNCH_CHECK("codebook_prepare_kernel");
    if (L.has_e16) {
        codebook_shadow_kernel<<<(K + 255) / 256, 256, 0, s>>>(E, K, D, static_cast<unsigned char*>(pack), L);
        VQB_LAUNCH_CHECK("codebook_shadow_kernel");
    }
    if (L.has_pairs) {
        codebook_image_kernel<<<(L.Kpad + 127) / 128, 128, 0, s>>>(E, K, D, static_cast<unsigned char*>(pack), L);
        VQB_LAUNCH_CHECK("codebook_image_kernel");
    }
    return VQB_OK;
}

}  // namespace vqb
