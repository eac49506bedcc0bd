// Codebook pre-pass: half norms, NaN scan, pair-interleaved fp32 rows (D <= 16)
// and the bf16 hi/lo split for the tensor path.  Replaces
// `torch.sum(self.embedding.weight ** 2, dim=1)` (quantizer.py:70 of the
// reference).  HBM-bound and tiny: reads K*D*4 bytes once.
#include "vqb_common.cuh"

namespace vqb {

__global__ void prepare_header_kernel(int* header, int K, int D) {
    if (threadIdx.x == 0) {
        header[0] = K;  // first NaN code (atomicMin below)
        header[1] = K;
        header[2] = D;
        header[4] = 0;  // bits of max_k 0.5|e_k|^2 (atomicMax below; non-negative floats order as ints)
    }
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
    const bool live = k < K;

    float sq = 0.f;
    bool bad = false;
    for (int d = lane; d < D; d += 32) {
        const float v = live ? E[(size_t)k * D + d] : 0.f;
        sq = fmaf(v, v, sq);
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
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    bad = __any_sync(0xffffffffu, bad);
    if (lane == 0) {
        half_norm[k] = live ? 0.5f * sq : INFINITY;
        if (live && (bad || sq != sq)) atomicMin(&header[0], k);
        if (live && sq == sq) atomicMax(&header[4], __float_as_int(0.5f * sq));
    }
}

int launch_codebook_prepare(const float* E, int K, int D, void* pack, cudaStream_t s) {
    const PackLayout L = pack_layout(K, D);
    prepare_header_kernel<<<1, 32, 0, s>>>(reinterpret_cast<int*>(pack), K, D);
    VQB_LAUNCH_CHECK("prepare_header_kernel");
    const int warps = 8;
    const int blocks = (L.Kpad + warps - 1) / warps;
    codebook_prepare_kernel<<<blocks, warps * 32, 0, s>>>(E, K, D, static_cast<unsigned char*>(pack), L);
    VQB_LAUNCH_CHECK("codebook_prepare_kernel");
    return VQB_OK;
}

}  // namespace vqb
