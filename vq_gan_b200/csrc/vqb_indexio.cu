// Compact index maps: the quantizer's int64 indices
// (quantizer.py:101, consumed by VQVAE.encode_to_indices / decode_from_indices, vq_vae.py:162-190)
// narrowed to the smallest unsigned type that holds K-1, and widened back.  HBM-bound streaming
// kernels: 8 + w bytes per token (w = 1, 2 or 4), eight tokens per thread, 16-byte accesses.
#include "vqb_common.cuh"

namespace vqb {

template <typename T>
__global__ void __launch_bounds__(256)
    narrow_kernel(const int64_t* __restrict__ idx, int64_t n, int K, T* __restrict__ out, int* __restrict__ err_flag) {
    const int64_t base = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 8;
    if (base >= n) return;
    bool bad = false;
    if (base + 8 <= n && (reinterpret_cast<uintptr_t>(out + base) & (8 * sizeof(T) - 1)) == 0 &&
        (reinterpret_cast<uintptr_t>(idx + base) & 15) == 0) {
        int64_t v[8];
        const longlong2* src = reinterpret_cast<const longlong2*>(idx + base);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const longlong2 t = __ldg(src + i);
            v[2 * i] = t.x;
            v[2 * i + 1] = t.y;
        }
        T w[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            bad |= (v[i] < 0 || v[i] >= K);
            w[i] = (T)v[i];
        }
        if constexpr (sizeof(T) == 1) {
            *reinterpret_cast<uint2*>(out + base) = *reinterpret_cast<const uint2*>(w);
        } else if constexpr (sizeof(T) == 2) {
            *reinterpret_cast<uint4*>(out + base) = *reinterpret_cast<const uint4*>(w);
        } else {
            reinterpret_cast<uint4*>(out + base)[0] = reinterpret_cast<const uint4*>(w)[0];
            reinterpret_cast<uint4*>(out + base)[1] = reinterpret_cast<const uint4*>(w)[1];
        }
    } else {
        for (int64_t i = base; i < n && i < base + 8; ++i) {
            const int64_t v = idx[i];
            bad |= (v < 0 || v >= K);
            out[i] = (T)v;
        }
    }
    if (bad && err_flag) *err_flag = 1;
}

template <typename T>
__global__ void __launch_bounds__(256)
    widen_kernel(const T* __restrict__ codes, int64_t n, int64_t* __restrict__ out) {
    const int64_t base = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 8;
    if (base >= n) return;
    if (base + 8 <= n && (reinterpret_cast<uintptr_t>(codes + base) & (8 * sizeof(T) - 1)) == 0 &&
        (reinterpret_cast<uintptr_t>(out + base) & 15) == 0) {
        T w[8];
        if constexpr (sizeof(T) == 1) {
            *reinterpret_cast<uint2*>(w) = __ldg(reinterpret_cast<const uint2*>(codes + base));
        } else if constexpr (sizeof(T) == 2) {
            *reinterpret_cast<uint4*>(w) = __ldg(reinterpret_cast<const uint4*>(codes + base));
        } else {
            reinterpret_cast<uint4*>(w)[0] = __ldg(reinterpret_cast<const uint4*>(codes + base));
            reinterpret_cast<uint4*>(w)[1] = __ldg(reinterpret_cast<const uint4*>(codes + base) + 1);
        }
        longlong2* dst = reinterpret_cast<longlong2*>(out + base);
#pragma unroll
        for (int i = 0; i < 4; ++i) dst[i] = make_longlong2((long long)w[2 * i], (long long)w[2 * i + 1]);
    } else {
        for (int64_t i = base; i < n && i < base + 8; ++i) out[i] = (int64_t)codes[i];
    }
}

}  // namespace vqb

using namespace vqb;

extern "C" int vqb_index_bytes(int K) { return K <= 0 ? 0 : (K <= 256 ? 1 : (K <= 65536 ? 2 : 4)); }

extern "C" int vqb_indices_narrow(const int64_t* idx, int64_t n, int K, void* codes_out, int elem_bytes,
                                  int* err_flag, vqb_stream_t stream) {
    VQB_DEVICE_TRY();
    if (n < 0 || K <= 0 || (elem_bytes != 1 && elem_bytes != 2 && elem_bytes != 4) || elem_bytes < vqb_index_bytes(K)) {
        set_error("vqb_indices_narrow: invalid argument (n=%lld K=%d elem_bytes=%d)", (long long)n, K, elem_bytes);
        return VQB_ERR_INVALID_ARG;
    }
    if (n == 0) return VQB_OK;
    if (!idx || !codes_out) {
        set_error("vqb_indices_narrow: null pointer");
        return VQB_ERR_INVALID_ARG;
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const unsigned blocks = (unsigned)((n + 2047) / 2048);
    if (elem_bytes == 1)
        narrow_kernel<uint8_t><<<blocks, 256, 0, s>>>(idx, n, K, static_cast<uint8_t*>(codes_out), err_flag);
    else if (elem_bytes == 2)
        narrow_kernel<uint16_t><<<blocks, 256, 0, s>>>(idx, n, K, static_cast<uint16_t*>(codes_out), err_flag);
    else
        narrow_kernel<uint32_t><<<blocks, 256, 0, s>>>(idx, n, K, static_cast<uint32_t*>(codes_out), err_flag);
    VQB_LAUNCH_CHECK("narrow_kernel");
    return VQB_OK;
}

extern "C" int vqb_indices_widen(const void* codes, int64_t n, int elem_bytes, int64_t* idx_out, vqb_stream_t stream) {
    VQB_DEVICE_TRY();
    if (n < 0 || (elem_bytes != 1 && elem_bytes != 2 && elem_bytes != 4)) {
        set_error("vqb_indices_widen: invalid argument (n=%lld elem_bytes=%d)", (long long)n, elem_bytes);
        return VQB_ERR_INVALID_ARG;
    }
    if (n == 0) return VQB_OK;
    if (!codes || !idx_out) {
        set_error("vqb_indices_widen: null pointer");
        return VQB_ERR_INVALID_ARG;
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const unsigned blocks = (unsigned)((n + 2047) / 2048);
    if (elem_bytes == 1)
        widen_kernel<uint8_t><<<blocks, 256, 0, s>>>(static_cast<const uint8_t*>(codes), n, idx_out);
    else if (elem_bytes == 2)
        widen_kernel<uint16_t><<<blocks, 256, 0, s>>>(static_cast<const uint16_t*>(codes), n, idx_out);
    else
        widen_kernel<uint32_t><<<blocks, 256, 0, s>>>(static_cast<const uint32_t*>(codes), n, idx_out);
    VQB_LAUNCH_CHECK("widen_kernel");
    return VQB_OK;
}
