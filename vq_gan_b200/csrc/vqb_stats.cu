// Packing of the per-step codebook statistics for the data-parallel all-reduce:
// [ dE (K*D fp32) | scalars (fp32) | hist mod 4096 | hist div 4096 ] in ONE flat fp32 message, and the
// inverse with the DDP-style 1/world averaging of dE (train_vqgan.py:197-209 via accelerate).  The two
// histogram planes are integers below 4096 * world, so their fp32 sums are exact.  One launch each instead
// of a dozen elementwise torch kernels.
#include "vqb_common.cuh"

namespace vqb {

constexpr int kHistRadix = 4096;

__global__ void __launch_bounds__(256)
    stats_pack_kernel(const float* __restrict__ dE, int64_t n_dE, const float* __restrict__ scalars, int n_scalars,
                      const int64_t* __restrict__ hist, int n_hist, float* __restrict__ flat) {
    const int64_t total = n_dE + n_scalars + 2 * (int64_t)n_hist;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        float v;
        if (i < n_dE) {
            v = dE[i];
        } else if (i < n_dE + n_scalars) {
            v = scalars[i - n_dE];
        } else if (i < n_dE + n_scalars + n_hist) {
            v = (float)(hist[i - n_dE - n_scalars] % kHistRadix);
        } else {
            v = (float)(hist[i - n_dE - n_scalars - n_hist] / kHistRadix);
        }
        flat[i] = v;
    }
}

__global__ void __launch_bounds__(256)
    stats_unpack_kernel(const float* __restrict__ flat, int64_t n_dE, int n_scalars, int n_hist, float dE_scale,
                        float* __restrict__ dE_out, float* __restrict__ scalars_out, int64_t* __restrict__ hist_out) {
    const int64_t total = n_dE + n_scalars + (int64_t)n_hist;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        if (i < n_dE) {
            dE_out[i] = flat[i] * dE_scale;
        } else if (i < n_dE + n_scalars) {
            scalars_out[i - n_dE] = flat[i];
        } else {
            const int64_t h = i - n_dE - n_scalars;
            hist_out[h] = llrintf(flat[n_dE + n_scalars + h]) + llrintf(flat[n_dE + n_scalars + n_hist + h]) * kHistRadix;
        }
    }
}

}  // namespace vqb

using namespace vqb;

extern "C" int vqb_stats_pack(const float* dE, int64_t n_dE, const float* scalars, int n_scalars, const int64_t* hist,
                              int n_hist, float* flat_out, vqb_stream_t stream) {
    VQB_DEVICE_TRY();
    if (n_dE < 0 || n_scalars < 0 || n_hist < 0 || (n_dE && !dE) || (n_scalars && !scalars) || (n_hist && !hist) || !flat_out) {
        set_error("vqb_stats_pack: invalid argument");
        return VQB_ERR_INVALID_ARG;
    }
    const int64_t total = n_dE + n_scalars + 2 * (int64_t)n_hist;
    if (total == 0) return VQB_OK;
    int64_t blocks = (total + 255) / 256;
    if (blocks > (int64_t)sm_count() * 8) blocks = (int64_t)sm_count() * 8;
    stats_pack_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(dE, n_dE, scalars, n_scalars, hist,
                                                                                       n_hist, flat_out);
    VQB_LAUNCH_CHECK("stats_pack_kernel");
    return VQB_OK;
}

extern "C" int vqb_stats_unpack(const float* flat, int64_t n_dE, int n_scalars, int n_hist, float dE_scale, float* dE_out,
                                float* scalars_out, int64_t* hist_out, vqb_stream_t stream) {
    VQB_DEVICE_TRY();
    if (n_dE < 0 || n_scalars < 0 || n_hist < 0 || !flat || (n_dE && !dE_out) || (n_scalars && !scalars_out) ||
        (n_hist && !hist_out)) {
        set_error("vqb_stats_unpack: invalid argument");
        return VQB_ERR_INVALID_ARG;
    }
    const int64_t total = n_dE + n_scalars + (int64_t)n_hist;
    if (total == 0) return VQB_OK;
    int64_t blocks = (total + 255) / 256;
    if (blocks > (int64_t)sm_count() * 8) blocks = (int64_t)sm_count() * 8;
    stats_unpack_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(flat, n_dE, n_scalars, n_hist,
                                                                                         dE_scale, dE_out, scalars_out,
                                                                                         hist_out);
    VQB_LAUNCH_CHECK("stats_unpack_kernel");
    return VQB_OK;
}
