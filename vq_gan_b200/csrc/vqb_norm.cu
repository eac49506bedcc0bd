// Next row N2: the normalisation + activation in front of the convolutions that produce /
// consume the quantizer's tensors -- `h = norm_out(h); h = F.silu(h)` of the reference Encoder / Decoder
// (vqgan_ldm_baseline/models/encoder_decoder.py:166-167, 249-250; nn.GroupNorm(32, C, eps=1e-6, affine=True)).
// PyTorch runs this as GroupNorm (statistics pass + apply pass) followed by an elementwise SiLU: the activation is
// read three times and written twice.  Here one CTA owns one (image, group): in NCHW the C/G channels of a group are
// one CONTIGUOUS run of n = (C/G)*HW floats, so the CTA stages it in shared memory with float4 loads (when it fits:
// 64 KB at the encoder tail, C = 512, 32 x 32), takes mean and variance in two exact passes over shared memory, and
// writes y = silu((x - mean) * rstd * gamma_c + beta_c) -- one HBM read, one HBM write.  Groups that do not fit
// (decoder tail: 4 channels x 256 x 256 = 1 MB) take the same three passes over global memory (the re-reads hit L2).
// The backward recomputes xhat / u / sigmoid(u) from x and the saved (mean, rstd), stages x and dy the same way,
// and produces dx plus per-channel dgamma / dbeta (one atomicAdd per channel per image).
// The 3x3 convolutions themselves stay cuDNN calls (dense conv stacks are out of scope).
#include "vqb_common.cuh"

namespace vqb {

constexpr int kNormThreads = 512;
constexpr int kNormSmemFloats = 48 * 1024;  // 192 KB of staging (forward: x; backward: x and dy -> half each)

__device__ __forceinline__ float block_sum(float v, float* red) {  // all threads get the total; fixed order
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();  // red[] may still be read from the previous call
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < kNormThreads / 32; ++w) t += red[w];
    return t;
}

__device__ __forceinline__ float silu_f(float u) { return u / (1.f + __expf(-u)); }

// kStage: the group fits in shared memory (n4 <= kNormSmemFloats / 4 float4s)
template <bool kStage>
__global__ void __launch_bounds__(kNormThreads)
    groupnorm_silu_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                              int C, int64_t HW, int G, float eps, float* __restrict__ y, float* __restrict__ mean_out,
                              float* __restrict__ rstd_out) {
    extern __shared__ __align__(16) float stage[];
    __shared__ float red[kNormThreads / 32];
    const int cpg = C / G;
    const int64_t n = (int64_t)cpg * HW;
    const int64_t bg = blockIdx.x;  // image * G + group
    const int g = (int)(bg % G);
    const float* xp = x + bg * n;   // contiguous: channels [g*cpg, (g+1)*cpg) of one image
    float* yp = y + bg * n;
    const bool vec = (HW % 4 == 0) && ((reinterpret_cast<uintptr_t>(xp) | reinterpret_cast<uintptr_t>(yp)) & 15u) == 0;
    const float inv_n = 1.f / (float)n;
    // pass 1: load (and stage), sum
    float s = 0.f;
    if (vec) {
        const int64_t n4 = n / 4;
        for (int64_t i = threadIdx.x; i < n4; i += kNormThreads) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(xp) + i);
            if (kStage) reinterpret_cast<float4*>(stage)[i] = v;
            s += (v.x + v.y) + (v.z + v.w);
        }
    } else {
        for (int64_t i = threadIdx.x; i < n; i += kNormThreads) {
            const float v = __ldg(xp + i);
            if (kStage) stage[i] = v;
            s += v;
        }
    }
    const float mean = block_sum(s, red) * inv_n;
    // pass 2: centred sum of squares (two-pass variance: no cancellation)
    float q = 0.f;
    for (int64_t i = threadIdx.x; i < n; i += kNormThreads) {
        const float d = (kStage ? stage[i] : __ldg(xp + i)) - mean;
        q = fmaf(d, d, q);
    }
    const float var = block_sum(q, red) * inv_n;
    const float rstd = rsqrtf(var + eps);
    if (threadIdx.x == 0) {
        mean_out[bg] = mean;
        rstd_out[bg] = rstd;
    }
    // pass 3: normalise, affine, SiLU
    if (vec) {
        const int64_t n4 = n / 4, hw4 = HW / 4;
        for (int64_t i = threadIdx.x; i < n4; i += kNormThreads) {
            const int c = g * cpg + (int)(i / hw4);
            const float a = __ldg(gamma + c) * rstd, b = __ldg(beta + c) - mean * a;
            const float4 v = kStage ? reinterpret_cast<const float4*>(stage)[i] : __ldg(reinterpret_cast<const float4*>(xp) + i);
            float4 o;
            o.x = silu_f(fmaf(v.x, a, b));
            o.y = silu_f(fmaf(v.y, a, b));
            o.z = silu_f(fmaf(v.z, a, b));
            o.w = silu_f(fmaf(v.w, a, b));
            reinterpret_cast<float4*>(yp)[i] = o;
        }
    } else {
        for (int64_t i = threadIdx.x; i < n; i += kNormThreads) {
            const int c = g * cpg + (int)(i / HW);
            const float a = __ldg(gamma + c) * rstd, b = __ldg(beta + c) - mean * a;
            yp[i] = silu_f(fmaf(kStage ? stage[i] : __ldg(xp + i), a, b));
        }
    }
}

// dx, dgamma, dbeta of y = silu(xhat * gamma + beta), xhat = (x - mean) * rstd, per (image, group):
//   u = xhat*gamma + beta, s = sigmoid(u), gu = dy * s * (1 + u * (1 - s))          (gradient w.r.t. u)
//   dgamma_c += sum_hw gu * xhat,  dbeta_c += sum_hw gu
//   dxhat = gu * gamma_c;  dx = rstd * (dxhat - mean_n(dxhat) - xhat * mean_n(dxhat * xhat))
// Work items of pass 1 = (channel, 1024-element segment), dealt round-robin to the 16 warps: the per-channel sums are
// warp-local (shuffles) and meet in shared-memory atomics, the two group sums ride in per-thread registers until
// one block reduction -- no CTA barrier inside the sweep.
constexpr int kNormSeg = 1024;
constexpr int kNormMaxCpg = 128;  // per-channel sums in shared memory up to this many channels per group

template <bool kStage>
__global__ void __launch_bounds__(kNormThreads)
    groupnorm_silu_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ gamma,
                              const float* __restrict__ beta, const float* __restrict__ mean_in,
                              const float* __restrict__ rstd_in, int C, int64_t HW, int G, float* __restrict__ dx,
                              float* __restrict__ dgamma, float* __restrict__ dbeta) {
    extern __shared__ __align__(16) float stage[];  // [xhat (n) | dxhat (n)] when kStage
    __shared__ float red[kNormThreads / 32];
    __shared__ float s_dg[kNormMaxCpg], s_db[kNormMaxCpg];
    const int cpg = C / G;
    const int64_t n = (int64_t)cpg * HW;
    const int64_t bg = blockIdx.x;
    const int g = (int)(bg % G);
    const float* xp = x + bg * n;
    const float* dyp = dy + bg * n;
    float* dxp = dx + bg * n;
    const float mean = mean_in[bg], rstd = rstd_in[bg];
    float* s_xhat = stage;
    float* s_dxh = stage + n;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool smem_sums = cpg <= kNormMaxCpg;
    if (smem_sums)
        for (int i = threadIdx.x; i < cpg; i += kNormThreads) s_dg[i] = s_db[i] = 0.f;
    __syncthreads();
    const bool vec = (HW % 4 == 0) && ((reinterpret_cast<uintptr_t>(xp) | reinterpret_cast<uintptr_t>(dyp)) & 15u) == 0;
    const int64_t segs = (HW + kNormSeg - 1) / kNormSeg;
    const int64_t items = (int64_t)cpg * segs;
    float s1 = 0.f, s2 = 0.f;
    for (int64_t it = warp; it < items; it += kNormThreads / 32) {
        const int cc = (int)(it / segs);
        const int64_t lo = (it - (int64_t)cc * segs) * kNormSeg;
        const int64_t hi = lo + kNormSeg < HW ? lo + kNormSeg : HW;
        const int c = g * cpg + cc;
        const float ga = __ldg(gamma + c), be = __ldg(beta + c);
        float dg = 0.f, db = 0.f;
        auto one = [&](float xv, float dyv, int64_t j) {
            const float xh = (xv - mean) * rstd;
            const float u = fmaf(xh, ga, be);
            const float sg = 1.f / (1.f + __expf(-u));
            const float gu = dyv * sg * fmaf(u, 1.f - sg, 1.f);
            dg = fmaf(gu, xh, dg);
            db += gu;
            const float dxh = gu * ga;
            if (kStage) {
                s_xhat[j] = xh;
                s_dxh[j] = dxh;
            }
            s1 += dxh;
            s2 = fmaf(dxh, xh, s2);
        };
        const int64_t base = (int64_t)cc * HW;
        if (vec) {
            for (int64_t i = lo + 4 * lane; i < hi; i += 128) {
                const float4 xv = __ldg(reinterpret_cast<const float4*>(xp + base + i));
                const float4 dv = __ldg(reinterpret_cast<const float4*>(dyp + base + i));
                one(xv.x, dv.x, base + i);
                one(xv.y, dv.y, base + i + 1);
                one(xv.z, dv.z, base + i + 2);
                one(xv.w, dv.w, base + i + 3);
            }
        } else {
            for (int64_t i = lo + lane; i < hi; i += 32) one(__ldg(xp + base + i), __ldg(dyp + base + i), base + i);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            dg += __shfl_xor_sync(0xffffffffu, dg, o);
            db += __shfl_xor_sync(0xffffffffu, db, o);
        }
        if (lane == 0) {
            if (smem_sums) {
                atomicAdd(&s_dg[cc], dg);
                atomicAdd(&s_db[cc], db);
            } else {
                if (dgamma) atomicAdd(dgamma + c, dg);
                if (dbeta) atomicAdd(dbeta + c, db);
            }
        }
    }
    const float inv_n = 1.f / (float)n;
    const float m1 = block_sum(s1, red) * inv_n;  // (its barriers also publish the staged tiles and the channel sums)
    const float m2 = block_sum(s2, red) * inv_n;
    if (smem_sums)
        for (int i = threadIdx.x; i < cpg; i += kNormThreads) {
            if (dgamma) atomicAdd(dgamma + g * cpg + i, s_dg[i]);
            if (dbeta) atomicAdd(dbeta + g * cpg + i, s_db[i]);
        }
    if (kStage) {
        if ((n % 4 == 0) && (reinterpret_cast<uintptr_t>(dxp) & 15u) == 0) {
            for (int64_t j4 = threadIdx.x; j4 < n / 4; j4 += kNormThreads) {
                const float4 xh = reinterpret_cast<const float4*>(s_xhat)[j4];
                const float4 dh = reinterpret_cast<const float4*>(s_dxh)[j4];
                float4 o;
                o.x = rstd * (dh.x - m1 - xh.x * m2);
                o.y = rstd * (dh.y - m1 - xh.y * m2);
                o.z = rstd * (dh.z - m1 - xh.z * m2);
                o.w = rstd * (dh.w - m1 - xh.w * m2);
                reinterpret_cast<float4*>(dxp)[j4] = o;
            }
        } else {
            for (int64_t j = threadIdx.x; j < n; j += kNormThreads) dxp[j] = rstd * (s_dxh[j] - m1 - s_xhat[j] * m2);
        }
    } else {
        for (int64_t j = threadIdx.x; j < n; j += kNormThreads) {
            const int c = g * cpg + (int)(j / HW);
            const float ga = __ldg(gamma + c), be = __ldg(beta + c);
            const float xh = (__ldg(xp + j) - mean) * rstd;
            const float u = fmaf(xh, ga, be);
            const float sg = 1.f / (1.f + __expf(-u));
            const float dxh = __ldg(dyp + j) * sg * fmaf(u, 1.f - sg, 1.f) * ga;
            dxp[j] = rstd * (dxh - m1 - xh * m2);
        }
    }
}

}  // namespace vqb

using namespace vqb;

static int norm_check(int64_t B, int C, int64_t HW, int G) {
    if (B < 0 || C <= 0 || HW <= 0 || G <= 0 || C % G != 0) {
        set_error("groupnorm_silu: invalid shape B=%lld C=%d HW=%lld groups=%d", (long long)B, C, (long long)HW, G);
        return VQB_ERR_INVALID_ARG;
    }
    if (B * (int64_t)G >= (1LL << 31)) {
        set_error("groupnorm_silu: too many (image, group) pairs");
        return VQB_ERR_INVALID_ARG;
    }
    return VQB_OK;
}

extern "C" int vqb_groupnorm_silu_f32(const float* x, int64_t B, int C, int64_t HW, const float* gamma, const float* beta,
                                      int groups, float eps, float* y, float* mean_out, float* rstd_out,
                                      vqb_stream_t stream) {
    VQB_DEVICE_TRY();
    if (int rc = norm_check(B, C, HW, groups)) return rc;
    if (B == 0) return VQB_OK;
    if (!x || !gamma || !beta || !y || !mean_out || !rstd_out) {
        set_error("vqb_groupnorm_silu_f32: null pointer");
        return VQB_ERR_INVALID_ARG;
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int64_t n = (int64_t)(C / groups) * HW;
    const unsigned blocks = (unsigned)(B * groups);
    if (n <= kNormSmemFloats) {
        const size_t smem = sizeof(float) * (size_t)n;
        VQB_CUDA_TRY(cudaFuncSetAttribute(groupnorm_silu_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)(sizeof(float) * kNormSmemFloats)));
        groupnorm_silu_fwd_kernel<true><<<blocks, kNormThreads, smem, s>>>(x, gamma, beta, C, HW, groups, eps, y, mean_out,
                                                                          rstd_out);
    } else {
        groupnorm_silu_fwd_kernel<false><<<blocks, kNormThreads, 0, s>>>(x, gamma, beta, C, HW, groups, eps, y, mean_out,
                                                                        rstd_out);
    }
    VQB_LAUNCH_CHECK("groupnorm_silu_fwd_kernel");
    return VQB_OK;
}

extern "C" int vqb_groupnorm_silu_backward_f32(const float* dy, const float* x, int64_t B, int C, int64_t HW,
                                               const float* gamma, const float* beta, int groups, const float* mean,
                                               const float* rstd, float* dx, float* dgamma_accum, float* dbeta_accum,
                                               vqb_stream_t stream) {
    VQB_DEVICE_TRY();
    if (int rc = norm_check(B, C, HW, groups)) return rc;
    if (B == 0) return VQB_OK;
    if (!dy || !x || !gamma || !beta || !mean || !rstd || !dx) {
        set_error("vqb_groupnorm_silu_backward_f32: null pointer");
        return VQB_ERR_INVALID_ARG;
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int64_t n = (int64_t)(C / groups) * HW;
    const unsigned blocks = (unsigned)(B * groups);
    if (2 * n <= kNormSmemFloats) {
        const size_t smem = sizeof(float) * (size_t)(2 * n);
        VQB_CUDA_TRY(cudaFuncSetAttribute(groupnorm_silu_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)(sizeof(float) * kNormSmemFloats)));
        groupnorm_silu_bwd_kernel<true><<<blocks, kNormThreads, smem, s>>>(dy, x, gamma, beta, mean, rstd, C, HW, groups, dx,
                                                                          dgamma_accum, dbeta_accum);
    } else {
        groupnorm_silu_bwd_kernel<false><<<blocks, kNormThreads, 0, s>>>(dy, x, gamma, beta, mean, rstd, C, HW, groups, dx,
                                                                        dgamma_accum, dbeta_accum);
    }
    VQB_LAUNCH_CHECK("groupnorm_silu_bwd_kernel");
    return VQB_OK;
}
