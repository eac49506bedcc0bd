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

// ---------------------------------------------------------------------------
// Small-share variants (round 2): 256 threads, <= 40 KB of staging per CTA, so four or more CTAs share an SM.  CL = 1
// is the product path for groups that fit (16 x 16 latents: forward 0.088 -> 0.068 ms, backward 0.149 -> 0.126 ms on
// [256, 512, 16, 16]).  CL > 1 -- a thread-block cluster per (image, group), each CTA staging 1/CL of the run, the group
// sums meeting through distributed shared memory in rank order -- was built to get the same residency for the 64 KB
// groups of the encoder tail and is SLOWER there (forward 0.81 -> 1.03 ms, backward 2.46 -> 3.11 ms on [1024, 512, 32,
// 32], profiles/r02_groupnorm_silu_bench.txt): measurement build only (vqb_tune "norm_cluster" 2).  What the one-CTA
// backward lacked was bytes in flight, not residency -- see groupnorm_silu_bwd_reg_kernel.
// ---------------------------------------------------------------------------
constexpr int kNormClThreads = 256;
constexpr int kNormClSmemBytes = 40 * 1024;  // per-CTA staging budget of the cluster kernels

__device__ __forceinline__ uint32_t norm_cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void norm_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float norm_ld_peer(const float* p, uint32_t rank) {
    uint32_t ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"((uint32_t)__cvta_generic_to_shared(p)), "r"(rank));
    float v;
    asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(ra) : "memory");
    return v;
}

// CTA sum (fixed order), then cluster sum in rank order; all threads of all CTAs of the cluster get the same total.
// slot: a float in static shared memory (same offset in every CTA of the cluster), exclusive to this call site.
template <int CL>
__device__ __forceinline__ float cluster_sum(float v, float* red, float* slot) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < kNormClThreads / 32; ++w) t += red[w];
    if constexpr (CL == 1) return t;
    if (threadIdx.x == 0) *slot = t;
    norm_cluster_sync();  // (also a CTA barrier) every partial sum is published
    float tot = 0.f;
#pragma unroll
    for (int r = 0; r < CL; ++r) tot += norm_ld_peer(slot, (uint32_t)r);
    return tot;
}

// this CTA's share [e0, e1) of the group's n elements: whole (channel, kNormSeg-element segment) items, dealt in
// contiguous runs of ceil(items / CL)
struct NormShare {
    int64_t i0, i1, e0, e1;
};
__host__ __device__ inline int64_t norm_item_elem(int64_t it, int64_t items, int64_t segs, int64_t HW, int64_t n) {
    if (it >= items) return n;
    const int64_t cc = it / segs;
    return cc * HW + (it - cc * segs) * kNormSeg;
}
__host__ __device__ inline NormShare norm_share(int rank, int CL, int cpg, int64_t HW) {
    const int64_t segs = (HW + kNormSeg - 1) / kNormSeg, items = (int64_t)cpg * segs, n = (int64_t)cpg * HW;
    const int64_t ipc = (items + CL - 1) / CL;
    NormShare sh;
    sh.i0 = (int64_t)rank * ipc < items ? (int64_t)rank * ipc : items;
    sh.i1 = sh.i0 + ipc < items ? sh.i0 + ipc : items;
    sh.e0 = norm_item_elem(sh.i0, items, segs, HW, n);
    sh.e1 = norm_item_elem(sh.i1, items, segs, HW, n);
    return sh;
}

template <int CL>
__global__ void __launch_bounds__(kNormClThreads)
    groupnorm_silu_fwd_cl_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                                 int C, int64_t HW, int G, float eps, float* __restrict__ y, float* __restrict__ mean_out,
                                 float* __restrict__ rstd_out) {
    extern __shared__ __align__(16) float stage[];
    __shared__ float red[kNormClThreads / 32];
    __shared__ float slot[2];
    const int cpg = C / G;
    const int64_t n = (int64_t)cpg * HW;
    const int rank = CL > 1 ? (int)norm_cluster_rank() : 0;
    const int64_t bg = blockIdx.x / CL;  // image * G + group
    const int g = (int)(bg % G);
    const NormShare sh = norm_share(rank, CL, cpg, HW);
    const float* xp = x + bg * n + sh.e0;
    float* yp = y + bg * n + sh.e0;
    const int64_t m = sh.e1 - sh.e0;
    const bool vec = (HW % 4 == 0) && ((reinterpret_cast<uintptr_t>(xp) | reinterpret_cast<uintptr_t>(yp)) & 15u) == 0;
    const float inv_n = 1.f / (float)n;
    float s = 0.f;
    if (vec) {
        for (int64_t i = threadIdx.x; i < m / 4; i += kNormClThreads) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(xp) + i);
            reinterpret_cast<float4*>(stage)[i] = v;
            s += (v.x + v.y) + (v.z + v.w);
        }
    } else {
        for (int64_t i = threadIdx.x; i < m; i += kNormClThreads) {
            const float v = __ldg(xp + i);
            stage[i] = v;
            s += v;
        }
    }
    const float mean = cluster_sum<CL>(s, red, slot + 0) * inv_n;
    float q = 0.f;
    for (int64_t i = threadIdx.x; i < m; i += kNormClThreads) {
        const float d = stage[i] - mean;
        q = fmaf(d, d, q);
    }
    const float var = cluster_sum<CL>(q, red, slot + 1) * inv_n;
    const float rstd = rsqrtf(var + eps);
    if (rank == 0 && threadIdx.x == 0) {
        mean_out[bg] = mean;
        rstd_out[bg] = rstd;
    }
    if (vec) {
        const int64_t hw4 = HW / 4, b4 = sh.e0 / 4;
        for (int64_t i = threadIdx.x; i < m / 4; i += kNormClThreads) {
            const int c = g * cpg + (int)((b4 + i) / hw4);
            const float a = __ldg(gamma + c) * rstd, b = __ldg(beta + c) - mean * a;
            const float4 v = reinterpret_cast<const float4*>(stage)[i];
            float4 o;
            o.x = silu_f(fmaf(v.x, a, b));
            o.y = silu_f(fmaf(v.y, a, b));
            o.z = silu_f(fmaf(v.z, a, b));
            o.w = silu_f(fmaf(v.w, a, b));
            reinterpret_cast<float4*>(yp)[i] = o;
        }
    } else {
        for (int64_t i = threadIdx.x; i < m; i += kNormClThreads) {
            const int c = g * cpg + (int)((sh.e0 + i) / HW);
            const float a = __ldg(gamma + c) * rstd, b = __ldg(beta + c) - mean * a;
            yp[i] = silu_f(fmaf(stage[i], a, b));
        }
    }
    if constexpr (CL > 1) norm_cluster_sync();  // no CTA leaves while a peer may still read its partial sums
}

template <int CL>
__global__ void __launch_bounds__(kNormClThreads)
    groupnorm_silu_bwd_cl_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ gamma,
                                 const float* __restrict__ beta, const float* __restrict__ mean_in,
                                 const float* __restrict__ rstd_in, int C, int64_t HW, int G, float* __restrict__ dx,
                                 float* __restrict__ dgamma, float* __restrict__ dbeta) {
    extern __shared__ __align__(16) float stage[];  // [xhat (m) | dxhat (m)], m = this CTA's share
    __shared__ float red[kNormClThreads / 32];
    __shared__ float slot[2];
    __shared__ float s_dg[kNormMaxCpg], s_db[kNormMaxCpg];
    const int cpg = C / G;
    const int64_t n = (int64_t)cpg * HW;
    const int rank = CL > 1 ? (int)norm_cluster_rank() : 0;
    const int64_t bg = blockIdx.x / CL;
    const int g = (int)(bg % G);
    const NormShare sh = norm_share(rank, CL, cpg, HW);
    const int64_t m = sh.e1 - sh.e0;
    const float* xp = x + bg * n;
    const float* dyp = dy + bg * n;
    float* dxp = dx + bg * n + sh.e0;
    const float mean = mean_in[bg], rstd = rstd_in[bg];
    float* s_xhat = stage;
    float* s_dxh = stage + m;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool smem_sums = cpg <= kNormMaxCpg;
    if (smem_sums)
        for (int i = threadIdx.x; i < cpg; i += kNormClThreads) s_dg[i] = s_db[i] = 0.f;
    __syncthreads();
    const bool vec = (HW % 4 == 0) && ((reinterpret_cast<uintptr_t>(xp) | reinterpret_cast<uintptr_t>(dyp)) & 15u) == 0;
    const int64_t segs = (HW + kNormSeg - 1) / kNormSeg;
    float s1 = 0.f, s2 = 0.f;
    for (int64_t it = sh.i0 + warp; it < sh.i1; it += kNormClThreads / 32) {
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
            s_xhat[j] = xh;
            s_dxh[j] = dxh;
            s1 += dxh;
            s2 = fmaf(dxh, xh, s2);
        };
        const int64_t base = (int64_t)cc * HW;
        if (vec) {
            for (int64_t i = lo + 4 * lane; i < hi; i += 128) {
                const float4 xv = __ldg(reinterpret_cast<const float4*>(xp + base + i));
                const float4 dv = __ldg(reinterpret_cast<const float4*>(dyp + base + i));
                const int64_t j = base + i - sh.e0;
                one(xv.x, dv.x, j);
                one(xv.y, dv.y, j + 1);
                one(xv.z, dv.z, j + 2);
                one(xv.w, dv.w, j + 3);
            }
        } else {
            for (int64_t i = lo + lane; i < hi; i += 32) one(__ldg(xp + base + i), __ldg(dyp + base + i), base + i - sh.e0);
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
    const float m1 = cluster_sum<CL>(s1, red, slot + 0) * inv_n;  // (its barriers also publish the staged tiles and the channel sums)
    const float m2 = cluster_sum<CL>(s2, red, slot + 1) * inv_n;
    if (smem_sums && sh.i1 > sh.i0) {
        const int c_lo = (int)(sh.i0 / segs), c_hi = (int)((sh.i1 - 1) / segs);
        for (int i = c_lo + (int)threadIdx.x; i <= c_hi; i += kNormClThreads) {
            if (dgamma) atomicAdd(dgamma + g * cpg + i, s_dg[i]);
            if (dbeta) atomicAdd(dbeta + g * cpg + i, s_db[i]);
        }
    }
    if ((m % 4 == 0) && (reinterpret_cast<uintptr_t>(dxp) & 15u) == 0) {
        for (int64_t j4 = threadIdx.x; j4 < m / 4; j4 += kNormClThreads) {
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
        for (int64_t j = threadIdx.x; j < m; j += kNormClThreads) dxp[j] = rstd * (s_dxh[j] - m1 - s_xhat[j] * m2);
    }
    if constexpr (CL > 1) norm_cluster_sync();  // no CTA leaves while a peer may still read its partial sums
}

// Backward for groups of at most 16 (channel, 1024-element segment) items -- the encoder tail: 16 channels x 32 x 32 --
// with NO staging: warp w owns item w, a lane loads its 8 float4 of x and 8 of dy back to back (128 KB in flight per
// SM: the staged kernel had two loads per lane outstanding, 16 KB per SM, and ran at the latency of that, 0.40 of the
// HBM peak), keeps xhat and dxhat in registers across the one block reduction, and stores dx from registers.
constexpr int kNormRegItems = kNormThreads / 32;
__global__ void __launch_bounds__(kNormThreads, 1)
    groupnorm_silu_bwd_reg_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ gamma,
                                  const float* __restrict__ beta, const float* __restrict__ mean_in,
                                  const float* __restrict__ rstd_in, int C, int64_t HW, int G, float* __restrict__ dx,
                                  float* __restrict__ dgamma, float* __restrict__ dbeta) {
    __shared__ float red[kNormThreads / 32];
    const int cpg = C / G;
    const int64_t n = (int64_t)cpg * HW;
    const int64_t bg = blockIdx.x;
    const int g = (int)(bg % G);
    const float mean = mean_in[bg], rstd = rstd_in[bg];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t segs = (HW + kNormSeg - 1) / kNormSeg;
    const int64_t items = (int64_t)cpg * segs;
    const bool live = warp < items;
    const int cc = live ? (int)(warp / segs) : 0;
    const int64_t lo = live ? (warp - (int64_t)cc * segs) * kNormSeg : 0;
    const int64_t hi = live ? (lo + kNormSeg < HW ? lo + kNormSeg : HW) : 0;
    const int64_t base = bg * n + (int64_t)cc * HW;
    const int c = g * cpg + cc;
    float4 xh[8], dh[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        const int64_t i = lo + 4 * lane + 128 * u;
        xh[u] = i < hi ? __ldg(reinterpret_cast<const float4*>(x + base + i)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        const int64_t i = lo + 4 * lane + 128 * u;
        dh[u] = i < hi ? __ldg(reinterpret_cast<const float4*>(dy + base + i)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const float ga = __ldg(gamma + c), be = __ldg(beta + c);
    float dg = 0.f, db = 0.f, s1 = 0.f, s2 = 0.f;
    auto one = [&](float& xv, float& dv) {  // (x, dy) -> (xhat, dxhat) in place
        const float xhat = (xv - mean) * rstd;
        const float u = fmaf(xhat, ga, be);
        const float sg = 1.f / (1.f + __expf(-u));
        const float gu = dv * sg * fmaf(u, 1.f - sg, 1.f);
        dg = fmaf(gu, xhat, dg);
        db += gu;
        const float dxh = gu * ga;
        s1 += dxh;
        s2 = fmaf(dxh, xhat, s2);
        xv = xhat;
        dv = dxh;
    };
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        if (lo + 4 * lane + 128 * u < hi) {
            one(xh[u].x, dh[u].x);
            one(xh[u].y, dh[u].y);
            one(xh[u].z, dh[u].z);
            one(xh[u].w, dh[u].w);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        dg += __shfl_xor_sync(0xffffffffu, dg, o);
        db += __shfl_xor_sync(0xffffffffu, db, o);
    }
    if (live && lane == 0) {
        if (dgamma) atomicAdd(dgamma + c, dg);
        if (dbeta) atomicAdd(dbeta + c, db);
    }
    const float inv_n = 1.f / (float)n;
    const float m1 = block_sum(s1, red) * inv_n;
    const float m2 = block_sum(s2, red) * inv_n;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        const int64_t i = lo + 4 * lane + 128 * u;
        if (i < hi) {
            float4 o;
            o.x = rstd * (dh[u].x - m1 - xh[u].x * m2);
            o.y = rstd * (dh[u].y - m1 - xh[u].y * m2);
            o.z = rstd * (dh[u].z - m1 - xh[u].z * m2);
            o.w = rstd * (dh[u].w - m1 - xh[u].w * m2);
            *reinterpret_cast<float4*>(dx + base + i) = o;
        }
    }
}

// Forward for the same groups (at most 16 items; the product path, vqb_tune "norm_fwd_reg" 0 selects the staged kernel):
// the group lives in registers (eight float4 per lane, all loads issued back to back), mean and the centred variance come
// from two block reductions over registers, nothing is staged; two 512-thread CTAs per SM.  [1024, 512, 32, 32]:
// 0.81 -> 0.70 ms (0.81 -> 0.95 of the HBM peak for its two passes).
__global__ void __launch_bounds__(kNormThreads, 2)
    groupnorm_silu_fwd_reg_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                                  int C, int64_t HW, int G, float eps, float* __restrict__ y, float* __restrict__ mean_out,
                                  float* __restrict__ rstd_out) {
    __shared__ float red[kNormThreads / 32];
    const int cpg = C / G;
    const int64_t n = (int64_t)cpg * HW;
    const int64_t bg = blockIdx.x;
    const int g = (int)(bg % G);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t segs = (HW + kNormSeg - 1) / kNormSeg;
    const int64_t items = (int64_t)cpg * segs;
    const bool live = warp < items;
    const int cc = live ? (int)(warp / segs) : 0;
    const int64_t lo = live ? (warp - (int64_t)cc * segs) * kNormSeg : 0;
    const int64_t hi = live ? (lo + kNormSeg < HW ? lo + kNormSeg : HW) : 0;
    const int64_t base = bg * n + (int64_t)cc * HW;
    const int c = g * cpg + cc;
    float4 v[8];
    float s = 0.f;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        const int64_t i = lo + 4 * lane + 128 * u;
        v[u] = i < hi ? __ldg(reinterpret_cast<const float4*>(x + base + i)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) s += (v[u].x + v[u].y) + (v[u].z + v[u].w);
    const float inv_n = 1.f / (float)n;
    const float mean = block_sum(s, red) * inv_n;
    float q = 0.f;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        if (lo + 4 * lane + 128 * u < hi) {
            const float a = v[u].x - mean, b = v[u].y - mean, c2 = v[u].z - mean, d = v[u].w - mean;
            q = fmaf(a, a, q);
            q = fmaf(b, b, q);
            q = fmaf(c2, c2, q);
            q = fmaf(d, d, q);
        }
    }
    const float var = block_sum(q, red) * inv_n;
    const float rstd = rsqrtf(var + eps);
    if (threadIdx.x == 0) {
        mean_out[bg] = mean;
        rstd_out[bg] = rstd;
    }
    const float a = __ldg(gamma + c) * rstd, b = __ldg(beta + c) - mean * a;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        const int64_t i = lo + 4 * lane + 128 * u;
        if (i < hi) {
            float4 o;
            o.x = silu_f(fmaf(v[u].x, a, b));
            o.y = silu_f(fmaf(v[u].y, a, b));
            o.z = silu_f(fmaf(v[u].z, a, b));
            o.w = silu_f(fmaf(v[u].w, a, b));
            *reinterpret_cast<float4*>(y + base + i) = o;
        }
    }
}

// The register-resident backward with TWO CTAs per SM (the product path; vqb_tune "norm_bwd2" 0 selects the one above):
// xhat goes to shared memory (64 KB per CTA at the encoder tail), only dxhat stays in registers (32 per thread), x and dy
// are loaded one after the other so that a thread never holds more than eight float4 -- the register budget of 64 that
// two 512-thread CTAs leave.  One CTA's loads now overlap the other's arithmetic and stores: [1024, 512, 32, 32]
// 1.52 -> 1.21 ms (0.65 -> 0.82 of the HBM peak for its three passes), bit-identical dx.
__global__ void __launch_bounds__(kNormThreads, 2)
    groupnorm_silu_bwd_reg2_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, const float* __restrict__ mean_in,
                                   const float* __restrict__ rstd_in, int C, int64_t HW, int G, float* __restrict__ dx,
                                   float* __restrict__ dgamma, float* __restrict__ dbeta) {
    extern __shared__ __align__(16) float stage[];  // xhat: [warp][8][32 lanes] float4
    __shared__ float red[kNormThreads / 32];
    const int cpg = C / G;
    const int64_t n = (int64_t)cpg * HW;
    const int64_t bg = blockIdx.x;
    const int g = (int)(bg % G);
    const float mean = mean_in[bg], rstd = rstd_in[bg];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t segs = (HW + kNormSeg - 1) / kNormSeg;
    const int64_t items = (int64_t)cpg * segs;
    const bool live = warp < items;
    const int cc = live ? (int)(warp / segs) : 0;
    const int64_t lo = live ? (warp - (int64_t)cc * segs) * kNormSeg : 0;
    const int64_t hi = live ? (lo + kNormSeg < HW ? lo + kNormSeg : HW) : 0;
    const int64_t base = bg * n + (int64_t)cc * HW;
    const int c = g * cpg + cc;
    float4* sx = reinterpret_cast<float4*>(stage) + (size_t)warp * 256 + lane;  // [u][lane]: conflict-free
    float4 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        const int64_t i = lo + 4 * lane + 128 * u;
        v[u] = i < hi ? __ldg(reinterpret_cast<const float4*>(x + base + i)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        float4 h;
        h.x = (v[u].x - mean) * rstd;
        h.y = (v[u].y - mean) * rstd;
        h.z = (v[u].z - mean) * rstd;
        h.w = (v[u].w - mean) * rstd;
        sx[32 * u] = h;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        const int64_t i = lo + 4 * lane + 128 * u;
        v[u] = i < hi ? __ldg(reinterpret_cast<const float4*>(dy + base + i)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const float ga = __ldg(gamma + c), be = __ldg(beta + c);
    float dg = 0.f, db = 0.f, s1 = 0.f, s2 = 0.f;
    auto one = [&](float xhat, float& dv) {  // dy -> dxhat in place
        const float u = fmaf(xhat, ga, be);
        const float sg = 1.f / (1.f + __expf(-u));
        const float gu = dv * sg * fmaf(u, 1.f - sg, 1.f);
        dg = fmaf(gu, xhat, dg);
        db += gu;
        const float dxh = gu * ga;
        s1 += dxh;
        s2 = fmaf(dxh, xhat, s2);
        dv = dxh;
    };
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        if (lo + 4 * lane + 128 * u < hi) {
            const float4 h = sx[32 * u];
            one(h.x, v[u].x);
            one(h.y, v[u].y);
            one(h.z, v[u].z);
            one(h.w, v[u].w);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        dg += __shfl_xor_sync(0xffffffffu, dg, o);
        db += __shfl_xor_sync(0xffffffffu, db, o);
    }
    if (live && lane == 0) {
        if (dgamma) atomicAdd(dgamma + c, dg);
        if (dbeta) atomicAdd(dbeta + c, db);
    }
    const float inv_n = 1.f / (float)n;
    const float m1 = block_sum(s1, red) * inv_n;
    const float m2 = block_sum(s2, red) * inv_n;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        const int64_t i = lo + 4 * lane + 128 * u;
        if (i < hi) {
            const float4 h = sx[32 * u];
            float4 o;
            o.x = rstd * (v[u].x - m1 - h.x * m2);
            o.y = rstd * (v[u].y - m1 - h.y * m2);
            o.z = rstd * (v[u].z - m1 - h.z * m2);
            o.w = rstd * (v[u].w - m1 - h.w * m2);
            *reinterpret_cast<float4*>(dx + base + i) = o;
        }
    }
}

// smallest cluster size whose largest per-CTA share, times `copies` staged floats per element, fits the budget; 0 = none
static int norm_cluster_size(int cpg, int64_t HW, int copies, int max_cl, int64_t* max_share) {
    for (int cl = 1; cl <= max_cl; cl *= 2) {
        int64_t mx = 0;
        for (int r = 0; r < cl; ++r) {
            const NormShare sh = norm_share(r, cl, cpg, HW);
            mx = sh.e1 - sh.e0 > mx ? sh.e1 - sh.e0 : mx;
        }
        if (mx * copies * (int64_t)sizeof(float) <= kNormClSmemBytes) {
            *max_share = mx;
            return cl;
        }
    }
    return 0;
}

// vqb_tune "norm_cluster": 0 = the round-1 one-CTA-per-group staged kernels only; 1 = product (register-resident backward,
// small-share kernels for groups that fit 40 KB); 2 = also clusters of 2-8 CTAs per group (measured slower)
VQB_KNOB g_norm_cluster = 1;
VQB_KNOB g_norm_fwd_reg = 1;  // vqb_tune "norm_fwd_reg": 1 (default) = register-resident forward for groups of at most 16 items
VQB_KNOB g_norm_bwd2 = 1;  // vqb_tune "norm_bwd2": 1 (default) = the two-CTAs-per-SM variant of the register-resident backward
#ifdef VQB_EXPERIMENTAL
void set_norm_cluster(int v) { g_norm_cluster = v; }
void set_norm_bwd2(int v) { g_norm_bwd2 = v; }
void set_norm_fwd_reg(int v) { g_norm_fwd_reg = v; }
#endif

template <typename Kern, typename... Args>
static cudaError_t norm_launch_cluster(Kern kern, unsigned blocks, int cl, size_t smem, cudaStream_t s, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(blocks);
    cfg.blockDim = dim3(kNormClThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)cl;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, args...);
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
    int64_t share = 0;
    const int cl = g_norm_cluster ? norm_cluster_size(C / groups, HW, 1, g_norm_cluster == 2 ? 8 : 1, &share) : 0;
    const int64_t f_items = (int64_t)(C / groups) * ((HW + kNormSeg - 1) / kNormSeg);
    if (g_norm_fwd_reg && g_norm_cluster == 1 && f_items <= kNormRegItems && HW % 4 == 0 && n >= 8192 &&
        ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15u) == 0) {
        groupnorm_silu_fwd_reg_kernel<<<blocks, kNormThreads, 0, s>>>(x, gamma, beta, C, HW, groups, eps, y, mean_out, rstd_out);
    } else if (cl > 0 && B * (int64_t)groups * cl < (1LL << 31)) {
        const size_t smem = sizeof(float) * (size_t)share;
#define VQB_NORM_FWD(c)                                                                                                  \
    VQB_CUDA_TRY(norm_launch_cluster(groupnorm_silu_fwd_cl_kernel<c>, blocks * (c), c, smem, s, x, gamma, beta, C, HW,    \
                                     groups, eps, y, mean_out, rstd_out))
        switch (cl) {
            case 1: VQB_NORM_FWD(1); break;
#ifdef VQB_EXPERIMENTAL
            case 2: VQB_NORM_FWD(2); break;
            case 4: VQB_NORM_FWD(4); break;
            default: VQB_NORM_FWD(8); break;
#endif
        }
#undef VQB_NORM_FWD
    } else if (n <= kNormSmemFloats) {
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
    int64_t share = 0;
    const int cl = g_norm_cluster ? norm_cluster_size(C / groups, HW, 2, g_norm_cluster == 2 ? 8 : 1, &share) : 0;
    const int64_t items = (int64_t)(C / groups) * ((HW + kNormSeg - 1) / kNormSeg);
    if (g_norm_cluster == 1 && items <= kNormRegItems && HW % 4 == 0 && n >= 8192 &&
        ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(dx)) & 15u) == 0) {
        if (g_norm_bwd2) {
            constexpr int kSm = kNormThreads * 8 * 16;  // 64 KB
            VQB_CUDA_TRY(cudaFuncSetAttribute(groupnorm_silu_bwd_reg2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSm));
            groupnorm_silu_bwd_reg2_kernel<<<blocks, kNormThreads, kSm, s>>>(dy, x, gamma, beta, mean, rstd, C, HW, groups, dx,
                                                                            dgamma_accum, dbeta_accum);
        } else {
            groupnorm_silu_bwd_reg_kernel<<<blocks, kNormThreads, 0, s>>>(dy, x, gamma, beta, mean, rstd, C, HW, groups, dx,
                                                                         dgamma_accum, dbeta_accum);
        }
    } else if (cl > 0 && B * (int64_t)groups * cl < (1LL << 31)) {
        const size_t smem = sizeof(float) * (size_t)(2 * share);
#define VQB_NORM_BWD(c)                                                                                                  \
    VQB_CUDA_TRY(norm_launch_cluster(groupnorm_silu_bwd_cl_kernel<c>, blocks * (c), c, smem, s, dy, x, gamma, beta, mean, \
                                     rstd, C, HW, groups, dx, dgamma_accum, dbeta_accum))
        switch (cl) {
            case 1: VQB_NORM_BWD(1); break;
#ifdef VQB_EXPERIMENTAL
            case 2: VQB_NORM_BWD(2); break;
            case 4: VQB_NORM_BWD(4); break;
            default: VQB_NORM_BWD(8); break;
#endif
        }
#undef VQB_NORM_BWD
    } else if (2 * n <= kNormSmemFloats) {
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
