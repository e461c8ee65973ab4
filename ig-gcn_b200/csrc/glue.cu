// Small fused pieces between the big kernels of the training step.  At the reference's batch sizes every one of these is a
// chain of 10-40 tiny torch kernels (profiles/r1_step_profile_graph_config2.txt: 241 elementwise + 36 BatchNorm launches
// out of 439 per step); each becomes one or two launches here.
//
//   igcn_bn_act_fwd/bwd   : mask * relu(BatchNorm1d(z)) in training mode, optionally over `groups` consecutive slices of the
//                           batch with their own statistics (the stacked plain/explain passes) -- the read-out heads of
//                           kernel/go_model.py:117-146 (conc_for_attention[1:], B, B_D, latent[1:4], latent[5:7]).
//   igcn_mask_loss_fwd/bwd: loss_probability of kernel/sgcn_img_snp.py:153-181 (L1 + binary entropy of sigmoid(prob),
//                           p_e and sigmoid(snps_prob)) as one reduction.
//   igcn_dot              : <a, b> with a fixed summation order (the Laplacian quadratic form of consist_loss).
#include <stdlib.h>

#include "common.cuh"

namespace igcn {

__device__ __forceinline__ float block_sum_256(float v, float* sm /* >= 9 floats */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();                       // protects sm against the previous call's readers
    if (lane == 0) sm[warp] = v;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += sm[w];
    return t;
}
// any block size that is a multiple of 32 (<= 1024); fixed order
__device__ __forceinline__ float block_sum_any(float v, float* sm /* >= 32 floats */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) sm[warp] = v;
    __syncthreads();
    float t = 0.f;
    for (int w = 0; w < nw; ++w) t += sm[w];
    return t;
}

// ---- BatchNorm1d (training) + ReLU + multiplicative mask ------------------------------------------------------------------
// z: (N, C, L) contiguous (L = 1 for a 2-D input).  One CTA per channel c; the groups are visited in order so the running
// statistics receive exactly the updates of `groups` successive module calls.
__global__ void __launch_bounds__(1024) bn_act_fwd_kernel(const float* __restrict__ z, const float* __restrict__ gamma,
                                                         const float* __restrict__ beta, const float* __restrict__ mask,
                                                         int N, int C, int L, int groups, float eps, float momentum, int relu,
                                                         float* __restrict__ running_mean, float* __restrict__ running_var,
                                                         long long* __restrict__ num_batches_tracked,
                                                         float* __restrict__ y, float* __restrict__ stats /* (groups, C, 2) */) {
    IGCN_PDL_SYNC();
    __shared__ float sm[33];
    const int nthr = blockDim.x;
    const int c = blockIdx.x, tid = threadIdx.x;
    const int ng = N / groups;
    const int cnt = ng * L;
    const float ga = gamma ? gamma[c] : 1.f, be = beta ? beta[c] : 0.f;
    float rm = running_mean ? running_mean[c] : 0.f, rv = running_var ? running_var[c] : 1.f;
    for (int g = 0; g < groups; ++g) {
        const int64_t base = ((int64_t)g * ng * C + c) * L;
        float s = 0.f;
        for (int e = tid; e < cnt; e += nthr) {
            const int n = e / L, l = e - n * L;
            s += z[base + (int64_t)n * C * L + l];
        }
        const float mean = block_sum_any(s, sm) / (float)cnt;
        float q = 0.f;
        for (int e = tid; e < cnt; e += nthr) {
            const int n = e / L, l = e - n * L;
            const float d = z[base + (int64_t)n * C * L + l] - mean;
            q += d * d;
        }
        const float var = block_sum_any(q, sm) / (float)cnt;
        const float rstd = rsqrtf(var + eps);
        for (int e = tid; e < cnt; e += nthr) {
            const int n = e / L, l = e - n * L;
            const int64_t i = base + (int64_t)n * C * L + l;
            float v = (z[i] - mean) * rstd * ga + be;
            if (relu) v = fmaxf(v, 0.f);
            if (mask) v *= mask[i];
            y[i] = v;
        }
        if (tid == 0) {
            stats[((int64_t)g * C + c) * 2 + 0] = mean;
            stats[((int64_t)g * C + c) * 2 + 1] = rstd;
        }
        rm = (1.f - momentum) * rm + momentum * mean;
        rv = (1.f - momentum) * rv + momentum * var * ((float)cnt / (float)max(cnt - 1, 1));
    }
    if (tid == 0) {
        if (running_mean) running_mean[c] = rm;
        if (running_var) running_var[c] = rv;
        if (num_batches_tracked && c == 0) *num_batches_tracked += groups;
    }
}

__global__ void __launch_bounds__(1024) bn_act_bwd_kernel(const float* __restrict__ z, const float* __restrict__ gamma,
                                                         const float* __restrict__ beta, const float* __restrict__ mask,
                                                         const float* __restrict__ stats, const float* __restrict__ gy,
                                                         int N, int C, int L, int groups, int relu,
                                                         float* __restrict__ dz, float* __restrict__ dgamma, float* __restrict__ dbeta) {
    IGCN_PDL_SYNC();
    __shared__ float sm[33];
    const int nthr = blockDim.x;
    const int c = blockIdx.x, tid = threadIdx.x;
    const int ng = N / groups;
    const int cnt = ng * L;
    const float ga = gamma ? gamma[c] : 1.f, be = beta ? beta[c] : 0.f;
    float dga = 0.f, dbe = 0.f;
    for (int g = 0; g < groups; ++g) {
        const int64_t base = ((int64_t)g * ng * C + c) * L;
        const float mean = stats[((int64_t)g * C + c) * 2 + 0], rstd = stats[((int64_t)g * C + c) * 2 + 1];
        float s1 = 0.f, s2 = 0.f;
        for (int e = tid; e < cnt; e += nthr) {
            const int n = e / L, l = e - n * L;
            const int64_t i = base + (int64_t)n * C * L + l;
            const float xh = (z[i] - mean) * rstd;
            float d = gy[i];
            if (mask) d *= mask[i];
            if (relu && !(xh * ga + be > 0.f)) d = 0.f;
            s1 += d;
            s2 += d * xh;
        }
        s1 = block_sum_any(s1, sm);
        s2 = block_sum_any(s2, sm);
        const float m1 = s1 / (float)cnt, m2 = s2 / (float)cnt;
        for (int e = tid; e < cnt; e += nthr) {
            const int n = e / L, l = e - n * L;
            const int64_t i = base + (int64_t)n * C * L + l;
            const float xh = (z[i] - mean) * rstd;
            float d = gy[i];
            if (mask) d *= mask[i];
            if (relu && !(xh * ga + be > 0.f)) d = 0.f;
            dz[i] = ga * rstd * (d - m1 - xh * m2);
        }
        dga += s2;
        dbe += s1;
    }
    if (tid == 0) {
        if (dgamma) dgamma[c] = dga;
        if (dbeta) dbeta[c] = dbe;
    }
}

// Register-resident variants for the reference's batch sizes (<= BN_EPT values per thread and group): z (and gy, mask) are read from
// global memory ONCE and the three passes of the kernels above run out of registers -- at a few thousand values per channel the
// kernels are nothing but dependent trips to L2 (CUPTI: 8.2 us each, ten launches per step).  Same element-to-thread mapping and the
// same reduction order as the streaming kernels, so the results are bit identical.
constexpr int BN_EPT = 8;
// offset of element e = n * L + l of a channel inside a (N, C, L) tensor, relative to the channel's first element, without an integer
// division in the common cases (L = 1, L a power of two): three divisions per element were ~half of these kernels' instructions
__device__ __forceinline__ int bn_off(int e, int L, int CL) {
    if (L == 1) return e * CL;
    if ((L & (L - 1)) == 0) {
        const int sh = __ffs(L) - 1;
        return (e >> sh) * CL + (e & (L - 1));
    }
    const int n = e / L;
    return n * CL + (e - n * L);
}

__global__ void __launch_bounds__(1024) bn_act_fwd_reg_kernel(const float* __restrict__ z, const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, const float* __restrict__ mask,
                                                             int N, int C, int L, int groups, float eps, float momentum, int relu,
                                                             float* __restrict__ running_mean, float* __restrict__ running_var,
                                                             long long* __restrict__ num_batches_tracked,
                                                             float* __restrict__ y, float* __restrict__ stats) {
    IGCN_PDL_SYNC();
    __shared__ float sm[33];
    const int nthr = blockDim.x, c = blockIdx.x, tid = threadIdx.x;
    const int ng = N / groups, cnt = ng * L;
    const float ga = gamma ? gamma[c] : 1.f, be = beta ? beta[c] : 0.f;
    float rm = running_mean ? running_mean[c] : 0.f, rv = running_var ? running_var[c] : 1.f;
    for (int g = 0; g < groups; ++g) {
        const int64_t base = ((int64_t)g * ng * C + c) * L;
        float zv[BN_EPT], mv[BN_EPT];
        float s = 0.f;
#pragma unroll
        for (int u = 0; u < BN_EPT; ++u) {
            const int e = tid + u * nthr;
            zv[u] = 0.f; mv[u] = 1.f;
            if (e < cnt) {
                const int n = e / L, l = e - n * L;
                const int64_t i = base + (int64_t)n * C * L + l;
                zv[u] = z[i];
                if (mask) mv[u] = mask[i];
                s += zv[u];
            }
        }
        const float mean = block_sum_any(s, sm) / (float)cnt;
        float q = 0.f;
#pragma unroll
        for (int u = 0; u < BN_EPT; ++u)
            if (tid + u * nthr < cnt) {
                const float d = zv[u] - mean;
                q += d * d;
            }
        const float var = block_sum_any(q, sm) / (float)cnt;
        const float rstd = rsqrtf(var + eps);
#pragma unroll
        for (int u = 0; u < BN_EPT; ++u) {
            const int e = tid + u * nthr;
            if (e < cnt) {
                const int n = e / L, l = e - n * L;
                float v = (zv[u] - mean) * rstd * ga + be;
                if (relu) v = fmaxf(v, 0.f);
                if (mask) v *= mv[u];
                y[base + (int64_t)n * C * L + l] = v;
            }
        }
        if (tid == 0) {
            stats[((int64_t)g * C + c) * 2 + 0] = mean;
            stats[((int64_t)g * C + c) * 2 + 1] = rstd;
        }
        rm = (1.f - momentum) * rm + momentum * mean;
        rv = (1.f - momentum) * rv + momentum * var * ((float)cnt / (float)max(cnt - 1, 1));
    }
    if (tid == 0) {
        if (running_mean) running_mean[c] = rm;
        if (running_var) running_var[c] = rv;
        if (num_batches_tracked && c == 0) *num_batches_tracked += groups;
    }
}

__global__ void __launch_bounds__(1024) bn_act_bwd_reg_kernel(const float* __restrict__ z, const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, const float* __restrict__ mask,
                                                             const float* __restrict__ stats, const float* __restrict__ gy,
                                                             int N, int C, int L, int groups, int relu,
                                                             float* __restrict__ dz, float* __restrict__ dgamma, float* __restrict__ dbeta) {
    IGCN_PDL_SYNC();
    __shared__ float sm[33];
    const int nthr = blockDim.x, c = blockIdx.x, tid = threadIdx.x;
    const int ng = N / groups, cnt = ng * L;
    const float ga = gamma ? gamma[c] : 1.f, be = beta ? beta[c] : 0.f;
    float dga = 0.f, dbe = 0.f;
    for (int g = 0; g < groups; ++g) {
        const int64_t base = ((int64_t)g * ng * C + c) * L;
        const float mean = stats[((int64_t)g * C + c) * 2 + 0], rstd = stats[((int64_t)g * C + c) * 2 + 1];
        float xh[BN_EPT], dv[BN_EPT];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int u = 0; u < BN_EPT; ++u) {
            const int e = tid + u * nthr;
            xh[u] = 0.f; dv[u] = 0.f;
            if (e < cnt) {
                const int n = e / L, l = e - n * L;
                const int64_t i = base + (int64_t)n * C * L + l;
                xh[u] = (z[i] - mean) * rstd;
                float d = gy[i];
                if (mask) d *= mask[i];
                if (relu && !(xh[u] * ga + be > 0.f)) d = 0.f;
                dv[u] = d;
                s1 += d;
                s2 += d * xh[u];
            }
        }
        s1 = block_sum_any(s1, sm);
        s2 = block_sum_any(s2, sm);
        const float m1 = s1 / (float)cnt, m2 = s2 / (float)cnt;
#pragma unroll
        for (int u = 0; u < BN_EPT; ++u) {
            const int e = tid + u * nthr;
            if (e < cnt) {
                const int n = e / L, l = e - n * L;
                dz[base + (int64_t)n * C * L + l] = ga * rstd * (dv[u] - m1 - xh[u] * m2);
            }
        }
        dga += s2;
        dbe += s1;
    }
    if (tid == 0) {
        if (dgamma) dgamma[c] = dga;
        if (dbeta) dbeta[c] = dbe;
    }
}

// Two stacked passes (groups = 2, the benchmarked step): the two groups of a channel are independent until the running-statistics
// update, so each half of the CTA takes one group and they run side by side -- half the serial length of the kernels above
// (CUPTI: 8.8 us -> see profiles/).  The halves share the block barriers; sums are per half, warp order fixed.
constexpr int BN_EPT2 = 16;

__device__ __forceinline__ float half_sum(float v, float* sm /* 32 floats */, int half, int wph /* warps per half */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) sm[warp] = v;
    __syncthreads();
    float t = 0.f;
    for (int w = 0; w < wph; ++w) t += sm[half * wph + w];
    return t;
}

__global__ void __launch_bounds__(1024) bn_act_fwd_pair_kernel(const float* __restrict__ z, const float* __restrict__ gamma,
                                                              const float* __restrict__ beta, const float* __restrict__ mask,
                                                              int N, int C, int L, float eps, float momentum, int relu,
                                                              float* __restrict__ running_mean, float* __restrict__ running_var,
                                                              long long* __restrict__ num_batches_tracked,
                                                              float* __restrict__ y, float* __restrict__ stats) {
    IGCN_PDL_SYNC();
    __shared__ float sm[33];
    __shared__ float res[4];                       // mean, var of group 0 ; mean, var of group 1
    const int nh = blockDim.x >> 1, c = blockIdx.x, g = threadIdx.x >= nh, tid = threadIdx.x - g * nh, wph = nh >> 5;
    const int ng = N >> 1, cnt = ng * L;
    const float ga = gamma ? gamma[c] : 1.f, be = beta ? beta[c] : 0.f;
    const int64_t base = ((int64_t)g * ng * C + c) * L;
    float zv[BN_EPT2], mv[BN_EPT2];
    int off[BN_EPT2];
    float s = 0.f;
#pragma unroll
    for (int u = 0; u < BN_EPT2; ++u) {
        const int e = tid + u * nh;
        zv[u] = 0.f; mv[u] = 1.f; off[u] = 0;
        if (e < cnt) {
            off[u] = bn_off(e, L, C * L);
            const int64_t i = base + off[u];
            zv[u] = z[i];
            if (mask) mv[u] = mask[i];
            s += zv[u];
        }
    }
    const float mean = half_sum(s, sm, g, wph) / (float)cnt;
    float q = 0.f;
#pragma unroll
    for (int u = 0; u < BN_EPT2; ++u)
        if (tid + u * nh < cnt) {
            const float d = zv[u] - mean;
            q += d * d;
        }
    const float var = half_sum(q, sm, g, wph) / (float)cnt;
    const float rstd = rsqrtf(var + eps);
#pragma unroll
    for (int u = 0; u < BN_EPT2; ++u) {
        const int e = tid + u * nh;
        if (e < cnt) {
            float v = (zv[u] - mean) * rstd * ga + be;
            if (relu) v = fmaxf(v, 0.f);
            if (mask) v *= mv[u];
            y[base + off[u]] = v;
        }
    }
    if (tid == 0) {
        stats[((int64_t)g * C + c) * 2 + 0] = mean;
        stats[((int64_t)g * C + c) * 2 + 1] = rstd;
        res[2 * g] = mean;
        res[2 * g + 1] = var;
    }
    __syncthreads();
    if (threadIdx.x == 0) {                        // the running buffers receive the two updates in pass order
        float rm = running_mean ? running_mean[c] : 0.f, rv = running_var ? running_var[c] : 1.f;
        const float unb = (float)cnt / (float)max(cnt - 1, 1);
        for (int k = 0; k < 2; ++k) {
            rm = (1.f - momentum) * rm + momentum * res[2 * k];
            rv = (1.f - momentum) * rv + momentum * res[2 * k + 1] * unb;
        }
        if (running_mean) running_mean[c] = rm;
        if (running_var) running_var[c] = rv;
        if (num_batches_tracked && c == 0) *num_batches_tracked += 2;
    }
}

__global__ void __launch_bounds__(1024) bn_act_bwd_pair_kernel(const float* __restrict__ z, const float* __restrict__ gamma,
                                                              const float* __restrict__ beta, const float* __restrict__ mask,
                                                              const float* __restrict__ stats, const float* __restrict__ gy,
                                                              int N, int C, int L, int relu,
                                                              float* __restrict__ dz, float* __restrict__ dgamma, float* __restrict__ dbeta) {
    IGCN_PDL_SYNC();
    __shared__ float sm[33];
    __shared__ float res[4];
    const int nh = blockDim.x >> 1, c = blockIdx.x, g = threadIdx.x >= nh, tid = threadIdx.x - g * nh, wph = nh >> 5;
    const int ng = N >> 1, cnt = ng * L;
    const float ga = gamma ? gamma[c] : 1.f, be = beta ? beta[c] : 0.f;
    const int64_t base = ((int64_t)g * ng * C + c) * L;
    const float mean = stats[((int64_t)g * C + c) * 2 + 0], rstd = stats[((int64_t)g * C + c) * 2 + 1];
    float xh[BN_EPT2], dv[BN_EPT2];
    int off[BN_EPT2];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int u = 0; u < BN_EPT2; ++u) {
        const int e = tid + u * nh;
        xh[u] = 0.f; dv[u] = 0.f; off[u] = 0;
        if (e < cnt) {
            off[u] = bn_off(e, L, C * L);
            const int64_t i = base + off[u];
            xh[u] = (z[i] - mean) * rstd;
            float d = gy[i];
            if (mask) d *= mask[i];
            if (relu && !(xh[u] * ga + be > 0.f)) d = 0.f;
            dv[u] = d;
            s1 += d;
            s2 += d * xh[u];
        }
    }
    s1 = half_sum(s1, sm, g, wph);
    s2 = half_sum(s2, sm, g, wph);
    const float m1 = s1 / (float)cnt, m2 = s2 / (float)cnt;
#pragma unroll
    for (int u = 0; u < BN_EPT2; ++u) {
        const int e = tid + u * nh;
        if (e < cnt) dz[base + off[u]] = ga * rstd * (dv[u] - m1 - xh[u] * m2);
    }
    if (tid == 0) {
        res[2 * g] = s2;
        res[2 * g + 1] = s1;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (dgamma) dgamma[c] = res[0] + res[2];
        if (dbeta) dbeta[c] = res[1] + res[3];
    }
}

// ---- per-node read-out Linear fused with its BatchNorm (kernel/go_model.py:117-131: conc_for_attention, conc -> B, conc_D -> B_D) ----
// z[n][c][l] = sum_k x[n][c][k] W[l][k] feeds BatchNorm1d over the channel (GO term) axis c directly, so the CTA of a channel computes
// its z values from the channel's (N x K) slice of x, staged once in shared memory -- z is never written or read, and the backward
// recomputes it the same way.  One launch each way instead of two / three (the read-out of the attention tokens is on the step's
// critical path twice).  Two stacked passes (one per half of the CTA, as bn_act_*_pair_kernel), L = 1 or 32, K <= 8.
constexpr int LB_MAXK = 8, LB_EPT = 16, LB_MAXCL = 8;
// A channel's rows are split over the `CL` CTAs of a thread-block cluster (CL = 1, 2, 4 or 8: 19 channels alone would use 19 of 148
// SMs); the CTAs exchange their partial statistics through distributed shared memory and combine them in rank order, so every CTA of
// the cluster holds bit-identical (mean, rstd).  Rank r owns rows [r * rpr, (r + 1) * rpr) of EACH pass, rpr = ceil(N/2 / CL).
struct LbGeo {
    int c, rank, ng, rpr, r0, nl;
};

__device__ __forceinline__ LbGeo lb_geo(int N, int CL) {
    LbGeo q;
    q.c = blockIdx.x / CL;
    q.rank = blockIdx.x - q.c * CL;
    q.ng = N >> 1;
    q.rpr = (q.ng + CL - 1) / CL;
    q.r0 = min(q.ng, q.rank * q.rpr);
    q.nl = min(q.ng, q.r0 + q.rpr) - q.r0;
    return q;
}

__device__ __forceinline__ void lb_stage(const float* __restrict__ x, const float* __restrict__ W, int C, int L, int K, const LbGeo& q,
                                         float* Ws, float* xs) {
    for (int i = threadIdx.x; i < L * K; i += blockDim.x) Ws[i] = W[i];
    const int per = q.rpr * K, tot = 2 * per;                // this CTA's slice of x: pass-major, rows of K floats (row stride C * K)
    for (int i0 = 0; i0 < tot; i0 += 8 * blockDim.x) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int i = i0 + threadIdx.x + u * blockDim.x;
            v[u] = 0.f;
            if (i < tot) {
                const int g = i >= per, j = i - g * per, n = j / K, k = j - n * K;
                if (n < q.nl) v[u] = x[((int64_t)(g * q.ng + q.r0 + n) * C + q.c) * K + k];
            }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int i = i0 + threadIdx.x + u * blockDim.x;
            if (i < tot) xs[i] = v[u];
        }
    }
    __syncthreads();
}

__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// value `idx` of the exchange block `xch` (this CTA's shared memory) as CTA `r` of the cluster holds it
__device__ __forceinline__ float cluster_peek(const float* xch, int idx, int r) {
    const uint32_t local = (uint32_t)__cvta_generic_to_shared(xch + idx);
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(r));
    float v;
    asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(remote) : "memory");
    return v;
}

// EPT = elements per thread (2, 4, 8 or 16: the host picks the smallest that covers a CTA's share -- the L = 1 read-outs need 2, and
// with a fixed 16 every thread issued eight times the instructions it needed: 21 us instead of 9 for the backward, ncu)
template <int EPT>
__global__ void __launch_bounds__(1024) lin_bn_act_fwd_pair_kernel(const float* __restrict__ x, const float* __restrict__ W,
                                                                  const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                  const float* __restrict__ mask, int N, int C, int L, int K, int CL, float eps,
                                                                  float momentum, int relu, float* __restrict__ running_mean,
                                                                  float* __restrict__ running_var, long long* __restrict__ num_batches_tracked,
                                                                  float* __restrict__ y, float* __restrict__ stats) {
    IGCN_PDL_SYNC();
    extern __shared__ float lbs[];
    __shared__ float sm[33];
    __shared__ float res[4];
    __shared__ float xch[4];                       // (sum, M2 about the CTA's own mean) of pass 0 ; of pass 1
    float* Ws = lbs;
    float* xs = lbs + L * K;
    const LbGeo q = lb_geo(N, CL);
    const int nh = blockDim.x >> 1, c = q.c, g = threadIdx.x >= nh, tid = threadIdx.x - g * nh, wph = nh >> 5;
    const int cnt = q.ng * L, cl = q.nl * L, lsh = __ffs(L) - 1;
    lb_stage(x, W, C, L, K, q, Ws, xs);
    const float ga = gamma ? gamma[c] : 1.f, be = beta ? beta[c] : 0.f;
    const int64_t base = ((int64_t)(g * q.ng + q.r0) * C + c) * L;
    float zv[EPT], mv[EPT];
    float s = 0.f;
#pragma unroll
    for (int u = 0; u < EPT; ++u) {
        const int e = tid + u * nh;
        zv[u] = 0.f; mv[u] = 1.f;
        if (e < cl) {
            const int n = e >> lsh, l = e & (L - 1);
            const float* xr = xs + (g * q.rpr + n) * K;
            const float* wr = Ws + l * K;
            float z = 0.f;
            for (int k = 0; k < K; ++k) z = fmaf(xr[k], wr[k], z);
            zv[u] = z;
            if (mask) mv[u] = mask[base + (int64_t)n * C * L + l];
            s += z;
        }
    }
    const float s_own = half_sum(s, sm, g, wph);
    const float m_own = cl > 0 ? s_own / (float)cl : 0.f;
    float qq = 0.f;
#pragma unroll
    for (int u = 0; u < EPT; ++u)
        if (tid + u * nh < cl) {
            const float d = zv[u] - m_own;
            qq += d * d;
        }
    const float q_own = half_sum(qq, sm, g, wph);
    float mean = m_own, var = q_own / (float)cnt;
    if (CL > 1) {
        // Chan's combination of the CTAs' (count, mean, M2): one exchange, no cancellation
        if (tid == 0) { xch[2 * g] = s_own; xch[2 * g + 1] = q_own; }
        cluster_sync_all();
        float sr[LB_MAXCL], qr[LB_MAXCL];
#pragma unroll
        for (int r = 0; r < LB_MAXCL; ++r)
            if (r < CL) { sr[r] = cluster_peek(xch, 2 * g, r); qr[r] = cluster_peek(xch, 2 * g + 1, r); }
        float tot = 0.f;
#pragma unroll
        for (int r = 0; r < LB_MAXCL; ++r)
            if (r < CL) tot += sr[r];
        mean = tot / (float)cnt;
        float m2 = 0.f;
#pragma unroll
        for (int r = 0; r < LB_MAXCL; ++r)
            if (r < CL) {
                const int nr = (min(q.ng, (r + 1) * q.rpr) - min(q.ng, r * q.rpr)) * L;
                const float d = nr > 0 ? sr[r] / (float)nr - mean : 0.f;
                m2 += qr[r] + (float)nr * d * d;
            }
        var = m2 / (float)cnt;
    }
    const float rstd = rsqrtf(var + eps);
#pragma unroll
    for (int u = 0; u < EPT; ++u) {
        const int e = tid + u * nh;
        if (e < cl) {
            const int n = e >> lsh, l = e & (L - 1);
            float v = (zv[u] - mean) * rstd * ga + be;
            if (relu) v = fmaxf(v, 0.f);
            if (mask) v *= mv[u];
            y[base + (int64_t)n * C * L + l] = v;
        }
    }
    if (tid == 0) {
        res[2 * g] = mean;
        res[2 * g + 1] = var;
        if (q.rank == 0) {
            stats[((int64_t)g * C + c) * 2 + 0] = mean;
            stats[((int64_t)g * C + c) * 2 + 1] = rstd;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0 && q.rank == 0) {         // the running buffers receive the two updates in pass order
        float rm = running_mean ? running_mean[c] : 0.f, rv = running_var ? running_var[c] : 1.f;
        const float unb = (float)cnt / (float)max(cnt - 1, 1);
        for (int k = 0; k < 2; ++k) {
            rm = (1.f - momentum) * rm + momentum * res[2 * k];
            rv = (1.f - momentum) * rv + momentum * res[2 * k + 1] * unb;
        }
        if (running_mean) running_mean[c] = rm;
        if (running_var) running_var[c] = rv;
        if (num_batches_tracked && c == 0) *num_batches_tracked += 2;
    }
    if (CL > 1) cluster_sync_all();                // no CTA leaves while a peer may still read its exchange block
}

// dx[n][c][k] = sum_l dz[n][c][l] W[l][k] (complete inside the row's CTA); partial dW[l][k] of CTA (c, rank) = sum over its rows of
// dz[n][c][l] x[n][c][k]
template <int EPT>
__global__ void __launch_bounds__(1024) lin_bn_act_bwd_pair_kernel(const float* __restrict__ x, const float* __restrict__ W,
                                                                  const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                  const float* __restrict__ mask, const float* __restrict__ stats,
                                                                  const float* __restrict__ gy, int N, int C, int L, int K, int CL, int relu,
                                                                  float* __restrict__ dx, float* __restrict__ partials /* (C * CL, L*K) */,
                                                                  float* __restrict__ dgamma, float* __restrict__ dbeta) {
    IGCN_PDL_SYNC();
    extern __shared__ float lbs[];
    __shared__ float sm[33];
    __shared__ float res[4];
    __shared__ float xch[4];                       // (sum dy, sum dy * xhat) of pass 0 ; of pass 1
    const LbGeo q = lb_geo(N, CL);
    float* Ws = lbs;
    float* xs = lbs + L * K;
    float* red = xs + 2 * q.rpr * K;               // (K, blockDim.x): the threads' dW partials
    float* dzs = red + K * blockDim.x;             // L = 32 only: (2 * rpr, 33) dz, padded rows (conflict-free row walks)
    const int nt = blockDim.x, nh = nt >> 1, c = q.c, g = threadIdx.x >= nh, tid = threadIdx.x - g * nh, wph = nh >> 5;
    const int lane = threadIdx.x & 31;
    const int cnt = q.ng * L, cl = q.nl * L, lsh = __ffs(L) - 1;
    lb_stage(x, W, C, L, K, q, Ws, xs);
    const float ga = gamma ? gamma[c] : 1.f, be = beta ? beta[c] : 0.f;
    const int64_t base = ((int64_t)(g * q.ng + q.r0) * C + c) * L;
    const float mean = stats[((int64_t)g * C + c) * 2 + 0], rstd = stats[((int64_t)g * C + c) * 2 + 1];
    float xh[EPT], dv[EPT];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int u = 0; u < EPT; ++u) {
        const int e = tid + u * nh;
        xh[u] = 0.f; dv[u] = 0.f;
        if (e < cl) {
            const int n = e >> lsh, l = e & (L - 1);
            const float* xr = xs + (g * q.rpr + n) * K;
            const float* wr = Ws + l * K;
            float z = 0.f;
            for (int k = 0; k < K; ++k) z = fmaf(xr[k], wr[k], z);
            xh[u] = (z - mean) * rstd;
            const int64_t i = base + (int64_t)n * C * L + l;
            float d = gy[i];
            if (mask) d *= mask[i];
            if (relu && !(xh[u] * ga + be > 0.f)) d = 0.f;
            dv[u] = d;
            s1 += d;
            s2 += d * xh[u];
        }
    }
    s1 = half_sum(s1, sm, g, wph);
    s2 = half_sum(s2, sm, g, wph);
    if (CL > 1) {
        if (tid == 0) { xch[2 * g] = s1; xch[2 * g + 1] = s2; }
        cluster_sync_all();
        float a1[LB_MAXCL], a2[LB_MAXCL];
#pragma unroll
        for (int r = 0; r < LB_MAXCL; ++r)
            if (r < CL) { a1[r] = cluster_peek(xch, 2 * g, r); a2[r] = cluster_peek(xch, 2 * g + 1, r); }
        s1 = 0.f; s2 = 0.f;
#pragma unroll
        for (int r = 0; r < LB_MAXCL; ++r)
            if (r < CL) { s1 += a1[r]; s2 += a2[r]; }
    }
    const float m1 = s1 / (float)cnt, m2 = s2 / (float)cnt;
    float accw[LB_MAXK];
#pragma unroll
    for (int k = 0; k < LB_MAXK; ++k) accw[k] = 0.f;
    const int l = tid & (L - 1);                   // nh is a multiple of L: a thread's elements share l
#pragma unroll
    for (int u = 0; u < EPT; ++u) {
        const int e = tid + u * nh;
        const bool on = e < cl;                    // uniform per warp when L = 32 (cl is a multiple of 32)
        const int n = on ? (e >> lsh) : 0;
        const float dz = on ? ga * rstd * (dv[u] - m1 - xh[u] * m2) : 0.f;
        const float* xr = xs + (g * q.rpr + n) * K;
#pragma unroll
        for (int k = 0; k < LB_MAXK; ++k)
            if (k < K) {
                accw[k] = fmaf(dz, xr[k], accw[k]);
                if (L == 1 && on) dx[((int64_t)(g * q.ng + q.r0 + n) * C + c) * K + k] = dz * Ws[k];
            }
        if (L == 32 && on) dzs[(g * q.rpr + n) * 33 + l] = dz;
    }
    if (L == 32) {
        // dx[n][k] = sum_l dz[n][l] W[l][k]: a thread per (n, k) walks the 32 outputs of its row in shared memory (a warp-shuffle sum per
        // element and k was 800 instructions per thread: the first version of this kernel took 40 us)
        __syncthreads();
        const int per = q.rpr * K;
        for (int o = threadIdx.x; o < 2 * per; o += nt) {
            const int gg = o >= per, j = o - gg * per, n = j / K, k = j - n * K;
            if (n < q.nl) {
                const float* dr = dzs + (gg * q.rpr + n) * 33;
                float t = 0.f;
#pragma unroll 8
                for (int jj = 0; jj < 32; ++jj) t = fmaf(dr[jj], Ws[jj * K + k], t);
                dx[((int64_t)(gg * q.ng + q.r0 + n) * C + c) * K + k] = t;
            }
        }
    }
#pragma unroll
    for (int k = 0; k < LB_MAXK; ++k)
        if (k < K) {
            if (L == 1) {                          // every thread holds a partial of the same dW[0][k]: warp sums first
                const float v = warp_sum(accw[k]);
                if (lane == 0) red[k * nt + (threadIdx.x >> 5)] = v;
            } else {
                red[k * nt + threadIdx.x] = accw[k];
            }
        }
    if (tid == 0) {
        res[2 * g] = s2;
        res[2 * g + 1] = s1;
    }
    __syncthreads();
    for (int o = threadIdx.x; o < L * K; o += nt) {   // dW[l][k] of this CTA's rows: the partials with this l, in thread (warp) order
        const int lo = o / K, k = o - lo * K;
        float t = 0.f;
        if (L == 1)
            for (int w = 0; w < (nt >> 5); ++w) t += red[k * nt + w];
        else
            for (int j = lo; j < nt; j += L) t += red[k * nt + j];
        partials[(int64_t)blockIdx.x * L * K + o] = t;
    }
    if (threadIdx.x == 0 && q.rank == 0) {
        if (dgamma) dgamma[c] = res[0] + res[2];
        if (dbeta) dbeta[c] = res[1] + res[3];
    }
    if (CL > 1) cluster_sync_all();
}

// ---- loss_probability -----------------------------------------------------------------------------------------------------
// segments: 0 = sigmoid(prob) (n0), 1 = p_e (n1, already a probability), 2 = sigmoid(snps_prob) (n2)
// loss = sum_seg  [ c_l1[seg] * sum|p| + c_en[seg] * sum -(p log(p+eps) + (1-p) log(1-p+eps)) ] / n_seg
struct MaskLossArgs {
    const float* p[3];
    int64_t n[3];
    float c_l1[3], c_en[3];
    int logit[3];
    float eps;
};

__device__ __forceinline__ float mask_loss_term(const MaskLossArgs& a, int seg, int64_t i) {
    float p = a.p[seg][i];
    if (a.logit[seg]) p = 1.f / (1.f + expf(-p));
    const float en = -(p * logf(p + a.eps) + (1.f - p) * logf(1.f - p + a.eps));
    return (a.c_l1[seg] * fabsf(p) + a.c_en[seg] * en) / (float)a.n[seg];
}

__global__ void __launch_bounds__(256) mask_loss_partial_kernel(MaskLossArgs a, float* __restrict__ partials) {
    IGCN_PDL_SYNC();
    __shared__ float sm[9];
    float s = 0.f;
    const int64_t stride = (int64_t)gridDim.x * 256;
    for (int seg = 0; seg < 3; ++seg)
        for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < a.n[seg]; i += stride) s += mask_loss_term(a, seg, i);
    s = block_sum_256(s, sm);
    if (threadIdx.x == 0) partials[blockIdx.x] = s;
}

// out[0] = sum of partials[0..n) in index order; one block
__global__ void __launch_bounds__(256) sum_partials_kernel(const float* __restrict__ partials, int n, float scale, float* __restrict__ out) {
    IGCN_PDL_SYNC();
    __shared__ float sm[9];
    float s = 0.f;
    for (int i = threadIdx.x; i < n; i += 256) s += partials[i];
    s = block_sum_256(s, sm);
    if (threadIdx.x == 0) out[0] = s * scale;
}

// d p_raw[i] = g * dloss/dp * (dp/draw)
__global__ void __launch_bounds__(256) mask_loss_bwd_kernel(MaskLossArgs a, const float* __restrict__ g_out, float* __restrict__ d0,
                                                            float* __restrict__ d1, float* __restrict__ d2) {
    IGCN_PDL_SYNC();
    const float g = g_out[0];
    float* d[3] = {d0, d1, d2};
    const int64_t stride = (int64_t)gridDim.x * 256;
    for (int seg = 0; seg < 3; ++seg) {
        if (!d[seg]) continue;
        for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < a.n[seg]; i += stride) {
            float p = a.p[seg][i];
            if (a.logit[seg]) p = 1.f / (1.f + expf(-p));
            const float sgn = p > 0.f ? 1.f : (p < 0.f ? -1.f : 0.f);
            const float den = -(logf(p + a.eps) + p / (p + a.eps) - logf(1.f - p + a.eps) - (1.f - p) / (1.f - p + a.eps));
            float v = g * (a.c_l1[seg] * sgn + a.c_en[seg] * den) / (float)a.n[seg];
            if (a.logit[seg]) v *= p * (1.f - p);
            d[seg][i] = v;
        }
    }
}

// ---- dot product with a fixed summation order -----------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dot_partial_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t n,
                                                          float* __restrict__ partials) {
    IGCN_PDL_SYNC();
    __shared__ float sm[9];
    float s = 0.f;
    const int64_t stride = (int64_t)gridDim.x * 256 * 4;
    const int64_t n4 = n & ~(int64_t)3;
    for (int64_t i = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 4; i < n4; i += stride) {
        const float4 x = *reinterpret_cast<const float4*>(a + i), y = *reinterpret_cast<const float4*>(b + i);
        s += x.x * y.x + x.y * y.y + x.z * y.z + x.w * y.w;
    }
    if (blockIdx.x == 0 && threadIdx.x < (int)(n - n4)) s += a[n4 + threadIdx.x] * b[n4 + threadIdx.x];
    s = block_sum_256(s, sm);
    if (threadIdx.x == 0) partials[blockIdx.x] = s;
}

// out[i] = (a[i] + b[i]) + c[i]  (c may be null): the gradient of a tensor with several consumers, in one launch
__global__ void __launch_bounds__(256) sum3_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ c,
                                                   int64_t n, float* __restrict__ out) {
    IGCN_PDL_SYNC();
    const int64_t stride = (int64_t)gridDim.x * 256 * 4;
    const int64_t n4 = n & ~(int64_t)3;
    for (int64_t i = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 4; i < n4; i += stride) {
        const float4 x = *reinterpret_cast<const float4*>(a + i);
        const float4 y = *reinterpret_cast<const float4*>(b + i);
        float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c) z = *reinterpret_cast<const float4*>(c + i);
        float4 r;
        r.x = (x.x + y.x) + z.x; r.y = (x.y + y.y) + z.y; r.z = (x.z + y.z) + z.z; r.w = (x.w + y.w) + z.w;
        *reinterpret_cast<float4*>(out + i) = r;
    }
    if (blockIdx.x == 0 && threadIdx.x < (int)(n - n4)) {
        const int64_t i = n4 + threadIdx.x;
        out[i] = (a[i] + b[i]) + (c ? c[i] : 0.f);
    }
}

// out[i] = a[i] * (s[0] * scale)
__global__ void __launch_bounds__(256) scale_by_scalar_kernel(const float* __restrict__ a, const float* __restrict__ s, float scale,
                                                              int64_t n, float* __restrict__ out) {
    IGCN_PDL_SYNC();
    const float f = s[0] * scale;
    const int64_t stride = (int64_t)gridDim.x * 256 * 4;
    const int64_t n4 = n & ~(int64_t)3;
    for (int64_t i = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 4; i < n4; i += stride) {
        float4 x = *reinterpret_cast<const float4*>(a + i);
        x.x *= f; x.y *= f; x.z *= f; x.w *= f;
        *reinterpret_cast<float4*>(out + i) = x;
    }
    if (blockIdx.x == 0 && threadIdx.x < (int)(n - n4)) out[n4 + threadIdx.x] = a[n4 + threadIdx.x] * f;
}


// ---- skinny linear: z[r][l] = sum_k x[r][k] W[l][k], Kin <= 32, Lout <= 64 (the per-node read-out projections of the GO
//      network, kernel/go_model.py:117-131: 5 -> dim_snps_atten, 5 -> 1, 2 -> 1 over batch * nodes rows).  cuBLAS runs the weight
//      gradient of these (Lout x rows)(rows x Kin) shapes on one CTA (35 us at 9 728 rows).
constexpr int SK_MAXK = 32, SK_MAXL = 64, SK_ROWS = 64, SK_ACC = (SK_MAXK * SK_MAXL + 255) / 256;

__global__ void __launch_bounds__(256) skinny_linear_fwd_kernel(const float* __restrict__ x, const float* __restrict__ W, int64_t rows, int Kin,
                                                                int Lout, float* __restrict__ z) {
    IGCN_PDL_SYNC();
    __shared__ float Ws[SK_MAXL * SK_MAXK];
    for (int i = threadIdx.x; i < Lout * Kin; i += 256) Ws[i] = W[i];
    __syncthreads();
    const uint32_t total = (uint32_t)(rows * Lout);              // < 2^31 (checked by the host wrapper)
    for (uint32_t idx = blockIdx.x * 256u + threadIdx.x; idx < total; idx += gridDim.x * 256u) {
        const uint32_t r = idx / (uint32_t)Lout;                 // 32-bit: the 64-bit division was most of this kernel
        const int l = (int)(idx - r * (uint32_t)Lout);
        const float* xr = x + (int64_t)r * Kin;
        const float* wr = Ws + l * Kin;
        float v = 0.f;
        // the row's inputs are loaded 8 at a time before the FMAs (a load -> FMA loop over a runtime Kin costs one L2 trip per k)
        for (int k0 = 0; k0 < Kin; k0 += 8) {
            float xv[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) xv[u] = (k0 + u < Kin) ? xr[k0 + u] : 0.f;
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (k0 + u < Kin) v = fmaf(xv[u], wr[k0 + u], v);
        }
        z[idx] = v;
    }
}

// dx[r][k] = sum_l dz[r][l] W[l][k] ; partial dW[l][k] of this CTA = sum over its rows of dz[r][l] x[r][k]
__global__ void __launch_bounds__(256) skinny_linear_bwd_kernel(const float* __restrict__ x, const float* __restrict__ W,
                                                                const float* __restrict__ dz, int64_t rows, int Kin, int Lout,
                                                                float* __restrict__ dx, float* __restrict__ partials) {
    IGCN_PDL_SYNC();
    __shared__ float Ws[SK_MAXL * SK_MAXK];
    __shared__ float dzs[SK_ROWS * SK_MAXL];
    __shared__ float xs[SK_ROWS * SK_MAXK];
    const int tid = threadIdx.x, LK = Lout * Kin;
    for (int i = tid; i < LK; i += 256) Ws[i] = W[i];
    float acc[SK_ACC];                                       // dW entries tid, tid + 256, ...
#pragma unroll
    for (int u = 0; u < SK_ACC; ++u) acc[u] = 0.f;
    for (int64_t r0 = (int64_t)blockIdx.x * SK_ROWS; r0 < rows; r0 += (int64_t)gridDim.x * SK_ROWS) {
        const int nr = (int)min((int64_t)SK_ROWS, rows - r0);
        __syncthreads();
        {   // <= 8 values per thread and array (SK_ROWS * 32 / 256): all loads in flight before the first store
            float a[8], b[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int i = tid + u * 256;
                a[u] = i < nr * Lout ? dz[r0 * Lout + i] : 0.f;
                b[u] = i < nr * Kin ? x[r0 * Kin + i] : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int i = tid + u * 256;
                if (i < nr * Lout) dzs[i] = a[u];
                if (i < nr * Kin) xs[i] = b[u];
            }
            for (int i = tid + 8 * 256; i < nr * Lout; i += 256) dzs[i] = dz[r0 * Lout + i];      // Lout > 32 only
        }
        __syncthreads();
        if (dx)
            for (int i = tid; i < nr * Kin; i += 256) {
                const int r = i / Kin, k = i - r * Kin;
                float v = 0.f;
                for (int l = 0; l < Lout; ++l) v = fmaf(dzs[r * Lout + l], Ws[l * Kin + k], v);
                dx[r0 * Kin + i] = v;
            }
#pragma unroll
        for (int u = 0; u < SK_ACC; ++u) {
            const int i = tid + u * 256;
            if (i < LK) {
                const int l = i / Kin, k = i - l * Kin;
                float v = acc[u];
                for (int r = 0; r < nr; ++r) v = fmaf(dzs[r * Lout + l], xs[r * Kin + k], v);
                acc[u] = v;
            }
        }
    }
#pragma unroll
    for (int u = 0; u < SK_ACC; ++u) {
        const int i = tid + u * 256;
        if (i < LK) partials[(int64_t)blockIdx.x * LK + i] = acc[u];
    }
}

static int blocks_for(int64_t n, int per_block) {
    int64_t b = (n + per_block - 1) / per_block;
    const int64_t cap = (int64_t)sm_count() * 4;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

}  // namespace igcn

namespace igcn {
// ---- BatchNorm1d in EVAL mode (+ ReLU): a per-channel affine map with the running statistics -- the inference path of the read-out
//      heads (kernel/go_model.py:117-146 under model.eval(), as eval_acc / eval_loss / eval_scores run it) ------------------------
__global__ void __launch_bounds__(256) bn_eval_act_kernel(const float* __restrict__ z, const float* __restrict__ gamma,
                                                         const float* __restrict__ beta, const float* __restrict__ rm,
                                                         const float* __restrict__ rv, int64_t total, int C, int L, float eps, int relu,
                                                         const float* __restrict__ gy, float* __restrict__ y) {
    IGCN_PDL_SYNC();
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
        const int c = (int)((i / L) % C);
        const float sc = (gamma ? gamma[c] : 1.f) * rsqrtf(rv[c] + eps);
        const float v = fmaf(z[i] - rm[c], sc, beta ? beta[c] : 0.f);
        if (gy)                                   // backward: d z = g * scale where the unit is active
            y[i] = (!relu || v > 0.f) ? gy[i] * sc : 0.f;
        else
            y[i] = relu ? fmaxf(v, 0.f) : v;
    }
}
}  // namespace igcn

using namespace igcn;

/* y = act((z - running_mean) / sqrt(running_var + eps) * gamma + beta) for z (N, C, L); with g_y given, y receives d z instead. */
extern "C" int igcn_bn_eval_act(const float* z, const float* gamma, const float* beta, const float* running_mean, const float* running_var,
                                int64_t N, int64_t C, int64_t L, double eps, int64_t relu, const float* g_y, float* y, void* stream) {
    IGCN_REQUIRE(N >= 0 && C > 0 && L > 0, IGCN_ERR_BAD_ARG, "bn_eval_act: bad sizes");
    if (N == 0) return IGCN_OK;
    IGCN_REQUIRE(z && running_mean && running_var && y, IGCN_ERR_BAD_ARG, "bn_eval_act: null pointer");
    const int64_t total = N * C * L;
    igcn::launch_k(bn_eval_act_kernel, dim3((unsigned)blocks_for(total, 256)), dim3(256), 0, (cudaStream_t)stream, z, gamma, beta, running_mean, running_var, total,
                                                                                          (int)C, (int)L, (float)eps, (int)relu, g_y, y);
    IGCN_CHECK_LAUNCH("bn_eval_act");
    return IGCN_OK;
}

extern "C" int igcn_bn_act_fwd(const float* z, const float* gamma, const float* beta, const float* mask, int64_t N, int64_t C, int64_t L,
                               int64_t groups, double eps, double momentum, int64_t relu, float* running_mean, float* running_var,
                               long long* num_batches_tracked, float* y, float* stats, void* stream) {
    IGCN_REQUIRE(z && y && stats, IGCN_ERR_BAD_ARG, "bn_act_fwd: null pointer");
    IGCN_REQUIRE(N > 0 && C > 0 && L > 0 && groups > 0 && N % groups == 0, IGCN_ERR_BAD_ARG, "bn_act_fwd: bad sizes (N=%lld C=%lld L=%lld groups=%lld)",
                 (long long)N, (long long)C, (long long)L, (long long)groups);
    IGCN_REQUIRE((N / groups) * L > 1, IGCN_ERR_UNSUPPORTED, "bn_act_fwd: training-mode BatchNorm needs more than one value per channel");
    const int64_t cnt = (N / groups) * L;
    const int nthr = cnt >= 4096 ? 1024 : (cnt >= 1024 ? 512 : 256);
    if (groups == 2 && cnt <= (int64_t)BN_EPT2 * (nthr / 2) && nthr >= 128)
        igcn::launch_k(bn_act_fwd_pair_kernel, dim3((unsigned)C), dim3(nthr), 0, (cudaStream_t)stream, z, gamma, beta, mask, (int)N, (int)C, (int)L, (float)eps,
                                                                              (float)momentum, (int)relu, running_mean, running_var,
                                                                              num_batches_tracked, y, stats);
    else if (cnt <= (int64_t)BN_EPT * nthr)
        igcn::launch_k(bn_act_fwd_reg_kernel, dim3((unsigned)C), dim3(nthr), 0, (cudaStream_t)stream, z, gamma, beta, mask, (int)N, (int)C, (int)L, (int)groups, (float)eps,
                                                                             (float)momentum, (int)relu, running_mean, running_var,
                                                                             num_batches_tracked, y, stats);
    else
        igcn::launch_k(bn_act_fwd_kernel, dim3((unsigned)C), dim3(nthr), 0, (cudaStream_t)stream, z, gamma, beta, mask, (int)N, (int)C, (int)L, (int)groups, (float)eps,
                                                                         (float)momentum, (int)relu, running_mean, running_var,
                                                                         num_batches_tracked, y, stats);
    IGCN_CHECK_LAUNCH("bn_act_fwd");
    return IGCN_OK;
}

extern "C" int igcn_bn_act_bwd(const float* z, const float* gamma, const float* beta, const float* mask, const float* stats, const float* g_y,
                               int64_t N, int64_t C, int64_t L, int64_t groups, int64_t relu, float* dz, float* dgamma, float* dbeta,
                               void* stream) {
    IGCN_REQUIRE(z && stats && g_y && dz, IGCN_ERR_BAD_ARG, "bn_act_bwd: null pointer");
    IGCN_REQUIRE(N > 0 && C > 0 && L > 0 && groups > 0 && N % groups == 0, IGCN_ERR_BAD_ARG, "bn_act_bwd: bad sizes");
    const int64_t cnt = (N / groups) * L;
    const int nthr = cnt >= 4096 ? 1024 : (cnt >= 1024 ? 512 : 256);
    if (groups == 2 && cnt <= (int64_t)BN_EPT2 * (nthr / 2) && nthr >= 128)
        igcn::launch_k(bn_act_bwd_pair_kernel, dim3((unsigned)C), dim3(nthr), 0, (cudaStream_t)stream, z, gamma, beta, mask, stats, g_y, (int)N, (int)C, (int)L,
                                                                              (int)relu, dz, dgamma, dbeta);
    else if (cnt <= (int64_t)BN_EPT * nthr)
        igcn::launch_k(bn_act_bwd_reg_kernel, dim3((unsigned)C), dim3(nthr), 0, (cudaStream_t)stream, z, gamma, beta, mask, stats, g_y, (int)N, (int)C, (int)L,
                                                                             (int)groups, (int)relu, dz, dgamma, dbeta);
    else
        igcn::launch_k(bn_act_bwd_kernel, dim3((unsigned)C), dim3(nthr), 0, (cudaStream_t)stream, z, gamma, beta, mask, stats, g_y, (int)N, (int)C, (int)L, (int)groups,
                                                                         (int)relu, dz, dgamma, dbeta);
    IGCN_CHECK_LAUNCH("bn_act_bwd");
    return IGCN_OK;
}

static int lb_threads(int64_t cnt) { return cnt >= 4096 ? 1024 : (cnt >= 1024 ? 512 : 256); }

// CTAs per channel (cluster size): halve the rows per CTA while a CTA keeps >= 32 rows of each pass and the grid stays within ~2 CTAs per SM
static int lb_cluster(int64_t N, int64_t C, int64_t L) {
    // opt-in (IGCN_LIN_BN_CLUSTER=1): measured SLOWER inside the step (0.447 vs 0.401 ms on the same box) although the kernels alone are
    // not -- a cluster needs its 8 CTAs co-scheduled in one GPC, which the concurrent SGCN / attention kernels rarely leave free
    const char* e = getenv("IGCN_LIN_BN_CLUSTER");
    if (!(e && e[0] == '1') || L != 32) return 1;
    int cl = 1;
    while (cl < LB_MAXCL && (N / 2) / (cl * 2) >= 32 && C * cl * 2 <= (int64_t)sm_count() * 2) cl *= 2;
    return cl;
}

static int64_t lb_rows(int64_t N, int cl) { return (N / 2 + cl - 1) / cl; }

static int lb_ept(int64_t cnt, int nthr) {
    const int64_t need = (cnt + nthr / 2 - 1) / (nthr / 2);
    return need <= 2 ? 2 : (need <= 4 ? 4 : (need <= 8 ? 8 : LB_EPT));
}

template <typename... Params, typename... Args>
static void launch_cluster(void (*kernel)(Params...), unsigned grid, unsigned block, size_t smem, cudaStream_t st, int cl, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    int na = 0;
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = (unsigned)cl;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
    if (igcn::pdl_enabled()) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    cudaLaunchKernelEx(&cfg, kernel, static_cast<Params>(args)...);
}

extern "C" int64_t igcn_lin_bn_act_supported(int64_t N, int64_t C, int64_t L, int64_t K, int64_t groups) {
    if (groups != 2 || N < 2 || (N & 1) || C < 1 || K < 1 || K > LB_MAXK || !(L == 1 || L == 32)) return 0;
    const int cl = lb_cluster(N, C, L);
    const int64_t rpr = lb_rows(N, cl), cnt = rpr * L;
    const int nthr = lb_threads(cnt);
    if ((N / 2) * L <= 1 || cnt > (int64_t)LB_EPT * (nthr / 2)) return 0;
    return (L * K + 2 * rpr * K + K * nthr + (L == 32 ? 2 * rpr * 33 : 0)) * 4 <= 160 * 1024 ? 1 : 0;
}

/* rows of the `partials` workspace of igcn_lin_bn_act_bwd (one per CTA) */
extern "C" int64_t igcn_lin_bn_act_partial_rows(int64_t N, int64_t C, int64_t L, int64_t K) {
    (void)K;
    return C * lb_cluster(N, C, L);
}

extern "C" int igcn_lin_bn_act_fwd(const float* x, const float* W, const float* gamma, const float* beta, const float* mask, int64_t N,
                                   int64_t C, int64_t L, int64_t K, int64_t groups, double eps, double momentum, int64_t relu,
                                   float* running_mean, float* running_var, long long* num_batches_tracked, float* y, float* stats,
                                   void* stream) {
    IGCN_REQUIRE(x && W && y && stats, IGCN_ERR_BAD_ARG, "lin_bn_act_fwd: null pointer");
    IGCN_REQUIRE(igcn_lin_bn_act_supported(N, C, L, K, groups), IGCN_ERR_UNSUPPORTED,
                 "lin_bn_act_fwd: shape (N=%lld C=%lld L=%lld K=%lld groups=%lld) not supported (see igcn_lin_bn_act_supported)", (long long)N,
                 (long long)C, (long long)L, (long long)K, (long long)groups);
    const int cl = lb_cluster(N, C, L);
    const int64_t rpr = lb_rows(N, cl);
    const int nthr = lb_threads(rpr * L);
    const size_t smem = sizeof(float) * (size_t)(L * K + 2 * rpr * K);
    int rc = 0;
    if (smem > 48 * 1024) {
        if ((rc = allow_smem(lin_bn_act_fwd_pair_kernel<2>, smem, "lin_bn_act_fwd"))) return rc;
        if ((rc = allow_smem(lin_bn_act_fwd_pair_kernel<4>, smem, "lin_bn_act_fwd"))) return rc;
        if ((rc = allow_smem(lin_bn_act_fwd_pair_kernel<8>, smem, "lin_bn_act_fwd"))) return rc;
        if ((rc = allow_smem(lin_bn_act_fwd_pair_kernel<16>, smem, "lin_bn_act_fwd"))) return rc;
    }
    const int ept = lb_ept(rpr * L, nthr);
#define IGCN_LB_FWD(E)                                                                                                                       \
    launch_cluster(lin_bn_act_fwd_pair_kernel<E>, (unsigned)(C * cl), (unsigned)nthr, smem, (cudaStream_t)stream, cl, x, W, gamma, beta, mask, \
                   (int)N, (int)C, (int)L, (int)K, cl, (float)eps, (float)momentum, (int)relu, running_mean, running_var,                     \
                   num_batches_tracked, y, stats)
    if (ept == 2) IGCN_LB_FWD(2);
    else if (ept == 4) IGCN_LB_FWD(4);
    else if (ept == 8) IGCN_LB_FWD(8);
    else IGCN_LB_FWD(16);
#undef IGCN_LB_FWD
    IGCN_CHECK_LAUNCH("lin_bn_act_fwd");
    return IGCN_OK;
}

extern "C" int igcn_lin_bn_act_bwd(const float* x, const float* W, const float* gamma, const float* beta, const float* mask, const float* stats,
                                   const float* g_y, int64_t N, int64_t C, int64_t L, int64_t K, int64_t groups, int64_t relu, float* dx,
                                   float* partials, float* dW, float* dgamma, float* dbeta, void* stream) {
    IGCN_REQUIRE(x && W && stats && g_y && dx && partials && dW, IGCN_ERR_BAD_ARG, "lin_bn_act_bwd: null pointer");
    IGCN_REQUIRE(igcn_lin_bn_act_supported(N, C, L, K, groups), IGCN_ERR_UNSUPPORTED, "lin_bn_act_bwd: shape not supported");
    const int cl = lb_cluster(N, C, L);
    const int64_t rpr = lb_rows(N, cl);
    const int nthr = lb_threads(rpr * L);
    const size_t smem = sizeof(float) * (size_t)(L * K + 2 * rpr * K + K * nthr + (L == 32 ? 2 * rpr * 33 : 0));
    int rc = 0;
    if (smem > 48 * 1024) {
        if ((rc = allow_smem(lin_bn_act_bwd_pair_kernel<2>, smem, "lin_bn_act_bwd"))) return rc;
        if ((rc = allow_smem(lin_bn_act_bwd_pair_kernel<4>, smem, "lin_bn_act_bwd"))) return rc;
        if ((rc = allow_smem(lin_bn_act_bwd_pair_kernel<8>, smem, "lin_bn_act_bwd"))) return rc;
        if ((rc = allow_smem(lin_bn_act_bwd_pair_kernel<16>, smem, "lin_bn_act_bwd"))) return rc;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int ept = lb_ept(rpr * L, nthr);
#define IGCN_LB_BWD(E)                                                                                                                     \
    launch_cluster(lin_bn_act_bwd_pair_kernel<E>, (unsigned)(C * cl), (unsigned)nthr, smem, st, cl, x, W, gamma, beta, mask, stats, g_y, (int)N, \
                   (int)C, (int)L, (int)K, cl, (int)relu, dx, partials, dgamma, dbeta)
    if (ept == 2) IGCN_LB_BWD(2);
    else if (ept == 4) IGCN_LB_BWD(4);
    else if (ept == 8) IGCN_LB_BWD(8);
    else IGCN_LB_BWD(16);
#undef IGCN_LB_BWD
    IGCN_CHECK_LAUNCH("lin_bn_act_bwd");
    igcn::launch_k(reduce_partials_kernel, dim3((unsigned)((L * K + 31) / 32)), dim3(reduce_threads(C * cl)), 0, st, partials, (int)(C * cl),
                   (int)(L * K), dW);
    IGCN_CHECK_LAUNCH("lin_bn_act_reduce");
    return IGCN_OK;
}

static int fill_mask_loss(MaskLossArgs& a, const float* prob, int64_t n_prob, const float* p_e, int64_t n_e, const float* snps_prob,
                          int64_t n_snps, const float* host_coef, double eps) {
    IGCN_REQUIRE(host_coef, IGCN_ERR_BAD_ARG, "mask_loss: null host_coef");
    IGCN_REQUIRE(n_prob >= 0 && n_e >= 0 && n_snps >= 0, IGCN_ERR_BAD_ARG, "mask_loss: negative size");
    IGCN_REQUIRE((n_prob == 0 || prob) && (n_e == 0 || p_e) && (n_snps == 0 || snps_prob), IGCN_ERR_BAD_ARG, "mask_loss: null pointer");
    a.p[0] = prob; a.p[1] = p_e; a.p[2] = snps_prob;
    a.n[0] = n_prob; a.n[1] = n_e; a.n[2] = n_snps;
    a.logit[0] = 1; a.logit[1] = 0; a.logit[2] = 1;
    // host_coef = {lamda_x_l1, lamda_e_l1, lamda_x_ent, lamda_e_ent} (sgcn_hyperparameters.py:18-21)
    a.c_l1[0] = host_coef[0]; a.c_l1[1] = host_coef[1]; a.c_l1[2] = host_coef[0];
    a.c_en[0] = host_coef[2]; a.c_en[1] = host_coef[3]; a.c_en[2] = host_coef[2];
    a.eps = (float)eps;
    return IGCN_OK;
}

extern "C" int64_t igcn_reduce_blocks(int64_t n) { return blocks_for(n, 256 * 8); }

extern "C" int igcn_mask_loss_fwd(const float* prob, int64_t n_prob, const float* p_e, int64_t n_e, const float* snps_prob, int64_t n_snps,
                                  const float* host_coef, double eps, float* partials, int64_t n_partials, float* loss, void* stream) {
    MaskLossArgs a;
    int rc = fill_mask_loss(a, prob, n_prob, p_e, n_e, snps_prob, n_snps, host_coef, eps);
    if (rc) return rc;
    IGCN_REQUIRE(partials && loss && n_partials >= 1, IGCN_ERR_BAD_ARG, "mask_loss_fwd: null workspace");
    cudaStream_t st = (cudaStream_t)stream;
    igcn::launch_k(mask_loss_partial_kernel, dim3((unsigned)n_partials), dim3(256), 0, st, a, partials);
    IGCN_CHECK_LAUNCH("mask_loss_partial");
    igcn::launch_k(sum_partials_kernel, dim3(1), dim3(256), 0, st, partials, (int)n_partials, 1.f, loss);
    IGCN_CHECK_LAUNCH("sum_partials");
    return IGCN_OK;
}

extern "C" int igcn_mask_loss_bwd(const float* prob, int64_t n_prob, const float* p_e, int64_t n_e, const float* snps_prob, int64_t n_snps,
                                  const float* host_coef, double eps, const float* g_loss, float* d_prob, float* d_pe, float* d_snps_prob,
                                  void* stream) {
    MaskLossArgs a;
    int rc = fill_mask_loss(a, prob, n_prob, p_e, n_e, snps_prob, n_snps, host_coef, eps);
    if (rc) return rc;
    IGCN_REQUIRE(g_loss, IGCN_ERR_BAD_ARG, "mask_loss_bwd: null g_loss");
    const int64_t nmax = n_prob > n_e ? (n_prob > n_snps ? n_prob : n_snps) : (n_e > n_snps ? n_e : n_snps);
    igcn::launch_k(mask_loss_bwd_kernel, dim3(blocks_for(nmax, 256)), dim3(256), 0, (cudaStream_t)stream, a, g_loss, d_prob, d_pe, d_snps_prob);
    IGCN_CHECK_LAUNCH("mask_loss_bwd");
    return IGCN_OK;
}

extern "C" int igcn_dot(const float* a, const float* b, int64_t n, double scale, float* partials, int64_t n_partials, float* out, void* stream) {
    IGCN_REQUIRE(a && b && partials && out && n >= 0 && n_partials >= 1, IGCN_ERR_BAD_ARG, "dot: bad argument");
    IGCN_REQUIRE((((uintptr_t)a | (uintptr_t)b) & 15) == 0, IGCN_ERR_BAD_ARG, "dot: operands must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    igcn::launch_k(dot_partial_kernel, dim3((unsigned)n_partials), dim3(256), 0, st, a, b, n, partials);
    IGCN_CHECK_LAUNCH("dot_partial");
    igcn::launch_k(sum_partials_kernel, dim3(1), dim3(256), 0, st, partials, (int)n_partials, (float)scale, out);
    IGCN_CHECK_LAUNCH("sum_partials");
    return IGCN_OK;
}

extern "C" int igcn_sum3(const float* a, const float* b, const float* c, int64_t n, float* out, void* stream) {
    IGCN_REQUIRE(a && b && out && n >= 0, IGCN_ERR_BAD_ARG, "sum3: bad argument");
    IGCN_REQUIRE((((uintptr_t)a | (uintptr_t)b | (uintptr_t)c | (uintptr_t)out) & 15) == 0, IGCN_ERR_BAD_ARG, "sum3: operands must be 16-byte aligned");
    if (n == 0) return IGCN_OK;
    igcn::launch_k(sum3_kernel, dim3(blocks_for(n, 256 * 4)), dim3(256), 0, (cudaStream_t)stream, a, b, c, n, out);
    IGCN_CHECK_LAUNCH("sum3");
    return IGCN_OK;
}

extern "C" int igcn_scale_by_scalar(const float* a, const float* s, double scale, int64_t n, float* out, void* stream) {
    IGCN_REQUIRE(a && s && out && n >= 0, IGCN_ERR_BAD_ARG, "scale_by_scalar: bad argument");
    IGCN_REQUIRE((((uintptr_t)a | (uintptr_t)out) & 15) == 0, IGCN_ERR_BAD_ARG, "scale_by_scalar: operands must be 16-byte aligned");
    igcn::launch_k(scale_by_scalar_kernel, dim3(blocks_for(n, 256 * 4)), dim3(256), 0, (cudaStream_t)stream, a, s, (float)scale, n, out);
    IGCN_CHECK_LAUNCH("scale_by_scalar");
    return IGCN_OK;
}

static int skinny_check(const char* who, int64_t rows, int64_t Kin, int64_t Lout) {
    IGCN_REQUIRE(rows >= 0 && Kin >= 1 && Lout >= 1, IGCN_ERR_BAD_ARG, "%s: bad size", who);
    IGCN_REQUIRE(Kin <= SK_MAXK && Lout <= SK_MAXL, IGCN_ERR_UNSUPPORTED, "%s: in_features <= %d and out_features <= %d only (got %lld, %lld)", who,
                 SK_MAXK, SK_MAXL, (long long)Kin, (long long)Lout);
    IGCN_REQUIRE(rows * Lout < (int64_t)1 << 31 && rows * Kin < (int64_t)1 << 31, IGCN_ERR_UNSUPPORTED, "%s: more than 2^31 elements", who);
    return IGCN_OK;
}

extern "C" int64_t igcn_skinny_linear_bwd_ctas(int64_t rows) {
    int64_t n = (rows + SK_ROWS - 1) / SK_ROWS;
    const int64_t cap = (int64_t)sm_count() * 2;
    if (n > cap) n = cap;
    return n < 1 ? 1 : n;
}

extern "C" int igcn_skinny_linear_fwd(const float* x, const float* W, int64_t rows, int64_t Kin, int64_t Lout, float* z, void* stream) {
    int rc = skinny_check("skinny_linear_fwd", rows, Kin, Lout);
    if (rc) return rc;
    IGCN_REQUIRE(x && W && z, IGCN_ERR_BAD_ARG, "skinny_linear_fwd: null pointer");
    if (rows == 0) return IGCN_OK;
    igcn::launch_k(skinny_linear_fwd_kernel, dim3(blocks_for(rows * Lout, 256 * 4)), dim3(256), 0, (cudaStream_t)stream, x, W, rows, (int)Kin, (int)Lout, z);
    IGCN_CHECK_LAUNCH("skinny_linear_fwd");
    return IGCN_OK;
}

extern "C" int igcn_skinny_linear_bwd(const float* x, const float* W, const float* dz, int64_t rows, int64_t Kin, int64_t Lout, float* dx,
                                      float* partials, int64_t n_cta, float* dW, void* stream) {
    int rc = skinny_check("skinny_linear_bwd", rows, Kin, Lout);
    if (rc) return rc;
    IGCN_REQUIRE(x && W && dz && partials && dW, IGCN_ERR_BAD_ARG, "skinny_linear_bwd: null pointer");
    IGCN_REQUIRE(n_cta == igcn_skinny_linear_bwd_ctas(rows), IGCN_ERR_BAD_ARG, "skinny_linear_bwd: n_cta must be igcn_skinny_linear_bwd_ctas()");
    cudaStream_t st = (cudaStream_t)stream;
    const int LK = (int)(Lout * Kin);
    if (rows == 0) {
        cudaMemsetAsync(dW, 0, sizeof(float) * LK, st);
        return IGCN_OK;
    }
    igcn::launch_k(skinny_linear_bwd_kernel, dim3((unsigned)n_cta), dim3(256), 0, st, x, W, dz, rows, (int)Kin, (int)Lout, dx, partials);
    IGCN_CHECK_LAUNCH("skinny_linear_bwd");
    igcn::launch_k(reduce_partials_kernel, dim3((LK + 31) / 32), dim3(reduce_threads((int)n_cta)), 0, st, partials, (int)n_cta, LK, dW);
    IGCN_CHECK_LAUNCH("skinny_linear_reduce");
    return IGCN_OK;
}

// =============================================================================================================================
// Tail of the training step (kernel/train_eval_sgcn_img_snps.py:521-544, kernel/sgcn_img_snp.py:286-305): three more fused pieces.
// =============================================================================================================================
namespace igcn {

// ---- SNP mask of the stacked passes: rows [0,B) = snps, rows [B,2B) = snps * sigmoid(snps_prob)  (sgcn_img_snp.py:147-148) ----
__global__ void __launch_bounds__(256) snp_mask_pair_fwd_kernel(const float* __restrict__ snps, const float* __restrict__ p, int B, int S,
                                                                float* __restrict__ out) {
    IGCN_PDL_SYNC();
    // grid.y strides the rows, threads walk the columns: no integer division per element
    const int64_t n = (int64_t)B * S;
    for (int s = blockIdx.x * 256 + threadIdx.x; s < S; s += gridDim.x * 256) {
        const float sg = sigmoidf_(p[s]);
        for (int b = blockIdx.y; b < B; b += gridDim.y) {
            const float v = snps[(int64_t)b * S + s];
            out[(int64_t)b * S + s] = v;
            out[n + (int64_t)b * S + s] = v * sg;
        }
    }
}
// d p[s] = sig'(p[s]) * sum_b g[B + b][s] * snps[b][s]: 32 columns per CTA, 32 warps take every 32nd row, warp partials added in order
__global__ void __launch_bounds__(1024) snp_mask_pair_bwd_kernel(const float* __restrict__ snps, const float* __restrict__ p,
                                                                 const float* __restrict__ g, int B, int S, float* __restrict__ dp) {
    IGCN_PDL_SYNC();
    __shared__ float part[32][33];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int s = blockIdx.x * 32 + lane;
    const float* g2 = g + (int64_t)B * S;
    float acc = 0.f;
    if (s < S) {
        // 32 warps take every 32nd row, four rows (eight loads) in flight per thread: with 8 warps and a load -> FMA loop this
        // kernel was 32 dependent L2 trips long (8 us for 54 columns x 256 rows)
        int b = warp;
        for (; b + 96 < B; b += 128) {
            float gv[4], xv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                gv[u] = g2[(int64_t)(b + 32 * u) * S + s];
                xv[u] = snps[(int64_t)(b + 32 * u) * S + s];
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) acc = fmaf(gv[u], xv[u], acc);
        }
        for (; b < B; b += 32) acc = fmaf(g2[(int64_t)b * S + s], snps[(int64_t)b * S + s], acc);
    }
    part[warp][lane] = acc;
    __syncthreads();
    if (warp == 0 && s < S) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 32; ++w) t += part[w][lane];
        const float sg = sigmoidf_(p[s]);
        dp[s] = t * sg * (1.f - sg);
    }
}

// ---- output heads: logp = log_softmax(lin2(h1 * m1)), reg = lin2_regr(h2 * m2)  (sgcn_img_snp.py:290-291,300-301) -------------
constexpr int HEAD_MAXC = 8;
struct HeadArgs {
    const float *h1, *m1, *h2, *m2;      // (rows, K); masks may be null
    const float *W1, *b1, *W2, *b2;      // (C1, K), (C1), (C2, K), (C2)
    int rows, K, C1, C2;
};

// warp per row: lane l owns features l and l + 32 (K <= 64), the C <= 8 outputs are warp sums
__global__ void __launch_bounds__(256) heads_fwd_kernel(HeadArgs a, float* __restrict__ logp, float* __restrict__ reg) {
    IGCN_PDL_SYNC();
    const int K = a.K, C1 = a.C1, C2 = a.C2, lane = threadIdx.x & 31;
    const int warp_g = blockIdx.x * 8 + (threadIdx.x >> 5), nwarp = gridDim.x * 8;
    float w1[HEAD_MAXC][2], w2[HEAD_MAXC][2];
#pragma unroll
    for (int c = 0; c < HEAD_MAXC; ++c)
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int k = lane + 32 * u;
            w1[c][u] = (c < C1 && k < K) ? a.W1[c * K + k] : 0.f;
            w2[c][u] = (c < C2 && k < K) ? a.W2[c * K + k] : 0.f;
        }
    for (int r = warp_g; r < a.rows; r += nwarp) {
        float x1[2], x2[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int k = lane + 32 * u;
            const int64_t i = (int64_t)r * K + k;
            x1[u] = k < K ? (a.m1 ? a.h1[i] * a.m1[i] : a.h1[i]) : 0.f;
            x2[u] = k < K ? (a.m2 ? a.h2[i] * a.m2[i] : a.h2[i]) : 0.f;
        }
        float z1[HEAD_MAXC], z2[HEAD_MAXC];
#pragma unroll
        for (int c = 0; c < HEAD_MAXC; ++c) {
            z1[c] = c < C1 ? warp_sum(fmaf(x1[1], w1[c][1], x1[0] * w1[c][0])) + a.b1[c] : 0.f;
            z2[c] = c < C2 ? warp_sum(fmaf(x2[1], w2[c][1], x2[0] * w2[c][0])) + a.b2[c] : 0.f;
        }
        float mx = -INFINITY;
#pragma unroll
        for (int c = 0; c < HEAD_MAXC; ++c)
            if (c < C1) mx = fmaxf(mx, z1[c]);
        float den = 0.f;
#pragma unroll
        for (int c = 0; c < HEAD_MAXC; ++c)
            if (c < C1) den += expf(z1[c] - mx);
        const float lse = mx + logf(den);
#pragma unroll
        for (int c = 0; c < HEAD_MAXC; ++c) {
            if (c < C1 && lane == c) logp[(int64_t)r * C1 + c] = z1[c] - lse;
            if (c < C2 && lane == c) reg[(int64_t)r * C2 + c] = z2[c];
        }
    }
}

// dz1 = g_logp - softmax * sum(g_logp) ; dz2 = g_reg ; dh = m * (dz W) ; per-CTA partial of [dW1 | db1 | dW2 | db2].
// Warp per row; a lane accumulates d W[c][l], d W[c][l + 32] of the rows its warp sees, the 8 warps are added in order.
__global__ void __launch_bounds__(256) heads_bwd_kernel(HeadArgs a, const float* __restrict__ logp, const float* __restrict__ g_logp,
                                                        const float* __restrict__ g_reg, float* __restrict__ dh1, float* __restrict__ dh2,
                                                        float* __restrict__ partials) {
    IGCN_PDL_SYNC();
    __shared__ float red[8][2 * HEAD_MAXC * 64 + 2 * HEAD_MAXC];
    const int K = a.K, C1 = a.C1, C2 = a.C2, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int warp_g = blockIdx.x * 8 + warp, nwarp = gridDim.x * 8;
    const int P = C1 * K + C1 + C2 * K + C2;
    float w1[HEAD_MAXC][2], w2[HEAD_MAXC][2], aw1[HEAD_MAXC][2], aw2[HEAD_MAXC][2], ab1[HEAD_MAXC], ab2[HEAD_MAXC];
#pragma unroll
    for (int c = 0; c < HEAD_MAXC; ++c) {
        ab1[c] = ab2[c] = 0.f;
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int k = lane + 32 * u;
            w1[c][u] = (c < C1 && k < K) ? a.W1[c * K + k] : 0.f;
            w2[c][u] = (c < C2 && k < K) ? a.W2[c * K + k] : 0.f;
            aw1[c][u] = aw2[c][u] = 0.f;
        }
    }
    for (int r = warp_g; r < a.rows; r += nwarp) {
        float dz1[HEAD_MAXC], dz2[HEAD_MAXC], gs = 0.f;
#pragma unroll
        for (int c = 0; c < HEAD_MAXC; ++c) {
            dz1[c] = (c < C1 && g_logp) ? g_logp[(int64_t)r * C1 + c] : 0.f;
            dz2[c] = (c < C2 && g_reg) ? g_reg[(int64_t)r * C2 + c] : 0.f;
            gs += dz1[c];
        }
#pragma unroll
        for (int c = 0; c < HEAD_MAXC; ++c)
            if (c < C1) dz1[c] -= expf(logp[(int64_t)r * C1 + c]) * gs;
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int k = lane + 32 * u;
            if (k >= K) continue;
            const int64_t i = (int64_t)r * K + k;
            const float m1 = a.m1 ? a.m1[i] : 1.f, m2 = a.m2 ? a.m2[i] : 1.f;
            const float x1 = a.h1[i] * m1, x2 = a.h2[i] * m2;
            float v1 = 0.f, v2 = 0.f;
#pragma unroll
            for (int c = 0; c < HEAD_MAXC; ++c) {
                v1 = fmaf(dz1[c], w1[c][u], v1);
                v2 = fmaf(dz2[c], w2[c][u], v2);
                aw1[c][u] = fmaf(dz1[c], x1, aw1[c][u]);
                aw2[c][u] = fmaf(dz2[c], x2, aw2[c][u]);
            }
            if (dh1) dh1[i] = v1 * m1;
            if (dh2) dh2[i] = v2 * m2;
        }
#pragma unroll
        for (int c = 0; c < HEAD_MAXC; ++c) {
            ab1[c] += dz1[c];
            ab2[c] += dz2[c];
        }
    }
#pragma unroll
    for (int c = 0; c < HEAD_MAXC; ++c) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int k = lane + 32 * u;
            if (c < C1 && k < K) red[warp][c * K + k] = aw1[c][u];
            if (c < C2 && k < K) red[warp][C1 * K + C1 + c * K + k] = aw2[c][u];
        }
        if (lane == 0) {
            if (c < C1) red[warp][C1 * K + c] = ab1[c];
            if (c < C2) red[warp][C1 * K + C1 + C2 * K + c] = ab2[c];
        }
    }
    __syncthreads();
    for (int e = threadIdx.x; e < P; e += 256) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += red[w][e];
        partials[(int64_t)blockIdx.x * P + e] = t;
    }
}

// ---- the scalar train() back-propagates (its four data terms): ---------------------------------------------------------------
// loss = c_reg * mean((reg - target)^2) + c_rec * sum((xhat - snps)^2) + c_prob * loss_prob + c_clu * quad
// reg (2, n_reg), target (n_reg) broadcast over the two passes; xhat (2, n_rec), snps (n_rec) broadcast.  One CTA, fixed order.
__global__ void __launch_bounds__(1024) step_loss_fwd_kernel(const float* __restrict__ reg, const float* __restrict__ target, int64_t n_reg,
                                                             const float* __restrict__ xhat, const float* __restrict__ snps, int64_t n_rec,
                                                             const float* __restrict__ loss_prob, const float* __restrict__ quad, float c_reg,
                                                             float c_rec, float c_prob, float c_clu, float* __restrict__ out) {
    IGCN_PDL_SYNC();
    __shared__ float sm[33];
    float s1 = 0.f, s2 = 0.f;
    for (int pass = 0; pass < 2; ++pass) {
        const float* r = reg + pass * n_reg;
        const float* x = xhat + pass * n_rec;
#pragma unroll 4
        for (int64_t i = threadIdx.x; i < n_reg; i += 1024) {
            const float d = r[i] - target[i];
            s1 = fmaf(d, d, s1);
        }
        for (int64_t i0 = 0; i0 < n_rec; i0 += 8 * 1024) {         // 16 loads per thread in flight, then the arithmetic: this kernel is
            float xa[8], sa[8];                                     // one CTA on the step's critical path (12 us as a load -> FMA loop)
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int64_t i = i0 + threadIdx.x + u * 1024;
                xa[u] = i < n_rec ? x[i] : 0.f;
                sa[u] = i < n_rec ? snps[i] : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const float d = xa[u] - sa[u];
                s2 = fmaf(d, d, s2);
            }
        }
    }
    s1 = block_sum_any(s1, sm);
    s2 = block_sum_any(s2, sm);
    if (threadIdx.x == 0)
        out[0] = c_reg * s1 / (float)(2 * n_reg) + c_rec * s2 + c_prob * (loss_prob ? loss_prob[0] : 0.f) + c_clu * (quad ? quad[0] : 0.f);
}
__global__ void __launch_bounds__(256) step_loss_bwd_kernel(const float* __restrict__ reg, const float* __restrict__ target, int64_t n_reg,
                                                            const float* __restrict__ xhat, const float* __restrict__ snps, int64_t n_rec,
                                                            const float* __restrict__ g, float c_reg, float c_rec, float c_prob, float c_clu,
                                                            float* __restrict__ d_reg, float* __restrict__ d_xhat, float* __restrict__ d_prob,
                                                            float* __restrict__ d_quad) {
    IGCN_PDL_SYNC();
    const float gv = g[0];
    const int64_t stride = (int64_t)gridDim.x * 256, t0 = (int64_t)blockIdx.x * 256 + threadIdx.x;
    const float k1 = gv * c_reg * 2.f / (float)(2 * n_reg), k2 = gv * c_rec * 2.f;
    for (int pass = 0; pass < 2; ++pass) {
        for (int64_t i = t0; i < n_reg; i += stride) d_reg[pass * n_reg + i] = k1 * (reg[pass * n_reg + i] - target[i]);
        for (int64_t i = t0; i < n_rec; i += stride) d_xhat[pass * n_rec + i] = k2 * (xhat[pass * n_rec + i] - snps[i]);
    }
    if (t0 == 0) {
        if (d_prob) d_prob[0] = gv * c_prob;
        if (d_quad) d_quad[0] = gv * c_clu;
    }
}

}  // namespace igcn

extern "C" int igcn_snp_mask_pair_fwd(const float* snps, const float* snps_prob, int64_t B, int64_t S, float* out, void* stream) {
    IGCN_REQUIRE(snps && snps_prob && out && B >= 0 && S > 0, IGCN_ERR_BAD_ARG, "snp_mask_pair_fwd: bad argument");
    if (B == 0) return IGCN_OK;
    const unsigned gx = (unsigned)((S + 255) / 256);
    unsigned gy = (unsigned)(sm_count() * 2 / gx);
    if (gy < 1) gy = 1;
    if (gy > (unsigned)B) gy = (unsigned)B;
    igcn::launch_k(snp_mask_pair_fwd_kernel, dim3(dim3(gx, gy)), dim3(256), 0, (cudaStream_t)stream, snps, snps_prob, (int)B, (int)S, out);
    IGCN_CHECK_LAUNCH("snp_mask_pair_fwd");
    return IGCN_OK;
}
extern "C" int igcn_snp_mask_pair_bwd(const float* snps, const float* snps_prob, const float* g_out, int64_t B, int64_t S, float* d_snps_prob,
                                      void* stream) {
    IGCN_REQUIRE(snps && snps_prob && g_out && d_snps_prob && B >= 0 && S > 0, IGCN_ERR_BAD_ARG, "snp_mask_pair_bwd: bad argument");
    igcn::launch_k(snp_mask_pair_bwd_kernel, dim3((unsigned)((S + 31) / 32)), dim3(1024), 0, (cudaStream_t)stream, snps, snps_prob, g_out, (int)B, (int)S, d_snps_prob);
    IGCN_CHECK_LAUNCH("snp_mask_pair_bwd");
    return IGCN_OK;
}

static int heads_fill(HeadArgs& a, const char* who, const float* h1, const float* m1, const float* h2, const float* m2, const float* W1,
                      const float* b1, const float* W2, const float* b2, int64_t rows, int64_t K, int64_t C1, int64_t C2) {
    IGCN_REQUIRE(h1 && h2 && W1 && b1 && W2 && b2 && rows >= 0, IGCN_ERR_BAD_ARG, "%s: null pointer", who);
    IGCN_REQUIRE(K >= 1 && K <= 64 && C1 >= 1 && C1 <= HEAD_MAXC && C2 >= 1 && C2 <= HEAD_MAXC, IGCN_ERR_UNSUPPORTED,
                 "%s: hidden <= 64 and <= %d outputs per head only (K=%lld, C1=%lld, C2=%lld)", who, HEAD_MAXC, (long long)K, (long long)C1,
                 (long long)C2);
    a.h1 = h1; a.m1 = m1; a.h2 = h2; a.m2 = m2; a.W1 = W1; a.b1 = b1; a.W2 = W2; a.b2 = b2;
    a.rows = (int)rows; a.K = (int)K; a.C1 = (int)C1; a.C2 = (int)C2;
    return IGCN_OK;
}
extern "C" int64_t igcn_heads_bwd_ctas(int64_t rows) {
    int64_t n = (rows + 31) / 32;              // 8 warps per CTA, ~4 rows per warp
    if (n > sm_count()) n = sm_count();
    return n < 1 ? 1 : n;
}
extern "C" int igcn_heads_fwd(const float* h1, const float* m1, const float* h2, const float* m2, const float* W1, const float* b1,
                              const float* W2, const float* b2, int64_t rows, int64_t K, int64_t C1, int64_t C2, float* logp, float* reg,
                              void* stream) {
    HeadArgs a;
    int rc = heads_fill(a, "heads_fwd", h1, m1, h2, m2, W1, b1, W2, b2, rows, K, C1, C2);
    if (rc) return rc;
    IGCN_REQUIRE(logp && reg, IGCN_ERR_BAD_ARG, "heads_fwd: null output");
    if (rows == 0) return IGCN_OK;
    igcn::launch_k(heads_fwd_kernel, dim3((unsigned)igcn_heads_bwd_ctas(rows)), dim3(256), 0, (cudaStream_t)stream, a, logp, reg);
    IGCN_CHECK_LAUNCH("heads_fwd");
    return IGCN_OK;
}
extern "C" int igcn_heads_bwd(const float* h1, const float* m1, const float* h2, const float* m2, const float* W1, const float* b1,
                              const float* W2, const float* b2, const float* logp, const float* g_logp, const float* g_reg, int64_t rows,
                              int64_t K, int64_t C1, int64_t C2, float* dh1, float* dh2, float* partials, int64_t n_cta, float* grads,
                              void* stream) {
    HeadArgs a;
    int rc = heads_fill(a, "heads_bwd", h1, m1, h2, m2, W1, b1, W2, b2, rows, K, C1, C2);
    if (rc) return rc;
    IGCN_REQUIRE(logp && partials && grads, IGCN_ERR_BAD_ARG, "heads_bwd: null pointer");
    IGCN_REQUIRE(n_cta == igcn_heads_bwd_ctas(rows), IGCN_ERR_BAD_ARG, "heads_bwd: n_cta must be igcn_heads_bwd_ctas()");
    cudaStream_t st = (cudaStream_t)stream;
    const int P = (int)(C1 * K + C1 + C2 * K + C2);
    if (rows == 0) {
        cudaMemsetAsync(grads, 0, sizeof(float) * P, st);
        return IGCN_OK;
    }
    igcn::launch_k(heads_bwd_kernel, dim3((unsigned)n_cta), dim3(256), 0, st, a, logp, g_logp, g_reg, dh1, dh2, partials);
    IGCN_CHECK_LAUNCH("heads_bwd");
    igcn::launch_k(reduce_partials_kernel, dim3((P + 31) / 32), dim3(reduce_threads((int)n_cta)), 0, st, partials, (int)n_cta, P, grads);
    IGCN_CHECK_LAUNCH("heads_reduce");
    return IGCN_OK;
}

extern "C" int igcn_step_loss_fwd(const float* reg, const float* target, int64_t n_reg, const float* xhat, const float* snps, int64_t n_rec,
                                  const float* loss_prob, const float* quad, double c_reg, double c_rec, double c_prob, double c_clu,
                                  float* out, void* stream) {
    IGCN_REQUIRE(reg && target && xhat && snps && out && n_reg > 0 && n_rec > 0, IGCN_ERR_BAD_ARG, "step_loss_fwd: bad argument");
    igcn::launch_k(step_loss_fwd_kernel, dim3(1), dim3(1024), 0, (cudaStream_t)stream, reg, target, n_reg, xhat, snps, n_rec, loss_prob, quad, (float)c_reg, (float)c_rec,
                                                              (float)c_prob, (float)c_clu, out);
    IGCN_CHECK_LAUNCH("step_loss_fwd");
    return IGCN_OK;
}
extern "C" int igcn_step_loss_bwd(const float* reg, const float* target, int64_t n_reg, const float* xhat, const float* snps, int64_t n_rec,
                                  const float* g_loss, double c_reg, double c_rec, double c_prob, double c_clu, float* d_reg, float* d_xhat,
                                  float* d_loss_prob, float* d_quad, void* stream) {
    IGCN_REQUIRE(reg && target && xhat && snps && g_loss && d_reg && d_xhat && n_reg > 0 && n_rec > 0, IGCN_ERR_BAD_ARG, "step_loss_bwd: bad argument");
    igcn::launch_k(step_loss_bwd_kernel, dim3(blocks_for(2 * (n_reg > n_rec ? n_reg : n_rec), 256 * 4)), dim3(256), 0, (cudaStream_t)stream, 
        reg, target, n_reg, xhat, snps, n_rec, g_loss, (float)c_reg, (float)c_rec, (float)c_prob, (float)c_clu, d_reg, d_xhat, d_loss_prob, d_quad);
    IGCN_CHECK_LAUNCH("step_loss_bwd");
    return IGCN_OK;
}
