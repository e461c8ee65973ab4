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
    __shared__ float sm[9];
    float s = 0.f;
    for (int i = threadIdx.x; i < n; i += 256) s += partials[i];
    s = block_sum_256(s, sm);
    if (threadIdx.x == 0) out[0] = s * scale;
}

// d p_raw[i] = g * dloss/dp * (dp/draw)
__global__ void __launch_bounds__(256) mask_loss_bwd_kernel(MaskLossArgs a, const float* __restrict__ g_out, float* __restrict__ d0,
                                                            float* __restrict__ d1, float* __restrict__ d2) {
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

// out[i] = a[i] * (s[0] * scale)
__global__ void __launch_bounds__(256) scale_by_scalar_kernel(const float* __restrict__ a, const float* __restrict__ s, float scale,
                                                              int64_t n, float* __restrict__ out) {
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


// ---- skinny linear: z[r][l] = sum_k x[r][k] W[l][k], Kin <= 8, Lout <= 64 (the per-node read-out projections of the GO
//      network, kernel/go_model.py:117-131: 5 -> dim_snps_atten, 5 -> 1, 2 -> 1 over batch * nodes rows).  cuBLAS runs the weight
//      gradient of these (Lout x rows)(rows x Kin) shapes on one CTA (35 us at 9 728 rows).
constexpr int SK_MAXK = 8, SK_MAXL = 64, SK_ROWS = 64;

__global__ void __launch_bounds__(256) skinny_linear_fwd_kernel(const float* __restrict__ x, const float* __restrict__ W, int64_t rows, int Kin,
                                                                int Lout, float* __restrict__ z) {
    __shared__ float Ws[SK_MAXL * SK_MAXK];
    for (int i = threadIdx.x; i < Lout * Kin; i += 256) Ws[i] = W[i];
    __syncthreads();
    const int64_t total = rows * Lout;
    for (int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * 256) {
        const int64_t r = idx / Lout;
        const int l = (int)(idx - r * Lout);
        const float* xr = x + r * Kin;
        float v = 0.f;
        for (int k = 0; k < Kin; ++k) v = fmaf(xr[k], Ws[l * Kin + k], v);
        z[idx] = v;
    }
}

// dx[r][k] = sum_l dz[r][l] W[l][k] ; partial dW[l][k] of this CTA = sum over its rows of dz[r][l] x[r][k]
__global__ void __launch_bounds__(256) skinny_linear_bwd_kernel(const float* __restrict__ x, const float* __restrict__ W,
                                                                const float* __restrict__ dz, int64_t rows, int Kin, int Lout,
                                                                float* __restrict__ dx, float* __restrict__ partials) {
    __shared__ float Ws[SK_MAXL * SK_MAXK];
    __shared__ float dzs[SK_ROWS * SK_MAXL];
    __shared__ float xs[SK_ROWS * SK_MAXK];
    const int tid = threadIdx.x, LK = Lout * Kin;
    for (int i = tid; i < LK; i += 256) Ws[i] = W[i];
    float acc[2] = {0.f, 0.f};                               // dW entries tid and tid + 256 (LK <= 512)
    for (int64_t r0 = (int64_t)blockIdx.x * SK_ROWS; r0 < rows; r0 += (int64_t)gridDim.x * SK_ROWS) {
        const int nr = (int)min((int64_t)SK_ROWS, rows - r0);
        __syncthreads();
        for (int i = tid; i < nr * Lout; i += 256) dzs[i] = dz[r0 * Lout + i];
        for (int i = tid; i < nr * Kin; i += 256) xs[i] = x[r0 * Kin + i];
        __syncthreads();
        if (dx)
            for (int i = tid; i < nr * Kin; i += 256) {
                const int r = i / Kin, k = i - r * Kin;
                float v = 0.f;
                for (int l = 0; l < Lout; ++l) v = fmaf(dzs[r * Lout + l], Ws[l * Kin + k], v);
                dx[r0 * Kin + i] = v;
            }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
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
    for (int u = 0; u < 2; ++u) {
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

using namespace igcn;

extern "C" int igcn_bn_act_fwd(const float* z, const float* gamma, const float* beta, const float* mask, int64_t N, int64_t C, int64_t L,
                               int64_t groups, double eps, double momentum, int64_t relu, float* running_mean, float* running_var,
                               long long* num_batches_tracked, float* y, float* stats, void* stream) {
    IGCN_REQUIRE(z && y && stats, IGCN_ERR_BAD_ARG, "bn_act_fwd: null pointer");
    IGCN_REQUIRE(N > 0 && C > 0 && L > 0 && groups > 0 && N % groups == 0, IGCN_ERR_BAD_ARG, "bn_act_fwd: bad sizes (N=%lld C=%lld L=%lld groups=%lld)",
                 (long long)N, (long long)C, (long long)L, (long long)groups);
    IGCN_REQUIRE((N / groups) * L > 1, IGCN_ERR_UNSUPPORTED, "bn_act_fwd: training-mode BatchNorm needs more than one value per channel");
    const int64_t cnt = (N / groups) * L;
    const int nthr = cnt >= 4096 ? 1024 : (cnt >= 1024 ? 512 : 256);
    bn_act_fwd_kernel<<<(unsigned)C, nthr, 0, (cudaStream_t)stream>>>(z, gamma, beta, mask, (int)N, (int)C, (int)L, (int)groups, (float)eps,
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
    bn_act_bwd_kernel<<<(unsigned)C, nthr, 0, (cudaStream_t)stream>>>(z, gamma, beta, mask, stats, g_y, (int)N, (int)C, (int)L, (int)groups,
                                                                     (int)relu, dz, dgamma, dbeta);
    IGCN_CHECK_LAUNCH("bn_act_bwd");
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
    mask_loss_partial_kernel<<<(unsigned)n_partials, 256, 0, st>>>(a, partials);
    IGCN_CHECK_LAUNCH("mask_loss_partial");
    sum_partials_kernel<<<1, 256, 0, st>>>(partials, (int)n_partials, 1.f, loss);
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
    mask_loss_bwd_kernel<<<blocks_for(nmax, 256), 256, 0, (cudaStream_t)stream>>>(a, g_loss, d_prob, d_pe, d_snps_prob);
    IGCN_CHECK_LAUNCH("mask_loss_bwd");
    return IGCN_OK;
}

extern "C" int igcn_dot(const float* a, const float* b, int64_t n, double scale, float* partials, int64_t n_partials, float* out, void* stream) {
    IGCN_REQUIRE(a && b && partials && out && n >= 0 && n_partials >= 1, IGCN_ERR_BAD_ARG, "dot: bad argument");
    IGCN_REQUIRE((((uintptr_t)a | (uintptr_t)b) & 15) == 0, IGCN_ERR_BAD_ARG, "dot: operands must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    dot_partial_kernel<<<(unsigned)n_partials, 256, 0, st>>>(a, b, n, partials);
    IGCN_CHECK_LAUNCH("dot_partial");
    sum_partials_kernel<<<1, 256, 0, st>>>(partials, (int)n_partials, (float)scale, out);
    IGCN_CHECK_LAUNCH("sum_partials");
    return IGCN_OK;
}

extern "C" int igcn_scale_by_scalar(const float* a, const float* s, double scale, int64_t n, float* out, void* stream) {
    IGCN_REQUIRE(a && s && out && n >= 0, IGCN_ERR_BAD_ARG, "scale_by_scalar: bad argument");
    IGCN_REQUIRE((((uintptr_t)a | (uintptr_t)out) & 15) == 0, IGCN_ERR_BAD_ARG, "scale_by_scalar: operands must be 16-byte aligned");
    scale_by_scalar_kernel<<<blocks_for(n, 256 * 4), 256, 0, (cudaStream_t)stream>>>(a, s, (float)scale, n, out);
    IGCN_CHECK_LAUNCH("scale_by_scalar");
    return IGCN_OK;
}

static int skinny_check(const char* who, int64_t rows, int64_t Kin, int64_t Lout) {
    IGCN_REQUIRE(rows >= 0 && Kin >= 1 && Lout >= 1, IGCN_ERR_BAD_ARG, "%s: bad size", who);
    IGCN_REQUIRE(Kin <= SK_MAXK && Lout <= SK_MAXL, IGCN_ERR_UNSUPPORTED, "%s: in_features <= %d and out_features <= %d only (got %lld, %lld)", who,
                 SK_MAXK, SK_MAXL, (long long)Kin, (long long)Lout);
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
    skinny_linear_fwd_kernel<<<blocks_for(rows * Lout, 256 * 4), 256, 0, (cudaStream_t)stream>>>(x, W, rows, (int)Kin, (int)Lout, z);
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
    skinny_linear_bwd_kernel<<<(unsigned)n_cta, 256, 0, st>>>(x, W, dz, rows, (int)Kin, (int)Lout, dx, partials);
    IGCN_CHECK_LAUNCH("skinny_linear_bwd");
    reduce_partials_kernel<<<(LK + 31) / 32, 256, 0, st>>>(partials, (int)n_cta, LK, dW);
    IGCN_CHECK_LAUNCH("skinny_linear_reduce");
    return IGCN_OK;
}
