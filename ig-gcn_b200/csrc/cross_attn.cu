// Fused imaging <- genetics cross attention (reference: kernel/sgcn_img_snp.py:46,239-241 --
// nn.MultiheadAttention(E, 2, batch_first=True)(query = ROI tokens (B,R,E), key = value = GO tokens (B,M,E)) followed by relu).
//
// torch runs this as ~15 launches forward and ~25 backward; the weight gradients of the projections are (E x B*R) x (B*R x E)
// products that cuBLAS executes on ONE CTA (ncu/torch.profiler: 8 x 53 us per step at B=256 -- more than every igcn kernel
// together).  Here a CTA owns a graph: Q/K/V projections, per-head scores, row softmax, P.V, the output projection and the
// ReLU stay in shared memory (R*M <= 264*300 scores per head fit); projection-weight gradients are accumulated per CTA in shared
// memory across its graphs and reduced in a fixed order afterwards (deterministic, no float atomics).
// Shape-generic (runtime R, M, E, heads); fp32 FFMA.
#include "common.cuh"

namespace igcn {

struct AttnArgs {
    const float* x;     // (B, R, E) queries
    const float* a;     // (B, M, E) keys = values
    const float* Win;   // (3E, E)  in_proj_weight  [Wq ; Wk ; Wv]
    const float* bin;   // (3E)
    const float* Wo;    // (E, E)   out_proj.weight
    const float* bo;    // (E)
    float* y;           // (B, R, E)  fwd output (after ReLU when relu=1)
    const float* yout;  // bwd: forward output (ReLU mask)
    const float* gy;    // bwd: (B, R, E)
    float* dx;          // (B, R, E)
    float* da;          // (B, M, E)
    float* partials;    // (n_cta, P)  P = 3E*E + 3E + E*E + E : [dWin | dbin | dWo | dbo]
    int B, R, M, E, heads, relu, P;
    int Rc;             // query rows staged per chunk
};

struct AttnSmem {
    float *WinT, *WoT, *bin, *bo;   // WinT[k][3E] (k-major), WoT[k][E]
    float *X, *A, *Q, *K, *V, *Pm, *O;   // X, Q, Pm, O hold ONE CHUNK of Rc query rows
    float* tail;
};
__device__ __forceinline__ AttnSmem attn_carve(float* p, int Rc, int M, int E, int heads) {
    AttnSmem s;
    s.WinT = p;  p += 3 * E * E;
    s.WoT = p;   p += E * E;
    s.bin = p;   p += 3 * E;
    s.bo = p;    p += E;
    s.X = p;     p += Rc * E;
    s.A = p;     p += M * E;
    s.Q = p;     p += Rc * (E + 1);   // row stride E+1: rows are walked by different threads at the same column
    s.K = p;     p += M * E;
    s.V = p;     p += M * E;
    s.Pm = p;    p += heads * Rc * M;
    s.O = p;     p += Rc * E;
    s.tail = p;
    return s;
}
static size_t attn_common_floats(int Rc, int M, int E, int heads) {
    return (size_t)4 * E * E + 4 * E + 3 * (size_t)Rc * E + Rc + 3 * (size_t)M * E + (size_t)heads * Rc * M;
}

__device__ __forceinline__ void attn_load_params(const AttnArgs& a, const AttnSmem& s) {
    const int tid = threadIdx.x, nt = blockDim.x, E = a.E;
    for (int i = tid; i < 3 * E * E; i += nt) {
        const int f = i / E, k = i - f * E;
        s.WinT[k * 3 * E + f] = a.Win[i];
    }
    for (int i = tid; i < E * E; i += nt) {
        const int f = i / E, k = i - f * E;
        s.WoT[k * E + f] = a.Wo[i];
    }
    for (int i = tid; i < 3 * E; i += nt) s.bin[i] = a.bin[i];
    for (int i = tid; i < E; i += nt) s.bo[i] = a.bo[i];
    __syncthreads();
}

// K, V of one graph (all M key/value tokens)
__device__ __forceinline__ void attn_keys_values(const AttnArgs& a, const AttnSmem& s, int b) {
    const int tid = threadIdx.x, nt = blockDim.x, M = a.M, E = a.E;
    const float* ab = a.a + (int64_t)b * M * E;
    for (int i = tid; i < M * E; i += nt) s.A[i] = ab[i];
    __syncthreads();
    for (int idx = tid; idx < M * E; idx += nt) {
        const int j = idx / E, f = idx - j * E;
        float ak = s.bin[E + f], av = s.bin[2 * E + f];
#pragma unroll 8
        for (int k = 0; k < E; ++k) {
            const float v = s.A[j * E + k];
            ak = fmaf(v, s.WinT[k * 3 * E + E + f], ak);
            av = fmaf(v, s.WinT[k * 3 * E + 2 * E + f], av);
        }
        s.K[idx] = ak;
        s.V[idx] = av;
    }
    __syncthreads();
}

// Q, P, O for the query rows [r0, r0+rc) of one graph
__device__ __forceinline__ void attn_forward_chunk(const AttnArgs& a, const AttnSmem& s, int b, int r0, int rc) {
    const int tid = threadIdx.x, nt = blockDim.x, R = a.R, M = a.M, E = a.E, H = a.heads, hd = E / H;
    const float scale = rsqrtf((float)hd);
    const float* xb = a.x + ((int64_t)b * R + r0) * E;
    for (int i = tid; i < rc * E; i += nt) s.X[i] = xb[i];
    __syncthreads();
    for (int idx = tid; idx < rc * E; idx += nt) {
        const int i = idx / E, f = idx - i * E;
        float acc = s.bin[f];
#pragma unroll 8
        for (int k = 0; k < E; ++k) acc = fmaf(s.X[i * E + k], s.WinT[k * 3 * E + f], acc);
        s.Q[i * (E + 1) + f] = acc;
    }
    __syncthreads();
    for (int idx = tid; idx < H * rc; idx += nt) {          // one (head, query) row per thread: scores, softmax
        const int h = idx / rc, i = idx - h * rc;
        float* prow = s.Pm + (h * rc + i) * M;
        const float* q = s.Q + i * (E + 1) + h * hd;
        float mx = -INFINITY;
        for (int j = 0; j < M; ++j) {
            const float* kk = s.K + j * E + h * hd;
            float d = 0.f;
            for (int c = 0; c < hd; ++c) d = fmaf(q[c], kk[c], d);
            d *= scale;
            prow[j] = d;
            mx = fmaxf(mx, d);
        }
        float den = 0.f;
        for (int j = 0; j < M; ++j) {
            const float e = __expf(prow[j] - mx);
            prow[j] = e;
            den += e;
        }
        const float inv = 1.f / den;
        for (int j = 0; j < M; ++j) prow[j] *= inv;
    }
    __syncthreads();
    for (int idx = tid; idx < rc * E; idx += nt) {
        const int i = idx / E, f = idx - i * E, h = f / hd;
        const float* prow = s.Pm + (h * rc + i) * M;
        float acc = 0.f;
#pragma unroll 4
        for (int j = 0; j < M; ++j) acc = fmaf(prow[j], s.V[j * E + f], acc);
        s.O[idx] = acc;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(256) cross_attn_fwd_kernel(AttnArgs a) {
    extern __shared__ float smf[];
    const int tid = threadIdx.x, nt = blockDim.x, R = a.R, E = a.E, Rc = a.Rc;
    AttnSmem s = attn_carve(smf, Rc, a.M, E, a.heads);
    attn_load_params(a, s);
    for (int b = blockIdx.x; b < a.B; b += gridDim.x) {
        attn_keys_values(a, s, b);
        for (int r0 = 0; r0 < R; r0 += Rc) {
            const int rc = min(Rc, R - r0);
            attn_forward_chunk(a, s, b, r0, rc);
            float* yb = a.y + ((int64_t)b * R + r0) * E;
            for (int idx = tid; idx < rc * E; idx += nt) {
                const int i = idx / E, f = idx - i * E;
                float acc = s.bo[f];
#pragma unroll 8
                for (int k = 0; k < E; ++k) acc = fmaf(s.O[i * E + k], s.WoT[k * E + f], acc);
                yb[idx] = a.relu ? fmaxf(acc, 0.f) : acc;
            }
            __syncthreads();
        }
    }
}

__global__ void __launch_bounds__(512) cross_attn_bwd_kernel(AttnArgs a) {
    extern __shared__ float smf[];
    const int tid = threadIdx.x, nt = blockDim.x, R = a.R, M = a.M, E = a.E, H = a.heads, hd = E / H, Rc = a.Rc;
    const float scale = rsqrtf((float)hd);
    AttnSmem s = attn_carve(smf, Rc, M, E, H);
    float* p = s.tail;
    float* dY = p;   p += Rc * E;        // dY, later dQ
    float* dO = p;   p += Rc * (E + 1);  // padded like Q
    float* dK = p;   p += M * E;         // accumulated over the row chunks of a graph
    float* dV = p;   p += M * E;
    float* acc = p;  p += a.P;           // [dWin (3E,E) | dbin (3E) | dWo (E,E) | dbo (E)]
    float* WinO = p; p += 3 * E * E;     // row-major copies [f][k] for the transposed products of the backward
    float* WoO = p;  p += E * E;
    attn_load_params(a, s);
    for (int i = tid; i < 3 * E * E; i += nt) WinO[i] = a.Win[i];
    for (int i = tid; i < E * E; i += nt) WoO[i] = a.Wo[i];
    for (int i = tid; i < a.P; i += nt) acc[i] = 0.f;
    const int oBin = 3 * E * E, oWo = oBin + 3 * E, oBo = oWo + E * E;
    __syncthreads();
    for (int b = blockIdx.x; b < a.B; b += gridDim.x) {
        attn_keys_values(a, s, b);
        for (int i = tid; i < M * E; i += nt) {
            dK[i] = 0.f;
            dV[i] = 0.f;
        }
        for (int r0 = 0; r0 < R; r0 += Rc) {
            const int rc = min(Rc, R - r0);
            attn_forward_chunk(a, s, b, r0, rc);
            const float* gb = a.gy + ((int64_t)b * R + r0) * E;
            const float* yb = a.yout + ((int64_t)b * R + r0) * E;
            for (int i = tid; i < rc * E; i += nt) dY[i] = (!a.relu || yb[i] > 0.f) ? gb[i] : 0.f;
            __syncthreads();
            // out_proj: dO = dY Wo ; dWo += dY^T O ; dbo += colsum(dY)
            for (int idx = tid; idx < rc * E; idx += nt) {
                const int i = idx / E, k = idx - i * E;
                float v = 0.f;
#pragma unroll 8
                for (int f = 0; f < E; ++f) v = fmaf(dY[i * E + f], WoO[f * E + k], v);
                dO[i * (E + 1) + k] = v;
            }
            for (int idx = tid; idx < E * E; idx += nt) {
                const int f = idx / E, k = idx - f * E;
                float v = 0.f;
#pragma unroll 4
                for (int i = 0; i < rc; ++i) v = fmaf(dY[i * E + f], s.O[i * E + k], v);
                acc[oWo + idx] += v;
            }
            for (int f = tid; f < E; f += nt) {
                float v = 0.f;
                for (int i = 0; i < rc; ++i) v += dY[i * E + f];
                acc[oBo + f] += v;
            }
            __syncthreads();
            // dV += P^T dO (per head) ; dS = P * (dP - rowsum(P*dP)), dP = dO V^T   (dS overwrites P)
            for (int idx = tid; idx < M * E; idx += nt) {
                const int j = idx / E, f = idx - j * E, h = f / hd;
                float v = 0.f;
#pragma unroll 4
                for (int i = 0; i < rc; ++i) v = fmaf(s.Pm[(h * rc + i) * M + j], dO[i * (E + 1) + f], v);
                dV[idx] += v;
            }
            __syncthreads();
            for (int idx = tid; idx < H * rc; idx += nt) {
                const int h = idx / rc, i = idx - h * rc;
                float* prow = s.Pm + (h * rc + i) * M;
                const float* go = dO + i * (E + 1) + h * hd;
                float rowdot = 0.f;
                for (int j = 0; j < M; ++j) {
                    const float* vv = s.V + j * E + h * hd;
                    float d = 0.f;
                    for (int c = 0; c < hd; ++c) d = fmaf(go[c], vv[c], d);
                    rowdot = fmaf(prow[j], d, rowdot);
                }
                for (int j = 0; j < M; ++j) {
                    const float* vv = s.V + j * E + h * hd;
                    float d = 0.f;
                    for (int c = 0; c < hd; ++c) d = fmaf(go[c], vv[c], d);
                    prow[j] = prow[j] * (d - rowdot) * scale;       // d loss / d (q.k), scale folded in
                }
            }
            __syncthreads();
            // dQ = dS K ; dK += dS^T Q
            float* dQ = dY;
            for (int idx = tid; idx < rc * E; idx += nt) {
                const int i = idx / E, f = idx - i * E, h = f / hd;
                const float* srow = s.Pm + (h * rc + i) * M;
                float v = 0.f;
#pragma unroll 4
                for (int j = 0; j < M; ++j) v = fmaf(srow[j], s.K[j * E + f], v);
                dQ[idx] = v;
            }
            for (int idx = tid; idx < M * E; idx += nt) {
                const int j = idx / E, f = idx - j * E, h = f / hd;
                float v = 0.f;
#pragma unroll 4
                for (int i = 0; i < rc; ++i) v = fmaf(s.Pm[(h * rc + i) * M + j], s.Q[i * (E + 1) + f], v);
                dK[idx] += v;
            }
            __syncthreads();
            // query-side input gradient and projection gradients of this chunk
            float* dxb = a.dx + ((int64_t)b * R + r0) * E;
            for (int idx = tid; idx < rc * E; idx += nt) {
                const int i = idx / E, k = idx - i * E;
                float v = 0.f;
#pragma unroll 8
                for (int f = 0; f < E; ++f) v = fmaf(dQ[i * E + f], WinO[f * E + k], v);
                dxb[idx] = v;
            }
            for (int idx = tid; idx < E * E; idx += nt) {
                const int f = idx / E, k = idx - f * E;
                float v = 0.f;
#pragma unroll 4
                for (int i = 0; i < rc; ++i) v = fmaf(dQ[i * E + f], s.X[i * E + k], v);
                acc[idx] += v;
            }
            for (int f = tid; f < E; f += nt) {
                float v = 0.f;
                for (int i = 0; i < rc; ++i) v += dQ[i * E + f];
                acc[oBin + f] += v;
            }
            __syncthreads();
        }
        // key/value side: input gradient and projection gradients (after all row chunks)
        float* dab = a.da + (int64_t)b * M * E;
        for (int idx = tid; idx < M * E; idx += nt) {
            const int j = idx / E, k = idx - j * E;
            float v = 0.f;
#pragma unroll 8
            for (int f = 0; f < E; ++f) {
                v = fmaf(dK[j * E + f], WinO[(E + f) * E + k], v);
                v = fmaf(dV[j * E + f], WinO[(2 * E + f) * E + k], v);
            }
            dab[idx] = v;
        }
        for (int idx = tid; idx < 2 * E * E; idx += nt) {
            const int f2 = idx / E, k = idx - f2 * E;
            const float* dsrc = (f2 < E) ? dK + f2 : dV + (f2 - E);
            float v = 0.f;
#pragma unroll 4
            for (int j = 0; j < M; ++j) v = fmaf(dsrc[j * E], s.A[j * E + k], v);
            acc[E * E + idx] += v;
        }
        for (int f2 = tid; f2 < 2 * E; f2 += nt) {
            const float* dsrc = (f2 < E) ? dK + f2 : dV + (f2 - E);
            float v = 0.f;
            for (int j = 0; j < M; ++j) v += dsrc[j * E];
            acc[oBin + E + f2] += v;
        }
        __syncthreads();
    }
    float* prow = a.partials + (int64_t)blockIdx.x * a.P;
    for (int i = tid; i < a.P; i += nt) prow[i] = acc[i];
}

// query rows per chunk: the whole graph when it fits, else the largest chunk that keeps the backward under ~200 KB
static int attn_rows_per_chunk(int R, int M, int E, int H) {
    int rc = R;
    while (rc > 8) {
        const size_t fl = attn_common_floats(rc, M, E, H) + 2 * (size_t)rc * E + rc + 2 * (size_t)M * E + (size_t)(4 * E * E + 4 * E) +
                          4 * (size_t)E * E;
        if (4 * fl <= 200 * 1024) break;
        rc = (rc + 1) / 2;
    }
    return rc;
}
static size_t attn_fwd_smem(int Rc, int M, int E, int H) { return 4 * attn_common_floats(Rc, M, E, H); }
static size_t attn_bwd_smem(int Rc, int M, int E, int H, int P) {
    return 4 * (attn_common_floats(Rc, M, E, H) + 2 * (size_t)Rc * E + Rc + 2 * (size_t)M * E + P + 4 * (size_t)E * E);
}
static int attn_ctas(size_t smem, int64_t B, int nthreads = 256) {
    int per_sm = (int)((227 * 1024) / (smem + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 2048 / nthreads) per_sm = 2048 / nthreads;
    int64_t n = (int64_t)sm_count() * per_sm;
    if (n > B) n = B;
    return (int)(n < 1 ? 1 : n);
}
static int attn_fill(AttnArgs& a, const char* who, const float* x, const float* kv, const float* Win, const float* bin, const float* Wo,
                     const float* bo, int64_t B, int64_t R, int64_t M, int64_t E, int64_t heads, int64_t relu) {
    IGCN_REQUIRE(B >= 0 && R > 0 && M > 0 && E > 0 && heads > 0 && E % heads == 0, IGCN_ERR_BAD_ARG, "%s: bad size", who);
    IGCN_REQUIRE(E <= 128, IGCN_ERR_UNSUPPORTED, "%s: embed dim %lld > 128 not supported", who, (long long)E);
    IGCN_REQUIRE(x && kv && Win && bin && Wo && bo, IGCN_ERR_BAD_ARG, "%s: null pointer", who);
    a.x = x; a.a = kv; a.Win = Win; a.bin = bin; a.Wo = Wo; a.bo = bo;
    a.B = (int)B; a.R = (int)R; a.M = (int)M; a.E = (int)E; a.heads = (int)heads; a.relu = relu ? 1 : 0;
    a.P = (int)(4 * E * E + 4 * E);
    a.Rc = attn_rows_per_chunk((int)R, (int)M, (int)E, (int)heads);
    return IGCN_OK;
}

}  // namespace igcn

using namespace igcn;

extern "C" int64_t igcn_cross_attn_param_count(int64_t E) { return 4 * E * E + 4 * E; }
extern "C" int64_t igcn_cross_attn_bwd_ctas(int64_t B, int64_t R, int64_t M, int64_t E, int64_t heads) {
    return attn_ctas(attn_bwd_smem(attn_rows_per_chunk((int)R, (int)M, (int)E, (int)heads), (int)M, (int)E, (int)heads, (int)(4 * E * E + 4 * E)), B, 512);
}

extern "C" int igcn_cross_attn_fwd(const float* q_in, const float* kv_in, const float* in_proj_weight, const float* in_proj_bias,
                                   const float* out_proj_weight, const float* out_proj_bias, int64_t B, int64_t R, int64_t M, int64_t E,
                                   int64_t heads, int64_t relu, float* out, void* stream) {
    AttnArgs a{};
    int rc = attn_fill(a, "cross_attn_fwd", q_in, kv_in, in_proj_weight, in_proj_bias, out_proj_weight, out_proj_bias, B, R, M, E, heads, relu);
    if (rc) return rc;
    IGCN_REQUIRE(out, IGCN_ERR_BAD_ARG, "cross_attn_fwd: null output");
    if (B == 0) return IGCN_OK;
    a.y = out;
    size_t smem = attn_fwd_smem(a.Rc, a.M, a.E, a.heads);
    if ((rc = allow_smem(cross_attn_fwd_kernel, smem, "cross_attn_fwd"))) return rc;
    cross_attn_fwd_kernel<<<attn_ctas(smem, B), 256, smem, (cudaStream_t)stream>>>(a);
    IGCN_CHECK_LAUNCH("cross_attn_fwd");
    return IGCN_OK;
}

extern "C" int igcn_cross_attn_bwd(const float* q_in, const float* kv_in, const float* in_proj_weight, const float* in_proj_bias,
                                   const float* out_proj_weight, const float* out_proj_bias, const float* out, const float* g_out,
                                   int64_t B, int64_t R, int64_t M, int64_t E, int64_t heads, int64_t relu, float* d_q_in, float* d_kv_in,
                                   float* partials, int64_t n_cta, float* grads, void* stream) {
    AttnArgs a{};
    int rc = attn_fill(a, "cross_attn_bwd", q_in, kv_in, in_proj_weight, in_proj_bias, out_proj_weight, out_proj_bias, B, R, M, E, heads, relu);
    if (rc) return rc;
    IGCN_REQUIRE(out && g_out && d_q_in && d_kv_in && partials && grads, IGCN_ERR_BAD_ARG, "cross_attn_bwd: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if (B == 0) {
        cudaMemsetAsync(grads, 0, sizeof(float) * a.P, st);
        return IGCN_OK;
    }
    const int want = (int)igcn_cross_attn_bwd_ctas(B, R, M, E, heads);
    IGCN_REQUIRE(n_cta == want, IGCN_ERR_BAD_ARG, "cross_attn_bwd: n_cta=%lld, expected %d", (long long)n_cta, want);
    a.yout = out; a.gy = g_out; a.dx = d_q_in; a.da = d_kv_in; a.partials = partials;
    size_t smem = attn_bwd_smem(a.Rc, a.M, a.E, a.heads, a.P);
    if ((rc = allow_smem(cross_attn_bwd_kernel, smem, "cross_attn_bwd"))) return rc;
    cross_attn_bwd_kernel<<<want, 512, smem, st>>>(a);   // one graph per CTA, 16 warps: the per-graph chain is latency bound
    IGCN_CHECK_LAUNCH("cross_attn_bwd");
    reduce_partials_kernel<<<(a.P + 31) / 32, 256, 0, st>>>(partials, want, a.P, grads);
    IGCN_CHECK_LAUNCH("cross_attn_reduce_partials");
    return IGCN_OK;
}
