// Fused imaging <- genetics cross attention (reference: kernel/sgcn_img_snp.py:46,239-241 --
// nn.MultiheadAttention(E, 2, batch_first=True)(query = ROI tokens (B,R,E), key = value = GO tokens (B,M,E)) followed by relu).
//
// torch runs this as ~15 launches forward and ~25 backward; the weight gradients of the projections are (E x B*R) x (B*R x E)
// products that cuBLAS executes on ONE CTA (ncu/torch.profiler: 8 x 53 us per step at B=256 -- more than every igcn kernel
// together).  Here a CTA owns a graph: Q/K/V projections, per-head scores, row softmax, P.V, the output projection and the
// ReLU stay in shared memory (R*M <= 264*300 scores per head fit); projection-weight gradients are accumulated per CTA in shared
// memory across its graphs and reduced in a fixed order afterwards (deterministic, no float atomics).
// Shape-generic (runtime R, M, E, heads); fp32 FFMA.
#include <stdlib.h>

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
    int mix;            // 1: y = (x + relu(attn(x))) / 2 in one pass (kernel/sgcn_img_snp.py:239-242 + the fusion average); tensor-core kernels only
    int Rc;             // query rows staged per chunk
    // table-driven tensor-core path (cross_attn_mma2.cuh): per-graph K'/V'/c/K/V records written by the forward, per-(graph, row
    // chunk) gradient records written by the backward row kernel and consumed by the chain kernel
    const float* tab;   // (B, tab_sz)  read side
    float* tab_out;     // (B, tab_sz)  attn_tables_kernel output
    float* dtab;        // (B * nchunk, dtab_sz)
    int tab_sz, dtab_sz, nchunk;
};

// Register tiling: every thread produces 4 consecutive output features, so one broadcast LDS.32 + one LDS.128 feed 4 FMAs
// (the first version had 2 LDS per FMA and was ~3x slower).  Row-walked arrays (Q, K, V, dO) use a row stride of E+4 floats:
// 16-byte aligned for LDS.128 and bank-conflict free when consecutive threads read consecutive rows.
__device__ __forceinline__ float4 ld4s(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4s(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void fma4(float s, float4 v, float4& acc) {
    acc.x = fmaf(s, v.x, acc.x);
    acc.y = fmaf(s, v.y, acc.y);
    acc.z = fmaf(s, v.z, acc.z);
    acc.w = fmaf(s, v.w, acc.w);
}
__device__ __forceinline__ float dot4s(float4 a, float4 b, float acc) {
    acc = fmaf(a.x, b.x, acc);
    acc = fmaf(a.y, b.y, acc);
    acc = fmaf(a.z, b.z, acc);
    return fmaf(a.w, b.w, acc);
}

struct AttnSmem {
    float *WinT, *WoT, *bin, *bo;   // WinT[k][3E] (k-major), WoT[k][E]
    float *X, *A, *Q, *K, *V, *Pm, *O;   // X, Q, Pm, O hold ONE CHUNK of Rc query rows; Q/K/V rows have stride E+4
    float* tail;
};
__device__ __forceinline__ AttnSmem attn_carve(float* p, int Rc, int M, int E, int heads) {
    AttnSmem s;
    const int EP = E + 4;
    s.WinT = p;  p += 3 * E * E;
    s.WoT = p;   p += E * E;
    s.bin = p;   p += 3 * E;
    s.bo = p;    p += E;
    s.X = p;     p += Rc * E;
    s.A = p;     p += M * E;
    s.Q = p;     p += Rc * EP;
    s.K = p;     p += M * EP;
    s.V = p;     p += M * EP;
    s.O = p;     p += Rc * E;
    s.Pm = p;    p += (heads * Rc * M + 3) & ~3;
    s.tail = p;
    return s;
}
static size_t attn_common_floats(int Rc, int M, int E, int heads) {
    return (size_t)4 * E * E + 4 * E + 2 * (size_t)Rc * E + (size_t)Rc * (E + 4) + (size_t)M * E + 2 * (size_t)M * (E + 4) +
           (((size_t)heads * Rc * M + 3) & ~(size_t)3);
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

// out[r][4c..4c+3] = bias[4c..] + sum_k in[r*ldin + k] * Wt[k*ldw + 4c..]   for r < rows, c < E/4
__device__ __forceinline__ void proj4(const float* in, int ldin, const float* Wt, int ldw, const float* bias, int rows, int E,
                                      float* out, int ldout) {
    const int E4 = E >> 2;
    for (int idx = threadIdx.x; idx < rows * E4; idx += blockDim.x) {
        const int r = idx / E4, c = idx - r * E4;
        float4 acc = bias ? ld4s(bias + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
        const float* row = in + r * ldin;
#pragma unroll 8
        for (int k = 0; k < E; ++k) fma4(row[k], ld4s(Wt + k * ldw + 4 * c), acc);
        st4s(out + r * ldout + 4 * c, acc);
    }
}

// K, V of one graph (all M key/value tokens)
__device__ __forceinline__ void attn_keys_values(const AttnArgs& a, const AttnSmem& s, int b) {
    const int tid = threadIdx.x, nt = blockDim.x, M = a.M, E = a.E, EP = E + 4;
    const float* ab = a.a + (int64_t)b * M * E;
    for (int i = tid; i < M * E; i += nt) s.A[i] = ab[i];
    __syncthreads();
    proj4(s.A, E, s.WinT + E, 3 * E, s.bin + E, M, E, s.K, EP);
    proj4(s.A, E, s.WinT + 2 * E, 3 * E, s.bin + 2 * E, M, E, s.V, EP);
    __syncthreads();
}

// scores for one chunk: Pm[h][i][j] = scale * <Q_i^h, K_j^h>, one element per thread; then a row softmax per (h, i)
__device__ __forceinline__ void attn_scores_softmax(const AttnSmem& s, int rc, int M, int E, int H) {
    const int tid = threadIdx.x, nt = blockDim.x, hd = E / H, EP = E + 4;
    const float scale = rsqrtf((float)hd);
    for (int idx = tid; idx < H * rc * M; idx += nt) {
        const int j = idx % M, hi = idx / M, h = hi / rc, i = hi - h * rc;
        const float* q = s.Q + i * EP + h * hd;
        const float* kk = s.K + j * EP + h * hd;
        float d = 0.f;
        for (int c = 0; c < hd; c += 4) d = dot4s(ld4s(q + c), ld4s(kk + c), d);
        s.Pm[idx] = d * scale;
    }
    __syncthreads();
    for (int hi = tid; hi < H * rc; hi += nt) {
        float* prow = s.Pm + hi * M;
        float mx = -INFINITY;
        for (int j = 0; j < M; ++j) mx = fmaxf(mx, prow[j]);
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
}

// Q, P, O for the query rows [r0, r0+rc) of one graph
__device__ __forceinline__ void attn_forward_chunk(const AttnArgs& a, const AttnSmem& s, int b, int r0, int rc) {
    const int tid = threadIdx.x, nt = blockDim.x, R = a.R, M = a.M, E = a.E, H = a.heads, hd = E / H, EP = E + 4, E4 = E >> 2;
    const float* xb = a.x + ((int64_t)b * R + r0) * E;
    for (int i = tid; i < rc * E; i += nt) s.X[i] = xb[i];
    __syncthreads();
    proj4(s.X, E, s.WinT, 3 * E, s.bin, rc, E, s.Q, EP);
    __syncthreads();
    attn_scores_softmax(s, rc, M, E, H);
    for (int idx = tid; idx < rc * E4; idx += nt) {          // O = P V
        const int i = idx / E4, c = idx - i * E4, h = (4 * c) / hd;
        const float* prow = s.Pm + (h * rc + i) * M;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
        for (int j = 0; j < M; ++j) fma4(prow[j], ld4s(s.V + j * EP + 4 * c), acc);
        st4s(s.O + i * E + 4 * c, acc);
    }
    __syncthreads();
}

__global__ void __launch_bounds__(256) cross_attn_fwd_kernel(AttnArgs a) {
    IGCN_PDL_SYNC();
    extern __shared__ __align__(16) float smf[];
    const int tid = threadIdx.x, nt = blockDim.x, R = a.R, E = a.E, Rc = a.Rc, E4 = E >> 2;
    AttnSmem s = attn_carve(smf, Rc, a.M, E, a.heads);
    attn_load_params(a, s);
    for (int b = blockIdx.x; b < a.B; b += gridDim.x) {
        attn_keys_values(a, s, b);
        for (int r0 = 0; r0 < R; r0 += Rc) {
            const int rc = min(Rc, R - r0);
            attn_forward_chunk(a, s, b, r0, rc);
            float* yb = a.y + ((int64_t)b * R + r0) * E;
            for (int idx = tid; idx < rc * E4; idx += nt) {
                const int i = idx / E4, c = idx - i * E4;
                float4 acc = ld4s(s.bo + 4 * c);
                const float* orow = s.O + i * E;
#pragma unroll 8
                for (int k = 0; k < E; ++k) fma4(orow[k], ld4s(s.WoT + k * E + 4 * c), acc);
                if (a.relu) {
                    acc.x = fmaxf(acc.x, 0.f); acc.y = fmaxf(acc.y, 0.f); acc.z = fmaxf(acc.z, 0.f); acc.w = fmaxf(acc.w, 0.f);
                }
                *reinterpret_cast<float4*>(yb + i * E + 4 * c) = acc;
            }
            __syncthreads();
        }
    }
}

__global__ void __launch_bounds__(512) cross_attn_bwd_kernel(AttnArgs a) {
    IGCN_PDL_SYNC();
    extern __shared__ __align__(16) float smf[];
    const int tid = threadIdx.x, nt = blockDim.x, R = a.R, M = a.M, E = a.E, H = a.heads, hd = E / H, Rc = a.Rc;
    const int EP = E + 4, E4 = E >> 2;
    const float scale = rsqrtf((float)hd);
    AttnSmem s = attn_carve(smf, Rc, M, E, H);
    float* p = s.tail;
    float* dY = p;   p += Rc * E;           // dY, later dQ
    float* dO = p;   p += Rc * EP;
    float* dK = p;   p += M * E;            // accumulated over the row chunks of a graph
    float* dV = p;   p += M * E;
    float* dP = p;   p += (H * Rc * M + 3) & ~3;
    float* WinO = p; p += 3 * E * E;        // row-major copies [f][k] for the transposed products of the backward
    float* WoO = p;  p += E * E;
    float* acc = p;  p += a.P;              // [dWin (3E,E) | dbin (3E) | dWo (E,E) | dbo (E)]
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
            proj4(dY, E, WoO, E, nullptr, rc, E, dO, EP);
            for (int idx = tid; idx < E * E4; idx += nt) {
                const int f = idx / E4, c = idx - f * E4;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
                for (int i = 0; i < rc; ++i) fma4(dY[i * E + f], ld4s(s.O + i * E + 4 * c), v);
                float* d = acc + oWo + f * E + 4 * c;
                d[0] += v.x; d[1] += v.y; d[2] += v.z; d[3] += v.w;
            }
            for (int f = tid; f < E; f += nt) {
                float v = 0.f;
                for (int i = 0; i < rc; ++i) v += dY[i * E + f];
                acc[oBo + f] += v;
            }
            __syncthreads();
            // dP = dO V^T (one element per thread) ; dV += P^T dO
            for (int idx = tid; idx < H * rc * M; idx += nt) {
                const int j = idx % M, hi = idx / M, h = hi / rc, i = hi - h * rc;
                const float* go = dO + i * EP + h * hd;
                const float* vv = s.V + j * EP + h * hd;
                float d = 0.f;
                for (int c = 0; c < hd; c += 4) d = dot4s(ld4s(go + c), ld4s(vv + c), d);
                dP[idx] = d;
            }
            for (int idx = tid; idx < M * E4; idx += nt) {
                const int j = idx / E4, c = idx - j * E4, h = (4 * c) / hd;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
                for (int i = 0; i < rc; ++i) fma4(s.Pm[(h * rc + i) * M + j], ld4s(dO + i * EP + 4 * c), v);
                float* d = dV + j * E + 4 * c;
                d[0] += v.x; d[1] += v.y; d[2] += v.z; d[3] += v.w;
            }
            __syncthreads();
            // dS = P * (dP - rowsum(P * dP)) * scale   (overwrites P)
            for (int hi = tid; hi < H * rc; hi += nt) {
                float* prow = s.Pm + hi * M;
                const float* dprow = dP + hi * M;
                float rowdot = 0.f;
                for (int j = 0; j < M; ++j) rowdot = fmaf(prow[j], dprow[j], rowdot);
                for (int j = 0; j < M; ++j) prow[j] = prow[j] * (dprow[j] - rowdot) * scale;
            }
            __syncthreads();
            // dQ = dS K ; dK += dS^T Q
            float* dQ = dY;
            for (int idx = tid; idx < rc * E4; idx += nt) {
                const int i = idx / E4, c = idx - i * E4, h = (4 * c) / hd;
                const float* srow = s.Pm + (h * rc + i) * M;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
                for (int j = 0; j < M; ++j) fma4(srow[j], ld4s(s.K + j * EP + 4 * c), v);
                st4s(dQ + i * E + 4 * c, v);
            }
            for (int idx = tid; idx < M * E4; idx += nt) {
                const int j = idx / E4, c = idx - j * E4, h = (4 * c) / hd;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
                for (int i = 0; i < rc; ++i) fma4(s.Pm[(h * rc + i) * M + j], ld4s(s.Q + i * EP + 4 * c), v);
                float* d = dK + j * E + 4 * c;
                d[0] += v.x; d[1] += v.y; d[2] += v.z; d[3] += v.w;
            }
            __syncthreads();
            // query-side input gradient and projection gradients of this chunk
            float* dxb = a.dx + ((int64_t)b * R + r0) * E;
            for (int idx = tid; idx < rc * E4; idx += nt) {
                const int i = idx / E4, c = idx - i * E4;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                const float* row = dQ + i * E;
#pragma unroll 8
                for (int f = 0; f < E; ++f) fma4(row[f], ld4s(WinO + f * E + 4 * c), v);
                *reinterpret_cast<float4*>(dxb + i * E + 4 * c) = v;
            }
            for (int idx = tid; idx < E * E4; idx += nt) {
                const int f = idx / E4, c = idx - f * E4;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
                for (int i = 0; i < rc; ++i) fma4(dQ[i * E + f], ld4s(s.X + i * E + 4 * c), v);
                float* d = acc + f * E + 4 * c;
                d[0] += v.x; d[1] += v.y; d[2] += v.z; d[3] += v.w;
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
        for (int idx = tid; idx < M * E4; idx += nt) {
            const int j = idx / E4, c = idx - j * E4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
            for (int f = 0; f < E; ++f) {
                fma4(dK[j * E + f], ld4s(WinO + (E + f) * E + 4 * c), v);
                fma4(dV[j * E + f], ld4s(WinO + (2 * E + f) * E + 4 * c), v);
            }
            *reinterpret_cast<float4*>(dab + j * E + 4 * c) = v;
        }
        for (int idx = tid; idx < 2 * E * E4; idx += nt) {
            const int f2 = idx / E4, c = idx - f2 * E4;
            const float* dsrc = (f2 < E) ? dK + f2 : dV + (f2 - E);
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
            for (int j = 0; j < M; ++j) fma4(dsrc[j * E], ld4s(s.A + j * E + 4 * c), v);
            float* d = acc + E * E + f2 * E + 4 * c;
            d[0] += v.x; d[1] += v.y; d[2] += v.z; d[3] += v.w;
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
    if (const char* e = getenv("IGCN_ATTN_RC")) {     // tuning hook
        const int v = atoi(e);
        if (v >= 8 && v < rc) rc = v;
    }
    while (rc > 8) {
        const size_t fl = attn_common_floats(rc, M, E, H) + (size_t)rc * E + (size_t)rc * (E + 4) + 2 * (size_t)M * E +
                          (size_t)H * rc * M + 4 + (size_t)(4 * E * E + 4 * E) + 4 * (size_t)E * E;
        if (4 * fl <= 200 * 1024) break;
        rc = (rc + 1) / 2;
    }
    return rc;
}
static size_t attn_fwd_smem(int Rc, int M, int E, int H) { return 4 * attn_common_floats(Rc, M, E, H) + 16; }
static size_t attn_bwd_smem(int Rc, int M, int E, int H, int P) {
    return 4 * (attn_common_floats(Rc, M, E, H) + (size_t)Rc * E + (size_t)Rc * (E + 4) + 2 * (size_t)M * E + (size_t)H * Rc * M + 4 + P +
                4 * (size_t)E * E) + 16;
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
    IGCN_REQUIRE(E <= 128 && E % 4 == 0 && (E / heads) % 4 == 0, IGCN_ERR_UNSUPPORTED,
                 "%s: embed dim %lld / heads %lld: need E <= 128 and head_dim a multiple of 4", who, (long long)E, (long long)heads);
    IGCN_REQUIRE(((uintptr_t)x | (uintptr_t)kv) % 16 == 0, IGCN_ERR_BAD_ARG, "%s: inputs must be 16-byte aligned", who);
    IGCN_REQUIRE(x && kv && Win && bin && Wo && bo, IGCN_ERR_BAD_ARG, "%s: null pointer", who);
    a.x = x; a.a = kv; a.Win = Win; a.bin = bin; a.Wo = Wo; a.bo = bo;
    a.B = (int)B; a.R = (int)R; a.M = (int)M; a.E = (int)E; a.heads = (int)heads; a.relu = relu ? 1 : 0; a.mix = relu == 2 ? 1 : 0;
    a.P = (int)(4 * E * E + 4 * E);
    a.Rc = attn_rows_per_chunk((int)R, (int)M, (int)E, (int)heads);
    return IGCN_OK;
}

}  // namespace igcn

#include "cross_attn_rows.cuh"
#include "cross_attn_mma.cuh"
#include "cross_attn_mma2.cuh"
#include "cross_attn_mma3.cuh"

namespace igcn {

// the row-parallel kernels cover the reference's shape; everything else stays on the staged kernels above
static bool use_rows(int64_t R, int64_t M, int64_t E, int64_t heads) {
    if (getenv("IGCN_ATTN_STAGED")) return false;            // A/B hook
    if (E != rows::kE || M < 1 || M > 32 || R > 288 || heads < 1 || (E % heads) != 0 || ((E / heads) % 4) != 0) return false;
    return rows::fwd_geo((int)R, (int)M, (int)heads).smem <= 110 * 1024 && rows::bwd_geo((int)R, (int)M, (int)heads).smem <= 224 * 1024;
}
// tensor-core kernels (cross_attn_mma.cuh): same shapes as the row kernels; IGCN_ATTN_ROWS=1 keeps the FFMA row kernels (A/B hook)
static bool use_mma_fwd(int64_t R, int64_t M, int64_t E, int64_t heads) {
    if (getenv("IGCN_ATTN_ROWS")) return false;
    return use_rows(R, M, E, heads) && amma::fwd_geo((int)R, (int)M, (int)heads).smem <= 110 * 1024;
}
static bool use_mma_bwd(int64_t R, int64_t M, int64_t E, int64_t heads) {
    if (getenv("IGCN_ATTN_ROWS")) return false;
    return use_rows(R, M, E, heads) && amma::bwd_geo((int)R, (int)M, (int)heads).smem <= 227 * 1024;
}
// table-driven path (cross_attn_mma2.cuh); IGCN_ATTN_V1=1 keeps the single-kernel version (A/B hook)
static bool use_v2(int64_t R, int64_t M, int64_t E, int64_t heads) {
    static const bool off = getenv("IGCN_ATTN_V1") != nullptr;
    if (off || !use_mma_fwd(R, M, E, heads) || !use_mma_bwd(R, M, E, heads)) return false;
    return amma2::bwd2_geo((int)R, (int)M, (int)heads, 1 << 20).smem <= 110 * 1024 && amma2::chain_smem((int)M, ((int)M + 7) & ~7, (int)heads) <= 100 * 1024;
}
// token-side kernels on tensor cores (cross_attn_mma3.cuh) for 2 heads of 16; IGCN_ATTN_SCALAR_TOKENS=1 keeps the scalar versions
static bool use_mma_tokens(int64_t E, int64_t heads) {
    static const bool off = getenv("IGCN_ATTN_SCALAR_TOKENS") != nullptr;
    return !off && E == 32 && heads == 2;
}
template <int NT>
static void launch_bwd2(const AttnArgs& a, const amma2::GeoB2& g, int items, cudaStream_t st) {
    cudaFuncSetAttribute(amma2::attn_bwd2_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem);
    igcn::launch_k(amma2::attn_bwd2_kernel<NT>, dim3(items), dim3(g.nthreads), g.smem, st, a, g);
}
template <int NT>
static void launch_mma_bwd(const AttnArgs& a, const amma::GeoB& g, int ctas, cudaStream_t st) {
    cudaFuncSetAttribute(amma::attn_mma_bwd_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem);
    igcn::launch_k(amma::attn_mma_bwd_kernel<NT>, dim3(ctas), dim3(g.nthreads), g.smem, st, a, g);
}
template <int NT>
static void launch_mma_fwd(const AttnArgs& a, const amma::Geo& g, int ctas, cudaStream_t st) {
    cudaFuncSetAttribute(amma::attn_mma_fwd_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem);
    igcn::launch_k(amma::attn_mma_fwd_kernel<NT>, dim3(ctas), dim3(amma::kThreads), g.smem, st, a, g);
}
static int rows_ctas(const rows::Geo& g, int64_t B, int per_sm) {
    const int64_t groups = (B + g.gpc - 1) / g.gpc;
    int64_t n = (int64_t)sm_count() * per_sm;
    if (n > groups) n = groups;
    return (int)(n < 1 ? 1 : n);
}
template <int MC>
static void launch_rows_fwd(const AttnArgs& a, const rows::Geo& g, int ctas, cudaStream_t st) {
    cudaFuncSetAttribute(rows::attn_rows_fwd_kernel<MC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem);
    igcn::launch_k(rows::attn_rows_fwd_kernel<MC>, dim3(ctas), dim3(g.nthreads), g.smem, st, a, g);
}
template <int MC>
static void launch_rows_bwd(const AttnArgs& a, const rows::Geo& g, int ctas, cudaStream_t st) {
    cudaFuncSetAttribute(rows::attn_rows_bwd_kernel<MC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem);
    igcn::launch_k(rows::attn_rows_bwd_kernel<MC>, dim3(ctas), dim3(g.nthreads), g.smem, st, a, g);
}

}  // namespace igcn

using namespace igcn;

extern "C" int64_t igcn_cross_attn_param_count(int64_t E) { return 4 * E * E + 4 * E; }
extern "C" int64_t igcn_cross_attn_fused_average(int64_t R, int64_t M, int64_t E, int64_t heads) {
    return (use_mma_fwd(R, M, E, heads) && use_mma_bwd(R, M, E, heads)) ? 1 : 0;
}
extern "C" int64_t igcn_cross_attn_bwd_ctas(int64_t B, int64_t R, int64_t M, int64_t E, int64_t heads) {
    if (use_mma_bwd(R, M, E, heads)) {                       // one graph per pass, one CTA per SM
        int64_t n = sm_count();
        if (n > B) n = B;
        return n < 1 ? 1 : n;
    }
    if (use_rows(R, M, E, heads)) {
        const rows::Geo g = rows::bwd_geo((int)R, (int)M, (int)heads);
        return rows_ctas(g, B, g.smem <= 110 * 1024 ? 2 : 1);
    }
    return attn_ctas(attn_bwd_smem(attn_rows_per_chunk((int)R, (int)M, (int)E, (int)heads), (int)M, (int)E, (int)heads, (int)(4 * E * E + 4 * E)), B, 512);
}

extern "C" int64_t igcn_cross_attn_v2_supported(int64_t R, int64_t M, int64_t E, int64_t heads) { return use_v2(R, M, E, heads) ? 1 : 0; }
extern "C" int64_t igcn_cross_attn_v2_tab_floats(int64_t M, int64_t heads) {
    return amma2::tab_floats((int)M, ((int)M + 7) & ~7, (int)heads);
}
extern "C" int64_t igcn_cross_attn_v2_work_floats(int64_t B, int64_t R, int64_t M, int64_t heads) {
    const amma2::GeoB2 g = amma2::bwd2_geo((int)R, (int)M, (int)heads, B);
    return B * g.nchunk * (int64_t)amma2::dtab_floats(g.MP, (int)heads);
}
extern "C" int64_t igcn_cross_attn_v2_bwd_ctas(int64_t B) {
    return balanced_ctas((int64_t)sm_count() * 2, (B + amma2::kChainGraphs - 1) / amma2::kChainGraphs);
}

extern "C" int igcn_cross_attn_v2_fwd(const float* q_in, const float* kv_in, const float* in_proj_weight, const float* in_proj_bias,
                                      const float* out_proj_weight, const float* out_proj_bias, int64_t B, int64_t R, int64_t M, int64_t E,
                                      int64_t heads, int64_t relu, float* out, float* tab, void* stream) {
    AttnArgs a{};
    int rc = attn_fill(a, "cross_attn_v2_fwd", q_in, kv_in, in_proj_weight, in_proj_bias, out_proj_weight, out_proj_bias, B, R, M, E, heads, relu);
    if (rc) return rc;
    IGCN_REQUIRE(use_v2(R, M, E, heads), IGCN_ERR_UNSUPPORTED, "cross_attn_v2_fwd: shape outside the table-driven kernels (see igcn_cross_attn_v2_supported)");
    IGCN_REQUIRE(out && tab, IGCN_ERR_BAD_ARG, "cross_attn_v2_fwd: null output");
    IGCN_REQUIRE(((uintptr_t)q_in | (uintptr_t)out | (uintptr_t)tab) % 16 == 0, IGCN_ERR_BAD_ARG, "cross_attn_v2_fwd: buffers must be 16-byte aligned");
    if (B == 0) return IGCN_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const amma::Geo g = amma::fwd_geo((int)R, (int)M, (int)heads);
    a.y = out; a.tab_out = tab; a.tab = tab; a.tab_sz = amma2::tab_floats((int)M, g.MP, (int)heads);
    if (use_mma_tokens(E, heads)) {
        const size_t smem = amma3::tables_smem();
        if ((rc = allow_smem(amma3::attn_tables_mma_kernel, smem, "cross_attn_tables_mma"))) return rc;
        int64_t ctas = (B + amma3::kWarps - 1) / amma3::kWarps;
        if (ctas > (int64_t)sm_count() * 4) ctas = (int64_t)sm_count() * 4;
        igcn::launch_k(amma3::attn_tables_mma_kernel, dim3((int)ctas), dim3(amma3::kThreads), smem, st, a, g.MP);
        IGCN_CHECK_LAUNCH("cross_attn_tables_mma");
    } else {
        // the per-graph table area doubles as the 64 x 33 staging tile of the weight transposes
        const int per_area = amma2::kTabGraphs * g.per_sz > 2 * amma::kE * 33 ? amma2::kTabGraphs * g.per_sz : 2 * amma::kE * 33;
        const size_t smem = (size_t)4 * (4 * amma::kE * amma::kE + 4 * amma::kE + per_area) + 16;
        if ((rc = allow_smem(amma2::attn_tables_kernel, smem, "cross_attn_tables"))) return rc;
        int64_t ctas = (B + amma2::kTabGraphs - 1) / amma2::kTabGraphs;
        if (ctas > (int64_t)sm_count() * 4) ctas = (int64_t)sm_count() * 4;
        igcn::launch_k(amma2::attn_tables_kernel, dim3((int)ctas), dim3(amma2::kTabThreads), smem, st, a, g.per_sz, g.MP);
        IGCN_CHECK_LAUNCH("cross_attn_tables");
    }
    const int64_t groups = (B + g.gpc - 1) / g.gpc;
    int64_t ctas = (int64_t)sm_count() * 2;
    if (ctas > groups) ctas = groups;
    switch (g.MP / 8) {
        case 1: launch_mma_fwd<1>(a, g, (int)ctas, st); break;
        case 2: launch_mma_fwd<2>(a, g, (int)ctas, st); break;
        case 3: launch_mma_fwd<3>(a, g, (int)ctas, st); break;
        default: launch_mma_fwd<4>(a, g, (int)ctas, st); break;
    }
    IGCN_CHECK_LAUNCH("cross_attn_v2_fwd");
    return IGCN_OK;
}

extern "C" int igcn_cross_attn_v2_bwd(const float* q_in, const float* kv_in, const float* in_proj_weight, const float* in_proj_bias,
                                      const float* out_proj_weight, const float* out_proj_bias, const float* out, const float* g_out,
                                      const float* tab, int64_t B, int64_t R, int64_t M, int64_t E, int64_t heads, int64_t relu,
                                      float* d_q_in, float* d_kv_in, float* work, float* partials, int64_t n_cta, float* grads, void* stream) {
    AttnArgs a{};
    int rc = attn_fill(a, "cross_attn_v2_bwd", q_in, kv_in, in_proj_weight, in_proj_bias, out_proj_weight, out_proj_bias, B, R, M, E, heads, relu);
    if (rc) return rc;
    IGCN_REQUIRE(use_v2(R, M, E, heads), IGCN_ERR_UNSUPPORTED, "cross_attn_v2_bwd: shape outside the table-driven kernels (see igcn_cross_attn_v2_supported)");
    IGCN_REQUIRE(out && g_out && tab && d_q_in && d_kv_in && work && partials && grads, IGCN_ERR_BAD_ARG, "cross_attn_v2_bwd: null pointer");
    IGCN_REQUIRE(((uintptr_t)q_in | (uintptr_t)out | (uintptr_t)g_out | (uintptr_t)d_q_in | (uintptr_t)d_kv_in | (uintptr_t)tab | (uintptr_t)work |
                  (uintptr_t)partials) % 16 == 0, IGCN_ERR_BAD_ARG, "cross_attn_v2_bwd: buffers must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    if (B == 0) {
        cudaMemsetAsync(grads, 0, sizeof(float) * a.P, st);
        return IGCN_OK;
    }
    const int want = (int)igcn_cross_attn_v2_bwd_ctas(B);
    IGCN_REQUIRE(n_cta == want, IGCN_ERR_BAD_ARG, "cross_attn_v2_bwd: n_cta=%lld, expected %d", (long long)n_cta, want);
    const amma2::GeoB2 g = amma2::bwd2_geo((int)R, (int)M, (int)heads, B);
    a.yout = out; a.gy = g_out; a.dx = d_q_in; a.da = d_kv_in; a.partials = partials;
    a.tab = tab; a.tab_sz = amma2::tab_floats((int)M, g.MP, (int)heads);
    a.dtab = work; a.dtab_sz = amma2::dtab_floats(g.MP, (int)heads); a.nchunk = g.nchunk;
    IGCN_REQUIRE(B * g.nchunk < (int64_t)2147483647, IGCN_ERR_UNSUPPORTED, "cross_attn_v2_bwd: too many (graph, chunk) items");
    const int items = (int)(B * g.nchunk);
    switch (g.MP / 8) {
        case 1: launch_bwd2<1>(a, g, items, st); break;
        case 2: launch_bwd2<2>(a, g, items, st); break;
        case 3: launch_bwd2<3>(a, g, items, st); break;
        default: launch_bwd2<4>(a, g, items, st); break;
    }
    IGCN_CHECK_LAUNCH("cross_attn_v2_bwd_rows");
    if (use_mma_tokens(E, heads)) {
        const size_t csm = amma3::chain_smem();
        if ((rc = allow_smem(amma3::attn_chain_mma_kernel, csm, "cross_attn_chain_mma"))) return rc;
        igcn::launch_k(amma3::attn_chain_mma_kernel, dim3(want), dim3(amma3::kThreads), csm, st, a, g.MP);
        IGCN_CHECK_LAUNCH("cross_attn_chain_mma");
    } else {
        const size_t csm = amma2::chain_smem((int)M, g.MP, (int)heads);
        if ((rc = allow_smem(amma2::attn_chain_kernel, csm, "cross_attn_chain"))) return rc;
        igcn::launch_k(amma2::attn_chain_kernel, dim3(want), dim3(amma2::kChainThreads), csm, st, a, g.MP);
        IGCN_CHECK_LAUNCH("cross_attn_chain");
    }
    igcn::launch_k(reduce_partials_kernel, dim3((a.P + 31) / 32), dim3(reduce_threads(want)), 0, st, partials, want, a.P, grads);
    IGCN_CHECK_LAUNCH("cross_attn_reduce_partials");
    return IGCN_OK;
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
    IGCN_REQUIRE(!a.mix || use_mma_fwd(R, M, E, heads), IGCN_ERR_UNSUPPORTED, "cross_attn_fwd: relu=2 (fused average) needs the tensor-core shape");
    if (use_mma_fwd(R, M, E, heads)) {
        IGCN_REQUIRE(((uintptr_t)q_in | (uintptr_t)out) % 16 == 0, IGCN_ERR_BAD_ARG, "cross_attn_fwd: buffers must be 16-byte aligned");
        const amma::Geo g = amma::fwd_geo((int)R, (int)M, (int)heads);
        const int64_t groups = (B + g.gpc - 1) / g.gpc;
        int64_t ctas = (int64_t)sm_count() * 2;
        if (ctas > groups) ctas = groups;
        cudaStream_t st = (cudaStream_t)stream;
        switch (g.MP / 8) {
            case 1: launch_mma_fwd<1>(a, g, (int)ctas, st); break;
            case 2: launch_mma_fwd<2>(a, g, (int)ctas, st); break;
            case 3: launch_mma_fwd<3>(a, g, (int)ctas, st); break;
            default: launch_mma_fwd<4>(a, g, (int)ctas, st); break;
        }
        IGCN_CHECK_LAUNCH("cross_attn_mma_fwd");
        return IGCN_OK;
    }
    if (use_rows(R, M, E, heads)) {
        const rows::Geo g = rows::fwd_geo((int)R, (int)M, (int)heads);
        const int ctas = rows_ctas(g, B, 2);
        cudaStream_t st = (cudaStream_t)stream;
        if (M <= 8) launch_rows_fwd<8>(a, g, ctas, st);
        else if (M <= 16) launch_rows_fwd<16>(a, g, ctas, st);
        else if (M <= 24) launch_rows_fwd<24>(a, g, ctas, st);
        else launch_rows_fwd<32>(a, g, ctas, st);
        IGCN_CHECK_LAUNCH("cross_attn_rows_fwd");
        return IGCN_OK;
    }
    size_t smem = attn_fwd_smem(a.Rc, a.M, a.E, a.heads);
    if ((rc = allow_smem(cross_attn_fwd_kernel, smem, "cross_attn_fwd"))) return rc;
    igcn::launch_k(cross_attn_fwd_kernel, dim3(attn_ctas(smem, B)), dim3(256), smem, (cudaStream_t)stream, a);
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
    IGCN_REQUIRE(!a.mix || use_mma_bwd(R, M, E, heads), IGCN_ERR_UNSUPPORTED, "cross_attn_bwd: relu=2 (fused average) needs the tensor-core shape");
    if (use_mma_bwd(R, M, E, heads)) {
        IGCN_REQUIRE(((uintptr_t)q_in | (uintptr_t)out | (uintptr_t)g_out | (uintptr_t)d_q_in | (uintptr_t)d_kv_in) % 16 == 0, IGCN_ERR_BAD_ARG,
                     "cross_attn_bwd: buffers must be 16-byte aligned");
        const amma::GeoB g = amma::bwd_geo((int)R, (int)M, (int)heads);
        switch (g.MP / 8) {
            case 1: launch_mma_bwd<1>(a, g, want, st); break;
            case 2: launch_mma_bwd<2>(a, g, want, st); break;
            case 3: launch_mma_bwd<3>(a, g, want, st); break;
            default: launch_mma_bwd<4>(a, g, want, st); break;
        }
        IGCN_CHECK_LAUNCH("cross_attn_mma_bwd");
        igcn::launch_k(reduce_partials_kernel, dim3((a.P + 31) / 32), dim3(reduce_threads(want)), 0, st, partials, want, a.P, grads);
        IGCN_CHECK_LAUNCH("cross_attn_reduce_partials");
        return IGCN_OK;
    }
    if (use_rows(R, M, E, heads)) {
        IGCN_REQUIRE(((uintptr_t)out | (uintptr_t)g_out | (uintptr_t)d_q_in | (uintptr_t)d_kv_in) % 16 == 0, IGCN_ERR_BAD_ARG, "cross_attn_bwd: buffers must be 16-byte aligned");
        const rows::Geo g = rows::bwd_geo((int)R, (int)M, (int)heads);
        if (M <= 8) launch_rows_bwd<8>(a, g, want, st);
        else if (M <= 16) launch_rows_bwd<16>(a, g, want, st);
        else if (M <= 24) launch_rows_bwd<24>(a, g, want, st);
        else launch_rows_bwd<32>(a, g, want, st);
        IGCN_CHECK_LAUNCH("cross_attn_rows_bwd");
        igcn::launch_k(reduce_partials_kernel, dim3((a.P + 31) / 32), dim3(reduce_threads(want)), 0, st, partials, want, a.P, grads);
        IGCN_CHECK_LAUNCH("cross_attn_reduce_partials");
        return IGCN_OK;
    }
    size_t smem = attn_bwd_smem(a.Rc, a.M, a.E, a.heads, a.P);
    if ((rc = allow_smem(cross_attn_bwd_kernel, smem, "cross_attn_bwd"))) return rc;
    igcn::launch_k(cross_attn_bwd_kernel, dim3(want), dim3(512), smem, st, a);   // one graph per CTA, 16 warps: the per-graph chain is latency bound
    IGCN_CHECK_LAUNCH("cross_attn_bwd");
    igcn::launch_k(reduce_partials_kernel, dim3((a.P + 31) / 32), dim3(reduce_threads(want)), 0, st, partials, want, a.P, grads);
    IGCN_CHECK_LAUNCH("cross_attn_reduce_partials");
    return IGCN_OK;
}
