// Fused GATConv(in, out, heads=1, edge_dim=1) layer for batched equal-size graphs (reference call sites:
// kernel/sgcn.py:163-166 `GATConv(.., edge_dim=1)` inside SGCN_GAT; semantics = PyG 2.0.2 GATConv):
//   h = x W^T ;  remove self loops, add one self loop per node whose edge attribute is the MEAN of the node's incoming
//   (non-loop) edge attributes ;  logit_e = LeakyReLU_0.2(a_src.h_s + a_dst.h_t + ea_e * <W_e, a_e>) ;
//   softmax over the in-edges of every target (max shifted, +1e-16) ;  out_t = sum_e alpha_e h_s + bias.
// One CTA per graph: the edge scores, the segment softmax and the aggregation never leave shared memory; the backward's
// transposed scatter is a gather over the source-sorted list; parameter gradients go through per-CTA partial rows and a
// fixed-order reduction (no float atomics).  Shape-generic (runtime F_in, H); this layer is not on the benchmarked path.
#include "common.cuh"

namespace igcn {

struct GatArgs {
    const float* x;          // (N, Fin)
    const int32_t* rowptr_t; // CSR by target
    const int32_t* csr_src;
    const float* ea;         // (E) edge attribute per CSR slot
    const int32_t* rowptr_s; // source-sorted (bwd)
    const int32_t* csc_pos;
    const float* W;          // (H, Fin)
    const float* att_src;    // (H)
    const float* att_dst;    // (H)
    const float* lin_edge;   // (H)   Linear(1 -> H) weight
    const float* att_edge;   // (H)
    const float* bias;       // (H)
    float* out;              // (N, H)
    const float* g_out;      // (N, H)
    float* dx;               // (N, Fin)
    float* d_ea;             // (E)
    float* partials;         // (n_cta, P)   P = H*Fin + 5H : [dW | datt_src | datt_dst | dlin_edge | datt_edge | dbias]
    int B, R, Fin, H, maxEg, P;
    float slope;
};

// shared prologue of fwd and bwd for one graph: h, a_s, a_d, per-slot alpha (0 on removed self loops), alpha_self, and the
// LeakyReLU slope actually applied per slot / per self loop (needed by the backward)
__device__ __forceinline__ void gat_forward_graph(const GatArgs& a, int g, int e0, int Eg, float ce, const float* Wt, const float* asrc,
                                                  const float* adst, float* xs, float* hs, float* a_s, float* a_d, int* rp, int* esrc,
                                                  float* eav, float* alpha, float* eslope, float* mean, float* cnt, float* alpha_self,
                                                  float* slope_self) {
    const int tid = threadIdx.x, nt = blockDim.x, R = a.R, Fin = a.Fin, H = a.H;
    const int64_t node0 = (int64_t)g * R;
    for (int i = tid; i < R * Fin; i += nt) xs[i] = a.x[node0 * Fin + i];
    for (int i = tid; i <= R; i += nt) rp[i] = a.rowptr_t[node0 + i] - e0;
    for (int k = tid; k < Eg; k += nt) {
        esrc[k] = a.csr_src[e0 + k] - (int)node0;
        eav[k] = a.ea[e0 + k];
    }
    __syncthreads();
    for (int idx = tid; idx < R * H; idx += nt) {
        const int i = idx / H, f = idx - i * H;
        float acc = 0.f;
        for (int k = 0; k < Fin; ++k) acc = fmaf(xs[i * Fin + k], Wt[k * H + f], acc);
        hs[idx] = acc;
    }
    __syncthreads();
    for (int i = tid; i < R; i += nt) {
        float s1 = 0.f, s2 = 0.f;
        for (int f = 0; f < H; ++f) {
            s1 = fmaf(hs[i * H + f], asrc[f], s1);
            s2 = fmaf(hs[i * H + f], adst[f], s2);
        }
        a_s[i] = s1;
        a_d[i] = s2;
    }
    __syncthreads();
    for (int i = tid; i < R; i += nt) {
        float c = 0.f, sm = 0.f;
        for (int k = rp[i]; k < rp[i + 1]; ++k)
            if (esrc[k] != i) {
                c += 1.f;
                sm += eav[k];
            }
        const float mu = sm / fmaxf(c, 1.f);          // fill_value='mean'; 0 for a node without incoming edges
        mean[i] = mu;
        cnt[i] = c;
        // logits (slot value parked in alpha[]), running max
        float zself = a_s[i] + a_d[i] + mu * ce;
        slope_self[i] = zself > 0.f ? 1.f : a.slope;
        zself *= slope_self[i];
        float mx = zself;
        for (int k = rp[i]; k < rp[i + 1]; ++k) {
            const int s = esrc[k];
            if (s == i) {
                alpha[k] = 0.f;
                eslope[k] = 0.f;
                continue;
            }
            float z = a_s[s] + a_d[i] + eav[k] * ce;
            const float sl = z > 0.f ? 1.f : a.slope;
            z *= sl;
            eslope[k] = sl;
            alpha[k] = z;
            mx = fmaxf(mx, z);
        }
        float den = 0.f;
        for (int k = rp[i]; k < rp[i + 1]; ++k)
            if (esrc[k] != i) {
                const float ex = __expf(alpha[k] - mx);
                alpha[k] = ex;
                den += ex;
            }
        const float exs = __expf(zself - mx);
        den += exs;                                   // self loop is the LAST edge of the target, as PyG appends it
        const float inv = 1.f / (den + 1e-16f);
        for (int k = rp[i]; k < rp[i + 1]; ++k) alpha[k] *= inv;
        alpha_self[i] = exs * inv;
    }
    __syncthreads();
}

struct GatSmem {
    float *Wt, *asrc, *adst, *ledge, *aedge, *xs, *hs, *a_s, *a_d, *eav, *alpha, *eslope, *mean, *cnt, *alpha_self, *slope_self;
    int *rp, *esrc;
    float* tail;
};

__device__ __forceinline__ GatSmem gat_carve(float* smf, int R, int Fin, int H, int maxEg) {
    GatSmem s;
    float* p = smf;
    s.Wt = p;          p += Fin * H;
    s.asrc = p;        p += H;
    s.adst = p;        p += H;
    s.ledge = p;       p += H;
    s.aedge = p;       p += H;
    s.xs = p;          p += R * Fin;
    s.hs = p;          p += R * H;
    s.a_s = p;         p += R;
    s.a_d = p;         p += R;
    s.eav = p;         p += maxEg;
    s.alpha = p;       p += maxEg;
    s.eslope = p;      p += maxEg;
    s.mean = p;        p += R;
    s.cnt = p;         p += R;
    s.alpha_self = p;  p += R;
    s.slope_self = p;  p += R;
    s.rp = (int*)p;    p += R + 1;
    s.esrc = (int*)p;  p += maxEg;
    s.tail = p;
    return s;
}
static size_t gat_common_floats(int R, int Fin, int H, int maxEg) {
    return (size_t)Fin * H + 4 * H + (size_t)R * Fin + (size_t)R * H + 6 * (size_t)R + 4 * (size_t)maxEg + R + 1;
}

__device__ __forceinline__ float gat_load_params(const GatArgs& a, const GatSmem& s) {
    const int tid = threadIdx.x, nt = blockDim.x, Fin = a.Fin, H = a.H;
    for (int i = tid; i < H * Fin; i += nt) {
        const int f = i / Fin, k = i - f * Fin;
        s.Wt[k * H + f] = a.W[i];
    }
    for (int i = tid; i < H; i += nt) {
        s.asrc[i] = a.att_src[i];
        s.adst[i] = a.att_dst[i];
        s.ledge[i] = a.lin_edge[i];
        s.aedge[i] = a.att_edge[i];
    }
    __syncthreads();
    float ce = 0.f;
    for (int f = 0; f < H; ++f) ce = fmaf(s.ledge[f], s.aedge[f], ce);   // <Linear(1->H).weight, att_edge>
    return ce;
}

__global__ void __launch_bounds__(256) gat_fwd_kernel(GatArgs a) {
    extern __shared__ float smf[];
    const int tid = threadIdx.x, nt = blockDim.x, R = a.R, H = a.H;
    GatSmem s = gat_carve(smf, R, a.Fin, H, a.maxEg);
    const float ce = gat_load_params(a, s);
    for (int g = blockIdx.x; g < a.B; g += gridDim.x) {
        const int64_t node0 = (int64_t)g * R;
        const int e0 = a.rowptr_t[node0], Eg = a.rowptr_t[node0 + R] - e0;
        if (Eg > a.maxEg) __trap();
        gat_forward_graph(a, g, e0, Eg, ce, s.Wt, s.asrc, s.adst, s.xs, s.hs, s.a_s, s.a_d, s.rp, s.esrc, s.eav, s.alpha, s.eslope,
                          s.mean, s.cnt, s.alpha_self, s.slope_self);
        for (int idx = tid; idx < R * H; idx += nt) {
            const int i = idx / H, f = idx - i * H;
            float acc = 0.f;
            for (int k = s.rp[i]; k < s.rp[i + 1]; ++k) acc = fmaf(s.alpha[k], s.hs[s.esrc[k] * H + f], acc);
            acc = fmaf(s.alpha_self[i], s.hs[idx], acc);
            a.out[node0 * H + idx] = acc + a.bias[f];
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256) gat_bwd_kernel(GatArgs a) {
    extern __shared__ float smf[];
    const int tid = threadIdx.x, nt = blockDim.x, R = a.R, H = a.H, Fin = a.Fin, maxEg = a.maxEg;
    GatSmem s = gat_carve(smf, R, Fin, H, maxEg);
    float* p = s.tail;
    float* gs = p;            p += R * H;     // g_out tile
    float* dh = p;            p += R * H;
    float* dz = p;            p += maxEg;     // d logit (pre-LeakyReLU) per slot
    float* dzself = p;        p += R;
    float* das = p;           p += R;
    float* dad = p;           p += R;
    float* acc = p;           p += a.P;       // per-CTA parameter-gradient accumulators
    float* red = p;           p += 8;
    int* rps = (int*)p;       p += R + 1;
    int* spos = (int*)p;      p += maxEg;
    int* etgt = (int*)p;      p += maxEg;
    const float ce = gat_load_params(a, s);
    for (int i = tid; i < a.P; i += nt) acc[i] = 0.f;
    float dce_reg = 0.f;
    const int oW = 0, oAs = H * Fin, oAd = oAs + H, oLe = oAd + H, oAe = oLe + H, oB = oAe + H;
    __syncthreads();
    for (int g = blockIdx.x; g < a.B; g += gridDim.x) {
        const int64_t node0 = (int64_t)g * R;
        const int e0 = a.rowptr_t[node0], Eg = a.rowptr_t[node0 + R] - e0;
        if (Eg > maxEg) __trap();
        gat_forward_graph(a, g, e0, Eg, ce, s.Wt, s.asrc, s.adst, s.xs, s.hs, s.a_s, s.a_d, s.rp, s.esrc, s.eav, s.alpha, s.eslope,
                          s.mean, s.cnt, s.alpha_self, s.slope_self);
        for (int i = tid; i < R * H; i += nt) gs[i] = a.g_out[node0 * H + i];
        for (int i = tid; i <= R; i += nt) rps[i] = a.rowptr_s[node0 + i] - e0;
        for (int q = tid; q < Eg; q += nt) spos[q] = a.csc_pos[e0 + q] - e0;
        for (int i = tid; i < R; i += nt)
            for (int k = s.rp[i]; k < s.rp[i + 1]; ++k) etgt[k] = i;
        __syncthreads();
        // per target: d alpha, softmax backward, LeakyReLU backward, edge-attribute gradient
        for (int i = tid; i < R; i += nt) {
            float dself = 0.f;
            for (int f = 0; f < H; ++f) dself = fmaf(gs[i * H + f], s.hs[i * H + f], dself);
            float tsum = s.alpha_self[i] * dself;
            for (int k = s.rp[i]; k < s.rp[i + 1]; ++k) {
                const int sn = s.esrc[k];
                float d = 0.f;
                if (sn != i)
                    for (int f = 0; f < H; ++f) d = fmaf(gs[i * H + f], s.hs[sn * H + f], d);
                dz[k] = d;                                   // d alpha_k for now
                tsum = fmaf(s.alpha[k], d, tsum);
            }
            const float dzs = s.alpha_self[i] * (dself - tsum) * s.slope_self[i];
            dzself[i] = dzs;
            float sdz = dzs;
            const float dmean = dzs * ce / fmaxf(s.cnt[i], 1.f);
            dce_reg = fmaf(dzs, s.mean[i], dce_reg);
            for (int k = s.rp[i]; k < s.rp[i + 1]; ++k) {
                const float v = s.alpha[k] * (dz[k] - tsum) * s.eslope[k];   // 0 on removed self loops
                dz[k] = v;
                sdz += v;
                dce_reg = fmaf(v, s.eav[k], dce_reg);
                a.d_ea[e0 + k] = (s.esrc[k] != i) ? fmaf(v, ce, dmean) : 0.f;
            }
            dad[i] = sdz;
        }
        __syncthreads();
        for (int j = tid; j < R; j += nt) {                  // d a_src: over the out-edges of j, plus j's own self loop
            float sv = dzself[j];
            for (int q = rps[j]; q < rps[j + 1]; ++q) sv += dz[spos[q]];
            das[j] = sv;
        }
        __syncthreads();
        // dh[j] = sum_{out-edges} alpha_k g[t_k] + alpha_self_j g[j] + das_j att_src + dad_j att_dst
        for (int idx = tid; idx < R * H; idx += nt) {
            const int j = idx / H, f = idx - j * H;
            float v = s.alpha_self[j] * gs[idx];
            for (int q = rps[j]; q < rps[j + 1]; ++q) {
                const int k = spos[q];
                v = fmaf(s.alpha[k], gs[etgt[k] * H + f], v);
            }
            v = fmaf(das[j], s.asrc[f], v);
            v = fmaf(dad[j], s.adst[f], v);
            dh[idx] = v;
        }
        // d att_src, d att_dst, d bias
        for (int f = tid; f < H; f += nt) {
            float s1 = 0.f, s2 = 0.f, s3 = 0.f;
            for (int j = 0; j < R; ++j) {
                s1 = fmaf(das[j], s.hs[j * H + f], s1);
                s2 = fmaf(dad[j], s.hs[j * H + f], s2);
                s3 += gs[j * H + f];
            }
            acc[oAs + f] += s1;
            acc[oAd + f] += s2;
            acc[oB + f] += s3;
        }
        __syncthreads();
        for (int idx = tid; idx < H * Fin; idx += nt) {      // dW[f][k] += sum_j dh[j][f] x[j][k]
            const int f = idx / Fin, k = idx - f * Fin;
            float v = 0.f;
            for (int j = 0; j < R; ++j) v = fmaf(dh[j * H + f], s.xs[j * Fin + k], v);
            acc[oW + idx] += v;
        }
        for (int idx = tid; idx < R * Fin; idx += nt) {      // dx[j][k] = sum_f dh[j][f] W[f][k]
            const int j = idx / Fin, k = idx - j * Fin;
            float v = 0.f;
            for (int f = 0; f < H; ++f) v = fmaf(dh[j * H + f], s.Wt[k * H + f], v);
            a.dx[node0 * Fin + idx] = v;
        }
        __syncthreads();
    }
    // d<W_e, a_e> -> d lin_edge, d att_edge ; one partial row per CTA
    float v = warp_sum(dce_reg);
    if ((tid & 31) == 0) red[tid >> 5] = v;
    __syncthreads();
    float dce = 0.f;
    for (int w = 0; w < (nt >> 5); ++w) dce += red[w];
    for (int f = tid; f < H; f += nt) {
        acc[oLe + f] = dce * s.aedge[f];
        acc[oAe + f] = dce * s.ledge[f];
    }
    __syncthreads();
    float* prow = a.partials + (int64_t)blockIdx.x * a.P;
    for (int i = tid; i < a.P; i += nt) prow[i] = acc[i];
}

static size_t gat_fwd_smem(int R, int Fin, int H, int maxEg) { return 4 * gat_common_floats(R, Fin, H, maxEg); }
static size_t gat_bwd_smem(int R, int Fin, int H, int maxEg, int P) {
    return 4 * (gat_common_floats(R, Fin, H, maxEg) + 2 * (size_t)R * H + (size_t)maxEg + 3 * (size_t)R + P + 8 + R + 1 + 2 * (size_t)maxEg);
}
static int gat_ctas(size_t smem, int64_t B) {
    int per_sm = (int)((227 * 1024) / (smem + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 8) per_sm = 8;
    int64_t n = (int64_t)sm_count() * per_sm;
    if (n > B) n = B;
    return (int)(n < 1 ? 1 : n);
}

static int gat_fill(GatArgs& a, const char* who, const float* x, const int32_t* rowptr_t, const int32_t* csr_src, const float* ea,
                    const float* W, const float* att_src, const float* att_dst, const float* lin_edge, const float* att_edge,
                    const float* bias, int64_t B, int64_t R, int64_t Fin, int64_t H, int64_t max_eg, double slope) {
    IGCN_REQUIRE(B >= 0 && R > 0 && Fin > 0 && H > 0 && max_eg >= 0, IGCN_ERR_BAD_ARG, "%s: bad size", who);
    IGCN_REQUIRE(Fin <= 128 && H <= 128, IGCN_ERR_UNSUPPORTED, "%s: Fin/H above 128 not supported", who);
    IGCN_REQUIRE(x && rowptr_t && W && att_src && att_dst && lin_edge && att_edge && bias, IGCN_ERR_BAD_ARG, "%s: null pointer", who);
    IGCN_REQUIRE(max_eg == 0 || (csr_src && ea), IGCN_ERR_BAD_ARG, "%s: null edge arrays", who);
    a.x = x; a.rowptr_t = rowptr_t; a.csr_src = csr_src; a.ea = ea; a.W = W; a.att_src = att_src; a.att_dst = att_dst;
    a.lin_edge = lin_edge; a.att_edge = att_edge; a.bias = bias;
    a.B = (int)B; a.R = (int)R; a.Fin = (int)Fin; a.H = (int)H; a.maxEg = (int)max_eg; a.slope = (float)slope;
    a.P = (int)(H * Fin + 5 * H);
    return IGCN_OK;
}

}  // namespace igcn

using namespace igcn;

extern "C" int64_t igcn_gat_param_count(int64_t Fin, int64_t H) { return H * Fin + 5 * H; }
extern "C" int64_t igcn_gat_bwd_ctas(int64_t B, int64_t R, int64_t Fin, int64_t H, int64_t max_eg) {
    return gat_ctas(gat_bwd_smem((int)R, (int)Fin, (int)H, (int)max_eg, (int)(H * Fin + 5 * H)), B);
}

extern "C" int igcn_gat_layer_fwd(const float* x, const int32_t* rowptr_t, const int32_t* csr_src, const float* edge_attr,
                                  const float* W, const float* att_src, const float* att_dst, const float* lin_edge,
                                  const float* att_edge, const float* bias, int64_t B, int64_t R, int64_t Fin, int64_t H,
                                  int64_t max_eg, double negative_slope, float* out, void* stream) {
    GatArgs a{};
    int rc = gat_fill(a, "gat_layer_fwd", x, rowptr_t, csr_src, edge_attr, W, att_src, att_dst, lin_edge, att_edge, bias, B, R, Fin, H,
                      max_eg, negative_slope);
    if (rc) return rc;
    IGCN_REQUIRE(out, IGCN_ERR_BAD_ARG, "gat_layer_fwd: null output");
    if (B == 0) return IGCN_OK;
    a.out = out;
    size_t smem = gat_fwd_smem(a.R, a.Fin, a.H, a.maxEg);
    if ((rc = allow_smem(gat_fwd_kernel, smem, "gat_layer_fwd"))) return rc;
    gat_fwd_kernel<<<gat_ctas(smem, B), 256, smem, (cudaStream_t)stream>>>(a);
    IGCN_CHECK_LAUNCH("gat_layer_fwd");
    return IGCN_OK;
}

extern "C" int igcn_gat_layer_bwd(const float* x, const int32_t* rowptr_t, const int32_t* csr_src, const float* edge_attr,
                                  const int32_t* rowptr_s, const int32_t* csc_pos, const float* W, const float* att_src,
                                  const float* att_dst, const float* lin_edge, const float* att_edge, const float* bias,
                                  const float* g_out, int64_t B, int64_t R, int64_t Fin, int64_t H, int64_t max_eg,
                                  double negative_slope, float* dx, float* d_edge_attr, float* partials, int64_t n_cta, float* grads,
                                  void* stream) {
    GatArgs a{};
    int rc = gat_fill(a, "gat_layer_bwd", x, rowptr_t, csr_src, edge_attr, W, att_src, att_dst, lin_edge, att_edge, bias, B, R, Fin, H,
                      max_eg, negative_slope);
    if (rc) return rc;
    IGCN_REQUIRE(rowptr_s && g_out && dx && partials && grads && (max_eg == 0 || (csc_pos && d_edge_attr)), IGCN_ERR_BAD_ARG,
                 "gat_layer_bwd: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if (B == 0) {
        cudaMemsetAsync(grads, 0, sizeof(float) * a.P, st);
        return IGCN_OK;
    }
    const int want = (int)igcn_gat_bwd_ctas(B, R, Fin, H, max_eg);
    IGCN_REQUIRE(n_cta == want, IGCN_ERR_BAD_ARG, "gat_layer_bwd: n_cta=%lld, expected %d", (long long)n_cta, want);
    a.rowptr_s = rowptr_s; a.csc_pos = csc_pos; a.g_out = g_out; a.dx = dx; a.d_ea = d_edge_attr; a.partials = partials;
    size_t smem = gat_bwd_smem(a.R, a.Fin, a.H, a.maxEg, a.P);
    if ((rc = allow_smem(gat_bwd_kernel, smem, "gat_layer_bwd"))) return rc;
    gat_bwd_kernel<<<want, 256, smem, st>>>(a);
    IGCN_CHECK_LAUNCH("gat_layer_bwd");
    reduce_partials_kernel<<<(a.P + 31) / 32, 256, 0, st>>>(partials, want, a.P, grads);
    IGCN_CHECK_LAUNCH("gat_reduce_partials");
    return IGCN_OK;
}
