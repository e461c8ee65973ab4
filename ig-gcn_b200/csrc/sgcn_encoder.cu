// Fused SGCN brain-graph encoder for sm_100a: importance masks + self-loop merge + symmetric degree
// normalisation + L x (X.W, CSR SpMM, bias, ReLU) + concat, one persistent CTA per graph slab.
//
// Reference path being replaced (Houliang-Zhou/IG-GCN):
//   cal_probability            kernel/sgcn_img_snp.py:133-151
//   GCNConv x L (PyG 2.0.2)    kernel/sgcn_img_snp.py:218-221   (gcn_norm, Linear, propagate, bias)
//   relu / cat / to_dense_batch kernel/sgcn_img_snp.py:218-228
// and the autograd graph of all of the above for the backward kernel.
//
// Layout: every graph owns R consecutive nodes and a contiguous range of CSR slots, so one CTA stages
// a whole graph (features, masks, normalised adjacency, all layer activations) in shared memory and HBM
// sees each compulsory byte once: x, CSR (src,w), and the (B,R,L*H) output.  Parameter gradients are
// accumulated per CTA in shared memory across the graphs it owns, written as one partial row per CTA and
// reduced in a fixed order by a second tiny kernel (deterministic; no float atomics anywhere).
#include <stdlib.h>

#include "common.cuh"

namespace igcn {

struct EncArgs {
    const float* x;
    const int32_t* rowptr_t;
    const int32_t* csr_src;
    const float* csr_w;
    const int32_t* rowptr_s;
    const int32_t* csc_pos;
    const float* prob;       // (R,F0) or null
    const float* prob_bias;  // (2*F0) or null
    const float* wb;         // packed layer params
    const float* out;        // bwd: forward output
    const float* g_out;
    const float* g_pe;
    float* out_w;  // fwd: output
    float* pe_w;   // fwd: p_e
    float* dx;
    float* partials;
    int B, R, F0, H, L, maxEg;
    int P;     // param count (bwd)
    int relu;  // 1: relu after every layer (SGCN encoder); 0: plain GCNConv output (single-layer operator)
};

__host__ __device__ inline int layer_fin(int l, int F0, int H) { return l == 0 ? F0 : H; }
__host__ __device__ inline int layer_off(int l, int F0, int H) {  // offset of W_l in wb
    return l == 0 ? 0 : (H * F0 + H) + (l - 1) * (H * H + H);
}
__host__ __device__ inline int wb_size(int F0, int H, int L) { return L == 0 ? 0 : layer_off(L, F0, H); }

// ------------------------------------------------------------------------------------------------
// Shared prologue: masks, self-loop merge, degrees, normalised weights for ONE graph.
//   xs[i*F0+c]  masked features      enorm[k] normalised weight of CSR slot k (0 for self-loop slots)
//   dinv[i]     deg^-1/2 (0 if deg==0)    nii[i] = dinv^2 * loop weight
// Optional (bwd): ew[k]=w~ , epe[k]=p_e , ell[i] loop weight, lsl[i] slot of the loop-weight provider (-1: none)
// ------------------------------------------------------------------------------------------------
template <bool kKeep>
__device__ __forceinline__ void graph_prologue(const EncArgs& a, int g, int e0, int Eg, float* xs, const float* xraw_s,
                                               int* rp, int* esrc, float* enorm, float* dinv, float* nii, float* ew,
                                               float* epe, float* ell, int* lsl, const float* pb) {
    const int tid = threadIdx.x, nt = blockDim.x;
    const int R = a.R, F0 = a.F0;
    const bool explain = a.prob != nullptr;
    const int64_t node0 = (int64_t)g * R;
    for (int i = tid; i < R * F0; i += nt) {
        float v = a.x[node0 * F0 + i];
        if (kKeep) const_cast<float*>(xraw_s)[i] = v;
        xs[i] = explain ? v * a.prob[i] : v;
    }
    for (int i = tid; i <= R; i += nt) rp[i] = a.rowptr_t[node0 + i] - e0;
    for (int k = tid; k < Eg; k += nt) esrc[k] = a.csr_src[e0 + k] - (int)node0;
    __syncthreads();
    // pass 1: per target node: edge mask, masked weight, loop weight, degree
    for (int i = tid; i < R; i += nt) {
        float deg = 0.f, loopw = 1.f;
        int loop_slot = -1;
        for (int k = rp[i]; k < rp[i + 1]; ++k) {
            const int s = esrc[k];
            float wt = a.csr_w[e0 + k];
            float p = 1.f;
            if (explain) {
                float z = 0.f;
                for (int c = 0; c < F0; ++c) z += pb[c] * xs[s * F0 + c] + pb[F0 + c] * xs[i * F0 + c];
                p = sigmoidf_(z);
                wt *= p;
                if (a.pe_w) a.pe_w[e0 + k] = p;
            }
            if (kKeep) {
                ew[k] = wt;
                epe[k] = p;
            }
            enorm[k] = wt;
            if (s == i) {
                loopw = wt;  // last self loop wins (PyG add_remaining_self_loops)
                loop_slot = k;
            } else {
                deg += wt;
            }
        }
        deg += loopw;
        const float d = (deg == 0.f) ? 0.f : rsqrtf(deg);
        dinv[i] = d;
        nii[i] = d * d * loopw;
        if (kKeep) {
            ell[i] = loopw;
            lsl[i] = loop_slot;
        }
    }
    __syncthreads();
    // pass 2: normalised weights
    for (int i = tid; i < R; i += nt) {
        const float di = dinv[i];
        for (int k = rp[i]; k < rp[i + 1]; ++k) {
            const int s = esrc[k];
            enorm[k] = (s == i) ? 0.f : dinv[s] * enorm[k] * di;
        }
    }
    __syncthreads();
}

// U[i][f] = sum_k Hprev[i*ldh + k] * Wt[k*H + f]
__device__ __forceinline__ void dense_xw(const float* Hprev, int ldh, int Fin, const float* Wt, int H, int R, float* U) {
    for (int idx = threadIdx.x; idx < R * H; idx += blockDim.x) {
        const int i = idx / H, f = idx - i * H;
        const float* h = Hprev + i * ldh;
        float acc = 0.f;
        for (int k = 0; k < Fin; ++k) acc = fmaf(h[k], Wt[k * H + f], acc);
        U[idx] = acc;
    }
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sgcn_encoder_fwd_kernel(EncArgs a) {
    IGCN_PDL_SYNC();
    extern __shared__ float smf[];
    const int R = a.R, F0 = a.F0, H = a.H, L = a.L, LH = L * H, maxEg = a.maxEg;
    const int tid = threadIdx.x, nt = blockDim.x;
    // carve
    float* Hbuf = smf;                           // R*LH   (first: 16 B aligned for the float4 copy-out)
    float* U = Hbuf + ((R * LH + 3) & ~3);       // R*H
    float* Wt = U + R * H;                       // per layer: Fin*H (transposed [k][f]) + H bias
    float* pb = Wt + wb_size(F0, H, L);          // 2*F0
    float* xs = pb + 2 * F0;                     // R*F0
    float* dinv = xs + R * F0;                   // R
    float* nii = dinv + R;                       // R
    float* enorm = nii + R;                      // maxEg
    int* rp = (int*)(enorm + maxEg);             // R+1
    int* esrc = rp + R + 1;                      // maxEg

    for (int l = 0; l < L; ++l) {
        const int Fin = layer_fin(l, F0, H), off = layer_off(l, F0, H);
        for (int i = tid; i < H * Fin; i += nt) {
            const int f = i / Fin, k = i - f * Fin;
            Wt[off + k * H + f] = a.wb[off + i];
        }
        for (int i = tid; i < H; i += nt) Wt[off + H * Fin + i] = a.wb[off + H * Fin + i];
    }
    if (a.prob_bias)
        for (int i = tid; i < 2 * F0; i += nt) pb[i] = a.prob_bias[i];
    __syncthreads();

    for (int g = blockIdx.x; g < a.B; g += gridDim.x) {
        const int e0 = a.rowptr_t[(int64_t)g * R];
        const int Eg = a.rowptr_t[(int64_t)(g + 1) * R] - e0;
        if (Eg > maxEg) __trap();
        graph_prologue<false>(a, g, e0, Eg, xs, nullptr, rp, esrc, enorm, dinv, nii, nullptr, nullptr, nullptr, nullptr, pb);
        for (int l = 0; l < L; ++l) {
            const int Fin = layer_fin(l, F0, H), off = layer_off(l, F0, H);
            const float* Hprev = (l == 0) ? xs : Hbuf + (l - 1) * H;
            const int ldh = (l == 0) ? F0 : LH;
            dense_xw(Hprev, ldh, Fin, Wt + off, H, R, U);
            __syncthreads();
            const float* bias = Wt + off + H * Fin;
            for (int idx = tid; idx < R * H; idx += nt) {
                const int i = idx / H, f = idx - i * H;
                float acc = 0.f;
                for (int k = rp[i]; k < rp[i + 1]; ++k) acc = fmaf(enorm[k], U[esrc[k] * H + f], acc);
                acc = fmaf(nii[i], U[idx], acc);  // self loop last, as scatter_add sees it
                acc += bias[f];
                Hbuf[i * LH + l * H + f] = a.relu ? fmaxf(acc, 0.f) : acc;
            }
            __syncthreads();
        }
        if (L > 0) {
            float* o = a.out_w + (int64_t)g * R * LH;
            const int n = R * LH;
            if ((n & 3) == 0 && ((((uintptr_t)o) & 15) == 0)) {
                const float4* s4 = reinterpret_cast<const float4*>(Hbuf);
                float4* o4 = reinterpret_cast<float4*>(o);
                for (int i = tid; i < n / 4; i += nt) o4[i] = s4[i];
            } else {
                for (int i = tid; i < n; i += nt) o[i] = Hbuf[i];
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sgcn_encoder_bwd_kernel(EncArgs a) {
    IGCN_PDL_SYNC();
    extern __shared__ float smf[];
    const int R = a.R, F0 = a.F0, H = a.H, L = a.L, LH = L * H, maxEg = a.maxEg;
    const int tid = threadIdx.x, nt = blockDim.x;
    const bool explain = a.prob != nullptr;
    const int WB = wb_size(F0, H, L);
    const int Fmax = H > F0 ? H : F0;
    // carve (floats)
    float* Wt = smf;                    // WB   : transposed weights [k][f] + bias (bias unused here)
    float* acc = Wt + WB;               // P    : per-CTA gradient accumulators [wb layout | dprob | dpb]
    float* pb = acc + a.P;              // 2*F0
    float* xs = pb + 2 * F0;            // R*F0 masked
    float* xraw = xs + R * F0;          // R*F0 raw
    float* dinv = xraw + R * F0;        // R
    float* nii = dinv + R;              // R
    float* ell = nii + R;               // R
    float* dnii = ell + R;              // R
    float* ddeg = dnii + R;             // R
    float* enorm = ddeg + R;            // maxEg
    float* ew = enorm + maxEg;          // maxEg  masked weight w~
    float* epe = ew + maxEg;            // maxEg  p_e
    float* edn = epe + maxEg;           // maxEg  d loss / d norm_e   (later: dz_e)
    float* Hprev = edn + maxEg;         // R*Fmax
    float* U = Hprev + R * Fmax;        // R*H   (U, then dU)
    float* Gl = U + R * H;              // R*Fmax (G_l, then dH^{l-1} in [i*Fin+k])
    float* red = Gl + R * Fmax;         // 8 warps * 16
    int* rp = (int*)(red + 128);        // R+1
    int* rps = rp + R + 1;              // R+1
    int* esrc = rps + R + 1;            // maxEg
    int* etgt = esrc + maxEg;           // maxEg
    int* spos = etgt + maxEg;           // maxEg  CSR slot of the q-th out-edge (source sorted)
    int* lsl = spos + maxEg;            // R

    for (int l = 0; l < L; ++l) {
        const int Fin = layer_fin(l, F0, H), off = layer_off(l, F0, H);
        for (int i = tid; i < H * Fin; i += nt) {
            const int f = i / Fin, k = i - f * Fin;
            Wt[off + k * H + f] = a.wb[off + i];
        }
    }
    for (int i = tid; i < a.P; i += nt) acc[i] = 0.f;
    if (a.prob_bias)
        for (int i = tid; i < 2 * F0; i += nt) pb[i] = a.prob_bias[i];
    float dpb_reg[16];  // per-thread partial of d prob_bias (2*F0 <= 16)
#pragma unroll
    for (int c = 0; c < 16; ++c) dpb_reg[c] = 0.f;
    __syncthreads();

    for (int g = blockIdx.x; g < a.B; g += gridDim.x) {
        const int64_t node0 = (int64_t)g * R;
        const int e0 = a.rowptr_t[node0];
        const int Eg = a.rowptr_t[node0 + R] - e0;
        if (Eg > maxEg) __trap();
        graph_prologue<true>(a, g, e0, Eg, xs, xraw, rp, esrc, enorm, dinv, nii, ew, epe, ell, lsl, pb);
        for (int i = tid; i <= R; i += nt) rps[i] = a.rowptr_s[node0 + i] - e0;
        for (int q = tid; q < Eg; q += nt) {
            spos[q] = a.csc_pos[e0 + q] - e0;
            edn[q] = 0.f;
        }
        for (int i = tid; i < R; i += nt) {
            dnii[i] = 0.f;
            for (int k = rp[i]; k < rp[i + 1]; ++k) etgt[k] = i;
        }
        __syncthreads();

        const float* go = a.g_out + node0 * LH;
        const float* fo = a.out + node0 * LH;
        for (int l = L - 1; l >= 0; --l) {
            const int Fin = layer_fin(l, F0, H), off = layer_off(l, F0, H);
            // G_l = (g_out slot l + dH^l from the layer above) * relu'
            for (int idx = tid; idx < R * H; idx += nt) {
                const int i = idx / H, f = idx - i * H;
                float gsum = go[i * LH + l * H + f];
                if (l < L - 1) gsum += Gl[idx];  // dH^{l} left here (layout [i*H+f]) by the previous iteration
                Gl[idx] = (!a.relu || fo[i * LH + l * H + f] > 0.f) ? gsum : 0.f;
            }
            // H^{l-1}
            if (l == 0) {
                for (int i = tid; i < R * F0; i += nt) Hprev[i] = xs[i];
            } else {
                for (int idx = tid; idx < R * H; idx += nt) {
                    const int i = idx / H, f = idx - i * H;
                    Hprev[idx] = fo[i * LH + (l - 1) * H + f];
                }
            }
            __syncthreads();
            dense_xw(Hprev, Fin, Fin, Wt + off, H, R, U);  // U_l (recomputed)
            __syncthreads();
            if (explain) {
                // d norm_e += <G_l[t_e], U_l[s_e]> ; d n_ii += <G_l[i], U_l[i]>
                for (int k = tid; k < Eg; k += nt) {
                    const float* gp = Gl + etgt[k] * H;
                    const float* up = U + esrc[k] * H;
                    float d = 0.f;
                    for (int f = 0; f < H; ++f) d = fmaf(gp[f], up[f], d);
                    edn[k] += d;
                }
                for (int i = tid; i < R; i += nt) {
                    float d = 0.f;
                    for (int f = 0; f < H; ++f) d = fmaf(Gl[i * H + f], U[i * H + f], d);
                    dnii[i] += d;
                }
            }
            // d bias_l
            for (int f = tid; f < H; f += nt) {
                float s = 0.f;
                for (int i = 0; i < R; ++i) s += Gl[i * H + f];
                acc[off + H * Fin + f] += s;
            }
            __syncthreads();
            // dU[j][f] = sum_{out-edges of j} norm_e G_l[t_e][f] + n_jj G_l[j][f]   (transposed SpMM, source-sorted list)
            for (int idx = tid; idx < R * H; idx += nt) {
                const int j = idx / H, f = idx - j * H;
                float s = 0.f;
                for (int q = rps[j]; q < rps[j + 1]; ++q) {
                    const int k = spos[q];
                    s = fmaf(enorm[k], Gl[etgt[k] * H + f], s);
                }
                s = fmaf(nii[j], Gl[idx], s);
                U[idx] = s;
            }
            __syncthreads();
            // dW_l[f][k] += sum_i dU[i][f] Hprev[i][k]
            for (int idx = tid; idx < H * Fin; idx += nt) {
                const int f = idx / Fin, k = idx - f * Fin;
                float s = 0.f;
                for (int i = 0; i < R; ++i) s = fmaf(U[i * H + f], Hprev[i * Fin + k], s);
                acc[off + idx] += s;
            }
            // dH^{l-1}[i][k] = sum_f dU[i][f] W_l[f][k]   -> Gl buffer in [i*Fin+k] layout
            for (int idx = tid; idx < R * Fin; idx += nt) {
                const int i = idx / Fin, k = idx - i * Fin;
                float s = 0.f;
                for (int f = 0; f < H; ++f) s = fmaf(U[i * H + f], Wt[off + k * H + f], s);
                Gl[idx] = s;
            }
            __syncthreads();
        }
        // here Gl holds d x~ (R*F0) from the conv stack (zero-layer call: nothing)
        float* dxt = Gl;
        if (L == 0) {
            for (int i = tid; i < R * F0; i += nt) dxt[i] = 0.f;
            __syncthreads();
        }
        if (explain) {
            // gradient through the symmetric normalisation
            for (int i = tid; i < R; i += nt) {
                float dd = 0.f;
                for (int k = rp[i]; k < rp[i + 1]; ++k) {  // as target
                    const int s = esrc[k];
                    if (s != i) dd = fmaf(edn[k] * ew[k], dinv[s], dd);
                }
                for (int q = rps[i]; q < rps[i + 1]; ++q) {  // as source
                    const int k = spos[q];
                    const int t = etgt[k];
                    if (t != i) dd = fmaf(edn[k] * ew[k], dinv[t], dd);
                }
                const float di = dinv[i];
                dd = fmaf(2.f * di * ell[i], dnii[i], dd);
                ddeg[i] = -0.5f * di * di * di * dd;
            }
            __syncthreads();
            // d w~ -> d p_e -> d z_e ; accumulate d prob_bias, d x~ (target side)
            for (int i = tid; i < R; i += nt) {
                const float di = dinv[i];
                float dxi[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) dxi[c] = 0.f;
                for (int k = rp[i]; k < rp[i + 1]; ++k) {
                    const int s = esrc[k];
                    float dwt;
                    if (s != i)
                        dwt = dinv[s] * di * edn[k] + ddeg[i];
                    else
                        // every self-loop slot receives d l_i, also an overwritten duplicate: that is what autograd
                        // of PyG's `loop_attr[row[~mask]] = edge_attr[~mask]` (index_put_) hands back
                        dwt = di * di * dnii[i] + ddeg[i];
                    float dp = a.csr_w[e0 + k] * dwt;
                    if (a.g_pe) dp += a.g_pe[e0 + k];
                    const float p = epe[k];
                    const float dz = p * (1.f - p) * dp;
                    edn[k] = dz;  // reuse: d z_e
                    for (int c = 0; c < F0; ++c) {
                        dpb_reg[c] = fmaf(dz, xs[s * F0 + c], dpb_reg[c]);
                        dpb_reg[F0 + c] = fmaf(dz, xs[i * F0 + c], dpb_reg[F0 + c]);
                        dxi[c] = fmaf(dz, pb[F0 + c], dxi[c]);
                    }
                }
                for (int c = 0; c < F0; ++c) dxt[i * F0 + c] += dxi[c];
            }
            __syncthreads();
            // source side of d x~ ; then d prob and dx
            for (int i = tid; i < R; i += nt) {
                float sdz = 0.f;
                for (int q = rps[i]; q < rps[i + 1]; ++q) sdz += edn[spos[q]];
                for (int c = 0; c < F0; ++c) {
                    const float dv = fmaf(sdz, pb[c], dxt[i * F0 + c]);
                    acc[WB + i * F0 + c] += xraw[i * F0 + c] * dv;
                    a.dx[(node0 + i) * F0 + c] = a.prob[i * F0 + c] * dv;
                }
            }
        } else {
            for (int i = tid; i < R * F0; i += nt) a.dx[node0 * F0 + i] = dxt[i];
        }
        __syncthreads();
    }
    // CTA-level reduction of d prob_bias, then one partial row per CTA
    if (explain) {
        const int lane = tid & 31, warp = tid >> 5;
        for (int c = 0; c < 2 * F0; ++c) {
            float v = warp_sum(dpb_reg[c]);
            if (lane == 0) red[warp * 16 + c] = v;
        }
        __syncthreads();
        if (tid < 2 * F0) {
            float s = 0.f;
            for (int w = 0; w < (nt >> 5); ++w) s += red[w * 16 + tid];
            acc[WB + R * F0 + tid] = s;
        }
        __syncthreads();
    }
    float* prow = a.partials + (int64_t)blockIdx.x * a.P;
    for (int i = tid; i < a.P; i += nt) prow[i] = acc[i];
}

}  // namespace igcn
#include "sgcn_fast.cuh"
#include "sgcn_mma.cuh"
namespace igcn {

// second-generation kernels (sgcn_mma.cuh): reference shape only; IGCN_SGCN_MMA=0 keeps the register-tiled kernels (A/B hook)
static bool use_mma(int64_t R, int64_t F0, int64_t H, int64_t L, int64_t relu) {
    const char* e = getenv("IGCN_SGCN_MMA");
    if (e && e[0] == '0') return false;
    return F0 == kF0 && H == kH && L == 2 && relu && R <= mma::kMaxThreads;
}
static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

static bool use_fast(int64_t F0, int64_t H, int64_t L) {
    const char* e = getenv("IGCN_FORCE_GENERIC");   // test hook: exercise the shape-generic kernels
    if (e && e[0] == '1') return false;
    return F0 == kF0 && H == kH && L >= 1;
}

static size_t fwd_smem(int R, int F0, int H, int L, int maxEg) {
    size_t fl = (size_t)wb_size(F0, H, L) + 2 * F0 + (size_t)R * F0 + 2 * R + maxEg + (size_t)R * L * H + 4 + (size_t)R * H;
    size_t in = (size_t)R + 1 + maxEg;
    return 4 * (fl + in);
}
static size_t bwd_smem(int R, int F0, int H, int L, int maxEg, int P) {
    const int Fmax = H > F0 ? H : F0;
    size_t fl = (size_t)wb_size(F0, H, L) + P + 2 * F0 + 2 * (size_t)R * F0 + 5 * (size_t)R + 4 * (size_t)maxEg +
                (size_t)R * Fmax + (size_t)R * H + (size_t)R * Fmax + 128;
    size_t in = 2 * ((size_t)R + 1) + 3 * (size_t)maxEg + R;
    return 4 * (fl + in);
}

static int check_shapes(const char* who, int64_t B, int64_t R, int64_t F0, int64_t H, int64_t L, int64_t max_eg) {
    IGCN_REQUIRE(B >= 0 && R > 0 && F0 > 0 && H >= 0 && L >= 0 && max_eg >= 0, IGCN_ERR_BAD_ARG, "%s: negative size", who);
    IGCN_REQUIRE(F0 <= 8, IGCN_ERR_UNSUPPORTED, "%s: F0=%lld > 8 input features not supported", who, (long long)F0);
    IGCN_REQUIRE(L == 0 || (H > 0 && H <= 128), IGCN_ERR_UNSUPPORTED, "%s: hidden=%lld outside 1..128", who, (long long)H);
    IGCN_REQUIRE(B * R < (1ll << 31), IGCN_ERR_UNSUPPORTED, "%s: more than 2^31 nodes", who);
    return IGCN_OK;
}

}  // namespace igcn

using namespace igcn;

extern "C" int64_t igcn_sgcn_param_count(int64_t R, int64_t F0, int64_t H, int64_t L) {
    return (int64_t)wb_size((int)F0, (int)H, (int)L) + R * F0 + 2 * F0;
}

static int ctas_for(size_t smem, int64_t B) {
    int per_sm = (int)((227 * 1024) / (smem + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 8) per_sm = 8;   // 256 threads each -> 2048 threads/SM
    int64_t n = (int64_t)sm_count() * per_sm;
    if (n > B) n = B;
    if (n < 1) n = 1;
    return (int)n;
}

static bool use_fast_bwd(int64_t R, int64_t F0, int64_t H, int64_t L, int64_t max_eg) {
    return use_fast(F0, H, L) && L == 2 && bwd_fast_smem((int)R, (int)max_eg) <= 227 * 1024;
}
// tensor-core backward (sgcn_mma.cuh): also one CTA per SM, so igcn_sgcn_bwd_ctas() is the same for both L == 2 paths
static bool use_mma_bwd(int64_t R, int64_t F0, int64_t H, int64_t L, int64_t max_eg, bool explain) {
    return use_fast(F0, H, L) && use_mma(R, F0, H, L, 1) &&
           mma::bwd_mma_smem((int)R, (int)max_eg, mma::mma_bwd_threads((int)R), explain) <= 227 * 1024;
}

extern "C" int64_t igcn_sgcn_bwd_ctas(int64_t B, int64_t R, int64_t F0, int64_t H, int64_t L, int64_t max_eg) {
    if (use_mma_bwd(R, F0, H, L, max_eg, true)) {
        // as many CTAs per SM as shared memory, threads and the 96-register budget allow (3 at 90 ROIs, 1 at 264): a small batch
        // then runs one graph per CTA instead of queueing graphs behind each other
        const int nthr = mma::mma_bwd_threads((int)R);
        const size_t smem = mma::bwd_mma_smem((int)R, (int)max_eg, nthr, true);
        int per_sm = (int)((227 * 1024) / (smem + 1024));
        if (per_sm > 2048 / nthr) per_sm = 2048 / nthr;
        if (per_sm > 65536 / (96 * nthr)) per_sm = 65536 / (96 * nthr);
        if (per_sm < 1) per_sm = 1;
        return balanced_ctas((int64_t)sm_count() * per_sm, B);
    }
    if (use_fast_bwd(R, F0, H, L, max_eg)) {
        int64_t n = sm_count();
        if (n > B) n = B;
        return n < 1 ? 1 : n;
    }
    int P = (int)igcn_sgcn_param_count(R, F0, H, L);
    return ctas_for(bwd_smem((int)R, (int)F0, (int)H, (int)L, (int)max_eg, P), B);
}

extern "C" int igcn_sgcn_encoder_fwd(const float* x, const int32_t* rowptr_t, const int32_t* csr_src, const float* csr_w,
                                     const float* prob, const float* prob_bias, const float* wb, int64_t B, int64_t R,
                                     int64_t F0, int64_t H, int64_t L, int64_t max_eg, int64_t relu, float* out, float* p_e, void* stream) {
    int rc = check_shapes("sgcn_encoder_fwd", B, R, F0, H, L, max_eg);
    if (rc) return rc;
    IGCN_REQUIRE(x && rowptr_t, IGCN_ERR_BAD_ARG, "sgcn_encoder_fwd: null x/rowptr");
    IGCN_REQUIRE(max_eg == 0 || (csr_src && csr_w), IGCN_ERR_BAD_ARG, "sgcn_encoder_fwd: null CSR arrays");
    IGCN_REQUIRE((prob == nullptr) == (prob_bias == nullptr), IGCN_ERR_BAD_ARG, "sgcn_encoder_fwd: prob and prob_bias go together");
    IGCN_REQUIRE(L == 0 || (wb && out), IGCN_ERR_BAD_ARG, "sgcn_encoder_fwd: null weights/output");
    if (B == 0) return IGCN_OK;
    EncArgs a{};
    a.x = x; a.rowptr_t = rowptr_t; a.csr_src = csr_src; a.csr_w = csr_w; a.prob = prob; a.prob_bias = prob_bias; a.wb = wb;
    a.out_w = out; a.pe_w = p_e; a.relu = relu ? 1 : 0; a.B = (int)B; a.R = (int)R; a.F0 = (int)F0; a.H = (int)H; a.L = (int)L; a.maxEg = (int)max_eg;
    if (use_fast(F0, H, L) && use_mma(R, F0, H, L, relu) && aligned16(x) && aligned16(rowptr_t) && aligned16(csr_src) && aligned16(csr_w)) {
        const int nthr = mma::mma_threads(a.R);
        const size_t smem = mma::fwd_mma_smem(a.R, a.maxEg, nthr);
        if (smem <= 227 * 1024) {
            // CTAs per SM the register budget is compiled for: 3 (72 registers) or 2 (112); IGCN_SGCN_FWD_OCC=2|3 is the A/B hook
            const char* occ = getenv("IGCN_SGCN_FWD_OCC");
            const bool occ2 = occ && occ[0] == '2';
            auto kern = prob ? (occ2 ? mma::sgcn_fwd_mma_kernel<true, 2> : mma::sgcn_fwd_mma_kernel<true, 3>)
                             : (occ2 ? mma::sgcn_fwd_mma_kernel<false, 2> : mma::sgcn_fwd_mma_kernel<false, 3>);
            rc = allow_smem(kern, smem, "sgcn_fwd_mma");
            if (rc) return rc;
            int per_sm = (int)((227 * 1024) / (smem + 1024));
            if (occ2 && per_sm > 2 && nthr > 192) per_sm = 2;
            const int by_threads = 2048 / nthr;
            if (per_sm > by_threads) per_sm = by_threads;
            if (per_sm > 8) per_sm = 8;
            if (per_sm < 1) per_sm = 1;
            const int64_t grid = balanced_ctas((int64_t)sm_count() * per_sm, B);
            igcn::launch_k(kern, dim3((int)grid), dim3(nthr), smem, (cudaStream_t)stream, a);
            IGCN_CHECK_LAUNCH("sgcn_fwd_mma");
            return IGCN_OK;
        }
    }
    if (use_fast(F0, H, L)) {
        size_t smem = fwd_fast_smem(a.R, a.L, a.maxEg);
        if (smem <= 227 * 1024) {
            auto kern = (L == 2) ? sgcn_fwd_h16_kernel<true> : sgcn_fwd_h16_kernel<false>;
            rc = allow_smem(kern, smem, "sgcn_fwd_h16");
            if (rc) return rc;
            const int nthr = fast_threads(a.R);
            int per_sm = (int)((227 * 1024) / (smem + 1024));
            const int by_regs = 65536 / (128 * nthr);
            if (per_sm > by_regs) per_sm = by_regs;
            if (per_sm < 1) per_sm = 1;
            const int64_t grid = balanced_ctas((int64_t)sm_count() * per_sm, B);
            igcn::launch_k(kern, dim3((int)grid), dim3(nthr), smem, (cudaStream_t)stream, a);
            IGCN_CHECK_LAUNCH("sgcn_fwd_h16");
            return IGCN_OK;
        }
    }
    size_t smem = fwd_smem(a.R, a.F0, a.H, a.L, a.maxEg);
    rc = allow_smem(sgcn_encoder_fwd_kernel, smem, "sgcn_encoder_fwd");
    if (rc) return rc;
    igcn::launch_k(sgcn_encoder_fwd_kernel, dim3(ctas_for(smem, B)), dim3(256), smem, (cudaStream_t)stream, a);
    IGCN_CHECK_LAUNCH("sgcn_encoder_fwd");
    return IGCN_OK;
}

extern "C" int igcn_sgcn_encoder_bwd(const float* x, const int32_t* rowptr_t, const int32_t* csr_src, const float* csr_w,
                                     const int32_t* rowptr_s, const int32_t* csc_pos, const float* prob, const float* prob_bias,
                                     const float* wb, const float* out, const float* g_out, const float* g_pe, int64_t B,
                                     int64_t R, int64_t F0, int64_t H, int64_t L, int64_t max_eg, int64_t relu, float* dx, float* partials,
                                     int64_t n_cta, float* grads, void* stream) {
    int rc = check_shapes("sgcn_encoder_bwd", B, R, F0, H, L, max_eg);
    if (rc) return rc;
    IGCN_REQUIRE(x && rowptr_t && rowptr_s && dx && partials && grads, IGCN_ERR_BAD_ARG, "sgcn_encoder_bwd: null pointer");
    IGCN_REQUIRE(max_eg == 0 || (csr_src && csr_w && csc_pos), IGCN_ERR_BAD_ARG, "sgcn_encoder_bwd: null CSR arrays");
    IGCN_REQUIRE((prob == nullptr) == (prob_bias == nullptr), IGCN_ERR_BAD_ARG, "sgcn_encoder_bwd: prob and prob_bias go together");
    IGCN_REQUIRE(L == 0 || (wb && out && g_out), IGCN_ERR_BAD_ARG, "sgcn_encoder_bwd: null weights/activations");
    IGCN_REQUIRE(g_pe == nullptr || prob != nullptr, IGCN_ERR_BAD_ARG, "sgcn_encoder_bwd: g_pe without masks");
    EncArgs a{};
    a.x = x; a.rowptr_t = rowptr_t; a.csr_src = csr_src; a.csr_w = csr_w; a.rowptr_s = rowptr_s; a.csc_pos = csc_pos;
    a.prob = prob; a.prob_bias = prob_bias; a.wb = wb; a.out = out; a.g_out = g_out; a.g_pe = g_pe; a.dx = dx; a.partials = partials;
    a.B = (int)B; a.R = (int)R; a.F0 = (int)F0; a.H = (int)H; a.L = (int)L; a.maxEg = (int)max_eg;
    a.P = (int)igcn_sgcn_param_count(R, F0, H, L);
    a.relu = relu ? 1 : 0;
    const int want = (int)igcn_sgcn_bwd_ctas(B, R, F0, H, L, max_eg);
    IGCN_REQUIRE(B == 0 || n_cta == want, IGCN_ERR_BAD_ARG, "sgcn_encoder_bwd: n_cta=%lld, expected igcn_sgcn_bwd_ctas()=%d",
                 (long long)n_cta, want);
    cudaStream_t st = (cudaStream_t)stream;
    if (B == 0) {
        cudaMemsetAsync(grads, 0, sizeof(float) * a.P, st);
        return IGCN_OK;
    }
    const bool al = aligned16(x) && aligned16(rowptr_t) && aligned16(csr_src) && aligned16(csr_w) && aligned16(rowptr_s) &&
                    aligned16(csc_pos) && aligned16(out) && aligned16(g_out) && (g_pe == nullptr || aligned16(g_pe));
    if (relu && al && use_mma_bwd(R, F0, H, L, max_eg, prob != nullptr)) {
        const int nthr = mma::mma_bwd_threads(a.R);
        size_t smem = mma::bwd_mma_smem(a.R, a.maxEg, nthr, prob != nullptr);
        auto kern = prob ? mma::sgcn_bwd_mma_kernel<true> : mma::sgcn_bwd_mma_kernel<false>;
        rc = allow_smem(kern, smem, "sgcn_bwd_mma");
        if (rc) return rc;
        igcn::launch_k(kern, dim3(want), dim3(nthr), smem, st, a);
        IGCN_CHECK_LAUNCH("sgcn_bwd_mma");
    } else if (use_fast_bwd(R, F0, H, L, max_eg)) {
        size_t smem = bwd_fast_smem(a.R, a.maxEg);
        auto kern = prob ? sgcn_bwd_h16_kernel<true> : sgcn_bwd_h16_kernel<false>;
        rc = allow_smem(kern, smem, "sgcn_bwd_h16");
        if (rc) return rc;
        igcn::launch_k(kern, dim3(want), dim3(fast_threads_bwd(a.R)), smem, st, a);
        IGCN_CHECK_LAUNCH("sgcn_bwd_h16");
    } else {
        size_t smem = bwd_smem(a.R, a.F0, a.H, a.L, a.maxEg, a.P);
        rc = allow_smem(sgcn_encoder_bwd_kernel, smem, "sgcn_encoder_bwd");
        if (rc) return rc;
        igcn::launch_k(sgcn_encoder_bwd_kernel, dim3(want), dim3(256), smem, st, a);
        IGCN_CHECK_LAUNCH("sgcn_encoder_bwd");
    }
    igcn::launch_k(reduce_partials_kernel, dim3((a.P + 31) / 32), dim3(reduce_threads(want)), 0, st, partials, want, a.P, grads);
    IGCN_CHECK_LAUNCH("sgcn_reduce_partials");
    return IGCN_OK;
}
