// GO-hierarchy encoder kernels for sm_100a (reference: kernel/go_model.py).
//
//   go_spmm        SNP->GO encode / GO->SNP decode: out[b,r,c] = sum_k vals[c][k] * in[b,col[k]]   (go_model.py:208-215, 281-282)
//   go_layer       one hierarchy layer, fused per subject:
//                    ATTN    x_in = X W_a^T, x_s = X W_s^T, a_e = exp(tanh(u.[x_in_i | x_in_j])), row-normalise, aggregate,
//                            + x_s * sigmoid(v.x_s)                                       (go_model.py:226-244)
//                    UNIFORM weights 1/|row|, self term on rows >= self_off               (go_model.py:262-272)
//                  then LayerNorm over the NODE axis per (subject, channel), ReLU, node-dropout scale, and the
//                  hierarchical pooling slice [keep_from:]                                  (go_model.py:246-251, 273-275)
//
// The reference walks subjects in a Python loop with ~8 sparse-op launches each (go_model.py:236-244); here a CTA owns a
// subject, the shared DAG (CSR by row + CSC by column, built once on the host) is read through L2, every intermediate lives in
// shared memory, and the transposed scatter of the backward pass is a gather over the CSC -- no float atomics.
// Parameter gradients: per-thread registers -> per-CTA partial row -> fixed-order reduction kernel.
#include "common.cuh"

namespace igcn {

struct GoGraph {
    const int32_t* rowptr;  // (Mrow+1)
    const int32_t* col;     // (nnz)  column (input node) of every CSR slot
    const int32_t* colptr;  // (Min+1)
    const int32_t* crow;    // (nnz)  row of the q-th CSC entry
    const int32_t* cpos;    // (nnz)  CSR slot of the q-th CSC entry
    int Mrow, Min, nnz;
};

// ------------------------------------------------------------------------------------------------
// SpMM with learnable per-nnz values, C value channels sharing one pattern.
// ------------------------------------------------------------------------------------------------
// Row lengths of the SNP <-> GO incidence are bimodal: a few entries per row (a SNP sits in ~3 terms, a term owns ~15 SNPs) and ONE
// row with every SNP (the root term, snps_graph.py:247-248).  Short rows are walked by one thread each (32x the parallelism of a
// warp per row: at G = 2 000 / S = 10 000 the warp-per-row version spent 770 us in dependent rowptr -> col -> x loads); rows longer
// than kHeavy entries are queued in shared memory and summed by the whole CTA afterwards, in a fixed order.
constexpr int kHeavy = 64, kMaxHeavy = 32;

template <int C>
__global__ void __launch_bounds__(256) go_spmm_fwd_kernel(const float* __restrict__ in, const int32_t* __restrict__ rowptr,
                                                          const int32_t* __restrict__ col, const float* __restrict__ vals,
                                                          int B, int Nin, int Nrow, int nnz, float* __restrict__ out) {
    IGCN_PDL_SYNC();
    extern __shared__ float smf[];
    float* xin = smf;  // Nin
    __shared__ int heavy[kMaxHeavy];
    __shared__ int n_heavy;
    __shared__ float red[8 * C];
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5;
    for (int b = blockIdx.x; b < B; b += gridDim.x) {
        if (tid == 0) n_heavy = 0;
        for (int i = tid; i < Nin; i += nt) xin[i] = in[(int64_t)b * Nin + i];
        __syncthreads();
        for (int r = tid; r < Nrow; r += nt) {
            const int k0 = rowptr[r], k1 = rowptr[r + 1];
            if (k1 - k0 > kHeavy) {
                const int slot = atomicAdd(&n_heavy, 1);
                if (slot < kMaxHeavy) heavy[slot] = r;
                if (slot < kMaxHeavy) continue;
            }
            float acc[C];
#pragma unroll
            for (int c = 0; c < C; ++c) acc[c] = 0.f;
            int k = k0;
            for (; k + 3 < k1; k += 4) {        // four entries per trip: indices and values are loaded before the first use
                int cj[4];
                float vv[4][C];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    cj[u] = col[k + u];
#pragma unroll
                    for (int c = 0; c < C; ++c) vv[u][c] = vals[(int64_t)c * nnz + k + u];
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const float xv = xin[cj[u]];
#pragma unroll
                    for (int c = 0; c < C; ++c) acc[c] = fmaf(vv[u][c], xv, acc[c]);
                }
            }
            for (; k < k1; ++k) {
                const float xv = xin[col[k]];
#pragma unroll
                for (int c = 0; c < C; ++c) acc[c] = fmaf(vals[(int64_t)c * nnz + k], xv, acc[c]);
            }
#pragma unroll
            for (int c = 0; c < C; ++c) out[((int64_t)b * Nrow + r) * C + c] = acc[c];
        }
        __syncthreads();
        const int nh = min(n_heavy, kMaxHeavy);
        for (int hidx = 0; hidx < nh; ++hidx) {
            const int r = heavy[hidx];
            float acc[C];
#pragma unroll
            for (int c = 0; c < C; ++c) acc[c] = 0.f;
#pragma unroll 4
            for (int k = rowptr[r] + tid; k < rowptr[r + 1]; k += nt) {
                const float xv = xin[col[k]];
#pragma unroll
                for (int c = 0; c < C; ++c) acc[c] = fmaf(vals[(int64_t)c * nnz + k], xv, acc[c]);
            }
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const float s = warp_sum(acc[c]);
                if (lane == 0) red[warp * C + c] = s;
            }
            __syncthreads();
            if (tid < C) {
                float t = 0.f;
                for (int w = 0; w < (nt >> 5); ++w) t += red[w * C + tid];
                out[((int64_t)b * Nrow + r) * C + tid] = t;
            }
            __syncthreads();
        }
    }
}

// d_in[b,s] = sum_{q in col s} sum_c vals[c][cpos q] * g[b, crow q, c]      (same short / heavy split over the columns)
template <int C>
__global__ void __launch_bounds__(256) go_spmm_bwd_in_kernel(const float* __restrict__ g, const int32_t* __restrict__ colptr,
                                                             const int32_t* __restrict__ crow, const int32_t* __restrict__ cpos,
                                                             const float* __restrict__ vals, int B, int Nin, int Nrow, int nnz,
                                                             float* __restrict__ d_in) {
    IGCN_PDL_SYNC();
    extern __shared__ float smf[];
    float* gs = smf;  // Nrow*C
    __shared__ int heavy[kMaxHeavy];
    __shared__ int n_heavy;
    __shared__ float red[8];
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5;
    for (int b = blockIdx.x; b < B; b += gridDim.x) {
        if (tid == 0) n_heavy = 0;
        for (int i = tid; i < Nrow * C; i += nt) gs[i] = g[(int64_t)b * Nrow * C + i];
        __syncthreads();
        for (int s = tid; s < Nin; s += nt) {
            const int q0 = colptr[s], q1 = colptr[s + 1];
            if (q1 - q0 > kHeavy) {
                const int slot = atomicAdd(&n_heavy, 1);
                if (slot < kMaxHeavy) heavy[slot] = s;
                if (slot < kMaxHeavy) continue;
            }
            float acc = 0.f;
            int q = q0;
            for (; q + 3 < q1; q += 4) {        // four entries per trip (was two dependent trips per entry: cpos -> vals)
                int rr[4], kk[4];
                float vv[4][C];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    rr[u] = crow[q + u];
                    kk[u] = cpos[q + u];
                }
#pragma unroll
                for (int u = 0; u < 4; ++u)
#pragma unroll
                    for (int c = 0; c < C; ++c) vv[u][c] = vals[(int64_t)c * nnz + kk[u]];
#pragma unroll
                for (int u = 0; u < 4; ++u)
#pragma unroll
                    for (int c = 0; c < C; ++c) acc = fmaf(vv[u][c], gs[rr[u] * C + c], acc);
            }
            for (; q < q1; ++q) {
                const int r = crow[q], k = cpos[q];
#pragma unroll
                for (int c = 0; c < C; ++c) acc = fmaf(vals[(int64_t)c * nnz + k], gs[r * C + c], acc);
            }
            d_in[(int64_t)b * Nin + s] = acc;
        }
        __syncthreads();
        const int nh = min(n_heavy, kMaxHeavy);
        for (int hidx = 0; hidx < nh; ++hidx) {
            const int s = heavy[hidx];
            float acc = 0.f;
#pragma unroll 4
            for (int q = colptr[s] + tid; q < colptr[s + 1]; q += nt) {
                const int r = crow[q], k = cpos[q];
#pragma unroll
                for (int c = 0; c < C; ++c) acc = fmaf(vals[(int64_t)c * nnz + k], gs[r * C + c], acc);
            }
            const float ws = warp_sum(acc);
            if (lane == 0) red[warp] = ws;
            __syncthreads();
            if (tid == 0) {
                float t = 0.f;
                for (int w = 0; w < (nt >> 5); ++w) t += red[w];
                d_in[(int64_t)b * Nin + s] = t;
            }
            __syncthreads();
        }
    }
}

// d_vals[c][k] = sum_b g[b,row(k),c] * in[b,col[k]]  -- one WARP per nnz: lanes stride over the batch, fixed-order
// shuffle reduction (deterministic).  g and in are small (B x n_row x C, B x n_in) and stay L2 resident.
template <int C>
__global__ void __launch_bounds__(256) go_spmm_bwd_vals_kernel(const float* __restrict__ g, const float* __restrict__ in,
                                                               const int32_t* __restrict__ row_of, const int32_t* __restrict__ col,
                                                               int B, int Nin, int Nrow, int nnz, float* __restrict__ d_vals) {
    IGCN_PDL_SYNC();
    const int lane = threadIdx.x & 31;
    const int k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (k >= nnz) return;
    const int r = row_of[k], s = col[k];
    float acc[C];
#pragma unroll
    for (int c = 0; c < C; ++c) acc[c] = 0.f;
#pragma unroll 4
    for (int b = lane; b < B; b += 32) {      // unrolled: the loads of four iterations are in flight together (was one L2 trip per iteration)
        const float xv = in[(int64_t)b * Nin + s];
#pragma unroll
        for (int c = 0; c < C; ++c) acc[c] = fmaf(g[((int64_t)b * Nrow + r) * C + c], xv, acc[c]);
    }
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const float t = warp_sum(acc[c]);
        if (lane == 0) d_vals[(int64_t)c * nnz + k] = t;
    }
}


// ---- large-hierarchy path of d_vals: the batch-strided reads above cost one 32-byte sector per value (config 3: 40 K nnz x 256
//      subjects x 2 operands = 650 MB of sector traffic, 890 us).  With g and in transposed once (batch contiguous) every nnz is a
//      dot product of two contiguous B-vectors: coalesced, 10x less traffic.  Same lane partition and shuffle order as above, so
//      the results are bit identical.
__global__ void __launch_bounds__(256) transpose2_kernel(const float* __restrict__ a, int ra, int ca, float* __restrict__ at,
                                                         const float* __restrict__ b, int rb, int cb, float* __restrict__ bt) {
    IGCN_PDL_SYNC();
    __shared__ float tile[32][33];
    const float* src = blockIdx.y ? b : a;
    float* dst = blockIdx.y ? bt : at;
    const int R0 = blockIdx.y ? rb : ra, C0 = blockIdx.y ? cb : ca;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int tiles_c = (C0 + 31) / 32, tiles = tiles_c * ((R0 + 31) / 32);
    for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
        const int r0 = (t / tiles_c) * 32, c0 = (t % tiles_c) * 32;
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = r0 + ty + 8 * i, c = c0 + tx;
            tile[ty + 8 * i][tx] = (r < R0 && c < C0) ? src[(int64_t)r * C0 + c] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int c = c0 + ty + 8 * i, r = r0 + tx;
            if (r < R0 && c < C0) dst[(int64_t)c * R0 + r] = tile[tx][ty + 8 * i];
        }
    }
}

template <int C>
__global__ void __launch_bounds__(256) go_spmm_bwd_vals_t_kernel(const float* __restrict__ gT /* (Nrow*C, B) */,
                                                                 const float* __restrict__ inT /* (Nin, B) */,
                                                                 const int32_t* __restrict__ row_of, const int32_t* __restrict__ col, int B,
                                                                 int nnz, float* __restrict__ d_vals) {
    IGCN_PDL_SYNC();
    const int lane = threadIdx.x & 31;
    for (int k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); k < nnz; k += gridDim.x * (blockDim.x >> 5)) {
        const int r = row_of[k], s = col[k];
        const float* xi = inT + (int64_t)s * B;
        float acc[C];
#pragma unroll
        for (int c = 0; c < C; ++c) acc[c] = 0.f;
#pragma unroll 4
        for (int b = lane; b < B; b += 32) {
            const float xv = xi[b];
#pragma unroll
            for (int c = 0; c < C; ++c) acc[c] = fmaf(gT[((int64_t)r * C + c) * B + b], xv, acc[c]);
        }
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const float t = warp_sum(acc[c]);
            if (lane == 0) d_vals[(int64_t)c * nnz + k] = t;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// hierarchy layer
// ------------------------------------------------------------------------------------------------
struct GoLayerArgs {
    const float* x;      // (B, Min, DIN)
    const float* Wa;     // (DOUT, DIN)   w_inc / w_out
    const float* Ws;     // (DOUT, DIN)   w_s_loop / w_s_loop_out
    const float* u;      // (2*DOUT)      w_att_in  [row half | col half]   (ATTN)
    const float* v;      // (DOUT)        w_att_s                            (ATTN)
    const float* gamma;  // (Mrow)        LayerNorm over nodes
    const float* beta;   // (Mrow)
    const float* mask;   // (B, Mrow) dropout scale or null
    GoGraph gr;
    int B, self_off, keep_from;
    float* y;            // (B, Mrow-keep_from, DOUT)
    float* stats;        // (B, 2*DOUT): mean | rstd  (saved for bwd)
    // bwd
    const float* gy;     // (B, Mrow-keep_from, DOUT)
    float* dx;           // (B, Min, DIN)
    float* partials;     // (n_cta, P)
    int P;
};

template <int N>
__device__ __forceinline__ void block_sum(float (&v)[N], float* scratch /* 8*N */, float (&out)[N]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
    for (int c = 0; c < N; ++c) {
        float s = warp_sum(v[c]);
        if (lane == 0) scratch[warp * N + c] = s;
    }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < N; ++c) {
        float s = 0.f;
        for (int w = 0; w < nw; ++w) s += scratch[w * N + c];
        out[c] = s;
    }
    __syncthreads();
}

// Phases 1+2 shared by fwd and bwd: Xin, Xs, pre-norm output O (all in smem).
// With kKeepEdges the per-slot attention coefficient alpha_e and tanh value are stored (bwd).
template <int DIN, int DOUT, bool ATTN, bool kKeepEdges>
__device__ __forceinline__ void go_layer_pre(const GoLayerArgs& a, int b, const float* Wa_s, const float* Ws_s, const float* u_s,
                                             const float* v_s, float* Xin, float* Xs, float* O, float* e_alpha, float* e_th) {
    const int tid = threadIdx.x, nt = blockDim.x;
    const int Min = a.gr.Min, Mrow = a.gr.Mrow;
    const float* xb = a.x + (int64_t)b * Min * DIN;
    for (int j = tid; j < Min; j += nt) {
        float xv[DIN];
#pragma unroll
        for (int k = 0; k < DIN; ++k) xv[k] = xb[j * DIN + k];
#pragma unroll
        for (int f = 0; f < DOUT; ++f) {
            float s1 = 0.f, s2 = 0.f;
#pragma unroll
            for (int k = 0; k < DIN; ++k) {
                s1 = fmaf(xv[k], Wa_s[f * DIN + k], s1);
                s2 = fmaf(xv[k], Ws_s[f * DIN + k], s2);
            }
            Xin[j * DOUT + f] = s1;
            Xs[j * DOUT + f] = s2;
        }
    }
    __syncthreads();
    for (int i = tid; i < Mrow; i += nt) {
        float acc[DOUT];
#pragma unroll
        for (int f = 0; f < DOUT; ++f) acc[f] = 0.f;
        const int k0 = a.gr.rowptr[i], k1 = a.gr.rowptr[i + 1];
        if (ATTN) {
            float qi = 0.f;
#pragma unroll
            for (int f = 0; f < DOUT; ++f) qi = fmaf(u_s[f], Xin[i * DOUT + f], qi);
            float S = 0.f;
            auto edge = [&](int k, int j) {
                float q = qi;
#pragma unroll
                for (int f = 0; f < DOUT; ++f) q = fmaf(u_s[DOUT + f], Xin[j * DOUT + f], q);
                const float th = tanhf(q);
                const float ae = __expf(th);
                S += ae;
                if (kKeepEdges) {
                    e_alpha[k] = ae;
                    e_th[k] = th;
                }
#pragma unroll
                for (int f = 0; f < DOUT; ++f) acc[f] = fmaf(ae, Xin[j * DOUT + f], acc[f]);
            };
            // a GO term has <= 3 parents: the first four column indices of the row are loaded together (one trip to L2, not one per edge)
            int cj[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) cj[u] = (k0 + u < k1) ? a.gr.col[k0 + u] : 0;
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (k0 + u < k1) edge(k0 + u, cj[u]);
            for (int k = k0 + 4; k < k1; ++k) edge(k, a.gr.col[k]);
            const float inv = (k1 > k0) ? 1.f / S : 0.f;
#pragma unroll
            for (int f = 0; f < DOUT; ++f) acc[f] *= inv;
            if (kKeepEdges)
                for (int k = k0; k < k1; ++k) e_alpha[k] *= inv;
            // self influence with its scalar gate (square layer: row i is node i)
            float gz = 0.f;
#pragma unroll
            for (int f = 0; f < DOUT; ++f) gz = fmaf(v_s[f], Xs[i * DOUT + f], gz);
            const float gate = sigmoidf_(gz);
#pragma unroll
            for (int f = 0; f < DOUT; ++f) acc[f] = fmaf(Xs[i * DOUT + f], gate, acc[f]);
        } else {
            int cj[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) cj[u] = (k0 + u < k1) ? a.gr.col[k0 + u] : 0;
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (k0 + u < k1) {
#pragma unroll
                    for (int f = 0; f < DOUT; ++f) acc[f] += Xin[cj[u] * DOUT + f];
                }
            for (int k = k0 + 4; k < k1; ++k) {
                const int j = a.gr.col[k];
#pragma unroll
                for (int f = 0; f < DOUT; ++f) acc[f] += Xin[j * DOUT + f];
            }
            const float inv = (k1 > k0) ? 1.f / (float)(k1 - k0) : 0.f;
#pragma unroll
            for (int f = 0; f < DOUT; ++f) acc[f] *= inv;
            if (i >= a.self_off) {
#pragma unroll
                for (int f = 0; f < DOUT; ++f) acc[f] += Xs[(i - a.self_off) * DOUT + f];
            }
        }
#pragma unroll
        for (int f = 0; f < DOUT; ++f) O[i * DOUT + f] = acc[f];
    }
    __syncthreads();
}

template <int DIN, int DOUT, bool ATTN>
__global__ void __launch_bounds__(1024) go_layer_fwd_kernel(GoLayerArgs a) {
    IGCN_PDL_SYNC();
    extern __shared__ float smf[];
    const int Min = a.gr.Min, Mrow = a.gr.Mrow;
    const int tid = threadIdx.x, nt = blockDim.x;
    float* Wa_s = smf;                    // DOUT*DIN
    float* Ws_s = Wa_s + DOUT * DIN;      // DOUT*DIN
    float* u_s = Ws_s + DOUT * DIN;       // 2*DOUT
    float* v_s = u_s + 2 * DOUT;          // DOUT
    float* red = v_s + DOUT;              // 32*DOUT (one slot per warp, up to 32 warps)
    float* Xin = red + 32 * DOUT;         // Min*DOUT
    float* Xs = Xin + Min * DOUT;         // Min*DOUT
    float* O = Xs + Min * DOUT;           // Mrow*DOUT
    for (int i = tid; i < DOUT * DIN; i += nt) {
        Wa_s[i] = a.Wa[i];
        Ws_s[i] = a.Ws[i];
    }
    if (ATTN) {
        for (int i = tid; i < 2 * DOUT; i += nt) u_s[i] = a.u[i];
        for (int i = tid; i < DOUT; i += nt) v_s[i] = a.v[i];
    }
    __syncthreads();
    const int Mkeep = Mrow - a.keep_from;
    for (int b = blockIdx.x; b < a.B; b += gridDim.x) {
        go_layer_pre<DIN, DOUT, ATTN, false>(a, b, Wa_s, Ws_s, u_s, v_s, Xin, Xs, O, nullptr, nullptr);
        // LayerNorm over nodes, two pass
        float part[DOUT], mean[DOUT], var[DOUT];
#pragma unroll
        for (int f = 0; f < DOUT; ++f) part[f] = 0.f;
        for (int i = tid; i < Mrow; i += nt) {
#pragma unroll
            for (int f = 0; f < DOUT; ++f) part[f] += O[i * DOUT + f];
        }
        block_sum<DOUT>(part, red, mean);
#pragma unroll
        for (int f = 0; f < DOUT; ++f) {
            mean[f] /= (float)Mrow;
            part[f] = 0.f;
        }
        for (int i = tid; i < Mrow; i += nt) {
#pragma unroll
            for (int f = 0; f < DOUT; ++f) {
                const float d = O[i * DOUT + f] - mean[f];
                part[f] = fmaf(d, d, part[f]);
            }
        }
        block_sum<DOUT>(part, red, var);
        float rstd[DOUT];
#pragma unroll
        for (int f = 0; f < DOUT; ++f) rstd[f] = rsqrtf(var[f] / (float)Mrow + 1e-5f);
        if (tid == 0 && a.stats) {
#pragma unroll
            for (int f = 0; f < DOUT; ++f) {
                a.stats[(int64_t)b * 2 * DOUT + f] = mean[f];
                a.stats[(int64_t)b * 2 * DOUT + DOUT + f] = rstd[f];
            }
        }
        float* yb = a.y + (int64_t)b * Mkeep * DOUT;
        for (int i = a.keep_from + tid; i < Mrow; i += nt) {
            const float ga = a.gamma[i], be = a.beta[i];
            const float ms = a.mask ? a.mask[(int64_t)b * Mrow + i] : 1.f;
#pragma unroll
            for (int f = 0; f < DOUT; ++f) {
                const float yh = fmaf((O[i * DOUT + f] - mean[f]) * rstd[f], ga, be);
                yb[(i - a.keep_from) * DOUT + f] = fmaxf(yh, 0.f) * ms;
            }
        }
        __syncthreads();
    }
}

// parameter-gradient layout of one layer: [dWa (DOUT*DIN) | dWs (DOUT*DIN) | du (2*DOUT) | dv (DOUT) | dgamma (Mrow) | dbeta (Mrow)]
template <int DIN, int DOUT, bool ATTN>
__global__ void __launch_bounds__(512) go_layer_bwd_kernel(GoLayerArgs a) {
    IGCN_PDL_SYNC();
    extern __shared__ float smf[];
    const int Min = a.gr.Min, Mrow = a.gr.Mrow, nnz = a.gr.nnz;
    const int tid = threadIdx.x, nt = blockDim.x;
    constexpr int NW = 2 * DOUT * DIN + 3 * DOUT;
    float* Wa_s = smf;
    float* Ws_s = Wa_s + DOUT * DIN;
    float* u_s = Ws_s + DOUT * DIN;
    float* v_s = u_s + 2 * DOUT;
    float* red = v_s + DOUT;              // 16*NW  (one slot per warp, up to 16 warps; also used for the DOUT-wide reductions)
    float* Xin = red + 16 * NW;           // Min*DOUT   -> becomes dXin
    float* Xs = Xin + Min * DOUT;         // Min*DOUT   -> becomes dXs
    float* O = Xs + Min * DOUT;           // Mrow*DOUT  -> becomes dO
    float* dgam = O + Mrow * DOUT;        // Mrow  (accumulates across the CTA's subjects)
    float* dbet = dgam + Mrow;            // Mrow
    float* e_alpha = dbet + Mrow;         // nnz   (ATTN)
    float* e_dq = e_alpha + (ATTN ? nnz : 0);  // nnz (ATTN): tanh, then dq_e
    float* rowsum = e_dq + (ATTN ? nnz : 0);   // Min : per input node, sum_e dq_e of ITS row (ATTN) -- needed as u_row term
    for (int i = tid; i < DOUT * DIN; i += nt) {
        Wa_s[i] = a.Wa[i];
        Ws_s[i] = a.Ws[i];
    }
    if (ATTN) {
        for (int i = tid; i < 2 * DOUT; i += nt) u_s[i] = a.u[i];
        for (int i = tid; i < DOUT; i += nt) v_s[i] = a.v[i];
    }
    for (int i = tid; i < Mrow; i += nt) {
        dgam[i] = 0.f;
        dbet[i] = 0.f;
    }
    float gW[NW];  // per-thread partial of [dWa | dWs | du | dv]
#pragma unroll
    for (int i = 0; i < NW; ++i) gW[i] = 0.f;
    __syncthreads();
    const int Mkeep = Mrow - a.keep_from;
    for (int b = blockIdx.x; b < a.B; b += gridDim.x) {
        go_layer_pre<DIN, DOUT, ATTN, true>(a, b, Wa_s, Ws_s, u_s, v_s, Xin, Xs, O, e_alpha, e_dq);
        float mean[DOUT], rstd[DOUT];
#pragma unroll
        for (int f = 0; f < DOUT; ++f) {
            mean[f] = a.stats[(int64_t)b * 2 * DOUT + f];
            rstd[f] = a.stats[(int64_t)b * 2 * DOUT + DOUT + f];
        }
        // dY (through dropout scale and ReLU), LN reductions
        const float* gyb = a.gy + (int64_t)b * Mkeep * DOUT;
        float s1[DOUT], s2[DOUT], r1[DOUT], r2[DOUT];
#pragma unroll
        for (int f = 0; f < DOUT; ++f) s1[f] = s2[f] = 0.f;
        for (int i = a.keep_from + tid; i < Mrow; i += nt) {
            const float ga = a.gamma[i], be = a.beta[i];
            const float ms = a.mask ? a.mask[(int64_t)b * Mrow + i] : 1.f;
            float dg = 0.f, db = 0.f;
#pragma unroll
            for (int f = 0; f < DOUT; ++f) {
                const float xh = (O[i * DOUT + f] - mean[f]) * rstd[f];
                const float yh = fmaf(xh, ga, be);
                const float dy = (yh > 0.f) ? gyb[(i - a.keep_from) * DOUT + f] * ms : 0.f;
                dg = fmaf(dy, xh, dg);
                db += dy;
                s1[f] = fmaf(dy, ga, s1[f]);
                s2[f] = fmaf(dy * ga, xh, s2[f]);
            }
            dgam[i] += dg;
            dbet[i] += db;
        }
        block_sum<DOUT>(s1, red, r1);
        block_sum<DOUT>(s2, red, r2);
        const float invM = 1.f / (float)Mrow;
        // dO (in place over O), then the self-term and attention-row backward
        for (int i = tid; i < Mrow; i += nt) {
            const float ga = a.gamma[i], be = a.beta[i];
            const float ms = a.mask ? a.mask[(int64_t)b * Mrow + i] : 1.f;
            float dO[DOUT];
#pragma unroll
            for (int f = 0; f < DOUT; ++f) {
                const float xh = (O[i * DOUT + f] - mean[f]) * rstd[f];
                float dy = 0.f;
                if (i >= a.keep_from) {
                    const float yh = fmaf(xh, ga, be);
                    dy = (yh > 0.f) ? gyb[(i - a.keep_from) * DOUT + f] * ms : 0.f;
                }
                dO[f] = rstd[f] * (ga * dy - r1[f] * invM - xh * r2[f] * invM);
            }
#pragma unroll
            for (int f = 0; f < DOUT; ++f) O[i * DOUT + f] = dO[f];
        }
        __syncthreads();
        // row pass: per-slot dq_e (ATTN), gradient of the self term -> dXs (in place over Xs)
        if (ATTN) {
            for (int i = tid; i < Mrow; i += nt) {
                float dO[DOUT], xs[DOUT];
#pragma unroll
                for (int f = 0; f < DOUT; ++f) {
                    dO[f] = O[i * DOUT + f];
                    xs[f] = Xs[i * DOUT + f];
                }
                float gz = 0.f, dot = 0.f;
#pragma unroll
                for (int f = 0; f < DOUT; ++f) {
                    gz = fmaf(v_s[f], xs[f], gz);
                    dot = fmaf(dO[f], xs[f], dot);
                }
                const float gate = sigmoidf_(gz);
                const float dgz = dot * gate * (1.f - gate);
#pragma unroll
                for (int f = 0; f < DOUT; ++f) {
                    Xs[i * DOUT + f] = fmaf(dgz, v_s[f], dO[f] * gate);           // dXs_i
                    gW[2 * DOUT * DIN + 2 * DOUT + f] = fmaf(dgz, xs[f], gW[2 * DOUT * DIN + 2 * DOUT + f]);  // dv
                }
                const int k0 = a.gr.rowptr[i], k1 = a.gr.rowptr[i + 1];
                float t = 0.f;
                for (int k = k0; k < k1; ++k) {
                    const int j = a.gr.col[k];
                    float da = 0.f;
#pragma unroll
                    for (int f = 0; f < DOUT; ++f) da = fmaf(dO[f], Xin[j * DOUT + f], da);
                    t = fmaf(e_alpha[k], da, t);
                }
                float sdq = 0.f;
                for (int k = k0; k < k1; ++k) {
                    const int j = a.gr.col[k];
                    float da = 0.f;
#pragma unroll
                    for (int f = 0; f < DOUT; ++f) da = fmaf(dO[f], Xin[j * DOUT + f], da);
                    const float th = e_dq[k];
                    const float dq = e_alpha[k] * (da - t) * (1.f - th * th);
                    e_dq[k] = dq;
                    sdq += dq;
#pragma unroll
                    for (int f = 0; f < DOUT; ++f) {
                        gW[2 * DOUT * DIN + f] = fmaf(dq, Xin[i * DOUT + f], gW[2 * DOUT * DIN + f]);                // du_row
                        gW[2 * DOUT * DIN + DOUT + f] = fmaf(dq, Xin[j * DOUT + f], gW[2 * DOUT * DIN + DOUT + f]);  // du_col
                    }
                }
                rowsum[i] = sdq;
            }
        } else {
            // uniform decoder: dXs_{i-self_off} = dO_i ; rows < self_off have no self term
            __syncthreads();
            for (int j = tid; j < Min; j += nt) {
                const int i = j + a.self_off;
#pragma unroll
                for (int f = 0; f < DOUT; ++f) Xs[j * DOUT + f] = (i < Mrow) ? O[i * DOUT + f] : 0.f;
            }
        }
        __syncthreads();
        // column pass: dXin_j (gather over the CSC), then input/weight gradients
        float* dxb = a.dx + (int64_t)b * Min * DIN;
        const float* xb = a.x + (int64_t)b * Min * DIN;
        for (int j = tid; j < Min; j += nt) {
            float dxin[DOUT];
#pragma unroll
            for (int f = 0; f < DOUT; ++f) dxin[f] = 0.f;
            for (int q = a.gr.colptr[j]; q < a.gr.colptr[j + 1]; ++q) {
                const int i = a.gr.crow[q], k = a.gr.cpos[q];
                if (ATTN) {
                    const float al = e_alpha[k], dq = e_dq[k];
#pragma unroll
                    for (int f = 0; f < DOUT; ++f) dxin[f] += al * O[i * DOUT + f] + dq * u_s[DOUT + f];
                } else {
                    const float al = 1.f / (float)(a.gr.rowptr[i + 1] - a.gr.rowptr[i]);
#pragma unroll
                    for (int f = 0; f < DOUT; ++f) dxin[f] = fmaf(al, O[i * DOUT + f], dxin[f]);
                }
            }
            if (ATTN) {
                const float sdq = rowsum[j];   // square layer: node j is also row j
#pragma unroll
                for (int f = 0; f < DOUT; ++f) dxin[f] = fmaf(sdq, u_s[f], dxin[f]);
            }
            float xv[DIN], dxs[DOUT], dxv[DIN];
#pragma unroll
            for (int k = 0; k < DIN; ++k) {
                xv[k] = xb[j * DIN + k];
                dxv[k] = 0.f;
            }
#pragma unroll
            for (int f = 0; f < DOUT; ++f) dxs[f] = Xs[j * DOUT + f];
#pragma unroll
            for (int f = 0; f < DOUT; ++f) {
#pragma unroll
                for (int k = 0; k < DIN; ++k) {
                    gW[f * DIN + k] = fmaf(dxin[f], xv[k], gW[f * DIN + k]);
                    gW[DOUT * DIN + f * DIN + k] = fmaf(dxs[f], xv[k], gW[DOUT * DIN + f * DIN + k]);
                    dxv[k] += dxin[f] * Wa_s[f * DIN + k] + dxs[f] * Ws_s[f * DIN + k];
                }
            }
#pragma unroll
            for (int k = 0; k < DIN; ++k) dxb[j * DIN + k] = dxv[k];
        }
        __syncthreads();
    }
    // per-CTA partial row
    float tot[NW];
    block_sum<NW>(gW, red, tot);
    float* prow = a.partials + (int64_t)blockIdx.x * a.P;
    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < NW; ++i) prow[i] = tot[i];
    }
    for (int i = tid; i < Mrow; i += nt) {
        prow[NW + i] = dgam[i];
        prow[NW + Mrow + i] = dbet[i];
    }
}

}  // namespace igcn

#include "go_small.cuh"

namespace igcn {

static size_t go_fwd_smem(int din, int dout, int Min, int Mrow) {
    return 4 * ((size_t)2 * dout * din + 3 * dout + 32 * dout + 2 * (size_t)Min * dout + (size_t)Mrow * dout);
}
static size_t go_bwd_smem(int din, int dout, int Min, int Mrow, int nnz, bool attn) {
    const int NW = 2 * dout * din + 3 * dout;
    return 4 * ((size_t)2 * dout * din + 3 * dout + 16 * NW + 2 * (size_t)Min * dout + (size_t)Mrow * dout + 2 * (size_t)Mrow +
                (attn ? 2 * (size_t)nnz : 0) + (size_t)Min);
}

// threads per CTA: the row / column phases give one node to a thread, so a CTA wider than the node count only idles
// (large hierarchies: ncu at config-3 size showed 12 % of the warp slots and 13 cycles per warp instruction with 256 threads and
// 8 nodes per thread -- profiles/r2_ncu_go_kernels_c3.json -- so the forward takes up to 1024 threads, the backward, which holds
// the layer's parameter gradients in registers, up to 512)
static int go_threads(int m_in, int m_row, int cap = 512) {
    const int m = m_in > m_row ? m_in : m_row;
    int nt = (m + 31) / 32 * 32;
    if (nt < 64) nt = 64;
    if (nt > cap) nt = cap;
    return nt;
}

static int go_ctas(size_t smem, int64_t B, int nthreads = 256) {
    int per_sm = (int)((227 * 1024) / (smem + 1024));
    if (per_sm < 1) per_sm = 1;
    const int by_threads = 2048 / nthreads;
    if (per_sm > by_threads) per_sm = by_threads;
    if (per_sm > 16) per_sm = 16;
    int64_t n = (int64_t)sm_count() * per_sm;
    if (n > B) n = B;
    if (n < 1) n = 1;
    return (int)n;
}

template <int DIN, int DOUT, bool ATTN>
static int launch_go_fwd(const GoLayerArgs& a, cudaStream_t st) {
    const gosm::Plan sp = gosm::plan(DIN, DOUT, a.gr.Min, a.gr.Mrow, a.gr.nnz, a.keep_from, ATTN, false, a.B);
    if (sp.ok) {                              // small hierarchy: several subjects per CTA (go_small.cuh)
        auto ks = gosm::go_small_fwd_kernel<DIN, DOUT, ATTN>;
        int rc = allow_smem(ks, sp.smem, "go_small_fwd");
        if (rc) return rc;
        igcn::launch_k(ks, dim3(sp.n_cta), dim3(gosm::kThreads), sp.smem, st, a, sp.SUB, sp.Mp);
        IGCN_CHECK_LAUNCH("go_small_fwd");
        return IGCN_OK;
    }
    size_t smem = go_fwd_smem(DIN, DOUT, a.gr.Min, a.gr.Mrow);
    auto k = go_layer_fwd_kernel<DIN, DOUT, ATTN>;
    int rc = allow_smem(k, smem, "go_layer_fwd");
    if (rc) return rc;
    const int nthr = go_threads(a.gr.Min, a.gr.Mrow, 1024);
    igcn::launch_k(k, dim3(go_ctas(smem, a.B, nthr)), dim3(nthr), smem, st, a);
    IGCN_CHECK_LAUNCH("go_layer_fwd");
    return IGCN_OK;
}

template <int DIN, int DOUT, bool ATTN>
static int launch_go_bwd(const GoLayerArgs& a, int n_cta, float* grads, cudaStream_t st) {
    const gosm::Plan sp = gosm::plan(DIN, DOUT, a.gr.Min, a.gr.Mrow, a.gr.nnz, a.keep_from, ATTN, true, a.B);
    if (sp.ok) {
        auto ks = gosm::go_small_bwd_kernel<DIN, DOUT, ATTN>;
        int rc = allow_smem(ks, sp.smem, "go_small_bwd");
        if (rc) return rc;
        igcn::launch_k(ks, dim3(n_cta), dim3(gosm::kThreads), sp.smem, st, a, sp.SUB, sp.Mp);
        IGCN_CHECK_LAUNCH("go_small_bwd");
        igcn::launch_k(reduce_partials_kernel, dim3((a.P + 31) / 32), dim3(reduce_threads(n_cta)), 0, st, a.partials, n_cta, a.P, grads);
        IGCN_CHECK_LAUNCH("go_reduce_partials");
        return IGCN_OK;
    }
    size_t smem = go_bwd_smem(DIN, DOUT, a.gr.Min, a.gr.Mrow, a.gr.nnz, ATTN);
    auto k = go_layer_bwd_kernel<DIN, DOUT, ATTN>;
    int rc = allow_smem(k, smem, "go_layer_bwd");
    if (rc) return rc;
    igcn::launch_k(k, dim3(n_cta), dim3(go_threads(a.gr.Min, a.gr.Mrow)), smem, st, a);
    IGCN_CHECK_LAUNCH("go_layer_bwd");
    igcn::launch_k(reduce_partials_kernel, dim3((a.P + 31) / 32), dim3(reduce_threads(n_cta)), 0, st, a.partials, n_cta, a.P, grads);
    IGCN_CHECK_LAUNCH("go_reduce_partials");
    return IGCN_OK;
}

}  // namespace igcn

using namespace igcn;

extern "C" int igcn_go_spmm_fwd(const float* in, const int32_t* rowptr, const int32_t* col, const float* vals, int64_t B,
                                int64_t n_in, int64_t n_row, int64_t nnz, int64_t channels, float* out, void* stream) {
    IGCN_REQUIRE(B >= 0 && n_in > 0 && n_row > 0 && nnz >= 0, IGCN_ERR_BAD_ARG, "go_spmm_fwd: bad size");
    IGCN_REQUIRE(channels == 1 || channels == 2, IGCN_ERR_UNSUPPORTED, "go_spmm_fwd: channels=%lld (1 or 2 supported)", (long long)channels);
    if (B == 0) return IGCN_OK;
    IGCN_REQUIRE(in && rowptr && out && (nnz == 0 || (col && vals)), IGCN_ERR_BAD_ARG, "go_spmm_fwd: null pointer");
    size_t smem = 4 * (size_t)n_in;
    cudaStream_t st = (cudaStream_t)stream;
    int grid = (int)(B < (int64_t)sm_count() * 4 ? B : (int64_t)sm_count() * 4);
    int rc;
    if (channels == 1) {
        if ((rc = allow_smem(go_spmm_fwd_kernel<1>, smem, "go_spmm_fwd"))) return rc;
        igcn::launch_k(go_spmm_fwd_kernel<1>, dim3(grid), dim3(256), smem, st, in, rowptr, col, vals, (int)B, (int)n_in, (int)n_row, (int)nnz, out);
    } else {
        if ((rc = allow_smem(go_spmm_fwd_kernel<2>, smem, "go_spmm_fwd"))) return rc;
        igcn::launch_k(go_spmm_fwd_kernel<2>, dim3(grid), dim3(256), smem, st, in, rowptr, col, vals, (int)B, (int)n_in, (int)n_row, (int)nnz, out);
    }
    IGCN_CHECK_LAUNCH("go_spmm_fwd");
    return IGCN_OK;
}

extern "C" int igcn_go_spmm_bwd(const float* g_out, const float* in, const int32_t* row_of, const int32_t* col,
                                const int32_t* colptr, const int32_t* crow, const int32_t* cpos, const float* vals, int64_t B,
                                int64_t n_in, int64_t n_row, int64_t nnz, int64_t channels, float* d_in, float* d_vals,
                                float* workspace, void* stream) {
    IGCN_REQUIRE(B >= 0 && n_in > 0 && n_row > 0 && nnz >= 0, IGCN_ERR_BAD_ARG, "go_spmm_bwd: bad size");
    IGCN_REQUIRE(channels == 1 || channels == 2, IGCN_ERR_UNSUPPORTED, "go_spmm_bwd: channels=%lld (1 or 2 supported)", (long long)channels);
    IGCN_REQUIRE(d_vals && (B == 0 || (g_out && in && colptr)), IGCN_ERR_BAD_ARG, "go_spmm_bwd: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if (B == 0) {
        cudaMemsetAsync(d_vals, 0, sizeof(float) * channels * nnz, st);
        return IGCN_OK;
    }
    size_t smem = 4 * (size_t)n_row * channels;
    int grid = (int)(B < (int64_t)sm_count() * 4 ? B : (int64_t)sm_count() * 4);
    int rc;
    if (channels == 1) {
        if (d_in) {
            if ((rc = allow_smem(go_spmm_bwd_in_kernel<1>, smem, "go_spmm_bwd"))) return rc;
            igcn::launch_k(go_spmm_bwd_in_kernel<1>, dim3(grid), dim3(256), smem, st, g_out, colptr, crow, cpos, vals, (int)B, (int)n_in, (int)n_row, (int)nnz, d_in);
        }
        if (nnz && workspace) {
            float* gT = workspace;
            float* inT = workspace + (size_t)n_row * B;
            igcn::launch_k(transpose2_kernel, dim3(dim3(sm_count() * 4, 2)), dim3(256), 0, st, g_out, (int)B, (int)n_row, gT, in, (int)B, (int)n_in, inT);
            igcn::launch_k(go_spmm_bwd_vals_t_kernel<1>, dim3(sm_count() * 8), dim3(256), 0, st, gT, inT, row_of, col, (int)B, (int)nnz, d_vals);
        } else if (nnz)
            igcn::launch_k(go_spmm_bwd_vals_kernel<1>, dim3((int)((nnz + 7) / 8)), dim3(256), 0, st, g_out, in, row_of, col, (int)B, (int)n_in, (int)n_row, (int)nnz, d_vals);
    } else {
        if (d_in) {
            if ((rc = allow_smem(go_spmm_bwd_in_kernel<2>, smem, "go_spmm_bwd"))) return rc;
            igcn::launch_k(go_spmm_bwd_in_kernel<2>, dim3(grid), dim3(256), smem, st, g_out, colptr, crow, cpos, vals, (int)B, (int)n_in, (int)n_row, (int)nnz, d_in);
        }
        if (nnz && workspace) {
            float* gT = workspace;
            float* inT = workspace + (size_t)n_row * 2 * B;
            igcn::launch_k(transpose2_kernel, dim3(dim3(sm_count() * 4, 2)), dim3(256), 0, st, g_out, (int)B, (int)n_row * 2, gT, in, (int)B, (int)n_in, inT);
            igcn::launch_k(go_spmm_bwd_vals_t_kernel<2>, dim3(sm_count() * 8), dim3(256), 0, st, gT, inT, row_of, col, (int)B, (int)nnz, d_vals);
        } else if (nnz)
            igcn::launch_k(go_spmm_bwd_vals_kernel<2>, dim3((int)((nnz + 7) / 8)), dim3(256), 0, st, g_out, in, row_of, col, (int)B, (int)n_in, (int)n_row, (int)nnz, d_vals);
    }
    IGCN_CHECK_LAUNCH("go_spmm_bwd");
    return IGCN_OK;
}

static int fill_go_args(GoLayerArgs& a, const char* who, const float* x, const float* Wa, const float* Ws, const float* u,
                        const float* v, const float* gamma, const float* beta, const float* mask, const int32_t* rowptr,
                        const int32_t* col, const int32_t* colptr, const int32_t* crow, const int32_t* cpos, int64_t B,
                        int64_t m_in, int64_t m_row, int64_t nnz, int64_t attn, int64_t self_off, int64_t keep_from) {
    IGCN_REQUIRE(B >= 0 && m_in > 0 && m_row > 0 && nnz >= 0 && self_off >= 0 && keep_from >= 0 && keep_from < m_row,
                 IGCN_ERR_BAD_ARG, "%s: bad size", who);
    IGCN_REQUIRE(x && Wa && Ws && gamma && beta && rowptr && colptr, IGCN_ERR_BAD_ARG, "%s: null pointer", who);
    IGCN_REQUIRE(!attn || (u && v && m_in == m_row && self_off == 0), IGCN_ERR_BAD_ARG,
                 "%s: attention layers are square with u,v given", who);
    IGCN_REQUIRE(attn || (m_row - self_off == m_in), IGCN_ERR_BAD_ARG, "%s: decoder needs m_row - self_off == m_in", who);
    a.x = x; a.Wa = Wa; a.Ws = Ws; a.u = u; a.v = v; a.gamma = gamma; a.beta = beta; a.mask = mask;
    a.gr.rowptr = rowptr; a.gr.col = col; a.gr.colptr = colptr; a.gr.crow = crow; a.gr.cpos = cpos;
    a.gr.Mrow = (int)m_row; a.gr.Min = (int)m_in; a.gr.nnz = (int)nnz;
    a.B = (int)B; a.self_off = (int)self_off; a.keep_from = (int)keep_from;
    return IGCN_OK;
}

#define GO_DISPATCH(DIN_, DOUT_, ATT_, CALL)                                         \
    if (din == DIN_ && dout == DOUT_ && (bool)attn == ATT_) { constexpr int DI = DIN_, DO = DOUT_; constexpr bool AT = ATT_; CALL; }

extern "C" int64_t igcn_go_layer_param_count(int64_t din, int64_t dout, int64_t m_row) {
    return 2 * dout * din + 3 * dout + 2 * m_row;
}

extern "C" int64_t igcn_go_layer_bwd_ctas(int64_t B, int64_t din, int64_t dout, int64_t m_in, int64_t m_row, int64_t nnz, int64_t attn) {
    // keep_from only moves the gy staging buffer (smaller when > 0): planning with 0 gives the same SUB and CTA count
    const gosm::Plan sp = gosm::plan((int)din, (int)dout, (int)m_in, (int)m_row, (int)nnz, 0, attn != 0, true, B);
    if (sp.ok) return sp.n_cta;
    return go_ctas(go_bwd_smem((int)din, (int)dout, (int)m_in, (int)m_row, (int)nnz, attn != 0), B, go_threads((int)m_in, (int)m_row));
}

extern "C" int igcn_go_layer_fwd(const float* x, const float* Wa, const float* Ws, const float* u, const float* v,
                                 const float* gamma, const float* beta, const float* mask, const int32_t* rowptr,
                                 const int32_t* col, const int32_t* colptr, const int32_t* crow, const int32_t* cpos, int64_t B,
                                 int64_t m_in, int64_t m_row, int64_t nnz, int64_t din, int64_t dout, int64_t attn,
                                 int64_t self_off, int64_t keep_from, float* y, float* stats, void* stream) {
    GoLayerArgs a{};
    int rc = fill_go_args(a, "go_layer_fwd", x, Wa, Ws, u, v, gamma, beta, mask, rowptr, col, colptr, crow, cpos, B, m_in, m_row,
                          nnz, attn, self_off, keep_from);
    if (rc) return rc;
    IGCN_REQUIRE(y && stats, IGCN_ERR_BAD_ARG, "go_layer_fwd: null output");
    if (B == 0) return IGCN_OK;
    a.y = y; a.stats = stats;
    cudaStream_t st = (cudaStream_t)stream;
    GO_DISPATCH(2, 5, true, return (launch_go_fwd<DI, DO, AT>(a, st)))
    GO_DISPATCH(5, 5, true, return (launch_go_fwd<DI, DO, AT>(a, st)))
    GO_DISPATCH(5, 5, false, return (launch_go_fwd<DI, DO, AT>(a, st)))
    GO_DISPATCH(5, 2, false, return (launch_go_fwd<DI, DO, AT>(a, st)))
    set_error("go_layer_fwd: (din=%lld,dout=%lld,attn=%lld) not instantiated; supported: (2,5,attn) (5,5,attn) (5,5,uniform) (5,2,uniform)",
              (long long)din, (long long)dout, (long long)attn);
    return IGCN_ERR_UNSUPPORTED;
}

extern "C" int igcn_go_layer_bwd(const float* x, const float* Wa, const float* Ws, const float* u, const float* v,
                                 const float* gamma, const float* beta, const float* mask, const int32_t* rowptr,
                                 const int32_t* col, const int32_t* colptr, const int32_t* crow, const int32_t* cpos, int64_t B,
                                 int64_t m_in, int64_t m_row, int64_t nnz, int64_t din, int64_t dout, int64_t attn,
                                 int64_t self_off, int64_t keep_from, const float* stats, const float* g_y, float* dx,
                                 float* partials, int64_t n_cta, float* grads, void* stream) {
    GoLayerArgs a{};
    int rc = fill_go_args(a, "go_layer_bwd", x, Wa, Ws, u, v, gamma, beta, mask, rowptr, col, colptr, crow, cpos, B, m_in, m_row,
                          nnz, attn, self_off, keep_from);
    if (rc) return rc;
    IGCN_REQUIRE(stats && g_y && dx && partials && grads, IGCN_ERR_BAD_ARG, "go_layer_bwd: null pointer");
    a.P = (int)igcn_go_layer_param_count(din, dout, m_row);
    cudaStream_t st = (cudaStream_t)stream;
    if (B == 0) {
        cudaMemsetAsync(grads, 0, sizeof(float) * a.P, st);
        return IGCN_OK;
    }
    const int want = (int)igcn_go_layer_bwd_ctas(B, din, dout, m_in, m_row, nnz, attn);
    IGCN_REQUIRE(n_cta == want, IGCN_ERR_BAD_ARG, "go_layer_bwd: n_cta=%lld, expected %d", (long long)n_cta, want);
    a.stats = const_cast<float*>(stats); a.gy = g_y; a.dx = dx; a.partials = partials;
    GO_DISPATCH(2, 5, true, return (launch_go_bwd<DI, DO, AT>(a, want, grads, st)))
    GO_DISPATCH(5, 5, true, return (launch_go_bwd<DI, DO, AT>(a, want, grads, st)))
    GO_DISPATCH(5, 5, false, return (launch_go_bwd<DI, DO, AT>(a, want, grads, st)))
    GO_DISPATCH(5, 2, false, return (launch_go_bwd<DI, DO, AT>(a, want, grads, st)))
    set_error("go_layer_bwd: (din=%lld,dout=%lld,attn=%lld) not instantiated", (long long)din, (long long)dout, (long long)attn);
    return IGCN_ERR_UNSUPPORTED;
}
