// GCNConv for ARBITRARY graphs (any COO edge_index, any node count): the operator-level seam of the path.
//
// The SGCN models run their GCNConv stack through the fused per-graph encoder kernels (sgcn_encoder.cu, sgcn_mma.cuh), which need
// equally sized graphs that fit one CTA.  The reference's model files, however, call the PyG operator itself --
// `GCNConv(in, out)(x, edge_index, edge_weight)` (kernel/sgcn_img_snp.py:218-221, kernel/sgcn.py:281-284) -- on whatever graph
// they hold, and differentiate through edge_weight in the explain pass.  These kernels implement exactly that operator
// (PyG 2.0.2: add_remaining_self_loops with fill 1, symmetric normalisation, X W^T, scatter-add over targets, + bias) for one big
// graph in global memory: CSR by target built on the device (counting sort, made stable per row), warp-per-row SpMM with the
// feature dimension across the lanes, transposed SpMM over the CSC for the backward, gradients with respect to x, edge_weight,
// weight and bias.  Deterministic: integer atomics only (the sort), fixed summation orders, per-CTA partials for dW.
#include "common.cuh"

namespace igcn {
namespace gen {

// ---- CSR / CSC of an arbitrary COO graph ---------------------------------------------------------------------------------
__global__ void hist_kernel(const int64_t* __restrict__ ei, int64_t E, int64_t N, int* __restrict__ cnt_t, int* __restrict__ cnt_s, int* __restrict__ bad) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < E; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t s = ei[e], t = ei[E + e];
        if (s < 0 || s >= N || t < 0 || t >= N) {
            atomicExch(bad, 1);
            continue;
        }
        atomicAdd(&cnt_t[t], 1);
        atomicAdd(&cnt_s[s], 1);
    }
}
// exclusive scans of two count arrays (n entries each) -> ptr arrays (n + 1); one block, chunked
__global__ void __launch_bounds__(1024) scan2_kernel(const int* __restrict__ c0, const int* __restrict__ c1, int n, int* __restrict__ p0,
                                                     int* __restrict__ p1) {
    __shared__ int wsum[32];
    __shared__ int carry_s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int which = 0; which < 2; ++which) {
        const int* c = which ? c1 : c0;
        int* p = which ? p1 : p0;
        if (tid == 0) carry_s = 0;
        __syncthreads();
        for (int base = 0; base < n; base += 1024) {
            const int i = base + tid;
            const int v = i < n ? c[i] : 0;
            int inc = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += t;
            }
            if (lane == 31) wsum[warp] = inc;
            __syncthreads();
            if (warp == 0) {
                int w = wsum[lane];
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int t = __shfl_up_sync(0xffffffffu, w, o);
                    if (lane >= o) w += t;
                }
                wsum[lane] = w;
            }
            __syncthreads();
            const int carry = carry_s;
            const int before = carry + (warp ? wsum[warp - 1] : 0) + inc - v;
            if (i < n) p[i] = before;
            __syncthreads();
            if (tid == 1023) carry_s = carry + wsum[31];
            __syncthreads();
        }
        if (tid == 0) p[n] = carry_s;
        __syncthreads();
    }
}
__global__ void fill_kernel(const int64_t* __restrict__ ei, int64_t E, int64_t N, const int* __restrict__ rp_t, const int* __restrict__ rp_s,
                            int* __restrict__ cur_t, int* __restrict__ cur_s, int* __restrict__ ids_t, int* __restrict__ ids_s) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < E; e += (int64_t)gridDim.x * blockDim.x) {
        if (ei[e] < 0 || ei[e] >= N || ei[E + e] < 0 || ei[E + e] >= N) continue;      // flagged by hist_kernel; the host raises
        const int s = (int)ei[e], t = (int)ei[E + e];
        ids_t[rp_t[t] + atomicAdd(&cur_t[t], 1)] = (int)e;
        ids_s[rp_s[s] + atomicAdd(&cur_s[s], 1)] = (int)e;
    }
}
__device__ __forceinline__ void sort_ids(int* a, int n) {
    for (int i = 1; i < n; ++i) {
        const int v = a[i];
        int j = i - 1;
        while (j >= 0 && a[j] > v) {
            a[j + 1] = a[j];
            --j;
        }
        a[j + 1] = v;
    }
}
// ascending edge id inside every row == a stable sort by target / source; then the per-slot arrays
__global__ void rows_kernel(const int64_t* __restrict__ ei, int64_t E, int N, const int* __restrict__ rp_t, const int* __restrict__ rp_s,
                            int* __restrict__ ids_t, int* __restrict__ ids_s, int* __restrict__ csr_src, int* __restrict__ slot_of) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
        sort_ids(ids_t + rp_t[i], rp_t[i + 1] - rp_t[i]);
        sort_ids(ids_s + rp_s[i], rp_s[i + 1] - rp_s[i]);
        for (int k = rp_t[i]; k < rp_t[i + 1]; ++k) {
            const int e = ids_t[k];
            csr_src[k] = (int)ei[e];
            slot_of[e] = k;
        }
    }
}
__global__ void cscpos_kernel(int64_t E, const int* __restrict__ ids_s, const int* __restrict__ slot_of, int* __restrict__ csc_pos) {
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < E; q += (int64_t)gridDim.x * blockDim.x) csc_pos[q] = slot_of[ids_s[q]];
}

// ---- forward ----------------------------------------------------------------------------------------------------------------
// per node: self-loop merge (existing loop keeps its weight, last one wins, else 1), degree, d^-1/2, n_ii; per slot: raw weight, target
__global__ void norm1_kernel(int N, const int* __restrict__ rp, const int* __restrict__ src, const int* __restrict__ perm,
                             const float* __restrict__ ew, float* __restrict__ wslot, int* __restrict__ etgt, float* __restrict__ dinv,
                             float* __restrict__ nii, float* __restrict__ ell) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
        float deg = 0.f, loopw = 1.f;
        for (int k = rp[i]; k < rp[i + 1]; ++k) {
            const float w = ew ? ew[perm[k]] : 1.f;
            wslot[k] = w;
            etgt[k] = i;
            if (src[k] == i)
                loopw = w;
            else
                deg += w;
        }
        deg += loopw;
        const float d = deg == 0.f ? 0.f : rsqrtf(deg);
        dinv[i] = d;
        nii[i] = d * d * loopw;
        ell[i] = loopw;
    }
}
__global__ void norm2_kernel(int64_t E, const int* __restrict__ src, const int* __restrict__ etgt, const float* __restrict__ wslot,
                             const float* __restrict__ dinv, float* __restrict__ norm) {
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < E; k += (int64_t)gridDim.x * blockDim.x) {
        const int s = src[k], t = etgt[k];
        norm[k] = s == t ? 0.f : dinv[s] * wslot[k] * dinv[t];
    }
}
// y[i][o] = sum_c a[i][c] * Wm[o][c] (trans = 0: Wm is (O, C) row-major)  or  sum_c a[i][c] * Wm[c][o] (trans = 1: Wm is (C, O))
__global__ void __launch_bounds__(256) dense_rows_kernel(const float* __restrict__ a, const float* __restrict__ Wm, int N, int C, int O, int trans,
                                                         float* __restrict__ y) {
    extern __shared__ float wsm[];                           // (C, O) layout: wsm[c * O + o]
    for (int i = threadIdx.x; i < C * O; i += blockDim.x) {
        const int c = i / O, o = i - c * O;
        wsm[i] = trans ? Wm[i] : Wm[o * C + c];
    }
    __syncthreads();
    const int64_t total = (int64_t)N * O;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = idx / O;
        const int o = (int)(idx - i * O);
        const float* ar = a + i * C;
        float acc = 0.f;
        for (int c = 0; c < C; ++c) acc = fmaf(ar[c], wsm[c * O + o], acc);
        y[idx] = acc;
    }
}
// warp per row: out[i][:] = sum_k coef_k v[nbr_k][:] + nii[i] v[i][:] (+ bias); lanes walk the feature dimension
// gather = 0: in-edges (k = CSR slot, nbr = src[k], coef = norm[k]); gather = 1: out-edges (q over the CSC: slot = pos[q], nbr = etgt[slot])
__global__ void __launch_bounds__(256) spmm_rows_kernel(int N, int F, const int* __restrict__ rp, const int* __restrict__ nbr_or_pos,
                                                        const int* __restrict__ etgt, const float* __restrict__ norm, const float* __restrict__ nii,
                                                        const float* __restrict__ v, const float* __restrict__ bias, int transposed,
                                                        float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    for (int i = blockIdx.x * wpb + (threadIdx.x >> 5); i < N; i += gridDim.x * wpb) {
        const int k0 = rp[i], k1 = rp[i + 1];
        const float ns = nii[i];
        for (int f = lane; f < F; f += 32) {
            float acc = 0.f;
            for (int k = k0; k < k1; ++k) {
                const int slot = transposed ? nbr_or_pos[k] : k;
                const int j = transposed ? etgt[slot] : nbr_or_pos[k];
                acc = fmaf(norm[slot], v[(int64_t)j * F + f], acc);
            }
            acc = fmaf(ns, v[(int64_t)i * F + f], acc);
            out[(int64_t)i * F + f] = acc + (bias ? bias[f] : 0.f);
        }
    }
}

// ---- backward ---------------------------------------------------------------------------------------------------------------
// d norm_k = <g[t_k], u[s_k]>, d n_ii = <g[i], u[i]> : warp per target row
__global__ void __launch_bounds__(256) edge_dots_kernel(int N, int F, const int* __restrict__ rp, const int* __restrict__ src,
                                                        const float* __restrict__ g, const float* __restrict__ u, float* __restrict__ edn,
                                                        float* __restrict__ dnii) {
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    for (int i = blockIdx.x * wpb + (threadIdx.x >> 5); i < N; i += gridDim.x * wpb) {
        for (int k = rp[i]; k <= rp[i + 1]; ++k) {           // the extra iteration is the self term
            const int j = k < rp[i + 1] ? src[k] : i;
            float acc = 0.f;
            for (int f = lane; f < F; f += 32) acc = fmaf(g[(int64_t)i * F + f], u[(int64_t)j * F + f], acc);
            acc = warp_sum(acc);
            if (lane == 0) {
                if (k < rp[i + 1])
                    edn[k] = acc;
                else
                    dnii[i] = acc;
            }
        }
    }
}
__global__ void ddeg_kernel(int N, const int* __restrict__ rp, const int* __restrict__ src, const int* __restrict__ rps, const int* __restrict__ pos,
                            const int* __restrict__ etgt, const float* __restrict__ edn, const float* __restrict__ wslot,
                            const float* __restrict__ dinv, const float* __restrict__ ell, const float* __restrict__ dnii, float* __restrict__ ddeg) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
        float dd = 0.f;
        for (int k = rp[i]; k < rp[i + 1]; ++k) {
            const int s = src[k];
            if (s != i) dd = fmaf(edn[k] * wslot[k], dinv[s], dd);
        }
        for (int q = rps[i]; q < rps[i + 1]; ++q) {
            const int k = pos[q], t = etgt[k];
            if (t != i) dd = fmaf(edn[k] * wslot[k], dinv[t], dd);
        }
        const float di = dinv[i];
        dd = fmaf(2.f * di * ell[i], dnii[i], dd);
        ddeg[i] = -0.5f * di * di * di * dd;
    }
}
// d edge_weight in ORIGINAL edge order (every slot maps to one edge); every self-loop slot receives the loop gradient, as autograd's
// index_put backward does for duplicated loops
__global__ void dweight_kernel(int N, const int* __restrict__ rp, const int* __restrict__ src, const int* __restrict__ perm,
                               const float* __restrict__ edn, const float* __restrict__ dinv, const float* __restrict__ dnii,
                               const float* __restrict__ ddeg, float* __restrict__ d_ew) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
        const float di = dinv[i];
        for (int k = rp[i]; k < rp[i + 1]; ++k) {
            const int s = src[k];
            d_ew[perm[k]] = (s != i) ? dinv[s] * di * edn[k] + ddeg[i] : di * di * dnii[i] + ddeg[i];
        }
    }
}
// per-CTA partials of dW (O x C) = dU^T x and db (O) = column sums of g over the CTA's node chunk; fixed order
__global__ void __launch_bounds__(256) dweights_kernel(int N, int C, int O, const float* __restrict__ dU, const float* __restrict__ x,
                                                       const float* __restrict__ g, int chunk, float* __restrict__ partials) {
    const int i0 = blockIdx.x * chunk, i1 = min(N, i0 + chunk);
    float* prow = partials + (int64_t)blockIdx.x * (O * C + O);
    for (int p = threadIdx.x; p < O * C + O; p += blockDim.x) {
        float acc = 0.f;
        if (p < O * C) {
            const int o = p / C, c = p - o * C;
            for (int i = i0; i < i1; ++i) acc = fmaf(dU[(int64_t)i * O + o], x[(int64_t)i * C + c], acc);
        } else {
            const int o = p - O * C;
            for (int i = i0; i < i1; ++i) acc += g[(int64_t)i * O + o];
        }
        prow[p] = acc;
    }
}

static int blocks_for(int64_t n, int threads) {
    int64_t b = (n + threads - 1) / threads;
    const int64_t cap = (int64_t)sm_count() * 8;
    if (b > cap) b = cap;
    return (int)(b < 1 ? 1 : b);
}
constexpr int kChunk = 512;      // nodes per CTA in the weight-gradient partials

}  // namespace gen
}  // namespace igcn

using namespace igcn;
using namespace igcn::gen;

extern "C" int64_t igcn_graph_csr_work_ints(int64_t N, int64_t E) { return 2 * (N + 1) + 2 * E + 1; }

extern "C" int igcn_graph_csr(const int64_t* edge_index, int64_t N, int64_t E, int32_t* rowptr_t, int32_t* csr_src, int32_t* csr_perm,
                              int32_t* rowptr_s, int32_t* csc_pos, int32_t* work, void* stream) {
    IGCN_REQUIRE(N >= 0 && E >= 0, IGCN_ERR_BAD_ARG, "graph_csr: negative size");
    IGCN_REQUIRE(N < (1ll << 31) && E < (1ll << 31), IGCN_ERR_UNSUPPORTED, "graph_csr: more than 2^31 nodes or edges");
    IGCN_REQUIRE(rowptr_t && rowptr_s && work, IGCN_ERR_BAD_ARG, "graph_csr: null pointer");
    IGCN_REQUIRE(E == 0 || (edge_index && csr_src && csr_perm && csc_pos), IGCN_ERR_BAD_ARG, "graph_csr: null edge pointer");
    cudaStream_t st = (cudaStream_t)stream;
    int* cnt_t = work;                 // N + 1 (cursor afterwards)
    int* cnt_s = cnt_t + (N + 1);      // N + 1
    int* ids_s = cnt_s + (N + 1);      // E
    int* slot_of = ids_s + E;          // E
    int* bad = slot_of + E;            // 1 : set when an endpoint is outside [0, N); the caller reads work[2(N+1)+2E] afterwards
    cudaMemsetAsync(work, 0, sizeof(int) * (size_t)(2 * (N + 1) + 2 * E + 1), st);
    if (E) {
        hist_kernel<<<blocks_for(E, 256), 256, 0, st>>>(edge_index, E, N, cnt_t, cnt_s, bad);
        IGCN_CHECK_LAUNCH("graph_csr hist");
    }
    scan2_kernel<<<1, 1024, 0, st>>>(cnt_t, cnt_s, (int)N, rowptr_t, rowptr_s);
    IGCN_CHECK_LAUNCH("graph_csr scan");
    if (E) {
        cudaMemsetAsync(work, 0, sizeof(int) * (size_t)(2 * (N + 1)), st);
        fill_kernel<<<blocks_for(E, 256), 256, 0, st>>>(edge_index, E, N, rowptr_t, rowptr_s, cnt_t, cnt_s, csr_perm, ids_s);
        IGCN_CHECK_LAUNCH("graph_csr fill");
        rows_kernel<<<blocks_for(N, 128), 128, 0, st>>>(edge_index, E, (int)N, rowptr_t, rowptr_s, csr_perm, ids_s, csr_src, slot_of);
        IGCN_CHECK_LAUNCH("graph_csr rows");
        cscpos_kernel<<<blocks_for(E, 256), 256, 0, st>>>(E, ids_s, slot_of, csc_pos);
        IGCN_CHECK_LAUNCH("graph_csr cscpos");
    }
    return IGCN_OK;
}

/* floats of workspace the forward keeps for the backward: [u (N*O) | wslot (E) | norm (E) | dinv (N) | nii (N) | ell (N)] + etgt (E ints) */
extern "C" int64_t igcn_gcn_conv_saved_floats(int64_t N, int64_t E, int64_t O) { return N * O + 3 * E + 3 * N; }

extern "C" int igcn_gcn_conv_fwd(const float* x, const int32_t* rowptr_t, const int32_t* csr_src, const int32_t* csr_perm,
                                 const float* edge_weight, const float* weight, const float* bias, int64_t N, int64_t E, int64_t C, int64_t O,
                                 float* saved, float* out, void* stream) {
    IGCN_REQUIRE(N >= 0 && E >= 0 && C > 0 && O > 0, IGCN_ERR_BAD_ARG, "gcn_conv_fwd: bad size");
    IGCN_REQUIRE(C * O * 4 <= 160 * 1024, IGCN_ERR_UNSUPPORTED, "gcn_conv_fwd: in*out = %lld weights exceed the staged limit (40960)", (long long)(C * O));
    if (N == 0) return IGCN_OK;
    IGCN_REQUIRE(x && rowptr_t && weight && saved && out, IGCN_ERR_BAD_ARG, "gcn_conv_fwd: null pointer");
    IGCN_REQUIRE(E == 0 || (csr_src && csr_perm), IGCN_ERR_BAD_ARG, "gcn_conv_fwd: null CSR arrays");
    cudaStream_t st = (cudaStream_t)stream;
    float* u = saved;
    float* wslot = u + N * O;
    float* norm = wslot + E;
    float* dinv = norm + E;
    float* nii = dinv + N;
    float* ell = nii + N;
    int* etgt = reinterpret_cast<int*>(ell + N);
    norm1_kernel<<<blocks_for(N, 128), 128, 0, st>>>((int)N, rowptr_t, csr_src, csr_perm, edge_weight, wslot, etgt, dinv, nii, ell);
    IGCN_CHECK_LAUNCH("gcn_conv norm1");
    if (E) {
        norm2_kernel<<<blocks_for(E, 256), 256, 0, st>>>(E, csr_src, etgt, wslot, dinv, norm);
        IGCN_CHECK_LAUNCH("gcn_conv norm2");
    }
    const size_t wsm = sizeof(float) * (size_t)(C * O);
    int rc = allow_smem(dense_rows_kernel, wsm, "gcn_conv_fwd");
    if (rc) return rc;
    dense_rows_kernel<<<blocks_for(N * O, 256), 256, wsm, st>>>(x, weight, (int)N, (int)C, (int)O, 0, u);
    IGCN_CHECK_LAUNCH("gcn_conv xw");
    spmm_rows_kernel<<<blocks_for(N, 8), 256, 0, st>>>((int)N, (int)O, rowptr_t, csr_src, etgt, norm, nii, u, bias, 0, out);
    IGCN_CHECK_LAUNCH("gcn_conv spmm");
    return IGCN_OK;
}

extern "C" int64_t igcn_gcn_conv_bwd_ctas(int64_t N) { return N <= 0 ? 1 : (N + kChunk - 1) / kChunk; }
/* workspace floats of the backward: dU (N*O) | edn (E) | dnii (N) | ddeg (N) */
extern "C" int64_t igcn_gcn_conv_bwd_work_floats(int64_t N, int64_t E, int64_t O) { return N * O + E + 2 * N; }

extern "C" int igcn_gcn_conv_bwd(const float* x, const int32_t* rowptr_t, const int32_t* csr_src, const int32_t* csr_perm,
                                 const int32_t* rowptr_s, const int32_t* csc_pos, const float* weight, const float* saved, const float* g_out,
                                 int64_t N, int64_t E, int64_t C, int64_t O, float* work, float* dx, float* d_edge_weight, float* partials,
                                 int64_t n_cta, float* grads, void* stream) {
    IGCN_REQUIRE(N >= 0 && E >= 0 && C > 0 && O > 0, IGCN_ERR_BAD_ARG, "gcn_conv_bwd: bad size");
    IGCN_REQUIRE(C * O * 4 <= 160 * 1024, IGCN_ERR_UNSUPPORTED, "gcn_conv_bwd: in*out = %lld weights exceed the staged limit", (long long)(C * O));
    cudaStream_t st = (cudaStream_t)stream;
    const int P = (int)(O * C + O);
    IGCN_REQUIRE(grads, IGCN_ERR_BAD_ARG, "gcn_conv_bwd: null grads");
    if (N == 0) {
        cudaMemsetAsync(grads, 0, sizeof(float) * P, st);
        return IGCN_OK;
    }
    IGCN_REQUIRE(x && rowptr_t && rowptr_s && weight && saved && g_out && work && partials, IGCN_ERR_BAD_ARG, "gcn_conv_bwd: null pointer");
    IGCN_REQUIRE(E == 0 || (csr_src && csr_perm && csc_pos), IGCN_ERR_BAD_ARG, "gcn_conv_bwd: null CSR arrays");
    IGCN_REQUIRE(n_cta == igcn_gcn_conv_bwd_ctas(N), IGCN_ERR_BAD_ARG, "gcn_conv_bwd: n_cta=%lld, expected %lld", (long long)n_cta,
                 (long long)igcn_gcn_conv_bwd_ctas(N));
    const float* u = saved;
    const float* wslot = u + N * O;
    const float* norm = wslot + E;
    const float* dinv = norm + E;
    const float* nii = dinv + N;
    const float* ell = nii + N;
    const int* etgt = reinterpret_cast<const int*>(ell + N);
    float* dU = work;
    float* edn = dU + N * O;
    float* dnii = edn + E;
    float* ddeg = dnii + N;
    // dU = A_n^T g  (gather over the out-edges)
    spmm_rows_kernel<<<blocks_for(N, 8), 256, 0, st>>>((int)N, (int)O, rowptr_s, csc_pos, etgt, norm, nii, g_out, nullptr, 1, dU);
    IGCN_CHECK_LAUNCH("gcn_conv_bwd spmm_t");
    if (d_edge_weight) {
        edge_dots_kernel<<<blocks_for(N, 8), 256, 0, st>>>((int)N, (int)O, rowptr_t, csr_src, g_out, u, edn, dnii);
        IGCN_CHECK_LAUNCH("gcn_conv_bwd edge_dots");
        ddeg_kernel<<<blocks_for(N, 128), 128, 0, st>>>((int)N, rowptr_t, csr_src, rowptr_s, csc_pos, etgt, edn, wslot, dinv, ell, dnii, ddeg);
        IGCN_CHECK_LAUNCH("gcn_conv_bwd ddeg");
        dweight_kernel<<<blocks_for(N, 128), 128, 0, st>>>((int)N, rowptr_t, csr_src, csr_perm, edn, dinv, dnii, ddeg, d_edge_weight);
        IGCN_CHECK_LAUNCH("gcn_conv_bwd dweight");
    }
    dweights_kernel<<<(int)n_cta, 256, 0, st>>>((int)N, (int)C, (int)O, dU, x, g_out, kChunk, partials);
    IGCN_CHECK_LAUNCH("gcn_conv_bwd dweights");
    reduce_partials_kernel<<<(P + 31) / 32, 256, 0, st>>>(partials, (int)n_cta, P, grads);
    IGCN_CHECK_LAUNCH("gcn_conv_bwd reduce");
    if (dx) {
        const size_t wsm = sizeof(float) * (size_t)(C * O);
        int rc = allow_smem(dense_rows_kernel, wsm, "gcn_conv_bwd");
        if (rc) return rc;
        // dx[i][c] = sum_o dU[i][o] W[o][c] : W is (O, C) = the (contraction, output) layout
        dense_rows_kernel<<<blocks_for(N * C, 256), 256, wsm, st>>>(dU, weight, (int)N, (int)O, (int)C, 1, dx);
        IGCN_CHECK_LAUNCH("gcn_conv_bwd dx");
    }
    return IGCN_OK;
}
