// Register-tiled SGCN encoder kernels for the reference's default layer shape: F0 = 3 input features, hidden = 16.
// (ncu on the generic kernels showed them ISSUE bound -- ~64 K warp instructions per 264-node graph for ~3 K warp-FMAs --
// not memory bound, so the specialisation is about instructions per FMA, not about bytes.)
//
// Work decomposition: a thread task is (node i, feature group fg) with 4 consecutive output features; a warp covers
// 8 nodes x 4 groups.  The layer weight rows W[4fg..4fg+3][0..15] live in 64 registers, the node's input row is read
// with 4 broadcast LDS.128, so X.W costs 69 instructions per 64 FMA; the CSR SpMM reads (src, norm) as one LDS.64 and the
// neighbour's 4 features as one LDS.128.  The concatenated per-graph output slab is written with ONE TMA bulk store
// (cp.async.bulk.global.shared::cta) that overlaps the next graph's prologue.
#pragma once

namespace igcn {

constexpr int kH = 16;
constexpr int kF0 = 3;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void bulk_store_slab(float* gdst, const float* ssrc, uint32_t bytes) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float dot4(float4 a, float4 b, float acc) {
    acc = fmaf(a.x, b.x, acc);
    acc = fmaf(a.y, b.y, acc);
    acc = fmaf(a.z, b.z, acc);
    return fmaf(a.w, b.w, acc);
}
__device__ __forceinline__ void axpy4(float s, float4 v, float4& acc) {
    acc.x = fmaf(s, v.x, acc.x);
    acc.y = fmaf(s, v.y, acc.y);
    acc.z = fmaf(s, v.z, acc.z);
    acc.w = fmaf(s, v.w, acc.w);
}

// ---- cp.async (LDGSTS) staging of one graph's inputs: x slab, rowptr slice, CSR (src, w) slices -----------------
__device__ __forceinline__ void cp_async4(void* sdst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(sdst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

struct Stage {
    float* x;    // R*3 raw features
    int* rp;     // R+1 global CSR offsets
    int* src;    // maxEg global source ids
    float* w;    // maxEg raw edge weights
};

__device__ __forceinline__ void stage_issue(const EncArgs& a, int g, int e0, int Eg, const Stage& st) {
    const int tid = threadIdx.x, nt = blockDim.x, R = a.R;
    const int64_t node0 = (int64_t)g * R;
    const float* xg = a.x + node0 * kF0;
    for (int i = tid; i < R * kF0; i += nt) cp_async4(st.x + i, xg + i);
    for (int i = tid; i <= R; i += nt) cp_async4(st.rp + i, a.rowptr_t + node0 + i);
    for (int k = tid; k < Eg; k += nt) {
        cp_async4(st.src + k, a.csr_src + e0 + k);
        cp_async4(st.w + k, a.csr_w + e0 + k);
    }
    cp_async_commit();
}

// Per-graph prologue for the fast kernels (F0 = 3), reading the staged inputs: masks, self-loop merge, degrees,
// normalised weights.   edges[k] = (local src, bits of norm_e)   [norm 0 on self-loop slots]
template <bool kKeep>
__device__ __forceinline__ void fast_prologue(const EncArgs& a, int g, int e0, int Eg, const Stage& st, float* xs, int* rp,
                                              int2* edges, float* dinv, float* nii, float* ew, float* epe, float* ell,
                                              const float* pb) {
    const int tid = threadIdx.x, nt = blockDim.x;
    const int R = a.R;
    const bool explain = a.prob != nullptr;
    const int node0 = g * R;
    for (int i = tid; i < R * kF0; i += nt) xs[i] = explain ? st.x[i] * a.prob[i] : st.x[i];
    for (int i = tid; i <= R; i += nt) rp[i] = st.rp[i] - e0;
    __syncthreads();
    for (int i = tid; i < R; i += nt) {
        float deg = 0.f, loopw = 1.f;
        const float xi0 = xs[i * 3], xi1 = xs[i * 3 + 1], xi2 = xs[i * 3 + 2];
        const int k1 = rp[i + 1];
        for (int k = rp[i]; k < k1; ++k) {
            const int s = st.src[k] - node0;
            float wt = st.w[k];
            if (explain) {
                // same association as the reference's [x_src | x_dst] . prob_bias : pairs (src_c, dst_c) summed in order
                float z = pb[0] * xs[s * 3] + pb[3] * xi0;
                z += pb[1] * xs[s * 3 + 1] + pb[4] * xi1;
                z += pb[2] * xs[s * 3 + 2] + pb[5] * xi2;
                const float p = sigmoidf_(z);
                wt *= p;
                if (a.pe_w) a.pe_w[e0 + k] = p;
                if (kKeep) epe[k] = p;
            }
            if (kKeep) ew[k] = wt;
            edges[k] = make_int2(s, __float_as_int(wt));
            if (s == i)
                loopw = wt;  // last self loop wins
            else
                deg += wt;
        }
        deg += loopw;
        const float d = (deg == 0.f) ? 0.f : rsqrtf(deg);
        dinv[i] = d;
        nii[i] = d * d * loopw;
        if (kKeep) ell[i] = loopw;
    }
    __syncthreads();
    for (int i = tid; i < R; i += nt) {
        const float di = dinv[i];
        const int k1 = rp[i + 1];
        for (int k = rp[i]; k < k1; ++k) {
            const int2 e = edges[k];
            const float n = (e.x == i) ? 0.f : dinv[e.x] * __int_as_float(e.y) * di;
            edges[k].y = __float_as_int(n);
        }
    }
    __syncthreads();
}

// U[i][4fg..4fg+3] = H_prev[i][0..15] . W[4fg+a][0..15]     (layer >= 2)
__device__ __forceinline__ float4 xw16(const float4 h0, const float4 h1, const float4 h2, const float4 h3, const float4 (&w)[4][4]) {
    float4 r;
    r.x = dot4(h3, w[0][3], dot4(h2, w[0][2], dot4(h1, w[0][1], dot4(h0, w[0][0], 0.f))));
    r.y = dot4(h3, w[1][3], dot4(h2, w[1][2], dot4(h1, w[1][1], dot4(h0, w[1][0], 0.f))));
    r.z = dot4(h3, w[2][3], dot4(h2, w[2][2], dot4(h1, w[2][1], dot4(h0, w[2][0], 0.f))));
    r.w = dot4(h3, w[3][3], dot4(h2, w[3][2], dot4(h1, w[3][1], dot4(h0, w[3][0], 0.f))));
    return r;
}

// Y[i][4fg..] = sum_{k in row i} norm_k U[src_k][4fg..] + n_ii U[i][4fg..]   (edge order, self loop last)
__device__ __forceinline__ float4 spmm_row(const float* U, const int2* edges, const int* rp, const float* nii, int i, int fg) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int k = rp[i];
    const int k1 = rp[i + 1];
    // rows of GDC top-k graphs have 3 in-edges: fetch up to 4 (src,norm) pairs and their feature quads before the FMAs
    for (; k < k1; k += 4) {
        const int n = k1 - k;
        const int2 e0 = edges[k];
        const int2 e1 = (n > 1) ? edges[k + 1] : make_int2(i, 0);
        const int2 e2 = (n > 2) ? edges[k + 2] : make_int2(i, 0);
        const int2 e3 = (n > 3) ? edges[k + 3] : make_int2(i, 0);
        const float4 u0 = ld4(U + e0.x * kH + 4 * fg), u1 = ld4(U + e1.x * kH + 4 * fg);
        const float4 u2 = ld4(U + e2.x * kH + 4 * fg), u3 = ld4(U + e3.x * kH + 4 * fg);
        axpy4(__int_as_float(e0.y), u0, acc);
        if (n > 1) axpy4(__int_as_float(e1.y), u1, acc);
        if (n > 2) axpy4(__int_as_float(e2.y), u2, acc);
        if (n > 3) axpy4(__int_as_float(e3.y), u3, acc);
    }
    axpy4(nii[i], ld4(U + i * kH + 4 * fg), acc);
    return acc;
}

// kHoist: L == 2, both layers' weights stay in registers across all graphs of the CTA.
template <bool kHoist>
__global__ void __launch_bounds__(256, 2) sgcn_fwd_h16_kernel(EncArgs a) {
    extern __shared__ __align__(16) float smf[];
    const int R = a.R, L = a.L, LH = L * kH, maxEg = a.maxEg;
    const int tid = threadIdx.x, nt = blockDim.x;
    const int fg = tid & 3;
    // carve: 16-byte aligned regions first
    float* Hbuf = smf;                                  // R*LH
    float* U = Hbuf + R * LH;                           // R*16
    float* Wsm = U + R * kH;                            // wb layout, row-major [f][k] (+bias)
    const int WB = wb_size(kF0, kH, L);
    int2* edges = reinterpret_cast<int2*>(Wsm + ((WB + 3) & ~3));   // maxEg
    float* xs = reinterpret_cast<float*>(edges + maxEg);            // R*3
    float* dinv = xs + R * kF0;                         // R
    float* nii = dinv + R;                              // R
    float* pb = nii + R;                                // 8
    int* rp = reinterpret_cast<int*>(pb + 8);           // R+1
    float* stage_base = reinterpret_cast<float*>(rp + R + 1);
    const int stage_len = R * kF0 + R + 1 + 2 * maxEg;
    auto stage_of = [&](int s) {
        Stage q;
        float* p = stage_base + s * stage_len;
        q.x = p;
        q.rp = reinterpret_cast<int*>(p + R * kF0);
        q.src = q.rp + R + 1;
        q.w = reinterpret_cast<float*>(q.src + maxEg);
        return q;
    };
    for (int i = tid; i < WB; i += nt) Wsm[i] = a.wb[i];
    if (a.prob_bias && tid < 6) pb[tid] = a.prob_bias[tid];
    __syncthreads();
    const int ntask = R * 4;
    bool store_pending = false;

    float w0[4][3];
    float4 w1[4][4];
    float4 bias0, bias1;
    if (kHoist) {
        const int off1 = layer_off(1, kF0, kH);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
#pragma unroll
            for (int k = 0; k < 3; ++k) w0[q][k] = Wsm[(4 * fg + q) * 3 + k];
#pragma unroll
            for (int c = 0; c < 4; ++c) w1[q][c] = ld4(Wsm + off1 + (4 * fg + q) * kH + 4 * c);
        }
        bias0 = ld4(Wsm + kH * kF0 + 4 * fg);
        bias1 = ld4(Wsm + off1 + kH * kH + 4 * fg);
    }

    int g = blockIdx.x;
    int e0 = 0, Eg = 0, e0n = 0, Egn = 0;
    if (g < a.B) {
        e0 = a.rowptr_t[(int64_t)g * R];
        Eg = a.rowptr_t[(int64_t)(g + 1) * R] - e0;
        if (Eg > maxEg) __trap();
        stage_issue(a, g, e0, Eg, stage_of(0));
    }
    int cur = 0;
    for (; g < a.B; g += gridDim.x, cur ^= 1) {
        const int gn = g + gridDim.x;
        if (gn < a.B) {   // offsets of the NEXT graph: issued now, consumed after the prologue
            e0n = a.rowptr_t[(int64_t)gn * R];
            Egn = a.rowptr_t[(int64_t)(gn + 1) * R] - e0n;
        }
        cp_async_wait_all();
        __syncthreads();
        fast_prologue<false>(a, g, e0, Eg, stage_of(cur), xs, rp, edges, dinv, nii, nullptr, nullptr, nullptr, pb);
        if (gn < a.B) {
            if (Egn > maxEg) __trap();
            stage_issue(a, gn, e0n, Egn, stage_of(cur ^ 1));   // lands while this graph's layers run
        }
        for (int l = 0; l < L; ++l) {
            const int off = layer_off(l, kF0, kH);
            // ---- U = H_prev . W^T --------------------------------------------------------------------------
            if (l == 0) {
                if (!kHoist) {
#pragma unroll
                    for (int q = 0; q < 4; ++q)
#pragma unroll
                        for (int k = 0; k < 3; ++k) w0[q][k] = Wsm[off + (4 * fg + q) * 3 + k];
                }
                for (int t = tid; t < ntask; t += nt) {
                    const int i = t >> 2;
                    const float x0 = xs[i * 3], x1 = xs[i * 3 + 1], x2 = xs[i * 3 + 2];
                    float4 r;
                    r.x = fmaf(x2, w0[0][2], fmaf(x1, w0[0][1], x0 * w0[0][0]));
                    r.y = fmaf(x2, w0[1][2], fmaf(x1, w0[1][1], x0 * w0[1][0]));
                    r.z = fmaf(x2, w0[2][2], fmaf(x1, w0[2][1], x0 * w0[2][0]));
                    r.w = fmaf(x2, w0[3][2], fmaf(x1, w0[3][1], x0 * w0[3][0]));
                    st4(U + i * kH + 4 * fg, r);
                }
            } else {
                if (!kHoist) {
#pragma unroll
                    for (int q = 0; q < 4; ++q)
#pragma unroll
                        for (int c = 0; c < 4; ++c) w1[q][c] = ld4(Wsm + off + (4 * fg + q) * kH + 4 * c);
                }
                // two tasks per iteration: the second task's row loads are in flight during the first task's FMAs
                for (int t = tid; t < ntask; t += 2 * nt) {
                    const int i = t >> 2;
                    const bool two = (t + nt) < ntask;
                    const int j = two ? ((t + nt) >> 2) : i;
                    const float* hi = Hbuf + i * LH + (l - 1) * kH;
                    const float* hj = Hbuf + j * LH + (l - 1) * kH;
                    const float4 a0 = ld4(hi), a1 = ld4(hi + 4), a2 = ld4(hi + 8), a3 = ld4(hi + 12);
                    const float4 b0 = ld4(hj), b1 = ld4(hj + 4), b2 = ld4(hj + 8), b3 = ld4(hj + 12);
                    st4(U + i * kH + 4 * fg, xw16(a0, a1, a2, a3, w1));
                    if (two) st4(U + j * kH + 4 * fg, xw16(b0, b1, b2, b3, w1));
                }
            }
            if (l == 0 && store_pending && tid == 0) bulk_store_wait_read();   // previous graph's slab has left Hbuf
            __syncthreads();
            // ---- Y = A_norm U + bias ; relu -> concat slot l ---------------------------------------------------
            float4 bias;
            if (kHoist)
                bias = (l == 0) ? bias0 : bias1;
            else
                bias = ld4(Wsm + off + kH * layer_fin(l, kF0, kH) + 4 * fg);
            for (int t = tid; t < ntask; t += nt) {
                const int i = t >> 2;
                float4 acc = spmm_row(U, edges, rp, nii, i, fg);
                acc.x += bias.x; acc.y += bias.y; acc.z += bias.z; acc.w += bias.w;
                if (a.relu) {
                    acc.x = fmaxf(acc.x, 0.f); acc.y = fmaxf(acc.y, 0.f); acc.z = fmaxf(acc.z, 0.f); acc.w = fmaxf(acc.w, 0.f);
                }
                st4(Hbuf + i * LH + l * kH + 4 * fg, acc);
            }
            __syncthreads();
        }
        // ---- one TMA bulk store of the (R, L*16) slab ---------------------------------------------------------
        if (tid == 0) bulk_store_slab(a.out_w + (int64_t)g * R * LH, Hbuf, (uint32_t)(R * LH * sizeof(float)));
        store_pending = true;
        e0 = e0n;
        Eg = Egn;
    }
    if (store_pending && tid == 0) bulk_store_wait_all();
}

static size_t fwd_fast_smem(int R, int L, int maxEg) {
    const int WB = wb_size(kF0, kH, L);
    const size_t stage = (size_t)R * kF0 + R + 1 + 2 * (size_t)maxEg;
    return 4 * ((size_t)R * L * kH + (size_t)R * kH + ((WB + 3) & ~3) + 2 * (size_t)maxEg + (size_t)R * kF0 + 2 * (size_t)R + 8 + R + 1 +
                2 * stage) + 16;
}

// threads per CTA: a multiple of 32 from {128..256} that wastes the fewest (node, feature-group) task slots
static int fast_threads(int R) {
    const int ntask = 4 * R;
    int best = 256;
    double best_u = 0.0;
    for (int nt = 128; nt <= 256; nt += 32) {
        const int iters = (ntask + nt - 1) / nt;
        const double u = (double)ntask / ((double)iters * nt);
        if (u >= best_u - 1e-9) {
            best_u = u > best_u ? u : best_u;
            best = nt;
        }
    }
    return best;
}

}  // namespace igcn
