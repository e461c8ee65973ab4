// Register-tiled SGCN encoder kernels for the reference's default layer shape: F0 = 3 input features, hidden = 16.
// (ncu on the generic kernels showed them ISSUE bound -- ~64 K warp instructions per 264-node graph for ~3 K warp-FMAs --
// not memory bound, so the specialisation is about instructions per FMA, not about bytes.)
//
// Work decomposition: a thread task is (node i, feature group fg) with 4 consecutive output features; a warp covers
// 8 nodes x 4 groups.  The layer weight rows W[4fg..4fg+3][0..15] live in 64 registers, the node's input row is read
// with 4 broadcast LDS.128, so X.W costs 69 instructions per 64 FMA; the CSR SpMM reads (src, norm) as one LDS.64 and the
// neighbour's 4 features as one LDS.128.  The next graph's inputs (x slab, rowptr slice, CSR slices) are staged with cp.async
// while the current graph's layers run; layer outputs go straight to their concat slot in HBM with 128-bit streaming stores.
// (A first version staged the whole (R, L*16) slab in shared memory and wrote it with one TMA bulk store; ncu showed the
// 128 B row stride of that slab costing 8-way bank conflicts on every X.W row read, so the slab was dropped.)
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
__device__ __forceinline__ float4 spmm_row(const float* U, const int ld, const int2* edges, const int* rp, const float* nii, int i, int fg) {
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
        const float4 u0 = ld4(U + e0.x * ld + 4 * fg), u1 = ld4(U + e1.x * ld + 4 * fg);
        const float4 u2 = ld4(U + e2.x * ld + 4 * fg), u3 = ld4(U + e3.x * ld + 4 * fg);
        axpy4(__int_as_float(e0.y), u0, acc);
        if (n > 1) axpy4(__int_as_float(e1.y), u1, acc);
        if (n > 2) axpy4(__int_as_float(e2.y), u2, acc);
        if (n > 3) axpy4(__int_as_float(e3.y), u3, acc);
    }
    axpy4(nii[i], ld4(U + i * ld + 4 * fg), acc);
    return acc;
}

constexpr int kHP = 20;   // padded row stride (floats) of the staged hidden activations: 80 B rows are bank-conflict free
                          // for the 8-nodes-per-warp LDS.128 / STS.128 pattern (128 B rows give 8-way conflicts)

// kHoist: L == 2, both layers' weights stay in registers across all graphs of the CTA.
template <bool kHoist>
__global__ void __launch_bounds__(256, 2) sgcn_fwd_h16_kernel(EncArgs a) {
    IGCN_PDL_SYNC();
    extern __shared__ __align__(16) float smf[];
    const int R = a.R, L = a.L, LH = L * kH, maxEg = a.maxEg;
    const int tid = threadIdx.x, nt = blockDim.x;
    const int fg = tid & 3;
    // carve: 16-byte aligned regions first
    float* U = smf;                                     // R*16
    float* Hp = U + R * kH;                             // nHp * R*20 : previous layer's activations (ping-pong for L > 2)
    const int nHp = (L > 2) ? 2 : (L > 1 ? 1 : 0);
    float* Wsm = Hp + nHp * R * kHP;                    // wb layout, row-major [f][k] (+bias)
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
        float* og = a.out_w + (int64_t)g * R * LH;
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
                const float* Hprev = Hp + ((l - 1) & 1) * R * kHP;
                // two tasks per iteration: the second task's row loads are in flight during the first task's FMAs
                for (int t = tid; t < ntask; t += 2 * nt) {
                    const int i = t >> 2;
                    const bool two = (t + nt) < ntask;
                    const int j = two ? ((t + nt) >> 2) : i;
                    const float* hi = Hprev + i * kHP;
                    const float* hj = Hprev + j * kHP;
                    const float4 a0 = ld4(hi), a1 = ld4(hi + 4), a2 = ld4(hi + 8), a3 = ld4(hi + 12);
                    const float4 b0 = ld4(hj), b1 = ld4(hj + 4), b2 = ld4(hj + 8), b3 = ld4(hj + 12);
                    st4(U + i * kH + 4 * fg, xw16(a0, a1, a2, a3, w1));
                    if (two) st4(U + j * kH + 4 * fg, xw16(b0, b1, b2, b3, w1));
                }
            }
            __syncthreads();
            // ---- Y = A_norm U + bias ; relu -> concat slot l of the output (+ padded smem copy for the next layer) ----
            float4 bias;
            if (kHoist)
                bias = (l == 0) ? bias0 : bias1;
            else
                bias = ld4(Wsm + off + kH * layer_fin(l, kF0, kH) + 4 * fg);
            float* Hnext = (l + 1 < L) ? Hp + (l & 1) * R * kHP : nullptr;
            for (int t = tid; t < ntask; t += nt) {
                const int i = t >> 2;
                float4 acc = spmm_row(U, kH, edges, rp, nii, i, fg);
                acc.x += bias.x; acc.y += bias.y; acc.z += bias.z; acc.w += bias.w;
                if (a.relu) {
                    acc.x = fmaxf(acc.x, 0.f); acc.y = fmaxf(acc.y, 0.f); acc.z = fmaxf(acc.z, 0.f); acc.w = fmaxf(acc.w, 0.f);
                }
                if (Hnext) st4(Hnext + i * kHP + 4 * fg, acc);
                __stcs(reinterpret_cast<float4*>(og + i * LH + l * kH + 4 * fg), acc);   // streaming: written once, read by other kernels
            }
            __syncthreads();
        }
        e0 = e0n;
        Eg = Egn;
    }
}

static size_t fwd_fast_smem(int R, int L, int maxEg) {
    const int WB = wb_size(kF0, kH, L);
    const int nHp = (L > 2) ? 2 : (L > 1 ? 1 : 0);
    const size_t stage = (size_t)R * kF0 + R + 1 + 2 * (size_t)maxEg;
    return 4 * ((size_t)R * kH + (size_t)nHp * R * kHP + ((WB + 3) & ~3) + 2 * (size_t)maxEg + (size_t)R * kF0 + 2 * (size_t)R + 8 + R + 1 +
                2 * stage) + 16;
}

// threads per CTA: a multiple of 32 from {128..256} that wastes the fewest (node, feature-group) task slots
static int fast_threads(int R, int max_threads = 256) {
    const int ntask = 4 * R;
    int best = 256;
    double best_u = 0.0;
    for (int nt = 128; nt <= max_threads; nt += 32) {
        const int iters = (ntask + nt - 1) / nt;
        const double u = (double)ntask / ((double)iters * nt);
        if (u >= best_u - 1e-9) {
            best_u = u > best_u ? u : best_u;
            best = nt;
        }
    }
    return best;
}


// =====================================================================================================================
// Backward, L == 2, H == 16, F0 == 3.
//
// Uses the aggregate-first factorisation of a GCN layer,  Y = A_n (H W^T) = (A_n H) W^T = Z W^T :
//     dW = G^T Z        dZ = G W        dH = A_n^T dZ        d norm_e = <dZ[t_e], H[s_e]>
// so the mask gradient needs H (already staged) instead of a recomputed U = H W^T, and the two dense products of a task
// (dZ row-quad and the dW outer product) share one read of the G row: 128 FMA for 5 LDS.128.
// One CTA per SM (register budget: 64 weight + 64 dW accumulator registers per thread), thread task = (node, 4-wide group).
// =====================================================================================================================
struct StageB {
    float* x;     // R*3
    int* rp;      // R+1   rowptr_t slice (global offsets)
    int* src;     // maxEg
    float* w;     // maxEg
    int* rps;     // R+1   rowptr_s slice
    int* cpos;    // maxEg
    float* gpe;   // maxEg (optional)
};

// sum over the 4 lanes of a (node) quad; only the quad's own lanes are named in the mask, so quads of one warp may
// sit in loops of different trip counts
__device__ __forceinline__ float quad_sum(float v) {
    const unsigned m = 0xFu << (threadIdx.x & 28);
    v += __shfl_xor_sync(m, v, 1);
    v += __shfl_xor_sync(m, v, 2);
    return v;
}

template <bool kExplain>
__global__ void __launch_bounds__(384, 1) sgcn_bwd_h16_kernel(EncArgs a) {
    IGCN_PDL_SYNC();
    extern __shared__ __align__(16) float smf[];
    const int R = a.R, maxEg = a.maxEg;
    constexpr int LH = 2 * kH;
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarp = nt >> 5;
    const int fg = tid & 3;
    const int ntask = R * 4;
    // ---- carve ---------------------------------------------------------------------------------------------------
    const int WBc = wb_size(kF0, kH, 2);
    float* red = smf;                    // 12 warps * WB : end-of-kernel reduction scratch
    float* bufA = red + 12 * ((WBc + 3) & ~3);   // R*20  H^{1} (layer-2 input)
    float* bufB = bufA + R * kHP;        // R*20  G_l
    float* bufC = bufB + R * kHP;        // R*20  Z, then dH
    float* bufD = bufC + R * kHP;        // R*20  dZ
    float* Wsm = bufD + R * kHP;         // wb (row-major) padded to 4
    const int WB = wb_size(kF0, kH, 2);
    int2* edges = reinterpret_cast<int2*>(Wsm + ((WB + 3) & ~3));   // maxEg (src, norm) target-sorted
    int2* tedges = edges + maxEg;                                    // maxEg (tgt, norm) source-sorted
    float* ew = reinterpret_cast<float*>(tedges + maxEg);            // maxEg masked weight
    float* epe = ew + maxEg;             // maxEg p_e
    float* edn = epe + maxEg;            // maxEg d loss / d norm  (then dz)
    int* spos = reinterpret_cast<int*>(edn + maxEg);                 // maxEg CSR slot of q-th out-edge
    int* etgt = spos + maxEg;            // maxEg
    float* xs = reinterpret_cast<float*>(etgt + maxEg);              // R*3 masked
    float* z0 = xs + R * kF0;            // R*3
    float* dz0 = z0 + R * kF0;           // R*3
    float* dxt = dz0 + R * kF0;          // R*3
    float* dinv = dxt + R * kF0;         // R
    float* nii = dinv + R;               // R
    float* ell = nii + R;                // R
    float* dnii = ell + R;               // R
    float* ddeg = dnii + R;              // R
    float* dprob = ddeg + R;             // R*3 accumulates over the CTA's graphs
    float* pb = dprob + R * kF0;         // 8
    int* rp = reinterpret_cast<int*>(pb + 8);    // R+1
    int* rps = rp + R + 1;               // R+1
    float* stage_base = reinterpret_cast<float*>(rps + R + 1);
    const int stage_len = R * kF0 + 2 * (R + 1) + 4 * maxEg;
    auto stage_of = [&](int s) {
        StageB q;
        float* p = stage_base + s * stage_len;
        q.x = p;
        q.rp = reinterpret_cast<int*>(p + R * kF0);
        q.src = q.rp + R + 1;
        q.w = reinterpret_cast<float*>(q.src + maxEg);
        q.rps = reinterpret_cast<int*>(q.w + maxEg);
        q.cpos = q.rps + R + 1;
        q.gpe = reinterpret_cast<float*>(q.cpos + maxEg);
        return q;
    };
    auto stage_issue_b = [&](int g, int e0, int Eg, const StageB& st) {
        const int64_t node0 = (int64_t)g * R;
        const float* xg = a.x + node0 * kF0;
        for (int i = tid; i < R * kF0; i += nt) cp_async4(st.x + i, xg + i);
        for (int i = tid; i <= R; i += nt) {
            cp_async4(st.rp + i, a.rowptr_t + node0 + i);
            cp_async4(st.rps + i, a.rowptr_s + node0 + i);
        }
        for (int k = tid; k < Eg; k += nt) {
            cp_async4(st.src + k, a.csr_src + e0 + k);
            cp_async4(st.w + k, a.csr_w + e0 + k);
            cp_async4(st.cpos + k, a.csc_pos + e0 + k);
            if (kExplain && a.g_pe) cp_async4(st.gpe + k, a.g_pe + e0 + k);
        }
        cp_async_commit();
    };

    for (int i = tid; i < WB; i += nt) Wsm[i] = a.wb[i];
    if (kExplain) {
        if (tid < 6) pb[tid] = a.prob_bias[tid];
        for (int i = tid; i < R * kF0; i += nt) dprob[i] = 0.f;
    }
    __syncthreads();
    const int off1 = layer_off(1, kF0, kH);
    // weights in registers: wt[f] = W2[f][4fg..4fg+3]  (columns for dZ = G W2) ; w1q[a][c] = W1[4fg+a][c]
    const float* wt_s = Wsm + off1 + 4 * fg;     // wt_s[f*16 .. +3] = W2[f][4fg..4fg+3]: read per use (LDS.128, 4 distinct
                                                  // addresses per warp) -- keeping them in 64 registers forced one 7-warp CTA per SM
    float w1q[4][3];
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int c = 0; c < 3; ++c) w1q[q][c] = Wsm[(4 * fg + q) * 3 + c];
    // gradient accumulators (registers, across all graphs of this CTA)
    float4 acc2[16];     // dW2[f][4fg..4fg+3]
#pragma unroll
    for (int f = 0; f < 16; ++f) acc2[f] = make_float4(0.f, 0.f, 0.f, 0.f);
    float acc1[4][3];    // dW1[4fg+a][c]
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int c = 0; c < 3; ++c) acc1[q][c] = 0.f;
    float4 db2 = make_float4(0.f, 0.f, 0.f, 0.f), db1 = make_float4(0.f, 0.f, 0.f, 0.f);
    float dpb_reg[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};

    int g = blockIdx.x;
    int e0 = 0, Eg = 0, e0n = 0, Egn = 0;
    if (g < a.B) {
        e0 = a.rowptr_t[(int64_t)g * R];
        Eg = a.rowptr_t[(int64_t)(g + 1) * R] - e0;
        if (Eg > maxEg) __trap();
        stage_issue_b(g, e0, Eg, stage_of(0));
    }
    int cur = 0;
    for (; g < a.B; g += gridDim.x, cur ^= 1) {
        const int gn = g + gridDim.x;
        if (gn < a.B) {
            e0n = a.rowptr_t[(int64_t)gn * R];
            Egn = a.rowptr_t[(int64_t)(gn + 1) * R] - e0n;
        }
        const int64_t node0 = (int64_t)g * R;
        const float* go = a.g_out + node0 * LH;
        const float* fo = a.out + node0 * LH;
        // ---- G_2 and H^1 straight from HBM (independent of the prologue) --------------------------------------------
        for (int t = tid; t < ntask; t += nt) {
            const int i = t >> 2;
            const float4 gq = __ldg(reinterpret_cast<const float4*>(go + i * LH + kH + 4 * fg));
            const float4 oq = __ldg(reinterpret_cast<const float4*>(fo + i * LH + kH + 4 * fg));
            const float4 hq = __ldg(reinterpret_cast<const float4*>(fo + i * LH + 4 * fg));
            float4 gl;
            gl.x = (!a.relu || oq.x > 0.f) ? gq.x : 0.f;
            gl.y = (!a.relu || oq.y > 0.f) ? gq.y : 0.f;
            gl.z = (!a.relu || oq.z > 0.f) ? gq.z : 0.f;
            gl.w = (!a.relu || oq.w > 0.f) ? gq.w : 0.f;
            st4(bufB + i * kHP + 4 * fg, gl);
            st4(bufA + i * kHP + 4 * fg, hq);
        }
        cp_async_wait_all();
        __syncthreads();
        const StageB st = stage_of(cur);
        // ---- prologue: masks, degrees, norms (+ what the mask gradient needs) -----------------------------------------
        {
            const bool explain = kExplain;
            const int nd0 = (int)node0;
            for (int i = tid; i < R * kF0; i += nt) xs[i] = explain ? st.x[i] * a.prob[i] : st.x[i];
            for (int i = tid; i <= R; i += nt) {
                rp[i] = st.rp[i] - e0;
                rps[i] = st.rps[i] - e0;
            }
            for (int q = tid; q < Eg; q += nt) {
                spos[q] = st.cpos[q] - e0;
                edn[q] = 0.f;
            }
            __syncthreads();
            for (int i = tid; i < R; i += nt) {
                float deg = 0.f, loopw = 1.f;
                const float xi0 = xs[i * 3], xi1 = xs[i * 3 + 1], xi2 = xs[i * 3 + 2];
                const int k1 = rp[i + 1];
                for (int k = rp[i]; k < k1; ++k) {
                    const int s = st.src[k] - nd0;
                    float wt_ = st.w[k];
                    if (explain) {
                        float z = pb[0] * xs[s * 3] + pb[3] * xi0;
                        z += pb[1] * xs[s * 3 + 1] + pb[4] * xi1;
                        z += pb[2] * xs[s * 3 + 2] + pb[5] * xi2;
                        const float p = sigmoidf_(z);
                        wt_ *= p;
                        epe[k] = p;
                        ew[k] = wt_;
                    }
                    edges[k] = make_int2(s, __float_as_int(wt_));
                    etgt[k] = i;
                    if (s == i)
                        loopw = wt_;
                    else
                        deg += wt_;
                }
                deg += loopw;
                const float d = (deg == 0.f) ? 0.f : rsqrtf(deg);
                dinv[i] = d;
                nii[i] = d * d * loopw;
                ell[i] = loopw;
                dnii[i] = 0.f;
            }
            __syncthreads();
            for (int i = tid; i < R; i += nt) {
                const float di = dinv[i];
                const int k1 = rp[i + 1];
                float zx = 0.f, zy = 0.f, zz = 0.f;
                for (int k = rp[i]; k < k1; ++k) {
                    const int2 e = edges[k];
                    const float n = (e.x == i) ? 0.f : dinv[e.x] * __int_as_float(e.y) * di;
                    edges[k].y = __float_as_int(n);
                    zx = fmaf(n, xs[e.x * 3], zx);
                    zy = fmaf(n, xs[e.x * 3 + 1], zy);
                    zz = fmaf(n, xs[e.x * 3 + 2], zz);
                }
                const float ns = nii[i];
                z0[i * 3] = fmaf(ns, xs[i * 3], zx);        // Z0 = A_n x~ (layer-1 aggregate)
                z0[i * 3 + 1] = fmaf(ns, xs[i * 3 + 1], zy);
                z0[i * 3 + 2] = fmaf(ns, xs[i * 3 + 2], zz);
            }
            __syncthreads();
            for (int q = tid; q < Eg; q += nt) {
                const int k = spos[q];
                tedges[q] = make_int2(etgt[k], edges[k].y);
            }
        }
        if (gn < a.B) {
            if (Egn > maxEg) __trap();
            stage_issue_b(gn, e0n, Egn, stage_of(cur ^ 1));
        }
        // ---- layer 2, phase 1: Z = A_n H^1 ---------------------------------------------------------------------------
        for (int t = tid; t < ntask; t += nt) {
            const int i = t >> 2;
            st4(bufC + i * kHP + 4 * fg, spmm_row(bufA, kHP, edges, rp, nii, i, fg));
        }
        __syncthreads();
        // ---- layer 2, phase 2: dZ = G W2 ; dW2 += G^T Z ; db2 += G ------------------------------------------------------
        for (int t = tid; t < ntask; t += nt) {
            const int i = t >> 2;
            const float* gr = bufB + i * kHP;
            const float4 g0 = ld4(gr), g1 = ld4(gr + 4), g2 = ld4(gr + 8), g3 = ld4(gr + 12);
            const float4 zq = ld4(bufC + i * kHP + 4 * fg);
            const float4 gme = ld4(gr + 4 * fg);
            float4 dz = make_float4(0.f, 0.f, 0.f, 0.f);
#define IGCN_BWD_STEP(GV, F)                    \
    axpy4(GV, ld4(wt_s + (F) * kH), dz);        \
    axpy4(GV, zq, acc2[F]);
            IGCN_BWD_STEP(g0.x, 0) IGCN_BWD_STEP(g0.y, 1) IGCN_BWD_STEP(g0.z, 2) IGCN_BWD_STEP(g0.w, 3)
            IGCN_BWD_STEP(g1.x, 4) IGCN_BWD_STEP(g1.y, 5) IGCN_BWD_STEP(g1.z, 6) IGCN_BWD_STEP(g1.w, 7)
            IGCN_BWD_STEP(g2.x, 8) IGCN_BWD_STEP(g2.y, 9) IGCN_BWD_STEP(g2.z, 10) IGCN_BWD_STEP(g2.w, 11)
            IGCN_BWD_STEP(g3.x, 12) IGCN_BWD_STEP(g3.y, 13) IGCN_BWD_STEP(g3.z, 14) IGCN_BWD_STEP(g3.w, 15)
#undef IGCN_BWD_STEP
            db2.x += gme.x; db2.y += gme.y; db2.z += gme.z; db2.w += gme.w;
            st4(bufD + i * kHP + 4 * fg, dz);
        }
        __syncthreads();
        // ---- layer 2, phase 3: dH^1 = A_n^T dZ  (+ d norm_e, d n_ii) ; fused with layer 1: G_1, dW1, db1, dZ0 ---------------
        for (int t = tid; t < ntask; t += nt) {
            const int i = t >> 2;
            const float4 dzi = ld4(bufD + i * kHP + 4 * fg);
            float4 dh = make_float4(0.f, 0.f, 0.f, 0.f);
            const int q1 = rps[i + 1];
            for (int q = rps[i]; q < q1; ++q) {
                const int2 e = tedges[q];
                axpy4(__int_as_float(e.y), ld4(bufD + e.x * kHP + 4 * fg), dh);
            }
            axpy4(nii[i], dzi, dh);
            if (kExplain) {
                const int k1 = rp[i + 1];
                for (int k = rp[i]; k < k1; ++k) {
                    const float p = quad_sum(dot4(dzi, ld4(bufA + edges[k].x * kHP + 4 * fg), 0.f));
                    if (fg == 0) edn[k] += p;
                }
                const float p = quad_sum(dot4(dzi, ld4(bufA + i * kHP + 4 * fg), 0.f));
                if (fg == 0) dnii[i] += p;
            }
            // layer 1 on the same (node, quad): G_1 = (g_out slot 0 + dH^1) * relu'
            const float4 gq = __ldg(reinterpret_cast<const float4*>(go + i * LH + 4 * fg));
            const float4 oq = ld4(bufA + i * kHP + 4 * fg);      // H^1 = forward output slot 0
            float4 g1v;
            g1v.x = (!a.relu || oq.x > 0.f) ? gq.x + dh.x : 0.f;
            g1v.y = (!a.relu || oq.y > 0.f) ? gq.y + dh.y : 0.f;
            g1v.z = (!a.relu || oq.z > 0.f) ? gq.z + dh.z : 0.f;
            g1v.w = (!a.relu || oq.w > 0.f) ? gq.w + dh.w : 0.f;
            db1.x += g1v.x; db1.y += g1v.y; db1.z += g1v.z; db1.w += g1v.w;
            const float zc0 = z0[i * 3], zc1 = z0[i * 3 + 1], zc2 = z0[i * 3 + 2];
            const float gv[4] = {g1v.x, g1v.y, g1v.z, g1v.w};
            float p0 = 0.f, p1 = 0.f, p2 = 0.f;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                acc1[q][0] = fmaf(gv[q], zc0, acc1[q][0]);
                acc1[q][1] = fmaf(gv[q], zc1, acc1[q][1]);
                acc1[q][2] = fmaf(gv[q], zc2, acc1[q][2]);
                p0 = fmaf(gv[q], w1q[q][0], p0);
                p1 = fmaf(gv[q], w1q[q][1], p1);
                p2 = fmaf(gv[q], w1q[q][2], p2);
            }
            p0 = quad_sum(p0);
            p1 = quad_sum(p1);
            p2 = quad_sum(p2);
            if (fg == 0) {                                    // dZ0 = G_1 W1
                dz0[i * 3] = p0;
                dz0[i * 3 + 1] = p1;
                dz0[i * 3 + 2] = p2;
            }
        }
        __syncthreads();
        // ---- layer 1, phase 3: d x~ = A_n^T dZ0 (+ d norm_e, d n_ii) -- thread per node ---------------------------------
        for (int i = tid; i < R; i += nt) {
            const float d0 = dz0[i * 3], d1 = dz0[i * 3 + 1], d2 = dz0[i * 3 + 2];
            float s0 = 0.f, s1 = 0.f, s2 = 0.f;
            const int q1 = rps[i + 1];
            for (int q = rps[i]; q < q1; ++q) {
                const int2 e = tedges[q];
                const float n = __int_as_float(e.y);
                s0 = fmaf(n, dz0[e.x * 3], s0);
                s1 = fmaf(n, dz0[e.x * 3 + 1], s1);
                s2 = fmaf(n, dz0[e.x * 3 + 2], s2);
            }
            const float ns = nii[i];
            dxt[i * 3] = fmaf(ns, d0, s0);
            dxt[i * 3 + 1] = fmaf(ns, d1, s1);
            dxt[i * 3 + 2] = fmaf(ns, d2, s2);
            if (kExplain) {
                const int k1 = rp[i + 1];
                for (int k = rp[i]; k < k1; ++k) {
                    const int s = edges[k].x;
                    edn[k] += d0 * xs[s * 3] + d1 * xs[s * 3 + 1] + d2 * xs[s * 3 + 2];
                }
                dnii[i] += d0 * xs[i * 3] + d1 * xs[i * 3 + 1] + d2 * xs[i * 3 + 2];
            }
        }
        __syncthreads();
        float* dxg = a.dx + node0 * kF0;
        if (kExplain) {
            // ---- gradient through the symmetric normalisation and the masks (see the generic kernel for the algebra) ----
            for (int i = tid; i < R; i += nt) {
                float dd = 0.f;
                const int k1 = rp[i + 1];
                for (int k = rp[i]; k < k1; ++k) {
                    const int s = edges[k].x;
                    if (s != i) dd = fmaf(edn[k] * ew[k], dinv[s], dd);
                }
                const int q1 = rps[i + 1];
                for (int q = rps[i]; q < q1; ++q) {
                    const int k = spos[q];
                    const int t = tedges[q].x;
                    if (t != i) dd = fmaf(edn[k] * ew[k], dinv[t], dd);
                }
                const float di = dinv[i];
                dd = fmaf(2.f * di * ell[i], dnii[i], dd);
                ddeg[i] = -0.5f * di * di * di * dd;
            }
            __syncthreads();
            for (int i = tid; i < R; i += nt) {
                const float di = dinv[i];
                const float xi0 = xs[i * 3], xi1 = xs[i * 3 + 1], xi2 = xs[i * 3 + 2];
                float sdz = 0.f;
                const int k1 = rp[i + 1];
                for (int k = rp[i]; k < k1; ++k) {
                    const int s = edges[k].x;
                    const float dwt = (s != i) ? dinv[s] * di * edn[k] + ddeg[i] : di * di * dnii[i] + ddeg[i];
                    float dp = st.w[k] * dwt;
                    if (a.g_pe) dp += st.gpe[k];
                    const float p = epe[k];
                    const float dz = p * (1.f - p) * dp;
                    edn[k] = dz;
                    dpb_reg[0] = fmaf(dz, xs[s * 3], dpb_reg[0]);
                    dpb_reg[1] = fmaf(dz, xs[s * 3 + 1], dpb_reg[1]);
                    dpb_reg[2] = fmaf(dz, xs[s * 3 + 2], dpb_reg[2]);
                    dpb_reg[3] = fmaf(dz, xi0, dpb_reg[3]);
                    dpb_reg[4] = fmaf(dz, xi1, dpb_reg[4]);
                    dpb_reg[5] = fmaf(dz, xi2, dpb_reg[5]);
                    sdz += dz;
                }
                dxt[i * 3] = fmaf(sdz, pb[3], dxt[i * 3]);
                dxt[i * 3 + 1] = fmaf(sdz, pb[4], dxt[i * 3 + 1]);
                dxt[i * 3 + 2] = fmaf(sdz, pb[5], dxt[i * 3 + 2]);
            }
            __syncthreads();
            for (int i = tid; i < R; i += nt) {
                float sdz = 0.f;
                const int q1 = rps[i + 1];
                for (int q = rps[i]; q < q1; ++q) sdz += edn[spos[q]];
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const float dv = fmaf(sdz, pb[c], dxt[i * 3 + c]);
                    dprob[i * 3 + c] += st.x[i * 3 + c] * dv;
                    dxg[i * 3 + c] = a.prob[i * 3 + c] * dv;
                }
            }
        } else {
            for (int i = tid; i < R * kF0; i += nt) dxg[i] = dxt[i];
        }
        __syncthreads();
        e0 = e0n;
        Eg = Egn;
    }
    cp_async_wait_all();
    // ---- CTA-level reduction of the register accumulators, one partial row per CTA -------------------------------------
    // lanes with the same fg inside a warp: xor-shuffle over lane bits 2,3,4
    auto fgsum = [](float v) {
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        v += __shfl_xor_sync(0xffffffffu, v, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 16);
        return v;
    };
    __syncthreads();
    float* prow = a.partials + (int64_t)blockIdx.x * a.P;
    {
        // dW1[4fg+q][c] , db1[4fg+q] , dW2[f][4fg+j] , db2[4fg+j]   at their wb offsets
        float* rw = red + warp * WB;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float v = fgsum(acc1[q][c]);
                if (lane < 4) rw[(4 * fg + q) * 3 + c] = v;
            }
        }
        {
            const float v0 = fgsum(db1.x), v1 = fgsum(db1.y), v2 = fgsum(db1.z), v3 = fgsum(db1.w);
            if (lane < 4) {
                rw[kH * kF0 + 4 * fg] = v0; rw[kH * kF0 + 4 * fg + 1] = v1; rw[kH * kF0 + 4 * fg + 2] = v2; rw[kH * kF0 + 4 * fg + 3] = v3;
            }
        }
#pragma unroll
        for (int f = 0; f < 16; ++f) {
            const float v0 = fgsum(acc2[f].x), v1 = fgsum(acc2[f].y), v2 = fgsum(acc2[f].z), v3 = fgsum(acc2[f].w);
            if (lane < 4) {
                float* d = rw + off1 + f * kH + 4 * fg;
                d[0] = v0; d[1] = v1; d[2] = v2; d[3] = v3;
            }
        }
        {
            const float v0 = fgsum(db2.x), v1 = fgsum(db2.y), v2 = fgsum(db2.z), v3 = fgsum(db2.w);
            if (lane < 4) {
                float* d = rw + off1 + kH * kH + 4 * fg;
                d[0] = v0; d[1] = v1; d[2] = v2; d[3] = v3;
            }
        }
    }
    __syncthreads();
    for (int j = tid; j < WB; j += nt) {
        float s = 0.f;
        for (int w = 0; w < nwarp; ++w) s += red[w * WB + j];
        prow[j] = s;
    }
    if (kExplain) {
        for (int j = tid; j < R * kF0; j += nt) prow[WB + j] = dprob[j];
        float* r2 = bufB;                      // [nwarp][8]
#pragma unroll
        for (int c = 0; c < 6; ++c) {
            const float v = warp_sum(dpb_reg[c]);
            if (lane == 0) r2[warp * 8 + c] = v;
        }
        __syncthreads();
        if (tid < 6) {
            float s = 0.f;
            for (int w = 0; w < nwarp; ++w) s += r2[w * 8 + tid];
            prow[WB + R * kF0 + tid] = s;
        }
    } else {
        for (int j = tid; j < R * kF0 + 6; j += nt) prow[WB + j] = 0.f;
    }
}

static size_t bwd_fast_smem(int R, int maxEg) {
    const int WB = wb_size(kF0, kH, 2);
    const size_t stage = (size_t)R * kF0 + 2 * ((size_t)R + 1) + 4 * (size_t)maxEg;
    size_t fl = 12 * (size_t)((WB + 3) & ~3) + 4 * (size_t)R * kHP + ((WB + 3) & ~3) + 4 * (size_t)maxEg /* edges, tedges (int2) */ + 3 * (size_t)maxEg + 2 * (size_t)maxEg +
                4 * (size_t)R * kF0 + 5 * (size_t)R + (size_t)R * kF0 + 8 + 2 * ((size_t)R + 1) + 2 * stage;
    return 4 * fl + 16;
}

static int fast_threads_bwd(int R) { return fast_threads(R, 384); }   // one CTA per SM: more warps hide more latency

}  // namespace igcn
