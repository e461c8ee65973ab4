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

// Per-graph prologue for the fast kernels (F0 = 3): masks, self-loop merge, degrees, normalised weights.
//   edges[k] = (local src, bits of norm_e)   [norm 0 on self-loop slots]
template <bool kKeep>
__device__ __forceinline__ void fast_prologue(const EncArgs& a, int g, int e0, int Eg, float* xs, float* xraw, int* rp, int2* edges,
                                              float* dinv, float* nii, float* ew, float* epe, float* ell, const float* pb) {
    const int tid = threadIdx.x, nt = blockDim.x;
    const int R = a.R;
    const bool explain = a.prob != nullptr;
    const int64_t node0 = (int64_t)g * R;
    const float* xg = a.x + node0 * kF0;
    for (int i = tid; i < R * kF0; i += nt) {
        const float v = xg[i];
        if (kKeep) xraw[i] = v;
        xs[i] = explain ? v * a.prob[i] : v;
    }
    for (int i = tid; i <= R; i += nt) rp[i] = a.rowptr_t[node0 + i] - e0;
    for (int k = tid; k < Eg; k += nt) edges[k] = make_int2(a.csr_src[e0 + k] - (int)node0, __float_as_int(a.csr_w[e0 + k]));
    __syncthreads();
    for (int i = tid; i < R; i += nt) {
        float deg = 0.f, loopw = 1.f;
        const float xi0 = xs[i * 3], xi1 = xs[i * 3 + 1], xi2 = xs[i * 3 + 2];
        for (int k = rp[i]; k < rp[i + 1]; ++k) {
            const int2 e = edges[k];
            const int s = e.x;
            float wt = __int_as_float(e.y);
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
            edges[k].y = __float_as_int(wt);
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
        for (int k = rp[i]; k < rp[i + 1]; ++k) {
            const int2 e = edges[k];
            const float n = (e.x == i) ? 0.f : dinv[e.x] * __int_as_float(e.y) * di;
            edges[k].y = __float_as_int(n);
        }
    }
    __syncthreads();
}

// U[i][4fg..4fg+3] = H_prev[i][0..15] . W[4fg+a][0..15]     (layer >= 2)
__device__ __forceinline__ float4 xw16(const float* hrow, const float4 (&w)[4][4]) {
    const float4 h0 = ld4(hrow), h1 = ld4(hrow + 4), h2 = ld4(hrow + 8), h3 = ld4(hrow + 12);
    float4 r;
    r.x = dot4(h3, w[0][3], dot4(h2, w[0][2], dot4(h1, w[0][1], dot4(h0, w[0][0], 0.f))));
    r.y = dot4(h3, w[1][3], dot4(h2, w[1][2], dot4(h1, w[1][1], dot4(h0, w[1][0], 0.f))));
    r.z = dot4(h3, w[2][3], dot4(h2, w[2][2], dot4(h1, w[2][1], dot4(h0, w[2][0], 0.f))));
    r.w = dot4(h3, w[3][3], dot4(h2, w[3][2], dot4(h1, w[3][1], dot4(h0, w[3][0], 0.f))));
    return r;
}

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

    for (int i = tid; i < WB; i += nt) Wsm[i] = a.wb[i];
    if (a.prob_bias && tid < 6) pb[tid] = a.prob_bias[tid];
    __syncthreads();
    const int ntask = R * 4;
    bool store_pending = false;

    for (int g = blockIdx.x; g < a.B; g += gridDim.x) {
        const int e0 = a.rowptr_t[(int64_t)g * R];
        const int Eg = a.rowptr_t[(int64_t)(g + 1) * R] - e0;
        if (Eg > maxEg) __trap();
        fast_prologue<false>(a, g, e0, Eg, xs, nullptr, rp, edges, dinv, nii, nullptr, nullptr, nullptr, pb);
        for (int l = 0; l < L; ++l) {
            const int off = layer_off(l, kF0, kH);
            // ---- U = H_prev . W^T --------------------------------------------------------------------------
            if (l == 0) {
                float w0[4][3];
#pragma unroll
                for (int q = 0; q < 4; ++q)
#pragma unroll
                    for (int k = 0; k < 3; ++k) w0[q][k] = Wsm[off + (4 * fg + q) * 3 + k];
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
                float4 w[4][4];
#pragma unroll
                for (int q = 0; q < 4; ++q)
#pragma unroll
                    for (int c = 0; c < 4; ++c) w[q][c] = ld4(Wsm + off + (4 * fg + q) * kH + 4 * c);
                for (int t = tid; t < ntask; t += nt) {
                    const int i = t >> 2;
                    st4(U + i * kH + 4 * fg, xw16(Hbuf + i * LH + (l - 1) * kH, w));
                }
            }
            if (l == 0 && store_pending && tid == 0) bulk_store_wait_read();   // previous graph's slab has left Hbuf
            __syncthreads();
            // ---- Y = A_norm U + bias ; relu -> concat slot l ---------------------------------------------------
            const float4 bias = ld4(Wsm + off + kH * layer_fin(l, kF0, kH) + 4 * fg);
            for (int t = tid; t < ntask; t += nt) {
                const int i = t >> 2;
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                const int k1 = rp[i + 1];
                for (int k = rp[i]; k < k1; ++k) {
                    const int2 e = edges[k];
                    axpy4(__int_as_float(e.y), ld4(U + e.x * kH + 4 * fg), acc);
                }
                axpy4(nii[i], ld4(U + i * kH + 4 * fg), acc);   // self loop last, as scatter_add sees it
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
    }
    if (store_pending && tid == 0) bulk_store_wait_all();
}

static size_t fwd_fast_smem(int R, int L, int maxEg) {
    const int WB = wb_size(kF0, kH, L);
    return 4 * ((size_t)R * L * kH + (size_t)R * kH + ((WB + 3) & ~3) + 2 * (size_t)maxEg + (size_t)R * kF0 + 2 * (size_t)R + 8 + R + 1) + 16;
}

}  // namespace igcn
