// SGCN encoder, reference shape (F0 = 3 input features, hidden = 16, L = 2 layers, ReLU): second-generation kernels.
//
// ncu on the register-tiled kernels of sgcn_fast.cuh (profiles/r1_ncu_sgcn_kernels_config4.json) showed them ISSUE bound at
// 0.22 (forward) / 0.12-0.18 (backward) of the HBM roofline: 16.9 K warp instructions per 264-node graph, half of them integer /
// address work, ~8 block barriers per graph, inputs staged with 4-byte LDGSTS.  These kernels cut the instruction count ~3.5x:
//
//   * aggregate first:  Y = A_n (H W^T) = (A_n H) W^T.  Layer 1 gathers 3-wide rows instead of 16-wide ones, and every dense
//     transform becomes a [nodes x K] x [K x 16] product with the weights as the constant operand;
//   * those products run on the tensor cores (warp-level mma.sync m16n8k4 / m16n8k8, TF32 split in three: hi*hi + hi*lo + lo*hi,
//     fp32 accumulate -- measured ~1e-6 from fp64): 12 MMAs + 2 ldmatrix replace 128 FFMA + 40 LDS per 16-node tile; the weight
//     fragments (hi and lo) and the biases (as the accumulator's initial value) live in 28 registers for the whole kernel;
//   * a warp OWNS 32 consecutive nodes of the graph through all phases, so the hand-off from the gather (one thread per node or
//     per (node, 4 features)) to the MMA fragment layout goes through a 1.3 KB per-warp tile and __syncwarp, not a block barrier;
//   * inputs of graph g+1 (x slab, rowptr slice, CSR source / weight slices) arrive by FOUR bulk async copies (cp.async.bulk,
//     TMA engine, one elected thread) on an mbarrier ring while graph g computes; outputs leave as 8-byte stores straight from
//     the accumulator fragments (32 B sectors fully written).
//
// Reference semantics: PyG 2.0.2 gcn_norm + GCNConv as called at kernel/sgcn_img_snp.py:218-221.
#pragma once
#include "mma_util.cuh"

namespace igcn {
namespace mma {

constexpr int kMaxThreads = 288;          // forward: 9 warps x 32 nodes: graphs of up to 288 ROIs (the reference uses 90 and 264)
constexpr int kBwdMaxThreads = 576;       // backward: 18 warps x 16 nodes (one CTA per SM: the extra warps are its latency hiding)
constexpr int kTile = 16 * kHP;           // per-warp hand-off tile: 16 rows x 20 floats (80 B rows: conflict-free ldmatrix)
constexpr long long kSpin = 4000000000LL;

using namespace igcn::mmau;

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    const long long t0 = clock64();
    while (true) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok)
                     : "r"(bar), "r"(parity)
                     : "memory");
        if (ok) break;
        if (clock64() - t0 > kSpin) {
            printf("igcn sgcn_mma: input stage never arrived (block %d)\n", blockIdx.x);
            __trap();
        }
    }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}

// A slice [i0, i0+n) of a 4-byte array arrives as its enclosing 16-byte aligned window (bulk copies move whole 16 B units);
// element j of the slice sits at word `shift + j` of the destination.  The window never starts before the array (base pointers
// are 16 B aligned, checked on the host) and is clipped to the last whole 16 B unit of the array; the <= 3 elements behind that
// (possible only for the slice that ends the array) are fetched with plain loads by the issuing thread, before its arrive.
struct Window {
    int shift;        // words between the window start and element i0
    uint32_t bytes;   // bulk-copy size (multiple of 16, may be 0)
    int64_t a0;       // window start, bytes from the array base
    int tail0, tail1; // [tail0, tail1): slice-relative elements to fetch by hand
};
__device__ __forceinline__ Window make_window(int64_t i0, int n, int64_t total) {
    Window w;
    const int64_t b0 = i0 * 4, b1 = (i0 + n) * 4;
    w.a0 = b0 & ~(int64_t)15;
    int64_t a1 = (b1 + 15) & ~(int64_t)15;
    const int64_t lim = (total * 4) & ~(int64_t)15;
    if (a1 > lim) a1 = lim;
    if (a1 < w.a0) a1 = w.a0;
    w.shift = (int)((b0 - w.a0) >> 2);
    w.bytes = (uint32_t)(a1 - w.a0);
    const int64_t covered = (a1 - b0) >> 2;          // slice elements inside the bulk window
    w.tail0 = (int)(covered < 0 ? 0 : (covered > n ? n : covered));
    w.tail1 = n;
    return w;
}
__device__ __forceinline__ void issue_window(const void* base, const Window& w, uint32_t* sdst, uint32_t bar) {
    if (w.bytes) bulk_g2s(smem_u32(sdst), reinterpret_cast<const char*>(base) + w.a0, w.bytes, bar);
}
__device__ __forceinline__ void fetch_tail(const void* base, int64_t i0, const Window& w, uint32_t* sdst) {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(base);
    for (int j = w.tail0; j < w.tail1; ++j) sdst[w.shift + j] = __ldg(src + i0 + j);
}

// Y[i][4fg..] = sum_{k in row i} norm_k U[src_k][4fg..] + n_ii U[i][4fg..]; `ed` is indexed with GLOBAL CSR slots (the caller
// passes the local array shifted by -e0, so the staged rowptr slice is used as it arrived).  GDC top-k graphs have exactly 3
// in-edges per node: when the whole warp agrees (one vote) the row is straight-line code with all loads issued before the FMAs.
__device__ __forceinline__ float4 spmm_row3(const float* U, const int ld, const int2* ed, const int* grp, const float* nii, int i, int fg,
                                            bool valid) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int k = 0, k1 = 0;
    if (valid) {
        k = grp[i];
        k1 = grp[i + 1];
    }
    if (__all_sync(0xffffffffu, !valid || k1 - k == 3)) {
        if (valid) {
            const int2 e0 = ed[k], e1 = ed[k + 1], e2 = ed[k + 2];
            const float4 u0 = ld4(U + e0.x * ld + 4 * fg), u1 = ld4(U + e1.x * ld + 4 * fg), u2 = ld4(U + e2.x * ld + 4 * fg);
            const float4 us = ld4(U + i * ld + 4 * fg);
            axpy4(__int_as_float(e0.y), u0, acc);
            axpy4(__int_as_float(e1.y), u1, acc);
            axpy4(__int_as_float(e2.y), u2, acc);
            axpy4(nii[i], us, acc);
        }
    } else if (valid) {
        for (; k < k1; ++k) {
            const int2 e = ed[k];
            axpy4(__int_as_float(e.y), ld4(U + e.x * ld + 4 * fg), acc);
        }
        axpy4(nii[i], ld4(U + i * ld + 4 * fg), acc);
    }
    return acc;
}

struct StageBuf {
    uint32_t* x;     // R*3 floats
    uint32_t* rp;    // R+1 ints
    uint32_t* src;   // maxEg ints
    uint32_t* w;     // maxEg floats
};
__host__ __device__ inline int win_words(int n) { return ((n + 3) & ~3) + 8; }          // slice + up to 3 lead-in words, rounded

// ---------------------------------------------------------------------------------------------------------------------
// Forward.  Block = ceil(R/32) warps; warp w owns nodes [32w, 32w+32).  Per graph:
//   P0  x~ = x * prob (explain), local rowptr                                                   | barrier
//   P1  thread per node: edge probabilities p_e (-> HBM), masked weights, self-loop merge, degree, d^-1/2, n_ii  | barrier
//   P2  thread per node: normalised edge weights, Z1 = A_n x~ (3 wide) -> warp tile -> H1 = relu(Z1 W1^T + b1) on the tensor
//       cores -> smem copy for the layer-2 gather + concat slot 0 in HBM                         | barrier
//   P3  (node, 4 features): Z2 = A_n H1 -> warp tile -> H2 = relu(Z2 W2^T + b2) on the tensor cores -> concat slot 1 in HBM
// ---------------------------------------------------------------------------------------------------------------------
template <bool kExplain, int kMinBlocks>
__global__ void __launch_bounds__(kMaxThreads, kMinBlocks) sgcn_fwd_mma_kernel(EncArgs a) {
    IGCN_PDL_SYNC();
    extern __shared__ __align__(16) uint32_t smw[];
    const int R = a.R, maxEg = a.maxEg;
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5;
    const int gq = lane >> 2, tq = lane & 3;                    // MMA fragment coordinates: groupID, thread-in-group
    constexpr int LH = 2 * kH;
    // ---- carve (every region a multiple of 16 bytes) ----------------------------------------------------------------
    const int xw = win_words(R * kF0), rw = win_words(R + 1), ew = win_words(maxEg);
    const int stage_words = xw + rw + 2 * ew;
    auto stage = [&](int s) {                       // computed, not stored: an indexed array of pointers would live in local memory
        StageBuf q;
        q.x = smw + s * stage_words;
        q.rp = q.x + xw;
        q.src = q.rp + rw;
        q.w = q.src + ew;
        return q;
    };
    uint32_t* p = smw + 2 * stage_words;
    float* H1s = reinterpret_cast<float*>(p); p += ((R * kHP + 3) & ~3);
    float* tiles = reinterpret_cast<float*>(p); p += (nt >> 5) * kTile;
    int2* edges = reinterpret_cast<int2*>(p); p += 2 * ((maxEg + 1) & ~1);
    float* xs = reinterpret_cast<float*>(p); p += ((R * kF0 + 3) & ~3);
    float* dinv = reinterpret_cast<float*>(p); p += ((R + 3) & ~3);
    float* nii = reinterpret_cast<float*>(p); p += ((R + 3) & ~3);
    uint64_t* bars = reinterpret_cast<uint64_t*>(p);
    const uint32_t bar0 = smem_u32(bars);
    float* tile = tiles + warp * kTile;

    if (tid == 0) {
        mbar_init(bar0, 1);
        mbar_init(bar0 + 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // ---- constant operands: weight fragments (hi / lo), biases as accumulator seeds, prob_bias ----------------------------
    const float* W1 = a.wb;                          // (16, 3)
    const float* B1 = a.wb + kH * kF0;
    const float* W2 = a.wb + layer_off(1, kF0, kH);  // (16, 16)
    const float* B2 = W2 + kH * kH;
    uint32_t w1h[2], w1l[2], w2h[2][2][2], w2l[2][2][2];
    float c1[2][2], c2[2][2];
#pragma unroll
    for (int n = 0; n < 2; ++n) {
        split(tq < kF0 ? __ldg(W1 + (n * 8 + gq) * kF0 + tq) : 0.f, w1h[n], w1l[n]);
        c1[n][0] = __ldg(B1 + n * 8 + 2 * tq);
        c1[n][1] = __ldg(B1 + n * 8 + 2 * tq + 1);
        c2[n][0] = __ldg(B2 + n * 8 + 2 * tq);
        c2[n][1] = __ldg(B2 + n * 8 + 2 * tq + 1);
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
            split(__ldg(W2 + (n * 8 + gq) * kH + ks * 8 + tq), w2h[n][ks][0], w2l[n][ks][0]);
            split(__ldg(W2 + (n * 8 + gq) * kH + ks * 8 + tq + 4), w2h[n][ks][1], w2l[n][ks][1]);
        }
    }
    float pb[6];
#pragma unroll
    for (int c = 0; c < 6; ++c) pb[c] = kExplain ? __ldg(a.prob_bias + c) : 0.f;
    __syncthreads();

    const int64_t N = (int64_t)a.B * R;
    const int64_t Etot = a.rowptr_t[N];
    // stage issue (thread 0 only): e-range of graph gi -> four windows + hand-fetched tails, one expect_tx
    auto issue = [&](int gi, int e0i, int Egi, int s) {
        const int64_t n0 = (int64_t)gi * R;
        const Window wx = make_window(n0 * kF0, R * kF0, N * kF0), wr = make_window(n0, R + 1, N + 1);
        const Window ws = make_window(e0i, Egi, Etot);
        const StageBuf sb = stage(s);
        fetch_tail(a.x, n0 * kF0, wx, sb.x);
        fetch_tail(a.rowptr_t, n0, wr, sb.rp);
        fetch_tail(a.csr_src, e0i, ws, sb.src);
        fetch_tail(a.csr_w, e0i, ws, sb.w);
        const uint32_t bar = bar0 + 8 * s;
        mbar_expect_tx(bar, wx.bytes + wr.bytes + 2 * ws.bytes);
        issue_window(a.x, wx, sb.x, bar);
        issue_window(a.rowptr_t, wr, sb.rp, bar);
        issue_window(a.csr_src, ws, sb.src, bar);
        issue_window(a.csr_w, ws, sb.w, bar);
    };
    int g = blockIdx.x;
    int e_next0 = 0, e_next1 = 0;            // thread 0: edge range of the graph after the one in flight
    if (tid == 0 && g < a.B) {
        const int e0 = a.rowptr_t[(int64_t)g * R], e1 = a.rowptr_t[(int64_t)(g + 1) * R];
        if (e1 - e0 > maxEg) __trap();
        issue(g, e0, e1 - e0, 0);
        const int gn = g + gridDim.x;
        if (gn < a.B) {
            e_next0 = a.rowptr_t[(int64_t)gn * R];
            e_next1 = a.rowptr_t[(int64_t)(gn + 1) * R];
        }
    }
    for (int it = 0; g < a.B; g += gridDim.x, ++it) {
        const int s = it & 1;
        const int gn = g + gridDim.x;
        if (tid == 0 && gn < a.B) {
            // the other stage was last read in P0/P1 of the previous graph, several block barriers ago
            if (e_next1 - e_next0 > maxEg) __trap();
            issue(gn, e_next0, e_next1 - e_next0, s ^ 1);
            const int g2 = gn + gridDim.x;
            if (g2 < a.B) {
                e_next0 = a.rowptr_t[(int64_t)g2 * R];
                e_next1 = a.rowptr_t[(int64_t)(g2 + 1) * R];
            }
        }
        mbar_wait(bar0 + 8 * s, (it >> 1) & 1);
        const int64_t node0 = (int64_t)g * R;
        const int shx = (int)(((node0 * kF0) * 4 & 15) >> 2), shr = (int)((node0 * 4 & 15) >> 2);
        const StageBuf sb = stage(s);
        const int e0 = (int)sb.rp[shr];
        const int she = (int)(((int64_t)e0 * 4 & 15) >> 2);
        const float* sx = reinterpret_cast<const float*>(sb.x) + shx;
        const int* srp = reinterpret_cast<const int*>(sb.rp) + shr;
        const int* ssrc = reinterpret_cast<const int*>(sb.src) + she;
        const float* sw = reinterpret_cast<const float*>(sb.w) + she;
        // ---- P0 (explain only): x~ = x * prob.  The plain pass reads the staged slab and rowptr slice where they arrived ----------
        const float* xsp = sx;
        if (kExplain) {
            for (int j = tid; j < R * kF0; j += nt) xs[j] = sx[j] * __ldg(a.prob + j);
            xsp = xs;
            __syncthreads();
        }
        const int2* edg = edges - e0;                   // indexed with global CSR slots (what the staged rowptr holds)
        // ---- P1: thread per node ----------------------------------------------------------------------------------------------
        const int i = tid;                               // nt >= R (host guarantees)
        const int nd0 = (int)node0;
        int k0 = 0, k1 = 0;
        float xi0 = 0.f, xi1 = 0.f, xi2 = 0.f;
        if (i < R) {
            k0 = srp[i] - e0;
            k1 = srp[i + 1] - e0;
            xi0 = xsp[i * 3]; xi1 = xsp[i * 3 + 1]; xi2 = xsp[i * 3 + 2];
            float deg = 0.f, loopw = 1.f;
            for (int k = k0; k < k1; ++k) {
                const int sl = ssrc[k] - nd0;
                float wt = sw[k];
                if (kExplain) {
                    // same association as the reference's [x_src | x_dst] . prob_bias : pairs (src_c, dst_c) summed in order
                    float z = pb[0] * xsp[sl * 3] + pb[3] * xi0;
                    z += pb[1] * xsp[sl * 3 + 1] + pb[4] * xi1;
                    z += pb[2] * xsp[sl * 3 + 2] + pb[5] * xi2;
                    const float pe = sigmoidf_(z);
                    wt *= pe;
                    if (a.pe_w) a.pe_w[e0 + k] = pe;
                }
                edges[k] = make_int2(sl, __float_as_int(wt));
                if (sl == i)
                    loopw = wt;          // last self loop wins
                else
                    deg += wt;
            }
            deg += loopw;
            const float d = (deg == 0.f) ? 0.f : rsqrtf(deg);
            dinv[i] = d;
            nii[i] = d * d * loopw;
        }
        __syncthreads();
        // ---- P2: normalised weights, Z1 = A_n x~ (thread per node), H1 on the tensor cores ----------------------------------
        {
            float z0 = 0.f, z1 = 0.f, z2 = 0.f;
            if (i < R) {
                const float di = dinv[i];
                for (int k = k0; k < k1; ++k) {
                    const int2 e = edges[k];
                    const float n = (e.x == i) ? 0.f : dinv[e.x] * __int_as_float(e.y) * di;
                    edges[k].y = __float_as_int(n);
                    z0 = fmaf(n, xsp[e.x * 3], z0);
                    z1 = fmaf(n, xsp[e.x * 3 + 1], z1);
                    z2 = fmaf(n, xsp[e.x * 3 + 2], z2);
                }
                const float ns = nii[i];
                z0 = fmaf(ns, xi0, z0);
                z1 = fmaf(ns, xi1, z1);
                z2 = fmaf(ns, xi2, z2);
            }
            st4(tile + lane * 4, make_float4(z0, z1, z2, 0.f));          // rows = the warp's 32 nodes, 4 floats each
            __syncwarp();
            float* og = a.out_w + node0 * LH;
#pragma unroll
            for (int t = 0; t < 2; ++t) {
                const int base = warp * 32 + t * 16;
                if (base < R) {
                    uint32_t ah0, al0, ah1, al1;
                    split(tile[(t * 16 + gq) * 4 + tq], ah0, al0);
                    split(tile[(t * 16 + gq + 8) * 4 + tq], ah1, al1);
                    const int r0 = base + gq, r1 = r0 + 8;
#pragma unroll
                    for (int n = 0; n < 2; ++n) {
                        float c[4] = {c1[n][0], c1[n][1], c1[n][0], c1[n][1]};
                        mma_k4(c, al0, al1, w1h[n]);
                        mma_k4(c, ah0, ah1, w1l[n]);
                        mma_k4(c, ah0, ah1, w1h[n]);
                        const float2 v0 = make_float2(fmaxf(c[0], 0.f), fmaxf(c[1], 0.f));
                        const float2 v1 = make_float2(fmaxf(c[2], 0.f), fmaxf(c[3], 0.f));
                        if (r0 < R) {
                            *reinterpret_cast<float2*>(H1s + r0 * kHP + n * 8 + 2 * tq) = v0;
                            __stcs(reinterpret_cast<float2*>(og + r0 * LH + n * 8 + 2 * tq), v0);
                        }
                        if (r1 < R) {
                            *reinterpret_cast<float2*>(H1s + r1 * kHP + n * 8 + 2 * tq) = v1;
                            __stcs(reinterpret_cast<float2*>(og + r1 * LH + n * 8 + 2 * tq), v1);
                        }
                    }
                }
            }
        }
        __syncthreads();
        // ---- P3: Z2 = A_n H1 per 16-node tile (node, 4 features), H2 on the tensor cores ------------------------------------------
        {
            float* og = a.out_w + node0 * LH + kH;
            const uint32_t taddr = smem_u32(tile) + (uint32_t)(((lane & 7) + 8 * ((lane >> 3) & 1)) * kHP + 4 * (lane >> 4)) * 4u;
#pragma unroll 1
            for (int t = 0; t < 2; ++t) {
                const int base = warp * 32 + t * 16;
                if (base >= R) break;
                __syncwarp();                                  // the previous tile's fragments have been read
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int lr = h * 8 + gq;                 // local row 0..15
                    const int node = base + lr;
                    st4(tile + lr * kHP + 4 * tq, spmm_row3(H1s, kHP, edg, srp, nii, node, tq, node < R));
                }
                __syncwarp();
                uint32_t af[2][4];
                ldmatrix_a(af[0], taddr);
                ldmatrix_a(af[1], taddr + 32);
                uint32_t ah[2][4], al[2][4];
#pragma unroll
                for (int ks = 0; ks < 2; ++ks)
#pragma unroll
                    for (int q = 0; q < 4; ++q) split(__uint_as_float(af[ks][q]), ah[ks][q], al[ks][q]);
                const int r0 = base + gq, r1 = r0 + 8;
#pragma unroll
                for (int n = 0; n < 2; ++n) {
                    float c[4] = {c2[n][0], c2[n][1], c2[n][0], c2[n][1]};
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks) {
                        mma_k8(c, al[ks][0], al[ks][1], al[ks][2], al[ks][3], w2h[n][ks][0], w2h[n][ks][1]);
                        mma_k8(c, ah[ks][0], ah[ks][1], ah[ks][2], ah[ks][3], w2l[n][ks][0], w2l[n][ks][1]);
                        mma_k8(c, ah[ks][0], ah[ks][1], ah[ks][2], ah[ks][3], w2h[n][ks][0], w2h[n][ks][1]);
                    }
                    if (r0 < R) __stcs(reinterpret_cast<float2*>(og + r0 * LH + n * 8 + 2 * tq), make_float2(fmaxf(c[0], 0.f), fmaxf(c[1], 0.f)));
                    if (r1 < R) __stcs(reinterpret_cast<float2*>(og + r1 * LH + n * 8 + 2 * tq), make_float2(fmaxf(c[2], 0.f), fmaxf(c[3], 0.f)));
                }
            }
        }
        __syncthreads();          // edges / xs / H1s are rewritten by the next graph
    }
}

static size_t fwd_mma_smem(int R, int maxEg, int nthreads) {
    size_t words = 2 * ((size_t)win_words(R * kF0) + win_words(R + 1) + 2 * (size_t)win_words(maxEg));
    words += ((size_t)R * kHP + 3) & ~(size_t)3;
    words += (size_t)(nthreads >> 5) * kTile;
    words += 2 * (((size_t)maxEg + 1) & ~(size_t)1);
    words += ((size_t)R * kF0 + 3) & ~(size_t)3;
    words += 2 * (((size_t)R + 3) & ~(size_t)3);
    return 4 * words + 16 + 16;
}

static inline int mma_threads(int R) { return ((R + 31) / 32) * 32; }

// =====================================================================================================================
// Backward (same shape: F0 = 3, H = 16, L = 2, ReLU).  One CTA per SM with ceil(R/16) warps: warp w owns the 16-node tile
// [16w, 16w+16) of the graph in flight in the tensor-core phases; the thread-per-node phases run on threads 0..R-1.
// (ncu on a first version with 32 nodes per warp: 9 warps per SM, issue slots 27 % used, every stall a dependent shared-memory /
// MMA latency -- the kernel needs warps, and shared memory allows only one CTA per SM at 264 ROIs.)
//
//   Y_l = Z_l W_l^T + b_l,  Z_l = A_n H_{l-1}:   dW_l = G_l^T Z_l,  dZ_l = G_l W_l,  dH_{l-1} = A_n^T dZ_l,  d n_e = <dZ_l[t_e], H_{l-1}[s_e]>
//
// Everything dense is a tensor-core product (mma.sync TF32 x3): dZ_2 = G_2 W_2 and dZ_1 = G_1 W_1 with the weights as the
// constant operand (ldmatrix A fragments from the warp's 16 x 20 tile), dW_2 += G_2^T Z_2 and [dW_1 | db_1] += G_1^T [Z_1 | 1]
// with BOTH operands from the warp's tiles and the 16 x 16 / 16 x 4 accumulators living in 12 registers per thread for the whole
// kernel (summed over warps at the end, one partial row per CTA, reduced over CTAs in fp64 by reduce_partials_kernel: no atomics).
// HBM traffic is prefetched by bulk async copies one graph ahead: the forward output tile (33 KB at 264 ROIs; consumed in the
// first phase and refilled at once), the output-gradient tile (consumed by the two layer phases, refilled after them) and a
// double-buffered stage of the small inputs (x, CSR, CSC, g_pe).
// Phases per graph (| = block barrier):
//   A   x~, local index arrays; H1 rows and the ReLU mask of H2 out of the staged forward output                       |
//   P1  thread per node: p_e, masked weights, self-loop merge, degree, d^-1/2, n_ii                                    |
//   P2  thread per node: normalised edge weights, Z1 = A_n x~
//   L2a per 16-node tile: Z2 = A_n H1 and G2 = g_out[:,16:] * relu' -> tiles; dZ2 = G2 W2 -> smem; dW2 += G2^T Z2; db2   |
//   L2b per tile: dH1 = A_n^T dZ2 (+ d n_e, d n_ii); G1 = (g_out[:,:16] + dH1) * relu' -> tile; dZ1 = G1 W1; [dW1|db1]   |
//   X   thread per node: d x~ = A_n^T dZ1 (+ d n_e, d n_ii)                                                            |
//   M1-M3 (explain): gradient through the symmetric normalisation, the sigmoid edge mask and the node mask         | | |
// =====================================================================================================================
struct StageBwd {
    uint32_t *x, *rp, *src, *w, *rps, *cpos, *gpe;
};

template <bool kExplain>
__global__ void __launch_bounds__(kBwdMaxThreads, 1) sgcn_bwd_mma_kernel(EncArgs a) {
    IGCN_PDL_SYNC();
    extern __shared__ __align__(16) uint32_t smw[];
    const int R = a.R, maxEg = a.maxEg;
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarp = nt >> 5;
    const int gq = lane >> 2, tq = lane & 3;
    constexpr int LH = 2 * kH;
    constexpr int WB = kH * kF0 + kH + kH * kH + kH;            // 336: [W1 | b1 | W2 | b2]
    const bool has_gpe = kExplain && a.g_pe != nullptr;
    // ---- carve ---------------------------------------------------------------------------------------------------------
    float* raw_out = reinterpret_cast<float*>(smw);            // R x 32 forward output (bulk destination)
    float* raw_g = raw_out + R * LH;                           // R x 32 output gradient (bulk destination)
    uint32_t* stage_base = smw + 2 * R * LH;
    const int xw = win_words(R * kF0), rw = win_words(R + 1), ew = win_words(maxEg);
    const int stage_words = xw + 2 * rw + (kExplain ? 4 : 3) * ew;
    auto stage = [&](int s) {
        StageBwd q;
        q.x = stage_base + s * stage_words;
        q.rp = q.x + xw;
        q.src = q.rp + rw;
        q.w = q.src + ew;
        q.rps = q.w + ew;
        q.cpos = q.rps + rw;
        q.gpe = q.cpos + ew;
        return q;
    };
    uint32_t* p = stage_base + 2 * stage_words;
    float* H1s = reinterpret_cast<float*>(p); p += ((R * kHP + 3) & ~3);
    float* dZ2s = raw_g + kH;                  // dZ2 rows overwrite the (consumed) layer-2 half of the staged gradient rows: stride LH
    float* tiles = reinterpret_cast<float*>(p); p += nwarp * 2 * kTile;
    int2* edges = reinterpret_cast<int2*>(p); p += 2 * ((maxEg + 1) & ~1);
    int* etgt = reinterpret_cast<int*>(p); p += ((maxEg + 3) & ~3);
    int* spos = reinterpret_cast<int*>(p); p += ((maxEg + 3) & ~3);
    float* ewt = reinterpret_cast<float*>(p); p += kExplain ? ((maxEg + 3) & ~3) : 0;
    float* epe = reinterpret_cast<float*>(p); p += kExplain ? ((maxEg + 3) & ~3) : 0;
    float* edn = reinterpret_cast<float*>(p); p += kExplain ? ((maxEg + 3) & ~3) : 0;
    float* xs = reinterpret_cast<float*>(p); p += ((R * kF0 + 3) & ~3);
    float* z1s = reinterpret_cast<float*>(p); p += nwarp * 16 * 4;
    float* dZ1s = reinterpret_cast<float*>(p); p += nwarp * 16 * 4;
    float* dxt = reinterpret_cast<float*>(p); p += kExplain ? ((R * kF0 + 3) & ~3) : 0;
    float* dinv = reinterpret_cast<float*>(p); p += ((R + 3) & ~3);
    float* nii = reinterpret_cast<float*>(p); p += ((R + 3) & ~3);
    float* ell = reinterpret_cast<float*>(p); p += kExplain ? ((R + 3) & ~3) : 0;
    float* dnii = reinterpret_cast<float*>(p); p += kExplain ? ((R + 3) & ~3) : 0;
    float* ddeg = reinterpret_cast<float*>(p); p += kExplain ? ((R + 3) & ~3) : 0;
    float* dprob = reinterpret_cast<float*>(p); p += kExplain ? ((R * kF0 + 3) & ~3) : 0;
    uint8_t* m2 = reinterpret_cast<uint8_t*>(p); p += ((R + 3) & ~3);          // 4 bytes per node: relu' bits of H2, one byte per quad
    int* rp = reinterpret_cast<int*>(p); p += ((R + 1 + 3) & ~3);
    int* rps = reinterpret_cast<int*>(p); p += ((R + 1 + 3) & ~3);
    uint64_t* bars = reinterpret_cast<uint64_t*>(p);
    const uint32_t bar_small = smem_u32(bars), bar_out = bar_small + 16, bar_g = bar_small + 24;
    float* Gt = tiles + warp * 2 * kTile;      // G_l tile of the warp (16 x 20)
    float* Zt = Gt + kTile;                    // Z_2 tile

    if (tid == 0) {
        mbar_init(bar_small, 1);
        mbar_init(bar_small + 8, 1);
        mbar_init(bar_out, 1);
        mbar_init(bar_g, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // ---- constant operands --------------------------------------------------------------------------------------------------
    const float* W1 = a.wb;
    const float* W2 = a.wb + layer_off(1, kF0, kH);
    uint32_t w2h[2][2][2], w2l[2][2][2];       // dZ2 = G2 W2 : B[k = f_out][n = f_in] = W2[k][n]
    uint32_t w1h[2][2], w1l[2][2];             // dZ1 = G1 W1 : B[k = f_out][n = c]    = W1[k][c], c < 3
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
#pragma unroll
        for (int n = 0; n < 2; ++n) {
            split(__ldg(W2 + (ks * 8 + tq) * kH + n * 8 + gq), w2h[n][ks][0], w2l[n][ks][0]);
            split(__ldg(W2 + (ks * 8 + tq + 4) * kH + n * 8 + gq), w2h[n][ks][1], w2l[n][ks][1]);
        }
        split(gq < kF0 ? __ldg(W1 + (ks * 8 + tq) * kF0 + gq) : 0.f, w1h[ks][0], w1l[ks][0]);
        split(gq < kF0 ? __ldg(W1 + (ks * 8 + tq + 4) * kF0 + gq) : 0.f, w1h[ks][1], w1l[ks][1]);
    }
    float pb[6];
#pragma unroll
    for (int c = 0; c < 6; ++c) pb[c] = kExplain ? __ldg(a.prob_bias + c) : 0.f;
    // gradient accumulators (registers, across all graphs of this CTA)
    float accW2[2][4], accW1[4], db2_lo = 0.f, db2_hi = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) accW2[0][q] = accW2[1][q] = accW1[q] = 0.f;
    float dpb_reg[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (kExplain)
        for (int i = tid; i < R * kF0; i += nt) dprob[i] = 0.f;
    __syncthreads();

    const int64_t N = (int64_t)a.B * R;
    const int64_t Etot = a.rowptr_t[N];
    const uint32_t tile_bytes = (uint32_t)(R * LH * 4);
    auto issue_small = [&](int gi, int e0i, int Egi, int s) {
        const int64_t n0 = (int64_t)gi * R;
        const Window wx = make_window(n0 * kF0, R * kF0, N * kF0), wr = make_window(n0, R + 1, N + 1);
        const Window ws = make_window(e0i, Egi, Etot);
        const StageBwd sb = stage(s);
        fetch_tail(a.x, n0 * kF0, wx, sb.x);
        fetch_tail(a.rowptr_t, n0, wr, sb.rp);
        fetch_tail(a.rowptr_s, n0, wr, sb.rps);
        fetch_tail(a.csr_src, e0i, ws, sb.src);
        fetch_tail(a.csr_w, e0i, ws, sb.w);
        fetch_tail(a.csc_pos, e0i, ws, sb.cpos);
        if (has_gpe) fetch_tail(a.g_pe, e0i, ws, sb.gpe);
        const uint32_t bar = bar_small + 8 * s;
        mbar_expect_tx(bar, wx.bytes + 2 * wr.bytes + (has_gpe ? 4 : 3) * ws.bytes);
        issue_window(a.x, wx, sb.x, bar);
        issue_window(a.rowptr_t, wr, sb.rp, bar);
        issue_window(a.rowptr_s, wr, sb.rps, bar);
        issue_window(a.csr_src, ws, sb.src, bar);
        issue_window(a.csr_w, ws, sb.w, bar);
        issue_window(a.csc_pos, ws, sb.cpos, bar);
        if (has_gpe) issue_window(a.g_pe, ws, sb.gpe, bar);
    };
    auto issue_tile = [&](const float* src, float* dst, uint32_t bar) {
        // the tile buffers were read -- and the layer-2 half of the gradient tile WRITTEN (dZ2 in place) -- through the generic
        // proxy; the block barrier in front of every call ordered those accesses among the threads, this fence orders them
        // before the async-proxy write of the refill (without it a late generic store can land on top of the new tile)
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(bar, tile_bytes);
        bulk_g2s(smem_u32(dst), src, tile_bytes, bar);
    };
    int g = blockIdx.x;
    int e_next0 = 0, e_next1 = 0;
    if (tid == 0 && g < a.B) {
        const int e0 = a.rowptr_t[(int64_t)g * R], e1 = a.rowptr_t[(int64_t)(g + 1) * R];
        if (e1 - e0 > maxEg) __trap();
        issue_small(g, e0, e1 - e0, 0);
        issue_tile(a.out + (int64_t)g * R * LH, raw_out, bar_out);
        issue_tile(a.g_out + (int64_t)g * R * LH, raw_g, bar_g);
        const int gn = g + gridDim.x;
        if (gn < a.B) {
            e_next0 = a.rowptr_t[(int64_t)gn * R];
            e_next1 = a.rowptr_t[(int64_t)(gn + 1) * R];
        }
    }
    const uint32_t gt_addr = smem_u32(Gt) + (uint32_t)(((lane & 7) + 8 * ((lane >> 3) & 1)) * kHP + 4 * (lane >> 4)) * 4u;
    for (int it = 0; g < a.B; g += gridDim.x, ++it) {
        const int s = it & 1;
        const int gn = g + gridDim.x;
        if (tid == 0 && gn < a.B) {
            if (e_next1 - e_next0 > maxEg) __trap();
            issue_small(gn, e_next0, e_next1 - e_next0, s ^ 1);
            const int g2 = gn + gridDim.x;
            if (g2 < a.B) {
                e_next0 = a.rowptr_t[(int64_t)g2 * R];
                e_next1 = a.rowptr_t[(int64_t)(g2 + 1) * R];
            }
        }
        mbar_wait(bar_small + 8 * s, (it >> 1) & 1);
        const int64_t node0 = (int64_t)g * R;
        const int nd0 = (int)node0;
        const int shx = (int)(((node0 * kF0) * 4 & 15) >> 2), shr = (int)((node0 * 4 & 15) >> 2);
        const StageBwd sb = stage(s);
        const int e0 = (int)sb.rp[shr];
        const int she = (int)(((int64_t)e0 * 4 & 15) >> 2);
        const float* sx = reinterpret_cast<const float*>(sb.x) + shx;
        const int* srp = reinterpret_cast<const int*>(sb.rp) + shr;
        const int* srps = reinterpret_cast<const int*>(sb.rps) + shr;
        const int* ssrc = reinterpret_cast<const int*>(sb.src) + she;
        const float* sw = reinterpret_cast<const float*>(sb.w) + she;
        const int* scpos = reinterpret_cast<const int*>(sb.cpos) + she;
        const float* sgpe = reinterpret_cast<const float*>(sb.gpe) + she;
        const int Eg = srp[R] - e0;
        // ---- A: x~, local index arrays; H1 rows + relu' bits of H2 out of the staged forward output ---------------------------
        for (int i = tid; i < R * kF0; i += nt) xs[i] = kExplain ? sx[i] * __ldg(a.prob + i) : sx[i];
        for (int i = tid; i <= R; i += nt) {
            rp[i] = srp[i] - e0;
            rps[i] = srps[i] - e0;
        }
        for (int q = tid; q < Eg; q += nt) spos[q] = scpos[q] - e0;
        mbar_wait(bar_out, it & 1);
        {
            // the warp's 16 rows x 8 sixteen-byte chunks in memory order: conflict-free LDS.128 on the 128 B rows
#pragma unroll
            for (int ps = 0; ps < 4; ++ps) {
                const int row = warp * 16 + ps * 4 + (lane >> 3), ch = lane & 7;
                if (row < R) {
                    const float4 v = ld4(raw_out + row * LH + 4 * ch);
                    if (ch < 4) {
                        st4(H1s + row * kHP + 4 * ch, v);
                    } else {
                        m2[row * 4 + (ch - 4)] = (uint8_t)((v.x > 0.f ? 1 : 0) | (v.y > 0.f ? 2 : 0) | (v.z > 0.f ? 4 : 0) | (v.w > 0.f ? 8 : 0));
                    }
                }
            }
        }
        __syncthreads();
        if (tid == 0 && gn < a.B) issue_tile(a.out + (int64_t)gn * R * LH, raw_out, bar_out);      // consumed: refill for the next graph
        // ---- P1: thread per node ----------------------------------------------------------------------------------------------
        const int i = tid;
        int k0 = 0, k1 = 0;
        float xi0 = 0.f, xi1 = 0.f, xi2 = 0.f;
        if (i < R) {
            k0 = rp[i];
            k1 = rp[i + 1];
            xi0 = xs[i * 3]; xi1 = xs[i * 3 + 1]; xi2 = xs[i * 3 + 2];
            float deg = 0.f, loopw = 1.f;
            for (int k = k0; k < k1; ++k) {
                const int sl = ssrc[k] - nd0;
                float wt = sw[k];
                if (kExplain) {
                    float z = pb[0] * xs[sl * 3] + pb[3] * xi0;
                    z += pb[1] * xs[sl * 3 + 1] + pb[4] * xi1;
                    z += pb[2] * xs[sl * 3 + 2] + pb[5] * xi2;
                    const float pe = sigmoidf_(z);
                    wt *= pe;
                    epe[k] = pe;
                    ewt[k] = wt;
                    edn[k] = 0.f;
                }
                edges[k] = make_int2(sl, __float_as_int(wt));
                etgt[k] = i;
                if (sl == i)
                    loopw = wt;
                else
                    deg += wt;
            }
            deg += loopw;
            const float d = (deg == 0.f) ? 0.f : rsqrtf(deg);
            dinv[i] = d;
            nii[i] = d * d * loopw;
            if (kExplain) {
                ell[i] = loopw;
                dnii[i] = 0.f;
            }
        }
        __syncthreads();
        // ---- P2: normalised weights, Z1 = A_n x~ (the B operand of dW1; column 3 = 1 gives db1) ---------------------------------
        {
            float z0 = 0.f, z1 = 0.f, z2 = 0.f;
            if (i < R) {
                const float di = dinv[i];
                for (int k = k0; k < k1; ++k) {
                    const int2 e = edges[k];
                    const float n = (e.x == i) ? 0.f : dinv[e.x] * __int_as_float(e.y) * di;
                    edges[k].y = __float_as_int(n);
                    z0 = fmaf(n, xs[e.x * 3], z0);
                    z1 = fmaf(n, xs[e.x * 3 + 1], z1);
                    z2 = fmaf(n, xs[e.x * 3 + 2], z2);
                }
                const float ns = nii[i];
                z0 = fmaf(ns, xi0, z0);
                z1 = fmaf(ns, xi1, z1);
                z2 = fmaf(ns, xi2, z2);
            }
            if (i < nwarp * 16) st4(z1s + i * 4, make_float4(z0, z1, z2, i < R ? 1.f : 0.f));
        }
        __syncthreads();
        // ---- L2a: per 16-node tile of the warp -------------------------------------------------------------------------------------
        mbar_wait(bar_g, it & 1);
        const int base = warp * 16;                 // the warp's tile (base < R: the block has ceil(R/16) warps)
        {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int lr = h * 8 + gq, node = base + lr;
                const float4 z = spmm_row3(H1s, kHP, edges, rp, nii, node, tq, node < R);
                float4 gv = make_float4(0.f, 0.f, 0.f, 0.f);
                if (node < R) {
                    const float4 gr = ld4(raw_g + node * LH + kH + 4 * tq);
                    const unsigned mb = m2[node * 4 + tq];
                    gv.x = (mb & 1) ? gr.x : 0.f;
                    gv.y = (mb & 2) ? gr.y : 0.f;
                    gv.z = (mb & 4) ? gr.z : 0.f;
                    gv.w = (mb & 8) ? gr.w : 0.f;
                }
                st4(Zt + lr * kHP + 4 * tq, z);
                st4(Gt + lr * kHP + 4 * tq, gv);
            }
            __syncwarp();
            // dZ2 = G2 W2
            {
                uint32_t af[2][4], ah[2][4], al[2][4];
                ldmatrix_a(af[0], gt_addr);
                ldmatrix_a(af[1], gt_addr + 32);
#pragma unroll
                for (int ks = 0; ks < 2; ++ks)
#pragma unroll
                    for (int q = 0; q < 4; ++q) split(__uint_as_float(af[ks][q]), ah[ks][q], al[ks][q]);
                const int r0 = base + gq, r1 = r0 + 8;
#pragma unroll
                for (int n = 0; n < 2; ++n) {
                    float c[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks) {
                        mma_k8(c, al[ks][0], al[ks][1], al[ks][2], al[ks][3], w2h[n][ks][0], w2h[n][ks][1]);
                        mma_k8(c, ah[ks][0], ah[ks][1], ah[ks][2], ah[ks][3], w2l[n][ks][0], w2l[n][ks][1]);
                        mma_k8(c, ah[ks][0], ah[ks][1], ah[ks][2], ah[ks][3], w2h[n][ks][0], w2h[n][ks][1]);
                    }
                    if (r0 < R) *reinterpret_cast<float2*>(dZ2s + r0 * LH + n * 8 + 2 * tq) = make_float2(c[0], c[1]);
                    if (r1 < R) *reinterpret_cast<float2*>(dZ2s + r1 * LH + n * 8 + 2 * tq) = make_float2(c[2], c[3]);
                }
            }
            // dW2 += G2^T Z2 (A = G2^T: rows f_out, k = the tile's nodes), db2 += column sums of G2
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
                const float* ga = Gt + (ks * 8 + tq) * kHP + gq;
                const float a0 = ga[0], a1 = ga[8], a2 = ga[4 * kHP], a3 = ga[4 * kHP + 8];
                db2_lo += a0 + a2;
                db2_hi += a1 + a3;
                uint32_t ah[4], al[4];
                split(a0, ah[0], al[0]); split(a1, ah[1], al[1]); split(a2, ah[2], al[2]); split(a3, ah[3], al[3]);
#pragma unroll
                for (int n = 0; n < 2; ++n) {
                    const float* zb = Zt + (ks * 8 + tq) * kHP + n * 8 + gq;
                    uint32_t bh0, bl0, bh1, bl1;
                    split(zb[0], bh0, bl0);
                    split(zb[4 * kHP], bh1, bl1);
                    mma_k8(accW2[n], al[0], al[1], al[2], al[3], bh0, bh1);
                    mma_k8(accW2[n], ah[0], ah[1], ah[2], ah[3], bl0, bl1);
                    mma_k8(accW2[n], ah[0], ah[1], ah[2], ah[3], bh0, bh1);
                }
            }
        }
        __syncthreads();
        // ---- L2b + layer 1, per tile ------------------------------------------------------------------------------------------------
        {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int lr = h * 8 + gq, node = base + lr;
                float4 g1 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (node < R) {
                    const float4 dzi = ld4(dZ2s + node * LH + 4 * tq);
                    float4 dh = make_float4(0.f, 0.f, 0.f, 0.f);
                    const int q1 = rps[node + 1];
                    for (int q = rps[node]; q < q1; ++q) {
                        const int k = spos[q];
                        axpy4(__int_as_float(edges[k].y), ld4(dZ2s + etgt[k] * LH + 4 * tq), dh);
                    }
                    axpy4(nii[node], dzi, dh);
                    const float4 h1 = ld4(H1s + node * kHP + 4 * tq);
                    if (kExplain) {
                        const int kk1 = rp[node + 1];
                        for (int k = rp[node]; k < kk1; ++k) {
                            const float pr = quad_sum(dot4(dzi, ld4(H1s + edges[k].x * kHP + 4 * tq), 0.f));
                            if (tq == 0) edn[k] = pr;
                        }
                        const float pr = quad_sum(dot4(dzi, h1, 0.f));
                        if (tq == 0) dnii[node] = pr;
                    }
                    const float4 gr = ld4(raw_g + node * LH + 4 * tq);
                    g1.x = h1.x > 0.f ? gr.x + dh.x : 0.f;
                    g1.y = h1.y > 0.f ? gr.y + dh.y : 0.f;
                    g1.z = h1.z > 0.f ? gr.z + dh.z : 0.f;
                    g1.w = h1.w > 0.f ? gr.w + dh.w : 0.f;
                }
                st4(Gt + lr * kHP + 4 * tq, g1);
            }
            __syncwarp();
            // dZ1 = G1 W1 (3 valid columns)
            {
                uint32_t af[2][4], ah[2][4], al[2][4];
                ldmatrix_a(af[0], gt_addr);
                ldmatrix_a(af[1], gt_addr + 32);
#pragma unroll
                for (int ks = 0; ks < 2; ++ks)
#pragma unroll
                    for (int q = 0; q < 4; ++q) split(__uint_as_float(af[ks][q]), ah[ks][q], al[ks][q]);
                float c[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) {
                    mma_k8(c, al[ks][0], al[ks][1], al[ks][2], al[ks][3], w1h[ks][0], w1h[ks][1]);
                    mma_k8(c, ah[ks][0], ah[ks][1], ah[ks][2], ah[ks][3], w1l[ks][0], w1l[ks][1]);
                    mma_k8(c, ah[ks][0], ah[ks][1], ah[ks][2], ah[ks][3], w1h[ks][0], w1h[ks][1]);
                }
                if (tq < 2) {       // columns 2tq, 2tq+1 of the 8-wide tile: 0..3 hold (dZ1_0, dZ1_1, dZ1_2, 0)
                    *reinterpret_cast<float2*>(dZ1s + (base + gq) * 4 + 2 * tq) = make_float2(c[0], c[1]);
                    *reinterpret_cast<float2*>(dZ1s + (base + gq + 8) * 4 + 2 * tq) = make_float2(c[2], c[3]);
                }
            }
            // [dW1 | db1] += G1^T [Z1 | 1]
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
                const float* ga = Gt + (ks * 8 + tq) * kHP + gq;
                uint32_t ah[4], al[4];
                split(ga[0], ah[0], al[0]); split(ga[8], ah[1], al[1]); split(ga[4 * kHP], ah[2], al[2]); split(ga[4 * kHP + 8], ah[3], al[3]);
                uint32_t bh0, bl0, bh1, bl1;
                split(gq < 4 ? z1s[(base + ks * 8 + tq) * 4 + gq] : 0.f, bh0, bl0);
                split(gq < 4 ? z1s[(base + ks * 8 + tq + 4) * 4 + gq] : 0.f, bh1, bl1);
                mma_k8(accW1, al[0], al[1], al[2], al[3], bh0, bh1);
                mma_k8(accW1, ah[0], ah[1], ah[2], ah[3], bl0, bl1);
                mma_k8(accW1, ah[0], ah[1], ah[2], ah[3], bh0, bh1);
            }
        }
        __syncthreads();
        if (tid == 0 && gn < a.B) issue_tile(a.g_out + (int64_t)gn * R * LH, raw_g, bar_g);          // consumed: refill
        // ---- X: d x~ = A_n^T dZ1 (+ d n_e, d n_ii), thread per node -------------------------------------------------------------
        float* dxg = a.dx + node0 * kF0;
        if (i < R) {
            const float d0 = dZ1s[i * 4], d1 = dZ1s[i * 4 + 1], d2 = dZ1s[i * 4 + 2];
            float s0 = 0.f, s1 = 0.f, s2 = 0.f;
            const int q1 = rps[i + 1];
            for (int q = rps[i]; q < q1; ++q) {
                const int k = spos[q];
                const float n = __int_as_float(edges[k].y);
                const int tg = etgt[k];
                s0 = fmaf(n, dZ1s[tg * 4], s0);
                s1 = fmaf(n, dZ1s[tg * 4 + 1], s1);
                s2 = fmaf(n, dZ1s[tg * 4 + 2], s2);
            }
            const float ns = nii[i];
            s0 = fmaf(ns, d0, s0);
            s1 = fmaf(ns, d1, s1);
            s2 = fmaf(ns, d2, s2);
            if (kExplain) {
                dxt[i * 3] = s0; dxt[i * 3 + 1] = s1; dxt[i * 3 + 2] = s2;
                for (int k = k0; k < k1; ++k) {
                    const int sl = edges[k].x;
                    edn[k] += d0 * xs[sl * 3] + d1 * xs[sl * 3 + 1] + d2 * xs[sl * 3 + 2];
                }
                dnii[i] += d0 * xi0 + d1 * xi1 + d2 * xi2;
            } else {
                dxg[i * 3] = s0; dxg[i * 3 + 1] = s1; dxg[i * 3 + 2] = s2;
            }
        }
        __syncthreads();
        if (kExplain) {
            // ---- M1: gradient through the symmetric normalisation -> d deg ----------------------------------------------------------
            if (i < R) {
                float dd = 0.f;
                for (int k = k0; k < k1; ++k) {
                    const int sl = edges[k].x;
                    if (sl != i) dd = fmaf(edn[k] * ewt[k], dinv[sl], dd);
                }
                const int q1 = rps[i + 1];
                for (int q = rps[i]; q < q1; ++q) {
                    const int k = spos[q];
                    const int tg = etgt[k];
                    if (tg != i) dd = fmaf(edn[k] * ewt[k], dinv[tg], dd);
                }
                const float di = dinv[i];
                dd = fmaf(2.f * di * ell[i], dnii[i], dd);
                ddeg[i] = -0.5f * di * di * di * dd;
            }
            __syncthreads();
            // ---- M2: d w~ -> d p_e -> d z_e ; d prob_bias ; target half of d x~ ------------------------------------------------------
            if (i < R) {
                const float di = dinv[i];
                float sdz = 0.f;
                for (int k = k0; k < k1; ++k) {
                    const int sl = edges[k].x;
                    const float dwt = (sl != i) ? dinv[sl] * di * edn[k] + ddeg[i] : di * di * dnii[i] + ddeg[i];
                    float dp = sw[k] * dwt;
                    if (has_gpe) dp += sgpe[k];
                    const float pe = epe[k];
                    const float dz = pe * (1.f - pe) * dp;
                    edn[k] = dz;
                    dpb_reg[0] = fmaf(dz, xs[sl * 3], dpb_reg[0]);
                    dpb_reg[1] = fmaf(dz, xs[sl * 3 + 1], dpb_reg[1]);
                    dpb_reg[2] = fmaf(dz, xs[sl * 3 + 2], dpb_reg[2]);
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
            // ---- M3: source half, node mask, outputs -----------------------------------------------------------------------------------
            if (i < R) {
                float sdz = 0.f;
                const int q1 = rps[i + 1];
                for (int q = rps[i]; q < q1; ++q) sdz += edn[spos[q]];
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const float dv = fmaf(sdz, pb[c], dxt[i * 3 + c]);
                    dprob[i * 3 + c] += sx[i * 3 + c] * dv;
                    dxg[i * 3 + c] = __ldg(a.prob + i * 3 + c) * dv;
                }
            }
            __syncthreads();
        }
    }
    // ---- parameter gradients: warps -> CTA -> one partial row per CTA -------------------------------------------------------------
    __syncthreads();
    float* red = raw_out;                       // nwarp x WB scratch (the staged tiles are dead)
    {
        float* rwp = red + warp * WB;
        const int off1 = kH * kF0 + kH;         // W2 offset inside wb
        for (int j = lane; j < WB; j += 32) rwp[j] = 0.f;
        __syncwarp();
#pragma unroll
        for (int n = 0; n < 2; ++n) {
            rwp[off1 + gq * kH + n * 8 + 2 * tq] = accW2[n][0];
            rwp[off1 + gq * kH + n * 8 + 2 * tq + 1] = accW2[n][1];
            rwp[off1 + (gq + 8) * kH + n * 8 + 2 * tq] = accW2[n][2];
            rwp[off1 + (gq + 8) * kH + n * 8 + 2 * tq + 1] = accW2[n][3];
        }
        // accW1 columns 2tq, 2tq+1: 0..2 -> dW1[f][c], 3 -> db1[f]
        if (tq == 0) {
            rwp[gq * kF0] = accW1[0];
            rwp[gq * kF0 + 1] = accW1[1];
            rwp[(gq + 8) * kF0] = accW1[2];
            rwp[(gq + 8) * kF0 + 1] = accW1[3];
        } else if (tq == 1) {
            rwp[gq * kF0 + 2] = accW1[0];
            rwp[kH * kF0 + gq] = accW1[1];
            rwp[(gq + 8) * kF0 + 2] = accW1[2];
            rwp[kH * kF0 + gq + 8] = accW1[3];
        }
        float lo = db2_lo, hi = db2_hi;
        lo += __shfl_xor_sync(0xffffffffu, lo, 1);
        lo += __shfl_xor_sync(0xffffffffu, lo, 2);
        hi += __shfl_xor_sync(0xffffffffu, hi, 1);
        hi += __shfl_xor_sync(0xffffffffu, hi, 2);
        if (tq == 0) {
            rwp[off1 + kH * kH + gq] = lo;
            rwp[off1 + kH * kH + gq + 8] = hi;
        }
    }
    __syncthreads();
    float* prow = a.partials + (int64_t)blockIdx.x * a.P;
    for (int j = tid; j < WB; j += nt) {
        float sacc = 0.f;
        for (int w = 0; w < nwarp; ++w) sacc += red[w * WB + j];
        prow[j] = sacc;
    }
    if (kExplain) {
        for (int j = tid; j < R * kF0; j += nt) prow[WB + j] = dprob[j];
        float* r2 = red + nwarp * WB;
#pragma unroll
        for (int c = 0; c < 6; ++c) {
            const float v = warp_sum(dpb_reg[c]);
            if (lane == 0) r2[warp * 8 + c] = v;
        }
        __syncthreads();
        if (tid < 6) {
            float sacc = 0.f;
            for (int w = 0; w < nwarp; ++w) sacc += r2[w * 8 + tid];
            prow[WB + R * kF0 + tid] = sacc;
        }
    } else {
        for (int j = tid; j < R * kF0 + 6; j += nt) prow[WB + j] = 0.f;
    }
}

static size_t bwd_mma_smem(int R, int maxEg, int nthreads, bool explain) {
    auto r4 = [](size_t n) { return (n + 3) & ~(size_t)3; };
    size_t w = 2 * (size_t)R * 32;
    w += 2 * ((size_t)win_words(R * kF0) + 2 * win_words(R + 1) + (explain ? 4 : 3) * (size_t)win_words(maxEg));
    w += r4((size_t)R * kHP);
    w += (size_t)(nthreads >> 5) * 2 * kTile;
    w += 2 * (((size_t)maxEg + 1) & ~(size_t)1);
    w += 2 * r4(maxEg);
    if (explain) w += 3 * r4(maxEg);
    w += r4((size_t)R * kF0);
    w += 2 * (size_t)(nthreads >> 5) * 16 * 4;
    if (explain) w += r4((size_t)R * kF0);
    w += 2 * r4(R);
    if (explain) w += 3 * r4(R) + r4((size_t)R * kF0);
    w += r4(R);
    w += 2 * r4((size_t)R + 1);
    return 4 * w + 32 + 16;
}

static inline int mma_bwd_threads(int R) { return ((R + 15) / 16) * 32; }

}  // namespace mma
}  // namespace igcn
