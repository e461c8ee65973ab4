// Cross attention on the tensor cores for the reference's shape (E = 32 embed, 2 heads of 16, M <= 32 GO tokens per subject).
//
// Same algebra as cross_attn_rows.cuh -- the query-side projections fold into per-graph key / value tables,
//     S_h = X K'_h^T + c_h        Y = sum_h softmax(S_h) V'_h + bo        K'_h, V'_h : (M x 32)
// -- but the two products per head run as warp-level MMAs (mma.sync m16n8k8, TF32 split in three, fp32 accumulate) on 16-row tiles
// of the query rows instead of one thread per row on FFMA: ncu on the row kernels showed FMA pipe 15 %, issue slots 26 %, the
// shared-memory pipe at 54 % -- a broadcast LDS.128 per 4 FFMA was the ceiling (profiles/r1_ncu_lines_attn_rows_bwd_config4.txt).
// Here a 16 x 32 tile of X is 4 ldmatrix, the softmax runs on the accumulator fragments (a row lives in 4 lanes: 2 shuffles per
// reduction) and P feeds the second product straight from those registers: the accumulator layout (row g, columns 2t, 2t+1) is
// turned into an A fragment by PERMUTING the contraction index -- k slot t takes column 2t, k slot t+4 takes column 2t+1 -- and
// V' is read with the same permutation, so nothing is shuffled or staged.
// Reference: nn.MultiheadAttention(E, 2, batch_first=True)(q, kv, kv) as called at kernel/sgcn_img_snp.py:46,239-242.
#pragma once
#include "mma_util.cuh"

namespace igcn {
namespace amma {

using namespace igcn::mmau;
constexpr int kE = 32;
constexpr int TS = 36;        // row stride (floats) of every staged table / tile: 16 B aligned rows, conflict-free ldmatrix and fragment loads
constexpr int kThreads = 288;

struct Geo {
    int gpc;        // graphs per CTA pass
    int MP;         // tokens padded to a multiple of 8 (8, 16, 24 or 32)
    int per_sz;     // floats of per-graph tables
    int rows_pad;   // staged query rows per pass (graphs x R, + 16 so the last tile of the last graph stays inside the buffer)
    size_t smem;
};

// per-graph tables (floats): A (M x 32) | K (M x 32) | V (M x 32) | K' (H x MP x TS) | V' (H x MP x TS) | c (H x MP)
__host__ __device__ inline int fwd_per_sz(int M, int MP, int H) { return 3 * M * kE + 2 * H * MP * TS + H * MP; }

static Geo fwd_geo(int R, int M, int H) {
    Geo g;
    g.MP = (M + 7) & ~7;
    g.per_sz = (fwd_per_sz(M, g.MP, H) + 3) & ~3;
    int gpc = 288 / R;
    if (gpc < 1) gpc = 1;
    if (gpc > 8) gpc = 8;
    auto smem_of = [&](int n) { return (size_t)4 * (4 * kE * kE + 4 * kE + (size_t)n * g.per_sz + ((size_t)n * R + 16) * TS) + 16; };
    while (gpc > 1 && smem_of(gpc) > 100 * 1024) --gpc;
    g.gpc = gpc;
    g.rows_pad = gpc * R + 16;
    g.smem = smem_of(gpc);
    return g;
}

// K, V, K', V', c of the local graphs [0, ng); rows j >= M of K', V' and c are zero (padded tokens are masked in the softmax)
__device__ __forceinline__ void graph_tables(const AttnArgs& a, int b0, int ng, float* per, int per_sz, int MP, const float* WkvT,
                                             const float* Wq, const float* WoT, const float* bin) {
    const int tid = threadIdx.x, nt = blockDim.x, M = a.M, H = a.heads, hd = kE / H;
    const float scale = rsqrtf((float)hd);
    const int oKp = 3 * M * kE, oVp = oKp + H * MP * TS, oC = oVp + H * MP * TS;
    for (int i = tid; i < ng * M * kE; i += nt) {
        const int gl = i / (M * kE), r = i - gl * M * kE;
        per[gl * per_sz + r] = a.a[((int64_t)(b0 + gl) * M) * kE + r];
    }
    __syncthreads();
    // K = A Wk^T + bk, V = A Wv^T + bv : thread = (graph, token j, 4 output features of [K | V])
    for (int idx = tid; idx < ng * M * 16; idx += nt) {
        const int gl = idx / (M * 16), r = idx - gl * M * 16, j = r >> 4, c = r & 15;     // c < 8: K quad, else V quad
        const float* arow = per + gl * per_sz + j * kE;
        float4 acc = ld4s(bin + kE + 4 * c);
#pragma unroll 8
        for (int k = 0; k < kE; ++k) fma4(arow[k], ld4s(WkvT + k * 2 * kE + 4 * c), acc);
        float* dstp = per + gl * per_sz + M * kE + (c < 8 ? j * kE + 4 * c : M * kE + j * kE + 4 * (c - 8));
        st4s(dstp, acc);
    }
    __syncthreads();
    // K'_h[j][e] = scale sum_d K[j][h hd + d] Wq[h hd + d][e] ; V'_h[j][f] = sum_d V[j][h hd + d] WoT[h hd + d][f]
    for (int idx = tid; idx < ng * H * MP * 16; idx += nt) {
        const int gl = idx / (H * MP * 16), r = idx - gl * H * MP * 16, hj = r >> 4, c = r & 15, h = hj / MP, j = hj - h * MP;
        const float* base = per + gl * per_sz;
        const bool isK = c < 8;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        if (j < M) {
            const float* srow = base + M * kE + (isK ? 0 : M * kE) + j * kE + h * hd;
            const float* wmat = (isK ? Wq : WoT) + h * hd * kE + 4 * (c & 7);
            for (int d = 0; d < hd; ++d) fma4(srow[d], ld4s(wmat + d * kE), acc);
            if (isK) {
                acc.x *= scale; acc.y *= scale; acc.z *= scale; acc.w *= scale;
            }
        }
        st4s(per + gl * per_sz + (isK ? oKp : oVp) + hj * TS + 4 * (c & 7), acc);
    }
    for (int idx = tid; idx < ng * H * MP; idx += nt) {
        const int gl = idx / (H * MP), hj = idx - gl * H * MP, h = hj / MP, j = hj - h * MP;
        float v = 0.f;
        if (j < M) {
            const float* krow = per + gl * per_sz + M * kE + j * kE + h * hd;
            for (int d = 0; d < hd; ++d) v = fmaf(krow[d], bin[h * hd + d], v);
        }
        per[gl * per_sz + oC + hj] = v * scale;
    }
    __syncthreads();
}

// S = X K'^T + c for one head on one 16-row tile, softmax over the tokens in the accumulator fragments.
// p[n][0..1]: row g, tokens 8n+2t, 8n+2t+1 ; p[n][2..3]: row g+8.  Tokens >= M get probability 0.
template <int NT>
__device__ __forceinline__ void tile_softmax(const uint32_t (&xh)[4][4], const uint32_t (&xl)[4][4], const float* Kp, const float* cb, int M,
                                             int gq, int tq, float (&p)[NT][4]) {
#pragma unroll
    for (int n = 0; n < NT; ++n) {
        const float c0 = cb[8 * n + 2 * tq], c1 = cb[8 * n + 2 * tq + 1];
        p[n][0] = c0; p[n][1] = c1; p[n][2] = c0; p[n][3] = c1;
        const float* kr = Kp + (8 * n + gq) * TS + tq;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            uint32_t bh0, bl0, bh1, bl1;
            split(kr[8 * ks], bh0, bl0);
            split(kr[8 * ks + 4], bh1, bl1);
            mma_k8(p[n], xl[ks][0], xl[ks][1], xl[ks][2], xl[ks][3], bh0, bh1);
            mma_k8(p[n], xh[ks][0], xh[ks][1], xh[ks][2], xh[ks][3], bl0, bl1);
            mma_k8(p[n], xh[ks][0], xh[ks][1], xh[ks][2], xh[ks][3], bh0, bh1);
        }
    }
    float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
    for (int n = 0; n < NT; ++n) {
        const int j = 8 * n + 2 * tq;
        if (j >= M) p[n][0] = p[n][2] = -INFINITY;
        if (j + 1 >= M) p[n][1] = p[n][3] = -INFINITY;
        m0 = fmaxf(m0, fmaxf(p[n][0], p[n][1]));
        m1 = fmaxf(m1, fmaxf(p[n][2], p[n][3]));
    }
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1));
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
    float d0 = 0.f, d1 = 0.f;
#pragma unroll
    for (int n = 0; n < NT; ++n) {
        p[n][0] = __expf(p[n][0] - m0); p[n][1] = __expf(p[n][1] - m0);
        p[n][2] = __expf(p[n][2] - m1); p[n][3] = __expf(p[n][3] - m1);
        d0 += p[n][0] + p[n][1];
        d1 += p[n][2] + p[n][3];
    }
    d0 += __shfl_xor_sync(0xffffffffu, d0, 1);
    d0 += __shfl_xor_sync(0xffffffffu, d0, 2);
    d1 += __shfl_xor_sync(0xffffffffu, d1, 1);
    d1 += __shfl_xor_sync(0xffffffffu, d1, 2);
    const float i0 = 1.f / d0, i1 = 1.f / d1;
#pragma unroll
    for (int n = 0; n < NT; ++n) {
        p[n][0] *= i0; p[n][1] *= i0; p[n][2] *= i1; p[n][3] *= i1;
    }
}

// acc[n2] (16 x 8 feature tiles) += F (16 x 8*NT, accumulator layout, used as the A operand with the permuted contraction index) x T,
// T = table rows (8*NT x 32, stride TS): k slot t <-> row 8k+2t, k slot t+4 <-> row 8k+2t+1.
template <int NT>
__device__ __forceinline__ void frag_times_table(const float (&f)[NT][4], const float* T, int gq, int tq, float (&acc)[4][4]) {
#pragma unroll
    for (int k = 0; k < NT; ++k) {
        uint32_t ah[4], al[4];
        split(f[k][0], ah[0], al[0]);      // (row g,   k slot t)
        split(f[k][2], ah[1], al[1]);      // (row g+8, k slot t)
        split(f[k][1], ah[2], al[2]);      // (row g,   k slot t+4)
        split(f[k][3], ah[3], al[3]);      // (row g+8, k slot t+4)
        const float* tr = T + (8 * k + 2 * tq) * TS + gq;
#pragma unroll
        for (int n2 = 0; n2 < 4; ++n2) {
            uint32_t bh0, bl0, bh1, bl1;
            split(tr[8 * n2], bh0, bl0);
            split(tr[TS + 8 * n2], bh1, bl1);
            mma_k8(acc[n2], al[0], al[1], al[2], al[3], bh0, bh1);
            mma_k8(acc[n2], ah[0], ah[1], ah[2], ah[3], bl0, bl1);
            mma_k8(acc[n2], ah[0], ah[1], ah[2], ah[3], bh0, bh1);
        }
    }
}

// the four k-steps of a 16 x 32 row tile as split A fragments
__device__ __forceinline__ void load_rows_a(const float* tile_rows, int lane, uint32_t (&h)[4][4], uint32_t (&l)[4][4]) {
    const uint32_t addr = smem_addr(tile_rows) + (uint32_t)(((lane & 7) + 8 * ((lane >> 3) & 1)) * TS + 4 * (lane >> 4)) * 4u;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
        uint32_t af[4];
        ldmatrix_a(af, addr + 32u * ks);
#pragma unroll
        for (int q = 0; q < 4; ++q) split(__uint_as_float(af[q]), h[ks][q], l[ks][q]);
    }
}

template <int NT>
__global__ void __launch_bounds__(kThreads, 2) attn_mma_fwd_kernel(AttnArgs a, Geo geo) {
    IGCN_PDL_SYNC();
    extern __shared__ __align__(16) float smf[];
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarp = nt >> 5;
    const int gq = lane >> 2, tq = lane & 3;
    const int R = a.R, M = a.M, H = a.heads, MP = geo.MP;
    float* WkvT = smf;
    float* Wq = WkvT + 2 * kE * kE;
    float* WoT = Wq + kE * kE;
    float* bin = WoT + kE * kE;
    float* bo = bin + 3 * kE;
    float* per = bo + kE;
    float* Xs = per + geo.gpc * geo.per_sz;
    if (a.tab) {                                            // tables come from attn_tables_kernel: no weights needed here
        for (int i = tid; i < kE; i += nt) bo[i] = a.bo[i];
    } else {
        rows::load_weights(a, WkvT, Wq, WoT, bin, bo);
    }
    __syncthreads();
    const int T = (R + 15) >> 4;                            // 16-row tiles per graph (tiles never straddle graphs: the tables differ)
    const int oKp = 3 * M * kE, oVp = oKp + H * MP * TS, oC = oVp + H * MP * TS;
    const int nhead4 = (2 * H * MP * TS + H * MP) >> 2;     // K' | V' | c of a graph, contiguous, in float4 units
    for (int b0 = blockIdx.x * geo.gpc; b0 < a.B; b0 += gridDim.x * geo.gpc) {
        const int ng = min(geo.gpc, a.B - b0), rows = ng * R;
        const float* xg = a.x + (int64_t)b0 * R * kE;
        for (int i0 = 0; i0 < rows * 8; i0 += 4 * nt) {     // coalesced 16-byte loads -> padded rows; four per thread in flight
            float4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (i0 + tid + u * nt < rows * 8) v[u] = ld4s(xg + (int64_t)(i0 + tid + u * nt) * 4);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int idx = i0 + tid + u * nt;
                if (idx < rows * 8) st4s(Xs + (idx >> 3) * TS + 4 * (idx & 7), v[u]);
            }
        }
        if (a.tab) {
            for (int i0 = 0; i0 < ng * nhead4; i0 += 5 * nt) {
                float4 v[5];
#pragma unroll
                for (int u = 0; u < 5; ++u) {
                    const int idx = i0 + tid + u * nt;
                    if (idx < ng * nhead4) {
                        const int gl = idx / nhead4, i = idx - gl * nhead4;
                        v[u] = ld4s(a.tab + (int64_t)(b0 + gl) * a.tab_sz + 4 * i);
                    }
                }
#pragma unroll
                for (int u = 0; u < 5; ++u) {
                    const int idx = i0 + tid + u * nt;
                    if (idx < ng * nhead4) {
                        const int gl = idx / nhead4, i = idx - gl * nhead4;
                        st4s(per + gl * geo.per_sz + oKp + 4 * i, v[u]);
                    }
                }
            }
            __syncthreads();
        } else {
            graph_tables(a, b0, ng, per, geo.per_sz, MP, WkvT, Wq, WoT, bin);   // ends with a barrier (covers Xs too)
        }
        for (int item = warp; item < ng * T; item += nwarp) {
            const int gl = item / T, t = item - gl * T;
            const float* tb = per + gl * geo.per_sz;
            uint32_t xh[4][4], xl[4][4];
            load_rows_a(Xs + (gl * R + 16 * t) * TS, lane, xh, xl);
            float y[4][4];
#pragma unroll
            for (int n2 = 0; n2 < 4; ++n2) {
                const float c0 = bo[8 * n2 + 2 * tq], c1 = bo[8 * n2 + 2 * tq + 1];
                y[n2][0] = c0; y[n2][1] = c1; y[n2][2] = c0; y[n2][3] = c1;
            }
#pragma unroll 1
            for (int h = 0; h < H; ++h) {
                float p[NT][4];
                tile_softmax<NT>(xh, xl, tb + oKp + h * MP * TS, tb + oC + h * MP, M, gq, tq, p);
                frag_times_table<NT>(p, tb + oVp + h * MP * TS, gq, tq, y);
            }
            const int r0 = 16 * t + gq, r1 = r0 + 8;
            float* yg = a.y + ((int64_t)(b0 + gl) * R) * kE;
#pragma unroll
            for (int n2 = 0; n2 < 4; ++n2) {
                float2 v0 = make_float2(y[n2][0], y[n2][1]), v1 = make_float2(y[n2][2], y[n2][3]);
                if (a.relu) {
                    v0.x = fmaxf(v0.x, 0.f); v0.y = fmaxf(v0.y, 0.f); v1.x = fmaxf(v1.x, 0.f); v1.y = fmaxf(v1.y, 0.f);
                }
                if (a.mix) {                                 // out_z = (img + relu(attn)) / 2: the query row is in the staged tile
                    const float2 x0 = *reinterpret_cast<const float2*>(Xs + (gl * R + r0) * TS + 8 * n2 + 2 * tq);
                    const float2 x1 = *reinterpret_cast<const float2*>(Xs + (gl * R + r1) * TS + 8 * n2 + 2 * tq);
                    v0.x = 0.5f * (x0.x + v0.x); v0.y = 0.5f * (x0.y + v0.y);
                    v1.x = 0.5f * (x1.x + v1.x); v1.y = 0.5f * (x1.y + v1.y);
                }
                if (r0 < R) *reinterpret_cast<float2*>(yg + r0 * kE + 8 * n2 + 2 * tq) = v0;
                if (r1 < R) *reinterpret_cast<float2*>(yg + r1 * kE + 8 * n2 + 2 * tq) = v1;
            }
        }
        __syncthreads();
    }
}


// =====================================================================================================================
// Backward.  One graph per CTA pass, one CTA per SM.  Per head (outer loop), every warp walks its 16-row tiles of the graph:
//     S, P (recomputed) -> dP = dY V'^T -> dS = P (dP - <P, dP>) -> dX += dS K'            (row-wise: A operands = row tiles)
//     dV'^T (32 x M) += dY^T P        dK'^T (32 x M) += X^T dS        dc += column sums of dS        (reductions over the rows)
// The reductions contract over the ROWS, so P / dS are needed with the row index on the k slots of a B fragment -- the transpose
// of the accumulator layout: it is done with four shuffles per fragment register pair (no shared-memory staging); the A operands
// dY^T / X^T are read transposed from the staged row tiles.  Each warp accumulates its tiles in registers, the warps' partial
// fragments are summed in warp order through a scratch buffer (deterministic, no atomics) into the per-graph tables dK'_h, dV'_h,
// and the M x 32-sized chain back to K, V, the tokens and the six parameter tensors is the row kernels' code.
// dX: head 0 stores its part, head 1 adds to it (same lane, same address: program order).
// =====================================================================================================================
struct GeoB {
    int MP, per_sz, rows_pad, VS, nthreads;
    size_t smem;
};
// per-graph tables: A, K, V (M x 32 each) | K', V' (H x MP x TS) | c (H x MP) | dK', dV' (H x MP x TS; column 32 of dK' = dc) | dbo (32) | dK, dV (M x 32)
__host__ __device__ inline int bwd_per_sz(int M, int MP, int H) { return 3 * M * kE + 4 * H * MP * TS + H * MP + kE + 2 * M * kE; }
static GeoB bwd_geo(int R, int M, int H) {
    GeoB g;
    g.MP = (M + 7) & ~7;
    g.per_sz = (bwd_per_sz(M, g.MP, H) + 3) & ~3;
    g.rows_pad = ((R + 15) & ~15);
    g.VS = 18 * (g.MP / 8) + 8;
    g.nthreads = kThreads;
    g.smem = (size_t)4 * (8 * kE * kE + 4 * kE + (size_t)g.per_sz + 2 * (size_t)g.rows_pad * TS + (size_t)(kThreads / 32) * g.VS * 32) + 16;
    return g;
}

// B fragment (k = tile row, n = token) for k-step KS out of a 16-row x 8-token block held in ACCUMULATOR layout
// (f[0..1]: row g, tokens 2t, 2t+1; f[2..3]: row g+8):  b0 = F[8 KS + t][g], b1 = F[8 KS + t + 4][g].
// Element F[r][c] lives in lane (g = r & 7, t = c >> 1), register (c & 1) + 2 (r >> 3).
template <int KS>
__device__ __forceinline__ void frag_b_from_acc(const float (&f)[4], int gq, int tq, float& b0, float& b1) {
    const int src0 = (tq << 2) | (gq >> 1), src1 = ((tq + 4) << 2) | (gq >> 1);
    const float ev = f[2 * KS], od = f[2 * KS + 1];
    const float e0 = __shfl_sync(0xffffffffu, ev, src0), o0 = __shfl_sync(0xffffffffu, od, src0);
    const float e1 = __shfl_sync(0xffffffffu, ev, src1), o1 = __shfl_sync(0xffffffffu, od, src1);
    b0 = (gq & 1) ? o0 : e0;
    b1 = (gq & 1) ? o1 : e1;
}

// acc[mt][n] (feature m-tile x token n-tile) += Rows^T (32 x 16, read transposed from the staged tile) x F (16 x 8 NT, accumulator layout)
template <int NT>
__device__ __forceinline__ void rows_t_times_frag(const float* tile_rows, const float (&f)[NT][4], int gq, int tq, float (&acc)[2][NT][4]) {
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
        uint32_t bh[NT][2], bl[NT][2];
#pragma unroll
        for (int n = 0; n < NT; ++n) {
            float b0, b1;
            if (ks == 0) frag_b_from_acc<0>(f[n], gq, tq, b0, b1); else frag_b_from_acc<1>(f[n], gq, tq, b0, b1);
            split(b0, bh[n][0], bl[n][0]);
            split(b1, bh[n][1], bl[n][1]);
        }
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
            const float* ra = tile_rows + (8 * ks + tq) * TS + 16 * mt + gq;
            uint32_t ah[4], al[4];
            split(ra[0], ah[0], al[0]);
            split(ra[8], ah[1], al[1]);
            split(ra[4 * TS], ah[2], al[2]);
            split(ra[4 * TS + 8], ah[3], al[3]);
#pragma unroll
            for (int n = 0; n < NT; ++n) {
                mma_k8(acc[mt][n], al[0], al[1], al[2], al[3], bh[n][0], bh[n][1]);
                mma_k8(acc[mt][n], ah[0], ah[1], ah[2], ah[3], bl[n][0], bl[n][1]);
                mma_k8(acc[mt][n], ah[0], ah[1], ah[2], ah[3], bh[n][0], bh[n][1]);
            }
        }
    }
}

template <int NT>
__global__ void __launch_bounds__(kThreads, 1) attn_mma_bwd_kernel(AttnArgs a, GeoB geo) {
    IGCN_PDL_SYNC();
    extern __shared__ __align__(16) float smf[];
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarp = nt >> 5;
    const int gq = lane >> 2, tq = lane & 3;
    const int R = a.R, M = a.M, H = a.heads, hd = kE / H, MP = geo.MP, VS = geo.VS;
    const float scale = rsqrtf((float)hd);
    float* WkvT = smf;
    float* Wq = WkvT + 2 * kE * kE;
    float* WoT = Wq + kE * kE;
    float* bin = WoT + kE * kE;
    float* bo = bin + 3 * kE;
    float* WqT = bo + kE;
    float* WoO = WqT + kE * kE;
    float* WkvO = WoO + kE * kE;
    float* per = WkvO + 2 * kE * kE;
    float* Xs = per + geo.per_sz;
    float* Ys = Xs + geo.rows_pad * TS;
    float* scratch = Ys + geo.rows_pad * TS;
    const int oK = M * kE, oV = 2 * M * kE, oKp = 3 * M * kE, oVp = oKp + H * MP * TS, oC = oVp + H * MP * TS, oDKp = oC + H * MP;
    const int oDVp = oDKp + H * MP * TS, oDBO = oDVp + H * MP * TS, oDK = oDBO + kE, oDV = oDK + M * kE;
    float* accg = a.partials + (int64_t)blockIdx.x * a.P;
    const int oBin = 3 * kE * kE, oWo = oBin + 3 * kE, oBo = oWo + kE * kE;
    for (int i = tid; i < a.P; i += nt) accg[i] = 0.f;
    rows::load_weights(a, WkvT, Wq, WoT, bin, bo);
    for (int i = tid; i < kE * kE; i += nt) {
        const int f = i / kE, e = i - f * kE;
        WqT[e * kE + f] = a.Win[i];
        WoO[i] = a.Wo[i];
    }
    for (int i = tid; i < 2 * kE * kE; i += nt) WkvO[i] = a.Win[kE * kE + i];
    __syncthreads();
    const int T = (R + 15) >> 4;
    const int nbusy = T < nwarp ? T : nwarp;
    for (int b0 = blockIdx.x; b0 < a.B; b0 += gridDim.x) {
        const float* xg = a.x + (int64_t)b0 * R * kE;
        const float* gg = a.gy + (int64_t)b0 * R * kE;
        const float* yg = a.yout + (int64_t)b0 * R * kE;
        for (int idx = tid; idx < geo.rows_pad * 8; idx += nt) {
            const int row = idx >> 3, ch = idx & 7;
            float4 xv = make_float4(0.f, 0.f, 0.f, 0.f), g = xv;          // rows >= R of the last tile: zeros (they add nothing)
            if (row < R) {
                xv = ld4s(xg + (int64_t)idx * 4);
                g = ld4s(gg + (int64_t)idx * 4);
                if (a.relu) {
                    float4 yv = ld4s(yg + (int64_t)idx * 4);
                    if (a.mix) {        // saved output is (x + relu(y)) / 2: relu(y) = 2 out - x ; the attention branch sees half the gradient
                        yv.x = 2.f * yv.x - xv.x; yv.y = 2.f * yv.y - xv.y; yv.z = 2.f * yv.z - xv.z; yv.w = 2.f * yv.w - xv.w;
                        g.x *= 0.5f; g.y *= 0.5f; g.z *= 0.5f; g.w *= 0.5f;
                    }
                    if (!(yv.x > 0.f)) g.x = 0.f;
                    if (!(yv.y > 0.f)) g.y = 0.f;
                    if (!(yv.z > 0.f)) g.z = 0.f;
                    if (!(yv.w > 0.f)) g.w = 0.f;
                }
            }
            st4s(Xs + row * TS + 4 * ch, xv);
            st4s(Ys + row * TS + 4 * ch, g);
        }
        graph_tables(a, b0, 1, per, geo.per_sz, MP, WkvT, Wq, WoT, bin);      // ends with a barrier
        float* dxg = a.dx + (int64_t)b0 * R * kE;
#pragma unroll 1
        for (int h = 0; h < H; ++h) {
            const float* Kp = per + oKp + h * MP * TS;
            const float* Vp = per + oVp + h * MP * TS;
            float accV[2][NT][4], accK[2][NT][4], dcp[NT][2], dbp[8];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int n = 0; n < NT; ++n)
#pragma unroll
                    for (int r = 0; r < 4; ++r) accV[mt][n][r] = accK[mt][n][r] = 0.f;
#pragma unroll
            for (int n = 0; n < NT; ++n) dcp[n][0] = dcp[n][1] = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) dbp[i] = 0.f;
            for (int t = warp; t < T; t += nwarp) {
                const float* xt = Xs + 16 * t * TS;
                const float* yt = Ys + 16 * t * TS;
                float p[NT][4], ds[NT][4];
                {
                    uint32_t xh[4][4], xl[4][4];
                    load_rows_a(xt, lane, xh, xl);
                    tile_softmax<NT>(xh, xl, Kp, per + oC + h * MP, M, gq, tq, p);
                }
                {
                    // dP = dY V'^T  (B[k = feature][n = token] = V'[token][feature]: the access pattern of K' in the scores)
                    uint32_t gh[4][4], gl[4][4];
                    load_rows_a(yt, lane, gh, gl);
                    if (h == 0) {
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) {      // hi + lo is the exact value: column sums of dY for d bo
                            dbp[2 * ks] += (__uint_as_float(gh[ks][0]) + __uint_as_float(gl[ks][0])) + (__uint_as_float(gh[ks][1]) + __uint_as_float(gl[ks][1]));
                            dbp[2 * ks + 1] += (__uint_as_float(gh[ks][2]) + __uint_as_float(gl[ks][2])) + (__uint_as_float(gh[ks][3]) + __uint_as_float(gl[ks][3]));
                        }
                    }
#pragma unroll
                    for (int n = 0; n < NT; ++n) {
                        ds[n][0] = ds[n][1] = ds[n][2] = ds[n][3] = 0.f;
                        const float* vr = Vp + (8 * n + gq) * TS + tq;
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) {
                            uint32_t bh0, bl0, bh1, bl1;
                            split(vr[8 * ks], bh0, bl0);
                            split(vr[8 * ks + 4], bh1, bl1);
                            mma_k8(ds[n], gl[ks][0], gl[ks][1], gl[ks][2], gl[ks][3], bh0, bh1);
                            mma_k8(ds[n], gh[ks][0], gh[ks][1], gh[ks][2], gh[ks][3], bl0, bl1);
                            mma_k8(ds[n], gh[ks][0], gh[ks][1], gh[ks][2], gh[ks][3], bh0, bh1);
                        }
                    }
                }
                float rd0 = 0.f, rd1 = 0.f;
#pragma unroll
                for (int n = 0; n < NT; ++n) {
                    rd0 = fmaf(p[n][0], ds[n][0], fmaf(p[n][1], ds[n][1], rd0));
                    rd1 = fmaf(p[n][2], ds[n][2], fmaf(p[n][3], ds[n][3], rd1));
                }
                rd0 += __shfl_xor_sync(0xffffffffu, rd0, 1);
                rd0 += __shfl_xor_sync(0xffffffffu, rd0, 2);
                rd1 += __shfl_xor_sync(0xffffffffu, rd1, 1);
                rd1 += __shfl_xor_sync(0xffffffffu, rd1, 2);
#pragma unroll
                for (int n = 0; n < NT; ++n) {
                    ds[n][0] = p[n][0] * (ds[n][0] - rd0);
                    ds[n][1] = p[n][1] * (ds[n][1] - rd0);
                    ds[n][2] = p[n][2] * (ds[n][2] - rd1);
                    ds[n][3] = p[n][3] * (ds[n][3] - rd1);
                    dcp[n][0] += ds[n][0] + ds[n][2];
                    dcp[n][1] += ds[n][1] + ds[n][3];
                }
                // dX (this head's part) = dS K'
                {
                    float dx[4][4];
#pragma unroll
                    for (int n2 = 0; n2 < 4; ++n2) dx[n2][0] = dx[n2][1] = dx[n2][2] = dx[n2][3] = 0.f;
                    frag_times_table<NT>(ds, Kp, gq, tq, dx);
                    const int r0 = 16 * t + gq, r1 = r0 + 8;
#pragma unroll
                    for (int n2 = 0; n2 < 4; ++n2) {
                        float2* d0 = reinterpret_cast<float2*>(dxg + r0 * kE + 8 * n2 + 2 * tq);
                        float2* d1 = reinterpret_cast<float2*>(dxg + r1 * kE + 8 * n2 + 2 * tq);
                        if (r0 < R) {
                            float2 v = make_float2(dx[n2][0], dx[n2][1]);
                            if (h > 0) { const float2 o = *d0; v.x += o.x; v.y += o.y; }
                            else if (a.mix) {                  // the direct half of the average: d out / d x = 1/2
                                const float2 o = *reinterpret_cast<const float2*>(gg + r0 * kE + 8 * n2 + 2 * tq);
                                v.x = fmaf(0.5f, o.x, v.x); v.y = fmaf(0.5f, o.y, v.y);
                            }
                            *d0 = v;
                        }
                        if (r1 < R) {
                            float2 v = make_float2(dx[n2][2], dx[n2][3]);
                            if (h > 0) { const float2 o = *d1; v.x += o.x; v.y += o.y; }
                            else if (a.mix) {
                                const float2 o = *reinterpret_cast<const float2*>(gg + r1 * kE + 8 * n2 + 2 * tq);
                                v.x = fmaf(0.5f, o.x, v.x); v.y = fmaf(0.5f, o.y, v.y);
                            }
                            *d1 = v;
                        }
                    }
                }
                rows_t_times_frag<NT>(yt, p, gq, tq, accV);      // dV'^T += dY^T P
                rows_t_times_frag<NT>(xt, ds, gq, tq, accK);     // dK'^T += X^T dS
            }
            // ---- flush the warp's partial fragments, sum them in warp order into the tables ---------------------------------------
            if (warp < nbusy) {
                float* sc = scratch + (size_t)warp * VS * 32 + lane;
#pragma unroll
                for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                    for (int n = 0; n < NT; ++n)
#pragma unroll
                        for (int r = 0; r < 4; ++r) {
                            sc[((mt * NT + n) * 4 + r) * 32] = accV[mt][n][r];
                            sc[(8 * NT + (mt * NT + n) * 4 + r) * 32] = accK[mt][n][r];
                        }
#pragma unroll
                for (int n = 0; n < NT; ++n)
#pragma unroll
                    for (int b = 0; b < 2; ++b) {
                        float v = dcp[n][b];                    // rows g, g+8 of the warp's tiles: sum over g (lanes with the same t)
                        v += __shfl_xor_sync(0xffffffffu, v, 4);
                        v += __shfl_xor_sync(0xffffffffu, v, 8);
                        v += __shfl_xor_sync(0xffffffffu, v, 16);
                        sc[(16 * NT + 2 * n + b) * 32] = v;
                    }
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float v = dbp[i];
                    v += __shfl_xor_sync(0xffffffffu, v, 4);
                    v += __shfl_xor_sync(0xffffffffu, v, 8);
                    v += __shfl_xor_sync(0xffffffffu, v, 16);
                    sc[(18 * NT + i) * 32] = v;
                }
            }
            __syncthreads();
            for (int idx = tid; idx < VS * 32; idx += nt) {
                const int v = idx >> 5, l = idx & 31, g = l >> 2, t4 = l & 3;
                float sacc = 0.f;
                for (int w = 0; w < nbusy; ++w) sacc += scratch[((size_t)w * VS + v) * 32 + l];
                if (v < 16 * NT) {
                    const int vv = v < 8 * NT ? v : v - 8 * NT;
                    const int mt = vv / (4 * NT), n = (vv >> 2) - mt * NT, r = vv & 3;
                    const int f = 16 * mt + g + ((r & 2) ? 8 : 0), j = 8 * n + 2 * t4 + (r & 1);
                    per[(v < 8 * NT ? oDVp : oDKp) + (h * MP + j) * TS + f] = sacc;
                } else if (v < 18 * NT) {
                    if (g == 0) {
                        const int n = (v - 16 * NT) >> 1, b = (v - 16 * NT) & 1;
                        per[oDKp + (h * MP + 8 * n + 2 * t4 + b) * TS + 32] = sacc;
                    }
                } else if (g == 0 && h == 0) {
                    const int i = v - 18 * NT;
                    per[oDBO + 8 * (i >> 1) + t4 + 4 * (i & 1)] = sacc;
                }
            }
            __syncthreads();
        }
        // ---- chain back to K, V, the tokens and the parameters (M x 32 sized; the row kernels' code on this table layout) --------
        // dK[j][f] = scale (<dK'_h[j], Wq[f]> + dc_h[j] bq[f]) ; dV[j][f] = <dV'_h[j], WoT[f]>     (f = h hd + d)
        for (int idx = tid; idx < M * 16; idx += nt) {
            const int j = idx >> 4, c = idx & 15, fq = c & 7;
            const bool isK = c < 8;
            const int h = (4 * fq) / hd;
            const float* src = per + (isK ? oDKp : oDVp) + (h * MP + j) * TS;
            const float* wm = (isK ? WqT : WoO) + 4 * fq;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
            for (int e = 0; e < kE; ++e) fma4(src[e], ld4s(wm + e * kE), v);
            if (isK) {
                const float dc = src[32];
                const float4 bq = ld4s(bin + 4 * fq);
                v.x = scale * (v.x + dc * bq.x); v.y = scale * (v.y + dc * bq.y);
                v.z = scale * (v.z + dc * bq.z); v.w = scale * (v.w + dc * bq.w);
            }
            st4s(per + (isK ? oDK : oDV) + j * kE + 4 * fq, v);
        }
        // dWq[f][4q..] += scale sum_j K[j][f] dK'_h[j][4q..] ; dWo[f'][4q..] += sum_j dV'_h[j][f'] V[j][4q..] ; dbq ; dbo
        for (int idx = tid; idx < 2 * kE * 8 + kE; idx += nt) {
            if (idx < 2 * kE * 8) {
                const bool isQ = idx < kE * 8;
                const int i2 = isQ ? idx : idx - kE * 8, f = i2 >> 3, q = i2 & 7;
                const int h = isQ ? f / hd : (4 * q) / hd;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                const float* sc1 = isQ ? per + oK + f : per + oDVp + h * MP * TS + f;
                const int ld1 = isQ ? kE : TS;
                const float* vec = isQ ? per + oDKp + h * MP * TS + 4 * q : per + oV + 4 * q;
                const int ld2 = isQ ? TS : kE;
#pragma unroll 4
                for (int j = 0; j < M; ++j) fma4(sc1[j * ld1], ld4s(vec + j * ld2), v);
                float* d = accg + (isQ ? 0 : oWo) + f * kE + 4 * q;
                const float sf = isQ ? scale : 1.f;
                d[0] += sf * v.x; d[1] += sf * v.y; d[2] += sf * v.z; d[3] += sf * v.w;
            } else {
                const int f = idx - 2 * kE * 8, h = f / hd;
                float v = 0.f;
#pragma unroll 4
                for (int j = 0; j < M; ++j) v = fmaf(per[oK + j * kE + f], per[oDKp + (h * MP + j) * TS + 32], v);
                accg[oBin + f] += scale * v;
                accg[oBo + f] += per[oDBO + f];
            }
        }
        __syncthreads();
        // token gradient dA[j][k] = <dK[j], Wk[:,k]> + <dV[j], Wv[:,k]> ; dWk/dWv[f][k] += sum_j d{K,V}[j][f] A[j][k] ; dbk/dbv
        for (int idx = tid; idx < M * 8; idx += nt) {
            const int j = idx >> 3, kq = idx & 7;
            const float* dk = per + oDK + j * kE;
            const float* dv = per + oDV + j * kE;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
            for (int f = 0; f < kE; ++f) {
                fma4(dk[f], ld4s(WkvO + f * kE + 4 * kq), v);
                fma4(dv[f], ld4s(WkvO + (kE + f) * kE + 4 * kq), v);
            }
            *reinterpret_cast<float4*>(a.da + ((int64_t)b0 * M + j) * kE + 4 * kq) = v;
        }
        for (int idx = tid; idx < 2 * kE * 8 + 2 * kE; idx += nt) {
            if (idx < 2 * kE * 8) {
                const int f2 = idx >> 3, q = idx & 7;
                const int o = (f2 < kE) ? oDK + f2 : oDV + (f2 - kE);
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
                for (int j = 0; j < M; ++j) fma4(per[o + j * kE], ld4s(per + j * kE + 4 * q), v);
                float* d = accg + kE * kE + f2 * kE + 4 * q;
                d[0] += v.x; d[1] += v.y; d[2] += v.z; d[3] += v.w;
            } else {
                const int f2 = idx - 2 * kE * 8;
                const int o = (f2 < kE) ? oDK + f2 : oDV + (f2 - kE);
                float v = 0.f;
                for (int j = 0; j < M; ++j) v += per[o + j * kE];
                accg[oBin + kE + f2] += v;
            }
        }
        __syncthreads();
    }
}

}  // namespace amma
}  // namespace igcn
