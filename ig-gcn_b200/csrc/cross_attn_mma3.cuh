// The two token-side kernels of the table-driven cross attention (cross_attn_mma2.cuh) on warp-level tensor cores.
//
// ncu on the scalar versions at the benchmarked size (profiles/r2_ncu_attn_v2_c2.json, 512 graphs, M = 19 tokens): the tables kernel
// runs 14 K warp instructions per graph and the chain kernel 16.6 K for ~2.5 K warp-FMAs of arithmetic each -- every product is a
// dot-product loop of one broadcast LDS + one LDS.128 per four FMAs, plus a weight prologue per CTA.  Both are small dense products
// with the weights as one operand, so they are written as m16n8k8 TF32 MMAs (split in three, fp32 accumulate; mma_util.cuh) with the
// weight operands pre-split into hi / lo ONCE per CTA:
//
//   attn_tables_mma_kernel   a warp owns a graph:  [K | V] (M x 64) = A (M x 32) [Wk ; Wv]^T + b, then per head K'_h = scale K_h Wq_h,
//                            V'_h = V_h Wo_h^T, c_h = scale K_h bq_h -- the second products take K / V straight from the accumulator
//                            registers (the accumulator layout is an A fragment with the contraction index permuted: k slot t <->
//                            column 2t, slot t+4 <-> column 2t+1, and the weight rows are read with the same permutation).
//   attn_chain_mma_kernel    a CTA of four warps walks graphs: dK = scale (dK'_h Wq_h^T + dc_h bq_h), dV = dV'_h Wo_h (one warp per
//                            (K|V, head)); then each warp OWNS one of the four 32 x 32 parameter gradients (dWq, dWo, dWk, dWv) as
//                            MMA accumulators that live in registers over all graphs of the CTA -- products K_h^T dK'_h,
//                            dV'_h^T V_h, dK^T A, dV^T A contract over the tokens -- and a quarter of the token gradient
//                            dA = dK Wk + dV Wv.  One partial row per CTA, reduced in a fixed order (deterministic, no atomics).
// Reference: nn.MultiheadAttention(E, 2, batch_first=True), kernel/sgcn_img_snp.py:46,239-242.
#pragma once
#include "cross_attn_mma2.cuh"

namespace igcn {
namespace amma3 {

using namespace igcn::mmau;
using amma::kE;
using amma::TS;

constexpr int kWarps = 4, kThreads = 32 * kWarps;
constexpr int LW = 72;        // row stride of the 64-wide [Wk ; Wv]^T operand: conflict-free B fragments (8 t + g)
constexpr int LQ = 40;        // row stride of the 32-wide weight operands read as B[k = row 8ks + t][n = g]
constexpr int TILE = 32 * TS; // one staged 32-row table (rows >= M stay zero)

__device__ __forceinline__ void put_split(float v, float* hi, float* lo, int i) {
    uint32_t h, l;
    split(v, h, l);
    hi[i] = __uint_as_float(h);
    lo[i] = __uint_as_float(l);
}
// C += A B for one 8-step: A fragments hi / lo, B[k][n] from pre-split tables (rows k0 + t and k1 + t', column col)
__device__ __forceinline__ void mma3(float (&c)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4], const float* Bh, const float* Bl, int i0,
                                     int i1) {
    const uint32_t bh0 = __float_as_uint(Bh[i0]), bh1 = __float_as_uint(Bh[i1]);
    const uint32_t bl0 = __float_as_uint(Bl[i0]), bl1 = __float_as_uint(Bl[i1]);
    mma_k8(c, al[0], al[1], al[2], al[3], bh0, bh1);
    mma_k8(c, ah[0], ah[1], ah[2], ah[3], bl0, bl1);
    mma_k8(c, ah[0], ah[1], ah[2], ah[3], bh0, bh1);
}
// the same with a B operand that is NOT pre-split (per-graph data)
__device__ __forceinline__ void mma3_raw(float (&c)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4], float b0, float b1) {
    uint32_t bh0, bl0, bh1, bl1;
    split(b0, bh0, bl0);
    split(b1, bh1, bl1);
    mma_k8(c, al[0], al[1], al[2], al[3], bh0, bh1);
    mma_k8(c, ah[0], ah[1], ah[2], ah[3], bl0, bl1);
    mma_k8(c, ah[0], ah[1], ah[2], ah[3], bh0, bh1);
}
// accumulator tile (rows g, g+8; columns 2t, 2t+1) -> A fragment of one 8-step with the permuted contraction index
__device__ __forceinline__ void acc_as_a(const float (&acc)[4], uint32_t (&ah)[4], uint32_t (&al)[4]) {
    split(acc[0], ah[0], al[0]);      // (row g,   slot t)     <- column 2t
    split(acc[2], ah[1], al[1]);      // (row g+8, slot t)
    split(acc[1], ah[2], al[2]);      // (row g,   slot t + 4) <- column 2t + 1
    split(acc[3], ah[3], al[3]);      // (row g+8, slot t + 4)
}

// ---- tables ------------------------------------------------------------------------------------------------------------------------
static size_t tables_smem() { return (size_t)4 * (2 * 32 * LW + 4 * 32 * TS + 4 * kE + kWarps * TILE) + 16; }

__global__ void __launch_bounds__(kThreads) attn_tables_mma_kernel(AttnArgs a, int MP) {
    IGCN_PDL_SYNC();
    extern __shared__ __align__(16) float smf[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
    const int M = a.M, H = a.heads, hd = kE / H;
    const float scale = rsqrtf((float)hd);
    float* Wkv_h = smf;                      // [k][n] = [Wk ; Wv][n][k], stride LW
    float* Wkv_l = Wkv_h + 32 * LW;
    float* Wq_h = Wkv_l + 32 * LW;           // Wq[f][e], stride TS
    float* Wq_l = Wq_h + 32 * TS;
    float* Wo_h = Wq_l + 32 * TS;            // WoT[k][f] = Wo[f][k], stride TS
    float* Wo_l = Wo_h + 32 * TS;
    float* bin = Wo_l + 32 * TS;             // bq | bk | bv (+ 32 unused)
    float* As = bin + 4 * kE;                // kWarps staged token tiles; first used as the 64 x 33 transpose tile
    // every weight load of the prologue is issued before the first use (ncu: with load -> store loops the prologue was 70 % of this
    // kernel's samples, all of them waiting on one L2 round trip per iteration)
    constexpr int PER = kE * kE / kThreads;            // 8 elements of a 32 x 32 matrix per thread
    float wkv[2 * PER], wq[PER], wo[PER];
#pragma unroll
    for (int u = 0; u < 2 * PER; ++u) wkv[u] = __ldg(a.Win + kE * kE + tid + u * kThreads);
#pragma unroll
    for (int u = 0; u < PER; ++u) {
        wq[u] = __ldg(a.Win + tid + u * kThreads);
        wo[u] = __ldg(a.Wo + tid + u * kThreads);
    }
    if (tid < 3 * kE) bin[tid] = a.bin[tid];
#pragma unroll
    for (int u = 0; u < 2 * PER; ++u) {
        const int i = tid + u * kThreads;
        As[(i >> 5) * 33 + (i & 31)] = wkv[u];
    }
#pragma unroll
    for (int u = 0; u < PER; ++u) {
        const int i = tid + u * kThreads;
        put_split(wq[u], Wq_h, Wq_l, (i >> 5) * TS + (i & 31));
    }
    __syncthreads();
    for (int o = tid; o < 2 * kE * kE; o += kThreads) {
        const int k = o >> 6, n = o & 63;
        put_split(As[n * 33 + k], Wkv_h, Wkv_l, k * LW + n);
    }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < PER; ++u) {
        const int i = tid + u * kThreads;
        As[(i >> 5) * 33 + (i & 31)] = wo[u];
    }
    __syncthreads();
    for (int o = tid; o < kE * kE; o += kThreads) {
        const int k = o >> 5, f = o & 31;
        put_split(As[f * 33 + k], Wo_h, Wo_l, k * TS + f);
    }
    __syncthreads();
    for (int i = tid; i < kWarps * TILE; i += kThreads) As[i] = 0.f;
    __syncthreads();
    float* as = As + warp * TILE;
    const int head = amma2::tab_head(MP, H);
    for (int b = blockIdx.x * kWarps + warp; b < a.B; b += gridDim.x * kWarps) {      // warp-private from here on (no block barriers)
        const float* ag = a.a + (int64_t)b * M * kE;
        {
            float4 v[8];                                       // M <= 32 tokens: at most 8 float4 per lane, all in flight together
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int i = lane + 32 * u;
                if (i < M * 8) v[u] = ld4s(ag + 4 * i);
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int i = lane + 32 * u;
                if (i < M * 8) st4s(as + (i >> 3) * TS + 4 * (i & 7), v[u]);
            }
        }
        __syncwarp();
        float* tb = a.tab_out + (int64_t)b * a.tab_sz;
#pragma unroll 1
        for (int mt = 0; 16 * mt < MP; ++mt) {
            const int r0 = 16 * mt + g, r1 = r0 + 8;
            float acc[8][4];                                   // [K | V] rows r0, r1
            {
                uint32_t ah[4][4], al[4][4];
                amma::load_rows_a(as + 16 * mt * TS, lane, ah, al);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float b0 = bin[kE + 8 * j + 2 * t], b1 = bin[kE + 8 * j + 2 * t + 1];
                    acc[j][0] = b0; acc[j][1] = b1; acc[j][2] = b0; acc[j][3] = b1;
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)
                        mma3(acc[j], ah[ks], al[ks], Wkv_h, Wkv_l, (8 * ks + t) * LW + 8 * j + g, (8 * ks + t + 4) * LW + 8 * j + g);
                }
            }
            float* kv = tb + head;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float* d = kv + (j < 4 ? 0 : M * kE) + (8 * (j & 3) + 2 * t);
                if (r0 < M) *reinterpret_cast<float2*>(d + r0 * kE) = make_float2(acc[j][0], acc[j][1]);
                if (r1 < M) *reinterpret_cast<float2*>(d + r1 * kE) = make_float2(acc[j][2], acc[j][3]);
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                if (h < H) {
                    // c_h = scale K_h bq_h
                    float p0 = 0.f, p1 = 0.f;
#pragma unroll
                    for (int kk = 0; kk < 2; ++kk) {
                        const float q0 = bin[hd * h + 8 * kk + 2 * t], q1 = bin[hd * h + 8 * kk + 2 * t + 1];
                        p0 = fmaf(acc[2 * h + kk][0], q0, fmaf(acc[2 * h + kk][1], q1, p0));
                        p1 = fmaf(acc[2 * h + kk][2], q0, fmaf(acc[2 * h + kk][3], q1, p1));
                    }
                    p0 += __shfl_xor_sync(0xffffffffu, p0, 1);
                    p0 += __shfl_xor_sync(0xffffffffu, p0, 2);
                    p1 += __shfl_xor_sync(0xffffffffu, p1, 1);
                    p1 += __shfl_xor_sync(0xffffffffu, p1, 2);
                    if (t == 0) {
                        float* cd = tb + 2 * H * MP * TS + h * MP;
                        if (r0 < MP) cd[r0] = r0 < M ? scale * p0 : 0.f;
                        if (r1 < MP) cd[r1] = r1 < M ? scale * p1 : 0.f;
                    }
#pragma unroll
                    for (int tab = 0; tab < 2; ++tab) {            // 0: K'_h = scale K_h Wq_h ; 1: V'_h = V_h Wo_h^T
                        const float* Bh = tab ? Wo_h : Wq_h;
                        const float* Bl = tab ? Wo_l : Wq_l;
                        float out[4][4];
#pragma unroll
                        for (int n = 0; n < 4; ++n) out[n][0] = out[n][1] = out[n][2] = out[n][3] = 0.f;
#pragma unroll
                        for (int kk = 0; kk < 2; ++kk) {
                            uint32_t ah[4], al[4];
                            acc_as_a(acc[4 * tab + 2 * h + kk], ah, al);
                            const int rb = (hd * h + 8 * kk + 2 * t) * TS + g;
#pragma unroll
                            for (int n = 0; n < 4; ++n) mma3(out[n], ah, al, Bh, Bl, rb + 8 * n, rb + TS + 8 * n);
                        }
                        const float sc = tab ? 1.f : scale;
                        float* d = tb + tab * H * MP * TS + (h * MP) * TS + 2 * t;
#pragma unroll
                        for (int n = 0; n < 4; ++n) {
                            if (r0 < MP)
                                *reinterpret_cast<float2*>(d + r0 * TS + 8 * n) = r0 < M ? make_float2(sc * out[n][0], sc * out[n][1]) : make_float2(0.f, 0.f);
                            if (r1 < MP)
                                *reinterpret_cast<float2*>(d + r1 * TS + 8 * n) = r1 < M ? make_float2(sc * out[n][2], sc * out[n][3]) : make_float2(0.f, 0.f);
                        }
                    }
                }
            }
        }
        __syncwarp();
    }
}

// ---- chain -------------------------------------------------------------------------------------------------------------------------
// shared tables per CTA: weights hi / lo (WqT, Wo, Wk, Wv; stride LQ) | bq | A, K, V | dK'_0, dK'_1 (column 32 = dc) | dV'_0, dV'_1 | dK, dV | dbo
static size_t chain_smem() { return (size_t)4 * (8 * 32 * LQ + kE + 9 * TILE + kE) + 16; }

// accP[mt][n] += L^T R over the tokens: rows of the result = columns 16 mt + (g, g + 8) of L, columns = 8 n + (2t, 2t + 1) of R
template <int MT_LO, int MT_HI, int N_LO, int N_HI>
__device__ __forceinline__ void outer_acc(const float* L, int lcol0, const float* R, int rcol0, int ksteps, int g, int t, float (&accP)[2][4][4]) {
    for (int ks = 0; ks < ksteps; ++ks) {
        const float* lr = L + (8 * ks + t) * TS + lcol0 + g;
        const float* rr = R + (8 * ks + t) * TS + rcol0 + g;
        float b0[N_HI - N_LO], b1[N_HI - N_LO];
#pragma unroll
        for (int n = N_LO; n < N_HI; ++n) {
            b0[n - N_LO] = rr[8 * (n - N_LO)];
            b1[n - N_LO] = rr[4 * TS + 8 * (n - N_LO)];
        }
#pragma unroll
        for (int mt = MT_LO; mt < MT_HI; ++mt) {
            uint32_t ah[4], al[4];
            split(lr[16 * (mt - MT_LO)], ah[0], al[0]);
            split(lr[16 * (mt - MT_LO) + 8], ah[1], al[1]);
            split(lr[4 * TS + 16 * (mt - MT_LO)], ah[2], al[2]);
            split(lr[4 * TS + 16 * (mt - MT_LO) + 8], ah[3], al[3]);
#pragma unroll
            for (int n = N_LO; n < N_HI; ++n) mma3_raw(accP[mt][n], ah, al, b0[n - N_LO], b1[n - N_LO]);
        }
    }
}

__global__ void __launch_bounds__(kThreads) attn_chain_mma_kernel(AttnArgs a, int MP) {
    IGCN_PDL_SYNC();
    extern __shared__ __align__(16) float smf[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
    const int M = a.M, H = a.heads, hd = kE / H;
    const float scale = rsqrtf((float)hd);
    float* WqT_h = smf;                       // [e][f] = Wq[f][e]
    float* WqT_l = WqT_h + 32 * LQ;
    float* Wo_h = WqT_l + 32 * LQ;            // Wo[e][f]
    float* Wo_l = Wo_h + 32 * LQ;
    float* Wk_h = Wo_l + 32 * LQ;             // Wk[f][k]
    float* Wk_l = Wk_h + 32 * LQ;
    float* Wv_h = Wk_l + 32 * LQ;
    float* Wv_l = Wv_h + 32 * LQ;
    float* bq = Wv_l + 32 * LQ;
    float* As = bq + kE;
    float* Ks = As + TILE;
    float* Vs = Ks + TILE;
    float* dKp = Vs + TILE;                   // 2 heads x TILE
    float* dVp = dKp + 2 * TILE;              // 2 heads x TILE
    float* dKs = dVp + 2 * TILE;
    float* dVs = dKs + TILE;
    float* dbo_s = dVs + TILE;
    {
        constexpr int PER = kE * kE / kThreads;        // all 32 weight loads of a thread in flight before the first use
        float wq[PER], wo[PER], wk[PER], wv[PER];
#pragma unroll
        for (int u = 0; u < PER; ++u) {
            const int i = tid + u * kThreads;
            wq[u] = __ldg(a.Win + i);
            wo[u] = __ldg(a.Wo + i);
            wk[u] = __ldg(a.Win + kE * kE + i);
            wv[u] = __ldg(a.Win + 2 * kE * kE + i);
        }
        if (tid < kE) bq[tid] = a.bin[tid];
#pragma unroll
        for (int u = 0; u < PER; ++u) {
            const int i = tid + u * kThreads, r = i >> 5, c = i & 31;
            As[r * 33 + c] = wq[u];                                         // Wq, transposed below
            put_split(wo[u], Wo_h, Wo_l, r * LQ + c);
            put_split(wk[u], Wk_h, Wk_l, r * LQ + c);
            put_split(wv[u], Wv_h, Wv_l, r * LQ + c);
        }
    }
    __syncthreads();
    for (int o = tid; o < kE * kE; o += kThreads) {
        const int e = o >> 5, f = o & 31;
        put_split(As[f * 33 + e], WqT_h, WqT_l, e * LQ + f);
    }
    __syncthreads();
    for (int i = tid; i < 9 * TILE; i += kThreads) As[i] = 0.f;             // rows >= M (>= MP for the records) stay zero
    float accP[2][4][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int n = 0; n < 4; ++n) accP[mt][n][0] = accP[mt][n][1] = accP[mt][n][2] = accP[mt][n][3] = 0.f;
    float sbias = 0.f;                        // lane f: warp 0 dbq, warp 1 dbo, warp 2 dbk, warp 3 dbv
    const int head = amma2::tab_head(MP, H);
    const int nrec4 = H * MP * 8;             // float4 groups of dV' (and of dK') in a chunk record
    const int ksteps = MP >> 3;
    for (int b = blockIdx.x; b < a.B; b += gridDim.x) {
        __syncthreads();                       // the previous graph is done with the tables (first pass: the zero fill)
        {
            const float* ab = a.a + (int64_t)b * M * kE;
            const float* tbk = a.tab + (int64_t)b * a.tab_sz + head;
            {
                float4 va[2], vk[2], vv[2];                                  // M * 8 <= 256 float4 per table: two per thread, batched
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int i = tid + u * kThreads;
                    if (i < M * 8) {
                        va[u] = ld4s(ab + 4 * i);
                        vk[u] = ld4s(tbk + 4 * i);
                        vv[u] = ld4s(tbk + M * kE + 4 * i);
                    }
                }
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int i = tid + u * kThreads;
                    if (i < M * 8) {
                        const int o = (i >> 3) * TS + 4 * (i & 7);
                        st4s(As + o, va[u]);
                        st4s(Ks + o, vk[u]);
                        st4s(Vs + o, vv[u]);
                    }
                }
            }
            const float* d0 = a.dtab + (int64_t)b * a.nchunk * a.dtab_sz;
            for (int i0 = 0; i0 < 2 * nrec4; i0 += 4 * kThreads) {          // chunk records summed in chunk order, 4 groups per thread in flight
                float4 v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int i = i0 + tid + u * kThreads;
                    if (i < 2 * nrec4) v[u] = ld4s(d0 + 4 * i);
                }
                for (int c = 1; c < a.nchunk; ++c) {
                    float4 w[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int i = i0 + tid + u * kThreads;
                        if (i < 2 * nrec4) w[u] = ld4s(d0 + (int64_t)c * a.dtab_sz + 4 * i);
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        v[u].x += w[u].x; v[u].y += w[u].y; v[u].z += w[u].z; v[u].w += w[u].w;
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int i = i0 + tid + u * kThreads;
                    if (i < 2 * nrec4) {
                        const int q = i >= nrec4, r = q ? i - nrec4 : i, hj = r >> 3, c4 = r & 7, h = hj / MP, j = hj - h * MP;
                        st4s((q ? dKp : dVp) + h * TILE + j * TS + 4 * c4, v[u]);      // record order: dV' first, then dK'
                    }
                }
            }
            for (int i = tid; i < H * MP + kE; i += kThreads) {
                float v = d0[2 * H * MP * kE + i];
                for (int c = 1; c < a.nchunk; ++c) v += d0[(int64_t)c * a.dtab_sz + 2 * H * MP * kE + i];
                if (i < H * MP) {
                    const int h = i / MP, j = i - h * MP;
                    dKp[h * TILE + j * TS + 32] = v;
                } else {
                    dbo_s[i - H * MP] = v;
                }
            }
        }
        __syncthreads();
        // ---- phase 1: dK = scale (dK'_h Wq_h^T + dc_h bq_h), dV = dV'_h Wo_h ; warp = (K | V, head) ---------------------------------
        if ((warp & 1) < H) {
            const int q = warp >> 1, h = warp & 1;
            const float* src = (q ? dVp : dKp) + h * TILE;
            const float* Bh = q ? Wo_h : WqT_h;
            const float* Bl = q ? Wo_l : WqT_l;
            float* dst = q ? dVs : dKs;
#pragma unroll 1
            for (int mt = 0; 16 * mt < MP; ++mt) {
                uint32_t ah[4][4], al[4][4];
                amma::load_rows_a(src + 16 * mt * TS, lane, ah, al);
                const int r0 = 16 * mt + g, r1 = r0 + 8;
#pragma unroll
                for (int nn = 0; nn < 2; ++nn) {
                    float c[4] = {0.f, 0.f, 0.f, 0.f};
                    const int col = hd * h + 8 * nn + g;
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) mma3(c, ah[ks], al[ks], Bh, Bl, (8 * ks + t) * LQ + col, (8 * ks + t + 4) * LQ + col);
                    const int f = hd * h + 8 * nn + 2 * t;
                    if (q == 0) {
                        const float dc0 = src[r0 * TS + 32], dc1 = src[r1 * TS + 32];
                        c[0] = scale * fmaf(dc0, bq[f], c[0]); c[1] = scale * fmaf(dc0, bq[f + 1], c[1]);
                        c[2] = scale * fmaf(dc1, bq[f], c[2]); c[3] = scale * fmaf(dc1, bq[f + 1], c[3]);
                    }
                    *reinterpret_cast<float2*>(dst + r0 * TS + f) = make_float2(c[0], c[1]);
                    *reinterpret_cast<float2*>(dst + r1 * TS + f) = make_float2(c[2], c[3]);
                }
            }
        }
        __syncthreads();
        // ---- phase 2: the warp's parameter gradient (contraction over the tokens), its bias vector, a quarter of dA -----------------
        if (warp == 0) {                      // dWq[f][e] += sum_j K[j][f] dK'_h(f)[j][e]
            outer_acc<0, 1, 0, 4>(Ks, 0, dKp, 0, ksteps, g, t, accP);
            if (H > 1) outer_acc<1, 2, 0, 4>(Ks, hd, dKp + TILE, 0, ksteps, g, t, accP);
            float v = 0.f;
            const float* dcol = dKp + (lane / hd) * TILE + 32;
            for (int j = 0; j < M; ++j) v = fmaf(Ks[j * TS + lane], dcol[j * TS], v);
            sbias += v;
        } else if (warp == 1) {               // dWo[f'][e] += sum_j dV'_h(e)[j][f'] V[j][e]
            outer_acc<0, 2, 0, 2>(dVp, 0, Vs, 0, ksteps, g, t, accP);
            if (H > 1) outer_acc<0, 2, 2, 4>(dVp + TILE, 0, Vs, hd, ksteps, g, t, accP);
            sbias += dbo_s[lane];
        } else {                              // dWk / dWv [f][k] += sum_j d{K,V}[j][f] A[j][k]
            const float* dsrc = warp == 2 ? dKs : dVs;
            outer_acc<0, 2, 0, 4>(dsrc, 0, As, 0, ksteps, g, t, accP);
            float v = 0.f;
            for (int j = 0; j < M; ++j) v += dsrc[j * TS + lane];
            sbias += v;
        }
        {   // token gradient columns 8 warp .. 8 warp + 7: dA = dK Wk + dV Wv
#pragma unroll 1
            for (int mt = 0; 16 * mt < M; ++mt) {
                float c[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int s2 = 0; s2 < 2; ++s2) {
                    uint32_t ah[4][4], al[4][4];
                    amma::load_rows_a((s2 ? dVs : dKs) + 16 * mt * TS, lane, ah, al);
                    const float* Bh = s2 ? Wv_h : Wk_h;
                    const float* Bl = s2 ? Wv_l : Wk_l;
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) mma3(c, ah[ks], al[ks], Bh, Bl, (8 * ks + t) * LQ + 8 * warp + g, (8 * ks + t + 4) * LQ + 8 * warp + g);
                }
                const int r0 = 16 * mt + g, r1 = r0 + 8;
                float* da = a.da + (int64_t)b * M * kE + 8 * warp + 2 * t;
                if (r0 < M) *reinterpret_cast<float2*>(da + r0 * kE) = make_float2(c[0], c[1]);
                if (r1 < M) *reinterpret_cast<float2*>(da + r1 * kE) = make_float2(c[2], c[3]);
            }
        }
    }
    // ---- the CTA's partial row: [dWq ; dWk ; dWv | dbq dbk dbv | dWo | dbo] -----------------------------------------------------------
    float* accg = a.partials + (int64_t)blockIdx.x * a.P;
    const int oBin = 3 * kE * kE, oWo = oBin + 3 * kE, oBo = oWo + kE * kE;
    const int ofs = warp == 0 ? 0 : (warp == 1 ? oWo : (warp == 2 ? kE * kE : 2 * kE * kE));
    const float sc = warp == 0 ? scale : 1.f;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int n = 0; n < 4; ++n) {
            float* d = accg + ofs + (16 * mt + g) * kE + 8 * n + 2 * t;
            *reinterpret_cast<float2*>(d) = make_float2(sc * accP[mt][n][0], sc * accP[mt][n][1]);
            *reinterpret_cast<float2*>(d + 8 * kE) = make_float2(sc * accP[mt][n][2], sc * accP[mt][n][3]);
        }
    const int bofs = warp == 0 ? oBin : (warp == 1 ? oBo : (warp == 2 ? oBin + kE : oBin + 2 * kE));
    accg[bofs + lane] = sc * sbias;
}

}  // namespace amma3
}  // namespace igcn
