// Table-driven cross attention on the tensor cores (second version of cross_attn_mma.cuh's backward).
//
// ncu --set full on attn_mma_bwd_kernel at the benchmarked size (profiles/r2_ncu_attn_mma_c2.json, B=512 stacked graphs, R=90): 113 us,
// 54 K warp instructions per graph, of which the MMA row phase is 28 % -- the rest is per-graph serial work inside the one CTA that
// owns the graph: flushing / summing the warps' partial dK'/dV' fragments through shared memory (22 %), recomputing the K'/V'
// tables (19 % of samples), the M x 32 chain back to the parameters with read-modify-writes of a global accumulator row (14 %),
// the weight prologue, and one 288-thread CTA per SM with nothing to overlap its barriers.  The split here:
//
//   attn_tables_kernel  (forward)   K, V, K'_h, V'_h, c_h of every graph -> a global record per graph; the forward row kernel and
//                                   the backward read it, nobody recomputes it and the row kernels need no weights at all.
//   attn_bwd2_kernel    (backward)  one CTA per (graph, chunk of <= 6 row tiles), >= 3 CTAs per SM.  Per head: phase 1, one warp per
//                                   16-row tile: S, P -> dP -> dS -> dX (accumulated over the heads in registers, stored once);
//                                   P and dS go to shared memory as (rows x tokens); phase 2, one warp per (quantity, token tile):
//                                   dV'^T = dY^T P and dK'^T = X^T dS contract over ALL rows of the chunk inside one warp, so there
//                                   is no cross-warp reduction -- the results go straight to the chunk's gradient record.
//   attn_chain_kernel   (backward)  sums the chunk records of a graph and runs the M x 32 chain (dK, dV, token gradient, the six
//                                   parameter gradients) one graph per pass with the parameter gradients in registers over all
//                                   passes: one partial row per CTA, reduced in a fixed order afterwards.
// Everything is deterministic (fixed summation orders, no atomics).  Same algebra as cross_attn_mma.cuh; reference:
// nn.MultiheadAttention(E, 2, batch_first=True)(q, kv, kv) + relu (+ the fusion average), kernel/sgcn_img_snp.py:46,239-242.
#pragma once
#include "cross_attn_mma.cuh"

namespace igcn {
namespace amma2 {

using namespace igcn::mmau;
using amma::kE;
using amma::TS;

// per-graph record (floats): K' (H, MP, TS) | V' (H, MP, TS) | c (H, MP) | K (M, 32) | V (M, 32)
__host__ __device__ inline int tab_head(int MP, int H) { return 2 * H * MP * TS + H * MP; }
__host__ __device__ inline int tab_floats(int M, int MP, int H) { return tab_head(MP, H) + 2 * M * kE; }
// per-(graph, chunk) gradient record: dV' (H, MP, 32) | dK' (H, MP, 32) | dc (H, MP) | dbo (32)
__host__ __device__ inline int dtab_floats(int MP, int H) { return 2 * H * MP * kE + H * MP + kE; }

constexpr int kTabThreads = 256, kTabGraphs = 1;

// dst[c][r] (row stride `rows`) = src[r][c] for a (rows x 32) row-major global matrix, through a padded tile: coalesced global reads,
// conflict-free shared stores and loads (a direct transposing store is a 32-way bank conflict: 4 096 wavefronts per CTA prologue
// in profiles/r2_ncu_attn_mma_c2.json).  tmp: rows * 33 floats.  The caller's next barrier publishes dst.
__device__ __forceinline__ void stage_transposed(const float* __restrict__ src, int rows, float* dst, float* tmp) {
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int i = tid; i < rows * kE; i += nt) tmp[(i >> 5) * 33 + (i & 31)] = src[i];
    __syncthreads();
    for (int o = tid; o < rows * kE; o += nt) {
        const int c = o / rows, r = o - c * rows;
        dst[o] = tmp[r * 33 + c];
    }
    __syncthreads();
}

__global__ void __launch_bounds__(kTabThreads, 4) attn_tables_kernel(AttnArgs a, int per_sz, int MP) {
    IGCN_PDL_SYNC();
    extern __shared__ __align__(16) float smf[];
    const int tid = threadIdx.x, nt = blockDim.x, M = a.M, H = a.heads;
    float* WkvT = smf;
    float* Wq = WkvT + 2 * kE * kE;
    float* WoT = Wq + kE * kE;
    float* bin = WoT + kE * kE;
    float* per = bin + 4 * kE;
    // WkvT[k][f] (f < 2E) = Win[E + f][k] ; WoT[k][f] = Wo[f][k] ; Wq row-major
    stage_transposed(a.Win + kE * kE, 2 * kE, WkvT, per);
    stage_transposed(a.Wo, kE, WoT, per);
    for (int i = tid; i < kE * kE; i += nt) Wq[i] = a.Win[i];
    for (int i = tid; i < 3 * kE; i += nt) bin[i] = a.bin[i];
    __syncthreads();
    const int nhead4 = tab_head(MP, H) >> 2, nkv4 = (2 * M * kE) >> 2;
    for (int b0 = blockIdx.x * kTabGraphs; b0 < a.B; b0 += gridDim.x * kTabGraphs) {
        const int ng = min(kTabGraphs, a.B - b0);
        amma::graph_tables(a, b0, ng, per, per_sz, MP, WkvT, Wq, WoT, bin);           // ends with a barrier
        for (int idx = tid; idx < ng * (nhead4 + nkv4); idx += nt) {
            const int gl = idx / (nhead4 + nkv4), i = idx - gl * (nhead4 + nkv4);
            const float* src = per + gl * per_sz;
            float* dst = a.tab_out + (int64_t)(b0 + gl) * a.tab_sz;
            // `per` holds A | K | V | K' | V' | c : K' V' c are contiguous from 3 M E, K V from M E
            if (i < nhead4) st4s(dst + 4 * i, ld4s(src + 3 * M * kE + 4 * i));
            else st4s(dst + 4 * i, ld4s(src + M * kE + 4 * (i - nhead4)));
        }
        __syncthreads();
    }
}

struct GeoB2 {
    int MP, TPI, nchunk, PS, rows_pad, nthreads;
    size_t smem;
};
// row tiles per item: 6 (least table / record traffic per row) when that still gives every SM many items, else 3 -- at the
// benchmarked size (512 graphs of 6 tiles) whole-graph items are 3.5 per SM, i.e. a 4 : 3 imbalance between SMs
static GeoB2 bwd2_geo(int R, int M, int H, int64_t B) {
    GeoB2 g;
    g.MP = (M + 7) & ~7;
    const int T = (R + 15) >> 4;
    const int tmax = (B * ((T + 5) / 6) >= (int64_t)8 * sm_count()) ? 6 : 3;
    g.nchunk = (T + tmax - 1) / tmax;
    g.TPI = (T + g.nchunk - 1) / g.nchunk;
    static const int force = [] { const char* e = getenv("IGCN_ATTN_TPI"); return e ? atoi(e) : 0; }();
    if (force >= 1 && force <= 6) {                          // experiment switch: fixed tiles per item (the last item of a graph may be short)
        g.TPI = force < T ? force : T;
        g.nchunk = (T + g.TPI - 1) / g.TPI;
    }
    g.PS = g.MP == 8 ? 8 : (g.MP == 32 ? 40 : 24);          // row stride of the P / dS buffers: conflict-free as an MMA B operand
    g.rows_pad = 16 * g.TPI;
    g.nthreads = 32 * g.TPI;
    g.smem = (size_t)4 * (tab_head(g.MP, H) + 2 * g.rows_pad * TS + 2 * g.rows_pad * g.PS + g.TPI * g.MP + g.TPI * kE) + 16;
    return g;
}

template <int NT>
__global__ void __launch_bounds__(192, 3) attn_bwd2_kernel(AttnArgs a, GeoB2 geo) {
    IGCN_PDL_SYNC();
    extern __shared__ __align__(16) float smf[];
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarp = nt >> 5;
    const int gq = lane >> 2, tq = lane & 3;
    const int R = a.R, M = a.M, H = a.heads, MP = geo.MP, PS = geo.PS;
    const int item = blockIdx.x, b = item / geo.nchunk, ch = item - b * geo.nchunk;
    const int row0 = ch * geo.rows_pad, nrows = min(R - row0, geo.rows_pad), ntile = (nrows + 15) >> 4;
    float* Kp = smf;
    float* Vp = Kp + H * MP * TS;
    float* cb = Vp + H * MP * TS;
    float* Xs = cb + H * MP;
    float* Ys = Xs + geo.rows_pad * TS;
    float* Pb = Ys + geo.rows_pad * TS;
    float* Sb = Pb + geo.rows_pad * PS;
    float* dcw = Sb + geo.rows_pad * PS;
    float* dbw = dcw + geo.TPI * MP;
    {
        const float* tb = a.tab + (int64_t)b * a.tab_sz;
        const int nhead4 = tab_head(MP, H) >> 2;
        for (int i0 = 0; i0 < nhead4; i0 += 5 * nt) {            // five loads per thread in flight (a load -> store loop is one L2 trip per iteration)
            float4 v[5];
#pragma unroll
            for (int u = 0; u < 5; ++u)
                if (i0 + tid + u * nt < nhead4) v[u] = ld4s(tb + 4 * (i0 + tid + u * nt));
#pragma unroll
            for (int u = 0; u < 5; ++u)
                if (i0 + tid + u * nt < nhead4) st4s(smf + 4 * (i0 + tid + u * nt), v[u]);
        }
    }
    const int64_t gofs = ((int64_t)b * R + row0) * kE;
    const float* xg = a.x + gofs;
    const float* gg = a.gy + gofs;
    const float* yg = a.yout + gofs;
    // 4 float4 groups per thread (16 rows x 8 groups per tile, 32 threads per tile): all of a thread's loads are issued before the
    // first use -- one L2 round trip per item instead of four
    {
        float4 xv[4], gv[4], yv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int idx = tid + u * nt;
            xv[u] = gv[u] = yv[u] = make_float4(0.f, 0.f, 0.f, 0.f);          // rows past the graph's last row: zeros (they add nothing)
            if (idx < ntile * 16 * 8 && (idx >> 3) < nrows) {
                xv[u] = ld4s(xg + (int64_t)idx * 4);
                gv[u] = ld4s(gg + (int64_t)idx * 4);
                if (a.relu) yv[u] = ld4s(yg + (int64_t)idx * 4);
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int idx = tid + u * nt;
            if (idx >= ntile * 16 * 8) continue;
            const int row = idx >> 3, c4 = idx & 7;
            float4 g = gv[u];
            if (a.relu && row < nrows) {
                float4 y = yv[u];
                if (a.mix) {        // saved output is (x + relu(y)) / 2: relu(y) = 2 out - x ; the attention branch sees half the gradient
                    y.x = 2.f * y.x - xv[u].x; y.y = 2.f * y.y - xv[u].y; y.z = 2.f * y.z - xv[u].z; y.w = 2.f * y.w - xv[u].w;
                    g.x *= 0.5f; g.y *= 0.5f; g.z *= 0.5f; g.w *= 0.5f;
                }
                if (!(y.x > 0.f)) g.x = 0.f;
                if (!(y.y > 0.f)) g.y = 0.f;
                if (!(y.z > 0.f)) g.z = 0.f;
                if (!(y.w > 0.f)) g.w = 0.f;
            }
            st4s(Xs + row * TS + 4 * c4, xv[u]);
            st4s(Ys + row * TS + 4 * c4, g);
        }
    }
    __syncthreads();
    float* drec = a.dtab + (int64_t)item * a.dtab_sz;
    const bool has_tile = warp < ntile;
    const float* xt = Xs + 16 * warp * TS;
    const float* yt = Ys + 16 * warp * TS;
    float dx[4][4], dbp[8];
#pragma unroll
    for (int n2 = 0; n2 < 4; ++n2) dx[n2][0] = dx[n2][1] = dx[n2][2] = dx[n2][3] = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) dbp[i] = 0.f;
#pragma unroll 1
    for (int h = 0; h < H; ++h) {
        const float* Kh = Kp + h * MP * TS;
        const float* Vh = Vp + h * MP * TS;
        // ---- phase 1: the warp's 16-row tile ------------------------------------------------------------------------------------
        if (has_tile) {
            float p[NT][4], ds[NT][4];
            {
                uint32_t xh[4][4], xl[4][4];
                amma::load_rows_a(xt, lane, xh, xl);
                amma::tile_softmax<NT>(xh, xl, Kh, cb + h * MP, M, gq, tq, p);
            }
            {
                uint32_t gh[4][4], gl[4][4];
                amma::load_rows_a(yt, lane, gh, gl);
                if (h == 0) {
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {          // hi + lo is the exact value: column sums of dY for d bo
                        dbp[2 * ks] += (__uint_as_float(gh[ks][0]) + __uint_as_float(gl[ks][0])) + (__uint_as_float(gh[ks][1]) + __uint_as_float(gl[ks][1]));
                        dbp[2 * ks + 1] += (__uint_as_float(gh[ks][2]) + __uint_as_float(gl[ks][2])) + (__uint_as_float(gh[ks][3]) + __uint_as_float(gl[ks][3]));
                    }
                }
#pragma unroll
                for (int n = 0; n < NT; ++n) {                // dP = dY V'^T
                    ds[n][0] = ds[n][1] = ds[n][2] = ds[n][3] = 0.f;
                    const float* vr = Vh + (8 * n + gq) * TS + tq;
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
            float* pr = Pb + (16 * warp + gq) * PS + 2 * tq;
            float* sr = Sb + (16 * warp + gq) * PS + 2 * tq;
#pragma unroll
            for (int n = 0; n < NT; ++n) {
                ds[n][0] = p[n][0] * (ds[n][0] - rd0);
                ds[n][1] = p[n][1] * (ds[n][1] - rd0);
                ds[n][2] = p[n][2] * (ds[n][2] - rd1);
                ds[n][3] = p[n][3] * (ds[n][3] - rd1);
                *reinterpret_cast<float2*>(pr + 8 * n) = make_float2(p[n][0], p[n][1]);
                *reinterpret_cast<float2*>(pr + 8 * PS + 8 * n) = make_float2(p[n][2], p[n][3]);
                *reinterpret_cast<float2*>(sr + 8 * n) = make_float2(ds[n][0], ds[n][1]);
                *reinterpret_cast<float2*>(sr + 8 * PS + 8 * n) = make_float2(ds[n][2], ds[n][3]);
#pragma unroll
                for (int b2 = 0; b2 < 2; ++b2) {              // dc: column sums of dS over the tile's 16 rows
                    float v = ds[n][b2] + ds[n][b2 + 2];
                    v += __shfl_xor_sync(0xffffffffu, v, 4);
                    v += __shfl_xor_sync(0xffffffffu, v, 8);
                    v += __shfl_xor_sync(0xffffffffu, v, 16);
                    if (gq == 0) dcw[warp * MP + 8 * n + 2 * tq + b2] = v;
                }
            }
            amma::frag_times_table<NT>(ds, Kh, gq, tq, dx);   // dX += dS K'
        }
        __syncthreads();
        // ---- phase 2: dV'^T (32 x tokens) = dY^T P, dK'^T = X^T dS, contraction over all rows of the chunk inside one warp --------
        for (int u = warp; u < 2 * NT; u += nwarp) {
            const int q = u / NT, n = u - q * NT;
            const float* rowsA = q ? Xs : Ys;
            const float* F = (q ? Sb : Pb) + 8 * n + gq;
            float acc[2][4];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) acc[mt][0] = acc[mt][1] = acc[mt][2] = acc[mt][3] = 0.f;
            for (int ks = 0; ks < 2 * ntile; ++ks) {
                uint32_t bh0, bl0, bh1, bl1;
                split(F[(8 * ks + tq) * PS], bh0, bl0);
                split(F[(8 * ks + tq + 4) * PS], bh1, bl1);
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) {
                    const float* ra = rowsA + (8 * ks + tq) * TS + 16 * mt + gq;
                    uint32_t ah[4], al[4];
                    split(ra[0], ah[0], al[0]);
                    split(ra[8], ah[1], al[1]);
                    split(ra[4 * TS], ah[2], al[2]);
                    split(ra[4 * TS + 8], ah[3], al[3]);
                    mma_k8(acc[mt], al[0], al[1], al[2], al[3], bh0, bh1);
                    mma_k8(acc[mt], ah[0], ah[1], ah[2], ah[3], bl0, bl1);
                    mma_k8(acc[mt], ah[0], ah[1], ah[2], ah[3], bh0, bh1);
                }
            }
            float* dst = drec + (q ? H * MP * kE : 0) + (h * MP + 8 * n + 2 * tq) * kE + gq;     // [token][feature]
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                dst[16 * mt] = acc[mt][0];                    // token 2t,     feature 16 mt + g
                dst[kE + 16 * mt] = acc[mt][1];               // token 2t + 1
                dst[16 * mt + 8] = acc[mt][2];                // token 2t,     feature 16 mt + g + 8
                dst[kE + 16 * mt + 8] = acc[mt][3];
            }
        }
        if (tid < MP) {                                       // dc of this head: the tiles' partial sums in tile order
            float v = 0.f;
            for (int w = 0; w < ntile; ++w) v += dcw[w * MP + tid];
            drec[2 * H * MP * kE + h * MP + tid] = v;
        }
        __syncthreads();
    }
    if (has_tile) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float v = dbp[i];
            v += __shfl_xor_sync(0xffffffffu, v, 4);
            v += __shfl_xor_sync(0xffffffffu, v, 8);
            v += __shfl_xor_sync(0xffffffffu, v, 16);
            if (gq == 0) dbw[warp * kE + 8 * (i >> 1) + tq + 4 * (i & 1)] = v;
        }
        const int r0 = 16 * warp + gq, r1 = r0 + 8;
        float* dxg = a.dx + gofs;
#pragma unroll
        for (int n2 = 0; n2 < 4; ++n2) {
            if (r0 < nrows) {
                float2 v = make_float2(dx[n2][0], dx[n2][1]);
                if (a.mix) {                                  // the direct half of the average: d out / d x = 1/2
                    const float2 o = *reinterpret_cast<const float2*>(gg + r0 * kE + 8 * n2 + 2 * tq);
                    v.x = fmaf(0.5f, o.x, v.x); v.y = fmaf(0.5f, o.y, v.y);
                }
                *reinterpret_cast<float2*>(dxg + r0 * kE + 8 * n2 + 2 * tq) = v;
            }
            if (r1 < nrows) {
                float2 v = make_float2(dx[n2][2], dx[n2][3]);
                if (a.mix) {
                    const float2 o = *reinterpret_cast<const float2*>(gg + r1 * kE + 8 * n2 + 2 * tq);
                    v.x = fmaf(0.5f, o.x, v.x); v.y = fmaf(0.5f, o.y, v.y);
                }
                *reinterpret_cast<float2*>(dxg + r1 * kE + 8 * n2 + 2 * tq) = v;
            }
        }
    }
    __syncthreads();
    if (tid < kE) {
        float v = 0.f;
        for (int w = 0; w < ntile; ++w) v += dbw[w * kE + tid];
        drec[2 * H * MP * kE + H * MP + tid] = v;
    }
}

// ---- chain: chunk records -> dK, dV -> token gradient + parameter gradients ----------------------------------------------------------
constexpr int kChainThreads = 256, kChainGraphs = 1;
// per-graph shared tables (floats): A | K | V (M x 32) | dK' (H x MP x TS, column 32 = dc) | dV' (H x MP x TS) | dbo (32) | dK | dV (M x 32)
__host__ __device__ inline int chain_per_sz(int M, int MP, int H) { return 5 * M * kE + 2 * H * MP * TS + kE; }
static size_t chain_smem(int M, int MP, int H) { return (size_t)4 * (4 * kE * kE + kE + kChainGraphs * chain_per_sz(M, MP, H)) + 16; }

__global__ void __launch_bounds__(kChainThreads, 4) attn_chain_kernel(AttnArgs a, int MP) {
    IGCN_PDL_SYNC();
    extern __shared__ __align__(16) float smf[];
    const int tid = threadIdx.x, M = a.M, H = a.heads, hd = kE / H;
    const float scale = rsqrtf((float)hd);
    float* WqT = smf;                         // WqT[e][f] = Wq[f][e]
    float* WoO = WqT + kE * kE;               // Wo[e][f]
    float* WkvO = WoO + kE * kE;              // [Wk ; Wv] [f][k]
    float* bq = WkvO + 2 * kE * kE;
    float* per0 = bq + kE;
    const int per_sz = chain_per_sz(M, MP, H);
    const int oK = M * kE, oV = 2 * M * kE, oDKp = 3 * M * kE, oDVp = oDKp + H * MP * TS, oDBO = oDVp + H * MP * TS, oDK = oDBO + kE,
              oDV = oDK + M * kE;
    stage_transposed(a.Win, kE, WqT, per0);
    for (int i = tid; i < kE * kE; i += kChainThreads) WoO[i] = a.Wo[i];
    for (int i = tid; i < 2 * kE * kE; i += kChainThreads) WkvO[i] = a.Win[kE * kE + i];
    if (tid < kE) bq[tid] = a.bin[tid];
    float4 accQ = make_float4(0.f, 0.f, 0.f, 0.f), accO = accQ, accK = accQ, accV = accQ;
    float s_bq = 0.f, s_bo = 0.f, s_bkv = 0.f;
    const int nq4 = (H * MP * kE) >> 2;                      // float4 groups of dV' (and of dK') in a chunk record
    const int head = tab_head(MP, H);
    for (int b0 = blockIdx.x * kChainGraphs; b0 < a.B; b0 += gridDim.x * kChainGraphs) {
        const int ng = min(kChainGraphs, a.B - b0);
        __syncthreads();                                     // previous pass done with the tables (and the weights are staged)
        for (int gl = 0; gl < ng; ++gl) {
            float* per = per0 + gl * per_sz;
            const int b = b0 + gl;
            const float* ab = a.a + (int64_t)b * M * kE;
            const float* tb = a.tab + (int64_t)b * a.tab_sz + head;
            for (int i = tid; i < (M * kE) >> 2; i += kChainThreads) st4s(per + 4 * i, ld4s(ab + 4 * i));
            for (int i = tid; i < (2 * M * kE) >> 2; i += kChainThreads) st4s(per + oK + 4 * i, ld4s(tb + 4 * i));
            const float* d0 = a.dtab + (int64_t)b * a.nchunk * a.dtab_sz;
            for (int i = tid; i < 2 * nq4; i += kChainThreads) {          // chunk records summed in chunk order
                float4 v = ld4s(d0 + 4 * i);
                for (int c = 1; c < a.nchunk; ++c) {
                    const float4 w = ld4s(d0 + (int64_t)c * a.dtab_sz + 4 * i);
                    v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
                }
                const int q = i >= nq4, r = q ? i - nq4 : i, hj = r >> 3, c4 = r & 7;        // record order: dV' first, then dK'
                st4s(per + (q ? oDKp : oDVp) + hj * TS + 4 * c4, v);
            }
            for (int i = tid; i < H * MP + kE; i += kChainThreads) {
                float v = d0[2 * H * MP * kE + i];
                for (int c = 1; c < a.nchunk; ++c) v += d0[(int64_t)c * a.dtab_sz + 2 * H * MP * kE + i];
                if (i < H * MP) per[oDKp + i * TS + 32] = v;
                else per[oDBO + i - H * MP] = v;
            }
        }
        __syncthreads();
        // dK[j][f] = scale (<dK'_h[j], Wq[f]> + dc_h[j] bq[f]) ; dV[j][f] = <dV'_h[j], Wo[:, f]>     (f = h hd + d)
        for (int idx = tid; idx < ng * M * 16; idx += kChainThreads) {
            const int gl = idx / (M * 16), r = idx - gl * M * 16, j = r >> 4, c = r & 15, fq = c & 7;
            float* per = per0 + gl * per_sz;
            const bool isK = c < 8;
            const int h = (4 * fq) / hd;
            const float* src = per + (isK ? oDKp : oDVp) + (h * MP + j) * TS;
            const float* wm = (isK ? WqT : WoO) + 4 * fq;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
            for (int e = 0; e < kE; ++e) fma4(src[e], ld4s(wm + e * kE), v);
            if (isK) {
                const float dc = src[32];
                const float4 bq4 = ld4s(bq + 4 * fq);
                v.x = scale * (v.x + dc * bq4.x); v.y = scale * (v.y + dc * bq4.y);
                v.z = scale * (v.z + dc * bq4.z); v.w = scale * (v.w + dc * bq4.w);
            }
            st4s(per + (isK ? oDK : oDV) + j * kE + 4 * fq, v);
        }
        __syncthreads();
        const int f = tid >> 3, q = tid & 7;
        for (int gl = 0; gl < ng; ++gl) {
            const float* per = per0 + gl * per_sz;
            {   // dWq[f][4q..] += sum_j K[j][f] dK'_h[j][4q..]  (h = head of f; scaled at the end)
                const int h = f / hd;
                const float* sc1 = per + oK + f;
                const float* vec = per + oDKp + h * MP * TS + 4 * q;
#pragma unroll 4
                for (int j = 0; j < M; ++j) fma4(sc1[j * kE], ld4s(vec + j * TS), accQ);
            }
            {   // dWo[f][4q..] += sum_j dV'_h[j][f] V[j][4q..]  (h = head of the quad)
                const int h = (4 * q) / hd;
                const float* sc1 = per + oDVp + h * MP * TS + f;
                const float* vec = per + oV + 4 * q;
#pragma unroll 4
                for (int j = 0; j < M; ++j) fma4(sc1[j * TS], ld4s(vec + j * kE), accO);
            }
            {   // dWk[f][4q..] += sum_j dK[j][f] A[j][4q..] ; dWv likewise
                const float* vec = per + 4 * q;
#pragma unroll 4
                for (int j = 0; j < M; ++j) {
                    const float4 av = ld4s(vec + j * kE);
                    fma4(per[oDK + j * kE + f], av, accK);
                    fma4(per[oDV + j * kE + f], av, accV);
                }
            }
            if (tid < kE) {                                  // dbq, dbo
                const int h = tid / hd;
                float v = 0.f;
#pragma unroll 4
                for (int j = 0; j < M; ++j) v = fmaf(per[oK + j * kE + tid], per[oDKp + (h * MP + j) * TS + 32], v);
                s_bq += v;
                s_bo += per[oDBO + tid];
            }
            if (tid < 2 * kE) {                              // dbk | dbv
                const int o = (tid < kE) ? oDK + tid : oDV + (tid - kE);
                float v = 0.f;
                for (int j = 0; j < M; ++j) v += per[o + j * kE];
                s_bkv += v;
            }
        }
        // token gradient dA[j][4kq..] = <dK[j], Wk[:, 4kq..]> + <dV[j], Wv[:, 4kq..]>
        for (int idx = tid; idx < ng * M * 8; idx += kChainThreads) {
            const int gl = idx / (M * 8), r = idx - gl * M * 8, j = r >> 3, kq = r & 7;
            const float* per = per0 + gl * per_sz;
            const float* dk = per + oDK + j * kE;
            const float* dv = per + oDV + j * kE;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
            for (int e = 0; e < kE; ++e) {
                fma4(dk[e], ld4s(WkvO + e * kE + 4 * kq), v);
                fma4(dv[e], ld4s(WkvO + (kE + e) * kE + 4 * kq), v);
            }
            *reinterpret_cast<float4*>(a.da + ((int64_t)(b0 + gl) * M + j) * kE + 4 * kq) = v;
        }
    }
    // ---- the CTA's partial row: [dWq ; dWk ; dWv | dbq dbk dbv | dWo | dbo] -------------------------------------------------------
    float* accg = a.partials + (int64_t)blockIdx.x * a.P;
    const int oBin = 3 * kE * kE, oWo = oBin + 3 * kE, oBo = oWo + kE * kE;
    const int f = tid >> 3, q = tid & 7;
    accQ.x *= scale; accQ.y *= scale; accQ.z *= scale; accQ.w *= scale;
    st4s(accg + f * kE + 4 * q, accQ);
    st4s(accg + kE * kE + f * kE + 4 * q, accK);
    st4s(accg + 2 * kE * kE + f * kE + 4 * q, accV);
    st4s(accg + oWo + f * kE + 4 * q, accO);
    if (tid < kE) {
        accg[oBin + tid] = scale * s_bq;
        accg[oBo + tid] = s_bo;
    }
    if (tid < 2 * kE) accg[oBin + kE + tid] = s_bkv;
}

}  // namespace amma2
}  // namespace igcn
